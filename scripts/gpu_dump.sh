mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dump.py -m gpu -x -q 2>&1 | grep -v "^E    *+" | tail -25 > gpurun_out/dump_tests.log
tail -5 gpurun_out/dump_tests.log
timeout 600 python scripts/time_dump.py > gpurun_out/dump_timing.log 2>&1; tail -3 gpurun_out/dump_timing.log
