mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dump.py tests/test_gpu_host_classes.py tests/test_gpu_multibrick.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -40 > gpurun_out/dump_tests.log
tail -5 gpurun_out/dump_tests.log
timeout 600 python scripts/time_dump.py > gpurun_out/dump_timing.log 2>&1; tail -1 gpurun_out/dump_timing.log
