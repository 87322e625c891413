mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dump.py tests/test_gpu_host_classes.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -40 > gpurun_out/dump_tests.log
tail -5 gpurun_out/dump_tests.log
timeout 600 python scripts/time_dump.py > gpurun_out/dump_timing.log 2>&1; tail -3 gpurun_out/dump_timing.log
NO_REF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_dump_launches.csv python scripts/time_dump.py > gpurun_out/dump_ncu.log 2>&1; tail -2 gpurun_out/dump_ncu.log
