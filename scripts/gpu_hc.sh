mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_classes.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -60 > gpurun_out/hc_tests.log
tail -8 gpurun_out/hc_tests.log
