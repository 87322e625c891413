mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_classes.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -30 > gpurun_out/hc_tests.log
grep -n "AssertionError\|passed\|failed\|Error" gpurun_out/hc_tests.log | head
timeout 600 python scripts/time_host_classes.py 2>&1 | tail -2
