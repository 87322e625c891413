mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_classes.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -40 > gpurun_out/hc_tests.log
grep -n "AssertionError\|passed\|failed\|Error" gpurun_out/hc_tests.log | head
