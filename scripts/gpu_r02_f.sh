# round 2, call F: host-class tests (tracked offload mode) + deck-level timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host_classes.py -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r02f_tests.log
tail -6 gpurun_out/r02f_tests.log
timeout 900 python scripts/time_host_classes.py 2>&1 | tail -3
