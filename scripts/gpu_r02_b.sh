# round 2, call B: respa tests, N3L variants (REDs / TMA bulk reductions) parity + timing at 1 M sites
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host_classes.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "respa or min_post or newton or resident or post_force_order" 2>&1 | tail -40 > gpurun_out/r02b_tests.log
tail -5 gpurun_out/r02b_tests.log
VARIANTS='[{"N3L":0},{"N3L":1},{"N3L":2},{"N3L":0},{"N3L":2}]' timeout 300 python scripts/tune_pair.py > gpurun_out/r02b_n3l.log 2>&1
tail -6 gpurun_out/r02b_n3l.log
