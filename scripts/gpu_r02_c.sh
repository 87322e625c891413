# round 2, call C: FP32-prefilter neighbor build: identical rows + timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sizes.py tests/test_gpu_edge_cases.py tests/test_gpu_multibrick.py -m gpu -q --tb=short -k "neighbor or config_size or edge or brick" 2>&1 | tail -30 > gpurun_out/r02c_tests.log
tail -5 gpurun_out/r02c_tests.log
for f in 1 0; do UCGB200_BUILD_F32=$f UCGB200_BUILD_TRACE=0 timeout 300 python scripts/time_build.py 2>&1 | tail -2; done > gpurun_out/r02c_build.log 2>&1
cat gpurun_out/r02c_build.log
UCGB200_BUILD_TRACE=1 timeout 300 python scripts/time_build.py 2>&1 | grep "\[build\]" | tail -12 > gpurun_out/r02c_trace.log
cat gpurun_out/r02c_trace.log
