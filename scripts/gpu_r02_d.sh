# round 2, call D: tests after the k_images fix + ncu capture of the FP32-prefilter build kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_multibrick.py -m gpu -q --tb=short -k "neighbor or edge or brick" 2>&1 | tail -30 > gpurun_out/r02d_tests.log
tail -3 gpurun_out/r02d_tests.log
python scripts/profile_build.py > gpurun_out/r02d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_build_rows_tiled_f32" -s 1 -c 2 -o gpurun_out/r02d_prof_build python scripts/profile_build.py > gpurun_out/r02d_ncu.log 2>&1
tail -3 gpurun_out/r02d_ncu.log
