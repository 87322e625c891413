mkdir -p gpurun_out
REPS=2 python scripts/profile_build.py > gpurun_out/plain_build.log 2>&1 && \
REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_build_rows -s 1 -c 1 -o gpurun_out/prof_build python scripts/profile_build.py > gpurun_out/ncu_build2.log 2>&1
tail -2 gpurun_out/ncu_build2.log
