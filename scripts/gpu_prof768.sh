mkdir -p gpurun_out
STEPS=2 python scripts/profile_step.py > gpurun_out/plain2.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast" -s 1 -c 1 -o gpurun_out/prof_pair768 python scripts/profile_step.py > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log
