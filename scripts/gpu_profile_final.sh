mkdir -p gpurun_out
python scripts/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu1.log 2>&1
tail -1 gpurun_out/plain.log
STEPS=2 python scripts/profile_step.py > gpurun_out/plain2.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast|k_step_tail" -s 2 -c 3 -o gpurun_out/prof_pair python scripts/profile_step.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
