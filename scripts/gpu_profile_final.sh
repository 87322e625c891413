mkdir -p gpurun_out
# launch list of bench.py itself (short run) and of one full capture of the top kernels
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/plain_bench.log
STEPS=2 python scripts/profile_step.py > gpurun_out/plain2.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast|k_step_tail" -s 2 -c 2 -o gpurun_out/prof_pair python scripts/profile_step.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
