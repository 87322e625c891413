mkdir -p gpurun_out
python scripts/profile_build.py > gpurun_out/plain_build.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_build.csv python scripts/profile_build.py > gpurun_out/ncu_build.log 2>&1
cat gpurun_out/plain_build.log | tail -5
