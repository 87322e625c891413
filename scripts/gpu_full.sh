# GPU tests (all), then launch list and one full ncu capture of the pair kernel on the 1M-site deck
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests.log; cat gpurun_out/tests.log
python scripts/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/plain.log
STEPS=2 python scripts/profile_step.py > gpurun_out/plain2.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:k_pair_ucgld -s 1 -c 2 -o gpurun_out/prof_pair python scripts/profile_step.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
