# round 2, call A: GPU test suite (incl. the config-size parity tests), pair-kernel floors, N3L variant, short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r02a_tests.log
tail -3 gpurun_out/r02a_tests.log
timeout 300 ./lammps-ucg-dev_b200/ucg_microbench > gpurun_out/r02a_microbench.jsonl 2> gpurun_out/r02a_microbench.err
cat gpurun_out/r02a_microbench.jsonl | cut -c1-120
VARIANTS='[{"N3L":0},{"N3L":1},{"N3L":0}]' timeout 300 python scripts/tune_pair.py > gpurun_out/r02a_n3l.log 2>&1
tail -4 gpurun_out/r02a_n3l.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.json
