mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pair or neighbor" 2>&1 | tail -4
python scripts/tune_pair.py 2>&1 | tee gpurun_out/tune2.log | tail -30
python bench.py --steps 50 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench3.json; cat gpurun_out/bench3.json
