mkdir -p gpurun_out
python bench.py --steps 100 --warmup 10 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
