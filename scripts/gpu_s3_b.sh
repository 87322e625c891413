#!/bin/bash
# session 3: split step_host (initial_integrate in two parts under the uploads), config 0 on the GPU
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "step_host or trajectory or resident" 2>&1 | tail -5 > gpurun_out/s3b_tests.log
python scripts/run_config0.py gpu > gpurun_out/s3b_config0_gpu.json 2> gpurun_out/s3b_config0_gpu.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s3b_bench.json 2> gpurun_out/s3b_bench.err
UCGB200_E2E_SPLIT=0 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s3b_bench_nosplit.json 2> gpurun_out/s3b_bench_nosplit.err
tail -n 3 gpurun_out/s3b_tests.log; cat gpurun_out/s3b_config0_gpu.json; python -c "
import json
for f in ('s3b_bench','s3b_bench_nosplit'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'])"
