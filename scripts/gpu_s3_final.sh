#!/bin/bash
# session 3, final state: the whole GPU suite, smoke(), the contract bench line (with the CPU leg) and the reference arm's serial figure
python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/s3f_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s3f_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/s3f_bench.json 2> gpurun_out/s3f_bench.err
tail -n 3 gpurun_out/s3f_tests.log; tail -n 2 gpurun_out/s3f_smoke.log; python -c "
import json
d=json.load(open('gpurun_out/s3f_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"
