#!/bin/bash
# session 3: full GPU suite, secondary-style timing, contract bench line (no CPU leg)
python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/s3_tests.log
python scripts/time_styles.py > gpurun_out/s3_styles.json 2> gpurun_out/s3_styles.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
tail -n 3 gpurun_out/s3_tests.log; cat gpurun_out/s3_styles.json; python -c "
import json; d=json.load(open('gpurun_out/s3_bench.json')); print(d['ms_per_step'], d['e2e'])"
