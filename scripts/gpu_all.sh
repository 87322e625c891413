mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^E    *+" | tail -15 > gpurun_out/all_tests.log
tail -4 gpurun_out/all_tests.log
timeout 600 python scripts/time_host_classes.py 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_new.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_new.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'pair_ms',d['roofline']['kernel_ms'],d['roofline']['stage_ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
PY
