mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^E    *+" | tail -15
python bench.py --steps 100 --warmup 10 --no-cpu 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_fused.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_fused.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],d['roofline']['stage_ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])
PY
