mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "neighbor" 2>&1 | grep -v "^E    *+" | tail -15
python bench.py --steps 100 --warmup 10 --no-cpu 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_tiled.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tiled.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],d['roofline']['stage_ms_per_step'],'e2e',d['e2e']['value'],'rebuilds',d['thermo'])
PY
UCGB200_BUILD_TILED=0 python bench.py --steps 100 --warmup 10 --no-cpu 2>>gpurun_out/bench.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('untiled value',d['value'],d['roofline']['stage_ms_per_step'])"
