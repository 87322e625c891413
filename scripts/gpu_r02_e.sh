# round 2, call E: step_host parity, contract bench (pipelined e2e), reference arm
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "step_host" 2>&1 | tail -15 > gpurun_out/r02e_tests.log
tail -4 gpurun_out/r02e_tests.log
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().split("\n")[-1])
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches")}, "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], d["roofline"]["stage_ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["seconds"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
PY
UCGB200_E2E_PIPELINE=0 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('unpipelined e2e', d['e2e']['ms_per_step'], 'value', d['value'])"
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02e_reference.json 2>> gpurun_out/r02e_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_reference.json').read().strip().split("\n")[-1])
print(d["impl"], d["value"], d["steps"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["wall_s"], d["cpu_baseline"]["serial_1M"])
PY
