mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split("\n")[-1])
print({k:d[k] for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, d["roofline"]["frac"], d["roofline"].get("secondary"), d["cpu_baseline"]["value"], d["e2e"]["value"], d["clocks"])
PY
( time python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_reference.json').read().strip().split("\n")[-1])
print(d["impl"], d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["sample"][:80], d["cpu_baseline"]["wall_s"])
PY
