mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | grep -v "^E    *+" | tail -12 > gpurun_out/all_tests.log
tail -3 gpurun_out/all_tests.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().split("\n")[-1])
print({k:d[k] for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["cpu_baseline"]["value"], d["e2e"]["value"], d["clocks"], d["taps"])
PY
python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err
tail -c 400 gpurun_out/bench_reference.json
