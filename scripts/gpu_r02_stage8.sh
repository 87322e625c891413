N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N"
UCGB200_WEAK4M_NCELL=0 UCGB200_BENCH_PARITY=0 timeout -s KILL 300 $TR bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/r02st_$N.err | tail -1 > gpurun_out/r02st_$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r02st_$N.json").read())
print("N=$N value",d["value"],"ms/step",d["ms_per_step"],"stages",d["roofline"]["stage_ms_per_step"],"halo",d["halo"])
PY
python bench.py --steps 100 --warmup 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N=1 value', d['value'], 'ms/step', d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['thermo'])"
