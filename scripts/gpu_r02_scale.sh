# usage: gpu_r02_scale.sh N [steps] — bench.py --gpus N (peer-mapped halo) under a hard KILL timeout; 100-step and contract (20-step) runs
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N"
UCGB200_WEAK4M_NCELL=0 UCGB200_BENCH_PARITY=0 timeout -s KILL 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02scale_${N}_100.json 2> gpurun_out/r02scale_${N}_100.err; echo "bench 100 rc=$?"
timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02scale_${N}_contract.json 2> gpurun_out/r02scale_${N}_contract.err; echo "bench contract rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02scale_${N}_100.json", "gpurun_out/r02scale_${N}_contract.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value",d["value"],"ms/step",d["ms_per_step"],"halo",d["halo"]["transport"][:30],"parity ok",(d.get("parity") or {}).get("ok"),"weak4M",d.get("weak_4M_per_gpu"))
    except Exception as e:
        print("no bench line", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
