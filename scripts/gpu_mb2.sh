mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mb_check.py 2>&1 | grep -v "^W\|^\[W\|warn" | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 2>gpurun_out/scale2.err | tail -1 > gpurun_out/scale2_resident.json
UCGB200_MB_PYTHON=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 10 2>>gpurun_out/scale2.err | tail -1 > gpurun_out/scale2_python.json
python - <<'PY'
import json
for f in ("scale2_resident","scale2_python"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["halo"])
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/scale2.err
