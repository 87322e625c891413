# usage: gpu_scale.sh N   — one bench.py run on N GPUs under a hard KILL timeout, output to files only
N=$1
mkdir -p gpurun_out
setsid timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/scale_$N.out 2> gpurun_out/scale_$N.err
tail -1 gpurun_out/scale_$N.out > gpurun_out/scale_$N.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$N.json")); print("N=$N", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["halo"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale_$N.err").read()[-1500:])
PY
