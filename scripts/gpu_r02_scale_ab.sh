# usage: gpu_r02_scale_ab.sh N — bench.py --gpus N, 100 steps: halo/compute overlap on vs off, on the same box
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N"
for ov in 1 0 1 0; do
UCGB200_OVERLAP=$ov UCGB200_WEAK4M_NCELL=0 UCGB200_BENCH_PARITY=0 timeout -s KILL 300 $TR bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/r02ab_${N}_$ov.err | tail -1 > gpurun_out/r02ab_${N}_$ov.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02ab_${N}_$ov.json").read())
    print("overlap=$ov value",d["value"],"ms/step",d["ms_per_step"],"launches",d["gpu_launches"],"pair_ms",d["roofline"]["kernel_ms"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02ab_${N}_$ov.err").read()[-800:])
PY
done
