# round 2, multi-GPU call: resident multi-brick driver with the peer-mapped forward halo (and the NCCL fallback)
# against the Python-orchestrated driver and one brick; then bench.py --gpus N with its parity block.  usage: gpu_r02_mb.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
UCGB200_COMM_TRACE=1 timeout -s KILL 240 $TR scripts/mb_check.py > gpurun_out/r02mb_check_p2p_$N.log 2>&1; echo "mb_check p2p rc=$?"
grep -E "mb_check|peer-mapped|Error|error" gpurun_out/r02mb_check_p2p_$N.log | head -8
UCGB200_P2P=0 timeout -s KILL 240 $TR scripts/mb_check.py > gpurun_out/r02mb_check_nccl_$N.log 2>&1; echo "mb_check nccl rc=$?"
grep -E "mb_check|Error|error" gpurun_out/r02mb_check_nccl_$N.log | head -4
for p2p in 1 0; do
  UCGB200_P2P=$p2p UCGB200_WEAK4M_NCELL=0 timeout -s KILL 300 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02mb_bench_${N}_p2p$p2p.json 2> gpurun_out/r02mb_bench_${N}_p2p$p2p.err; echo "bench p2p=$p2p rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02mb_bench_${N}_p2p$p2p.json").read().strip().split("\n")[-1])
    print("value",d["value"],"ms/step",d["ms_per_step"],"halo",d["halo"]["transport"][:40],"parity",d["parity"])
except Exception as e:
    print("no bench line", e); print(open("gpurun_out/r02mb_bench_${N}_p2p$p2p.err").read()[-1500:])
PY
done
timeout -s KILL 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02mb_bench_${N}_contract.json 2> gpurun_out/r02mb_bench_${N}_contract.err; echo "bench contract rc=$?"
tail -c 900 gpurun_out/r02mb_bench_${N}_contract.json
