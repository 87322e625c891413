# usage: gpu_r02_final.sh N — what the driver runs at round end for N GPUs (bench.py --gpus N --steps 20 --warmup 5), plus,
# at N = 2, the 2-GPU checks of the resident multi-brick driver (cluster_switch with and without a group, density styles)
N=${1:-1}
mkdir -p gpurun_out
O=gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02final_1.json 2> $O/r02final_1.err; echo "rc=$?"
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02final_1_reference.json 2>> $O/r02final_1.err
  python bench.py --gpus 1 --steps 1000 --warmup 20 --no-cpu > $O/r02final_1_1000steps.json 2>> $O/r02final_1.err
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N"
  if [ "$N" = "2" ]; then
    for gb in 1 2; do GROUPBIT=$gb timeout -s KILL 200 $TR scripts/mb_cluster_check.py > $O/r02final_mbcluster_g$gb.log 2>&1; grep "mb_cluster_check" $O/r02final_mbcluster_g$gb.log; done
    timeout -s KILL 200 $TR scripts/mb_density_check.py > $O/r02final_mbdensity.log 2>&1; grep "_check" $O/r02final_mbdensity.log
    timeout -s KILL 200 $TR scripts/mb_check.py > $O/r02final_mbcheck.log 2>&1; grep "mb_check" $O/r02final_mbcheck.log
  fi
  timeout -s KILL 500 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r02final_$N.json 2> $O/r02final_$N.err; echo "rc=$?"
fi
python - <<PY
import json
for f in ("$O/r02final_$N.json", "$O/r02final_${N}_1000steps.json", "$O/r02final_${N}_reference.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", (d.get("parity") or {}).get("ok"), "4M", (d.get("weak_4M_per_gpu") or {}).get("value"))
    except Exception as e:
        pass
PY
