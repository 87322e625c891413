# session 3: 2-GPU contract run with the final library (uploads now go through their own stream); weak-4M leg skipped to save box time
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562"
UCGB200_WEAK4M_NCELL=0 timeout -s KILL 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s3_n2.json 2> gpurun_out/s3_n2.err; echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s3_n2.json").read().strip().split("\n")[-1])
    print("value",d["value"],"ms/step",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d.get("parity"))
except Exception as e:
    print("no bench line", e); print(open("gpurun_out/s3_n2.err").read()[-1500:])
PY
