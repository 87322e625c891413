N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2959$N"
UCGB200_COMM_TRACE=2 UCGB200_BUILD_TRACE=1 STEPS=45 timeout -s KILL 200 $TR scripts/mb_trace.py > gpurun_out/r02trace_$N.log 2>&1
grep -E "mb_rebuild|\[build\]|trace run" gpurun_out/r02trace_$N.log | tail -46
