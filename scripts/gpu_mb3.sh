# every multi-rank command runs in its own session under a hard KILL timeout and writes to a file
# (no pipe that a surviving worker could keep open)
run() { setsid timeout -s KILL $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 scripts/mb_check.py > gpurun_out/mbc_$1.log 2>&1; grep "mb_check" gpurun_out/mbc_$1.log || tail -5 gpurun_out/mbc_$1.log; }
mkdir -p gpurun_out
echo "steps=8 fused";  STEPS=8 run 29523 120
echo "steps=8 unfused"; STEPS=8 UCGB200_FUSED_TAIL=0 run 29524 120
echo "steps=60 fused"; STEPS=60 run 29525 150
