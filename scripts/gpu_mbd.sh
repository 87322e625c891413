mkdir -p gpurun_out
setsid timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/mb_density_check.py > gpurun_out/mbd.log 2>&1
grep "mb_density_check" gpurun_out/mbd.log || tail -25 gpurun_out/mbd.log
