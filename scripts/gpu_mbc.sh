mkdir -p gpurun_out
setsid timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/mb_cluster_check.py > gpurun_out/mbcl.log 2>&1
grep "mb_cluster_check" gpurun_out/mbcl.log || tail -25 gpurun_out/mbcl.log
