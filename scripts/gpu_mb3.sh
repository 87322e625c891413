# 2-GPU checks of the resident multi-brick driver; every multi-rank command runs in its own session under a
# hard KILL timeout and writes to a file (no pipe that a surviving worker could keep open)
run() { setsid timeout -s KILL $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 scripts/$1 > gpurun_out/$1.log 2>&1; grep "_check" gpurun_out/$1.log || tail -8 gpurun_out/$1.log; }
mkdir -p gpurun_out
STEPS=60 run mb_check.py 29525 150
run mb_density_check.py 29541 150
run mb_cluster_check.py 29551 150
