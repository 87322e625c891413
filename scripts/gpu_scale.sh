mkdir -p gpurun_out
N=${NG:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
tail -5 gpurun_out/scale_$N.err; cat gpurun_out/scale_$N.json
