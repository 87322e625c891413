# round 2, final code: launch list of the contract command (cold-cache, serialised per-launch times: shares only)
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu > $O/r02f_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu > $O/r02f_ncu_bench.log 2>&1
wc -l $O/r02_launches_bench.csv; tail -1 $O/r02f_plain_bench.log | cut -c1-200
