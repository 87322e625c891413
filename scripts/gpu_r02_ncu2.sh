# round 2: refresh of the ncu evidence after the last kernel changes: build kernel (level buckets), tail, pair; k_rle_* sweeps
mkdir -p gpurun_out
O=gpurun_out
STEPS=2 python scripts/profile_step.py > $O/r02n_plain_step.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast|k_step_tail|k_build_rows_tiled_f32" -s 1 -c 7 -o $O/r02n_step python scripts/profile_step.py > $O/r02n_ncu_step.log 2>&1
tail -1 $O/r02n_ncu_step.log
python scripts/ncu_summary.py $O/r02n_step.ncu-rep $O/r02_step_kernels_full > $O/r02n_sum_step.log 2>&1
python scripts/ncu_lines.py $O/r02n_step.ncu-rep k_build_rows_tiled_f32 $O/r02_build_rows_lines.json > /dev/null 2>&1
rm -f $O/r02n_step.ncu-rep
python scripts/time_styles.py > $O/r02n_plain_styles.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_rle_pair|k_rle_density|k_rle_back" -s 3 -c 6 -o $O/r02n_rle python scripts/time_styles.py > $O/r02n_ncu_rle.log 2>&1
tail -1 $O/r02n_ncu_rle.log
python scripts/ncu_summary.py $O/r02n_rle.ncu-rep $O/r02_rle_kernels_full > $O/r02n_sum_rle.log 2>&1
python scripts/ncu_lines.py $O/r02n_rle.ncu-rep k_rle_pair $O/r02_k_rle_pair_lines.json > /dev/null 2>&1
rm -f $O/r02n_rle.ncu-rep
cat $O/r02n_sum_step.log $O/r02n_sum_rle.log | cut -c1-120; tail -1 $O/r02n_plain_styles.log | cut -c1-300; du -sh $O
