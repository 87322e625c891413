# round 2, final code: --set full capture of the step kernels (pair, fused tail, row build) on the 1M-site liquid
mkdir -p gpurun_out
O=gpurun_out
STEPS=2 python scripts/profile_step.py > $O/r02f_plain_step.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast|k_step_tail|k_build_rows_tiled_f32" -s 1 -c 7 -o $O/r02f_step python scripts/profile_step.py > $O/r02f_ncu_step.log 2>&1
tail -1 $O/r02f_ncu_step.log
python scripts/ncu_summary.py $O/r02f_step.ncu-rep $O/r02_step_kernels_full > $O/r02f_sum_step.log 2>&1
python scripts/ncu_lines.py $O/r02f_step.ncu-rep k_build_rows_tiled_f32 $O/r02_build_rows_lines.json > /dev/null 2>&1
python scripts/ncu_lines.py $O/r02f_step.ncu-rep k_step_tail $O/r02_step_tail_lines.json > /dev/null 2>&1
rm -f $O/r02f_step.ncu-rep
cat $O/r02f_sum_step.log | cut -c1-140; du -sh $O
