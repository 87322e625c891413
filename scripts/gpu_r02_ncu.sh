# round 2: ncu evidence.  (1) --set full of the step's kernels, (2) of the other pair styles, (3) of the N3L variants,
# (4) launch list of the contract bench run.  Every ncu command follows a plain run of the same command line.  The
# reports are summarised on the box (scripts/ncu_summary.py, scripts/ncu_lines.py) and deleted: gpurun_out/ may carry 64 MiB.
mkdir -p gpurun_out
O=gpurun_out
STEPS=2 python scripts/profile_step.py > $O/r02n_plain_step.log 2>&1 && \
STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_fast|k_step_tail|k_build_rows_tiled_f32" -s 1 -c 7 -o $O/r02n_step python scripts/profile_step.py > $O/r02n_ncu_step.log 2>&1
tail -1 $O/r02n_ncu_step.log
python scripts/ncu_summary.py $O/r02n_step.ncu-rep $O/r02_step_kernels_full > $O/r02n_sum_step.log 2>&1
python scripts/ncu_lines.py $O/r02n_step.ncu-rep k_pair_ucgld_fast $O/r02_pair_lines.json > /dev/null 2>&1
python scripts/ncu_lines.py $O/r02n_step.ncu-rep k_step_tail $O/r02_step_tail_lines.json > /dev/null 2>&1
python scripts/ncu_lines.py $O/r02n_step.ncu-rep k_build_rows_tiled_f32 $O/r02_build_rows_lines.json > /dev/null 2>&1
rm -f $O/r02n_step.ncu-rep
python scripts/time_styles.py > $O/r02n_plain_styles.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_pair_bethe|k_bd_pair|k_rle_pair|k_bd_prior|k_bd_back|k_rle_density|k_rle_back" -s 3 -c 9 -o $O/r02n_styles python scripts/time_styles.py > $O/r02n_ncu_styles.log 2>&1
tail -1 $O/r02n_ncu_styles.log
python scripts/ncu_summary.py $O/r02n_styles.ncu-rep $O/r02_style_kernels_full > $O/r02n_sum_styles.log 2>&1
for k in k_pair_bethe k_bd_pair k_rle_pair; do python scripts/ncu_lines.py $O/r02n_styles.ncu-rep $k $O/r02_${k}_lines.json > /dev/null 2>&1; done
rm -f $O/r02n_styles.ncu-rep
for m in 1 2; do
UCGB200_N3L=$m STEPS=2 python scripts/profile_step.py > $O/r02n_plain_n3l$m.log 2>&1 && \
UCGB200_N3L=$m STEPS=2 ncu --set full --clock-control none --import-source on -k regex:"k_pair_ucgld_n3l" -s 1 -c 1 -o $O/r02n_n3l$m python scripts/profile_step.py > $O/r02n_ncu_n3l$m.log 2>&1
python scripts/ncu_summary.py $O/r02n_n3l$m.ncu-rep $O/r02_pair_n3l${m}_full > $O/r02n_sum_n3l$m.log 2>&1
python scripts/ncu_lines.py $O/r02n_n3l$m.ncu-rep k_pair_ucgld_n3l $O/r02_pair_n3l${m}_lines.json > /dev/null 2>&1
rm -f $O/r02n_n3l$m.ncu-rep
done
python bench.py --steps 20 --warmup 5 --no-cpu > $O/r02n_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-cpu > $O/r02n_ncu_bench.log 2>&1
wc -l $O/r02_launches_bench.csv; cat $O/r02n_sum_step.log | cut -c1-120; cat $O/r02n_sum_styles.log | cut -c1-120; du -sh $O
