# session 3: ncu --set full of the density-style sweeps after the log-sum and packed-record changes
O=gpurun_out
python scripts/time_styles.py > $O/s3n_plain_styles.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_bd_prior|k_bd_pair|k_bd_posc|k_cv_back_fast|k_rle_density|k_rle_pair|k_rle_posc" -s 8 -c 20 -o $O/s3n_styles python scripts/time_styles.py > $O/s3n_ncu_styles.log 2>&1
tail -1 $O/s3n_ncu_styles.log
python scripts/ncu_summary.py $O/s3n_styles.ncu-rep $O/r02_style_kernels_full_s3 > $O/s3n_sum.log 2>&1
python scripts/ncu_lines.py $O/s3n_styles.ncu-rep k_cv_back_fast $O/r02_k_cv_back_fast_lines.json > /dev/null 2>&1
rm -f $O/s3n_styles.ncu-rep
cut -c1-160 $O/s3n_sum.log | head -30
