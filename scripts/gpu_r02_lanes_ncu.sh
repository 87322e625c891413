# round 2: ncu capture of the lane-per-site row build (k_build_rows_lanes_f32) on the running 1M-site liquid
mkdir -p gpurun_out
O=gpurun_out
python scripts/profile_build.py > $O/r02l_plain_build.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_build_rows_lanes_f32" -s 1 -c 2 -o $O/r02l_build python scripts/profile_build.py > $O/r02l_ncu_build.log 2>&1
tail -1 $O/r02l_ncu_build.log
python scripts/ncu_summary.py $O/r02l_build.ncu-rep $O/r02_build_rows_lanes_full > $O/r02l_sum.log 2>&1
python scripts/ncu_lines.py $O/r02l_build.ncu-rep k_build_rows_lanes_f32 $O/r02_build_rows_lanes_lines.json > /dev/null 2>&1
rm -f $O/r02l_build.ncu-rep
cat $O/r02l_sum.log | cut -c1-160; du -sh $O
