mkdir -p gpurun_out
python scripts/profile_dump.py > gpurun_out/plain_dump.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_dump_choose|k_dump_pack|k_dump_format|k_dump_emit|k_parse_rows|k_update_by_tag" -c 6 -o gpurun_out/prof_dump python scripts/profile_dump.py > gpurun_out/ncu_dump.log 2>&1
tail -2 gpurun_out/plain_dump.log; tail -2 gpurun_out/ncu_dump.log; ls -la gpurun_out/prof_dump.ncu-rep
