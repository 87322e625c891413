mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_multibrick.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -25 > gpurun_out/spec_tests.log
grep -n "AssertionError\|passed\|failed\|Error" gpurun_out/spec_tests.log | head
python bench.py --steps 1000 --warmup 20 --no-cpu 2>gpurun_out/bench.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'rebuilds', d['thermo']['rebuilds_total'], d['roofline']['stage_ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_build_rows_tiled -c 4 --csv --log-file gpurun_out/build_launches.csv python bench.py --steps 40 --warmup 3 --no-cpu > /dev/null 2>&1
grep k_build_rows gpurun_out/build_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -4
