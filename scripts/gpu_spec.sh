mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multibrick.py tests/test_gpu_host_classes.py tests/test_gpu_edge_cases.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -25 > gpurun_out/spec_tests.log
grep -n "AssertionError\|passed\|failed\|Error" gpurun_out/spec_tests.log | head
for s in 1 0; do UCGB200_SPECULATE=$s python bench.py --steps 500 --warmup 20 --no-cpu 2>gpurun_out/bench.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('speculate=$s', 'value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'rebuilds', d['thermo']['rebuilds_total'])"; done
