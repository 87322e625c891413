mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_multibrick.py -m gpu -q 2>&1 | grep -v "^E    *+" | tail -25 > gpurun_out/spec_tests.log
grep -n "AssertionError\|passed\|failed\|Error" gpurun_out/spec_tests.log | head
python scripts/time_build.py 2>&1 | tail -5
python bench.py --steps 1000 --warmup 20 --no-cpu 2>gpurun_out/bench.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'rebuilds', d['thermo']['rebuilds_total'])"
