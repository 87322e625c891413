python -m pytest tests/test_gpu_parity.py -m gpu -q -k "neighbor" 2>&1 | tail -3
python scripts/profile_build.py 2>&1 | tail -2
