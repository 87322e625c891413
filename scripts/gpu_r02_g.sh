# round 2, call G: full GPU suite
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02g_tests.log
tail -6 gpurun_out/r02g_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
