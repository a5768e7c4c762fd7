mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests/test_gpu_mc.py -m gpu -x -q ) > gpurun_out/r2af_mc_tests.log 2>&1; tail -12 gpurun_out/r2af_mc_tests.log | cut -c1-300
( time timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2af_smoke.log 2>&1; tail -4 gpurun_out/r2af_smoke.log
