mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_baseline_sizes.py -m gpu -x -q ) > gpurun_out/r2ae_mc_tests.log 2>&1; tail -3 gpurun_out/r2ae_mc_tests.log
( timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2ae_mc512.log 2>&1; tail -1 gpurun_out/r2ae_mc512.log
( timeout -s KILL 300 python tools/mc_profile.py 256 ) > gpurun_out/r2ae_mc256.log 2>&1; tail -1 gpurun_out/r2ae_mc256.log
