# scratch script for `gpurun -- 'bash tools/_run.sh'`
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests/test_gpu_mc.py tests/test_gpu_baseline_sizes.py -m gpu -x -q ) > gpurun_out/r2e_mc_tests.log 2>&1; tail -5 gpurun_out/r2e_mc_tests.log
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 200 python tools/mc_profile.py 256 ) > gpurun_out/r2e_mcprof256.log 2>&1; tail -1 gpurun_out/r2e_mcprof256.log
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2e_mcprof512.log 2>&1; tail -1 gpurun_out/r2e_mcprof512.log
( timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2e_mc512.log 2>&1; tail -1 gpurun_out/r2e_mc512.log
( timeout -s KILL 300 python tools/mc_profile.py 256 ) > gpurun_out/r2e_mc256.log 2>&1; tail -1 gpurun_out/r2e_mc256.log
