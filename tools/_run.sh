# scratch script for `gpurun -- 'bash tools/_run.sh'`
mkdir -p gpurun_out
( timeout -s KILL 300 python tools/probe_large_batch.py 64 ) > gpurun_out/r2b_large64.log 2>&1; tail -8 gpurun_out/r2b_large64.log
( timeout -s KILL 200 python tools/probe_batch_scaling.py 32 ) > gpurun_out/r2b_scaling32.log 2>&1; cat gpurun_out/r2b_scaling32.log | tail -8
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 200 python tools/mc_profile.py 256 ) > gpurun_out/r2b_mcprof256.log 2>&1; tail -1 gpurun_out/r2b_mcprof256.log
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2b_mcprof512.log 2>&1; tail -1 gpurun_out/r2b_mcprof512.log
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2b_pytest_gpu.log
( time timeout -s KILL 600 python bench.py --steps 2 --warmup 1 ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cut -c1-600 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
