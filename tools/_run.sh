# scratch script for `gpurun -- 'bash tools/_run.sh'`
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests/test_gpu_mc.py -m gpu -x -q ) > gpurun_out/r2f_mc_tests.log 2>&1; tail -3 gpurun_out/r2f_mc_tests.log
( timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2f_mc512.log 2>&1; tail -1 gpurun_out/r2f_mc512.log
( timeout -s KILL 300 python tools/mc_profile.py 256 ) > gpurun_out/r2f_mc256.log 2>&1; tail -1 gpurun_out/r2f_mc256.log
( time timeout -s KILL 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
cut -c1-1500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
