# scratch script for `gpurun -- 'bash tools/_run.sh'`
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
( time timeout -s KILL 400 python tools/probe_large_batch.py 32 ) > gpurun_out/r2a_large32.log 2>&1; tail -30 gpurun_out/r2a_large32.log
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 200 python tools/mc_profile.py 256 ) > gpurun_out/r2a_mcprof256.log 2>&1; tail -2 gpurun_out/r2a_mcprof256.log
( SURFD_B200_LIB=$PWD/surfd_b200/_surfd_b200_mcprof.so timeout -s KILL 300 python tools/mc_profile.py 512 ) > gpurun_out/r2a_mcprof512.log 2>&1; tail -2 gpurun_out/r2a_mcprof512.log
( timeout -s KILL 300 python tools/sampler_profile.py ) > gpurun_out/r2a_sampler_profile.log 2>&1; tail -8 gpurun_out/r2a_sampler_profile.log
