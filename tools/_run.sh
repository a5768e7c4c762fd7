mkdir -p gpurun_out
( time timeout -s KILL 300 python tools/sampler_profile.py ) > gpurun_out/r2ac_sampler_profile.log 2>&1; head -2 gpurun_out/r2ac_sampler_profile.log | cut -c1-300; grep -n "GroupNorm unit" gpurun_out/r2ac_sampler_profile.log | cut -c1-400; tail -4 gpurun_out/r2ac_sampler_profile.log | cut -c1-400
