timeout -s KILL 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_mc.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_unet.log
cat gpurun_out/pytest_unet.log
timeout -s KILL 300 python tools/sampler_profile.py > gpurun_out/sampler_profile.log 2>&1
head -12 gpurun_out/sampler_profile.log
timeout -s KILL 200 python tools/mc_profile.py 256 > gpurun_out/mc_profile.log 2>&1
tail -1 gpurun_out/mc_profile.log
SURFD_MC_FLAGS=-DMC_PROFILE timeout -s KILL 300 python -m surfd_b200.build > gpurun_out/build_prof.log 2>&1
timeout -s KILL 200 python tools/mc_profile.py 256 > gpurun_out/mc_profile_cycles.log 2>&1
tail -1 gpurun_out/mc_profile_cycles.log
