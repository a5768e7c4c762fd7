timeout -s KILL 300 python -m pytest tests/test_gpu_mc.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_mc.log
timeout -s KILL 200 python tools/mc_profile.py 256 > gpurun_out/mc_profile.log 2>&1
timeout -s KILL 380 python tools/overlap_probe.py 256 > gpurun_out/overlap_probe.log 2>&1
SURFD_MC_FLAGS=-DMC_PROFILE timeout -s KILL 300 python -m surfd_b200.build > gpurun_out/build_prof.log 2>&1
timeout -s KILL 200 python tools/mc_profile.py 256 > gpurun_out/mc_profile_cycles.log 2>&1
cat gpurun_out/pytest_mc.log gpurun_out/mc_profile.log gpurun_out/mc_profile_cycles.log gpurun_out/overlap_probe.log
