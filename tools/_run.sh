mkdir -p gpurun_out
( timeout -s KILL 300 python -m pytest tests/test_gpu_unet.py -x -q ) > gpurun_out/pytest_unet.log 2>&1; tail -12 gpurun_out/pytest_unet.log
timeout -s KILL 200 python tools/sampler_profile.py 2>&1 | grep -v Warning | tail -6 > gpurun_out/sampler_debug.txt 2>&1
cat gpurun_out/sampler_debug.txt
