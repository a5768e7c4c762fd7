mkdir -p gpurun_out
( time timeout -s KILL 300 python tools/sampler_profile.py ) > gpurun_out/r2aj_sampler_profile.log 2>&1; head -3 gpurun_out/r2aj_sampler_profile.log | cut -c1-900; grep -n "first_cta" gpurun_out/r2aj_sampler_profile.log | cut -c1-400
( time timeout -s KILL 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -s ) > gpurun_out/r2aj_unet_tests.log 2>&1; grep -n "1000-step\|passed\|failed\|Error" gpurun_out/r2aj_unet_tests.log | head
