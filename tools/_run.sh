mkdir -p gpurun_out
( time timeout -s KILL 300 python tools/sampler_profile.py ) > gpurun_out/r2ai_sampler_profile.log 2>&1; head -2 gpurun_out/r2ai_sampler_profile.log | cut -c1-300; tail -5 gpurun_out/r2ai_sampler_profile.log | cut -c1-500
( time timeout -s KILL 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -s ) > gpurun_out/r2ai_unet_tests.log 2>&1; grep -n "1000-step\|passed\|failed\|Error" gpurun_out/r2ai_unet_tests.log | head
