timeout -s KILL 600 python -m pytest tests/test_gpu_unet.py -x -q 2>&1 | tail -4
timeout -s KILL 300 python tools/sampler_profile.py 2>&1 | grep -E "^B=|graph replay"
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
cat gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
