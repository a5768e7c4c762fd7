timeout -s KILL 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_zz_tensorcore.py -x -q 2>&1 | tail -8 > gpurun_out/pytest_dec.log
cat gpurun_out/pytest_dec.log
CHUNK=35840 timeout -s KILL 200 python tools/ncu_extract.py 256 2 2>&1 | tail -1
CHUNK=71680 timeout -s KILL 200 python tools/ncu_extract.py 256 2 2>&1 | tail -1
timeout -s KILL 300 python tools/sampler_profile.py 2>&1 | tail -4
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
cat gpurun_out/bench_tf32.json; tail -3 gpurun_out/bench_tf32.err
