mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
( time timeout -s KILL 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
cat gpurun_out/bench_tf32.json; tail -5 gpurun_out/bench_tf32.err
