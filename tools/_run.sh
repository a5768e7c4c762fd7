# scratch script for `gpurun -- 'bash tools/_run.sh'`: GPU tests, smoke, default bench (C3)
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -q ) > gpurun_out/tests_gpu.log 2>&1; grep -n "passed\|failed" gpurun_out/tests_gpu.log
( time timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; grep "smoke ok" gpurun_out/smoke.log
( time timeout -s KILL 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; cut -c1-400 gpurun_out/bench_C3.json
