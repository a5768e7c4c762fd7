# scratch script for `gpurun -- 'bash tools/_run.sh'`: GPU tests, smoke, bench
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|rror" gpurun_out/pytest_gpu.log | tail -3
( timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | grep -E "smoke|rror" | tail -3
( time timeout -s KILL 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
