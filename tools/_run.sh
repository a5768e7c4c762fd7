mkdir -p gpurun_out
timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -c 400 gpurun_out/bench_ncu.log; wc -l gpurun_out/launches_bench.csv
