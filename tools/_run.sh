mkdir -p gpurun_out
( time timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 ) > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
cut -c1-260 gpurun_out/bench_n2_final.json; tail -3 gpurun_out/bench_n2_final.err
