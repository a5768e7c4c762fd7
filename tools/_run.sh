mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tail -8 | tr '\n' ';'; echo
( time timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 1 --no-fp32 --no-cpu-baseline ) > gpurun_out/r2ao_bench_n8_C3.json 2> gpurun_out/r2ao_bench_n8_C3.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2ao_bench_n8_C3.json').read().strip().splitlines()[-1])
    print('N=8 C3 value', d['value'], 'n_gpus', d['n_gpus'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('stages_s_per_step'))
except Exception as e:
    print('failed', e)
PY
tail -3 gpurun_out/r2ao_bench_n8_C3.err
