mkdir -p gpurun_out
( time timeout -s KILL 600 python tools/probe_chain.py 256 ) > gpurun_out/r2x_probe_chain.log 2>&1; tail -10 gpurun_out/r2x_probe_chain.log
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2x_tests.log 2>&1; tail -4 gpurun_out/r2x_tests.log
( time timeout -s KILL 900 python bench.py --steps 3 --warmup 1 ) > gpurun_out/r2x_bench_C3.json 2> gpurun_out/r2x_bench_C3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_C3.json').read().strip().splitlines()[-1])
print('C3 value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('stages_s_per_step'), 'frac', d['roofline']['frac'], 'dec', d['roofline_decoder']['frac'], d['roofline_decoder']['ms_per_launch'], d['roofline_decoder'].get('isolated'), d['roofline_decoder'].get('launches'))
PY
tail -3 gpurun_out/r2x_bench_C3.err
