mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -s ) > gpurun_out/r2z_unet_tests.log 2>&1; grep -n "1000-step\|passed\|failed\|Error" gpurun_out/r2z_unet_tests.log | head
( time timeout -s KILL 900 python bench.py --config C5 --steps 3 --warmup 1 ) > gpurun_out/r2z_bench_C5.json 2> gpurun_out/r2z_bench_C5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_C5.json').read().strip().splitlines()[-1])
print('C5 value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('stages_s_per_step'), 'frac', d['roofline']['frac'], d['roofline'].get('ms_per_ddpm_step'))
PY
tail -3 gpurun_out/r2z_bench_C5.err
