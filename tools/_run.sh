mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
( time timeout -s KILL 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2ak_tests.log 2>&1; grep -n "passed\|failed" gpurun_out/r2ak_tests.log; grep -n "^FAILED\|Error" gpurun_out/r2ak_tests.log | head -5
( time timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2ak_smoke.log 2>&1; grep "smoke ok" gpurun_out/r2ak_smoke.log
for c in C3 C2 C4 C5; do
  ( time timeout -s KILL 900 python bench.py --config $c --steps 5 --warmup 3 ) > gpurun_out/r2ak_bench_$c.json 2> gpurun_out/r2ak_bench_$c.err
done
python - <<'PY'
import json
for c in ("C3","C2","C4","C5"):
    try:
        d=json.loads(open('gpurun_out/r2ak_bench_%s.json'%c).read().strip().splitlines()[-1])
        print(c, 'value', d['value'], 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('stages_s_per_step'), 'frac', d['roofline']['frac'], 'dec', d['roofline_decoder']['frac'], 'cls', d['roofline_mc_classify']['frac'], d.get('precision_fp32',{}).get('value'), 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(c, 'failed', e)
PY
