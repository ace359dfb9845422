mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_h.log | head; tail -3 gpurun_out/pytest_h.log
timeout 900 python bench.py --steps 20 --warmup 3 --dump-launches gpurun_out/launch_table_r2h.txt > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']; print('roof', r['kernel'], round(r['achieved'],1), round(r['frac'],3), r['measured_in'])
for k,v in r['families'].items(): print('  ',k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print('nce', r.get('nce',{}).get('frac')); print('gpu base', d.get('gpu_library_baseline')); print('cpu', d.get('cpu_baseline',{}).get('value'))
PY
