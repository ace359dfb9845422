# the driver's round-end sequence on one GPU: GPU tests, smoke, reference arm, own arm (default flags)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_final_reference.json
timeout 1200 python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final_1gpu.err; echo "ours rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('roofline', r['kernel'], round(r['achieved'],1), round(r['frac'],3), 'executed', round(r['executed_frac'],3), 'traffic', r['traffic'])
for k,v in r['families'].items(): print('  ', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print('nce', {k:(round(v,3) if isinstance(v,float) else v) for k,v in r['nce'].items() if k!='sweep'})
print('cpu', d.get('cpu_baseline')); print('gpu lib', d.get('gpu_library_baseline'))
PY
