mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_towers_gpu.py tests/test_conv_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_i.log | head; tail -3 gpurun_out/pytest_i.log
AVID_PROFILE_ALL=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2i.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))
r=d['roofline']
for k,v in sorted(r['families'].items(), key=lambda kv:-kv[1]['ms_per_step']): print('  %-18s %6.3f ms  n=%d'%(k,v['ms_per_step'],v['launches']))
PY
