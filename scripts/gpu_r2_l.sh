mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider -k stem > gpurun_out/pytest_l.log 2>&1; echo "pytest stem rc=$?"; grep -E "^(FAILED|ERROR)|assert|Error" gpurun_out/pytest_l.log | head -20; tail -3 gpurun_out/pytest_l.log
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_towers_gpu.py tests/test_graphs_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_l2.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_l2.log | head; tail -3 gpurun_out/pytest_l2.log
timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo "bench rc=$?"; grep -v "^$\|Warning\|warn" gpurun_out/bench_r2l.err | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2l.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'loss', d['last_loss'])
r=d['roofline']
for k,v in sorted(r['families'].items(), key=lambda kv:-kv[1]['ms_per_step']): print('  %-18s %6.3f ms  n=%d  %s'%(k,v['ms_per_step'],v['launches'], round(v.get('TFLOP/s',0),1)))
PY
