mkdir -p gpurun_out
AVID_WGRAD_STREAM=1 timeout 900 python -m pytest tests/test_towers_gpu.py tests/test_graphs_gpu.py tests/test_config2_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_ws.log
for w in 0 1 0 1; do
AVID_WGRAD_STREAM=$w timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_ws_$w.json 2> gpurun_out/bench_ws_$w.err; echo "bench wgrad_stream=$w rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ws_$w.json').read())
print('wgrad_stream=$w value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'loss', d['last_loss'])
PY
done
