mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_towers_gpu.py tests/test_graphs_gpu.py tests/test_train_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_v.log | head; tail -3 gpurun_out/pytest_v.log
for w in 1 0; do
AVID_WGRAD_STREAM=$w timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2v_$w.json 2> gpurun_out/bench_r2v_$w.err; echo "bench wgrad_stream=$w rc=$?"; grep -v "^$\|Warning\|warn\|run_backward" gpurun_out/bench_r2v_$w.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2v_$w.json').read())
print('wgrad_stream=$w value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'loss', d['last_loss'])
PY
done
