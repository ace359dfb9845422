mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graphs_gpu.py tests/test_train_gpu.py tests/test_dropin.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|Error|warn" gpurun_out/pytest_j.log | head; tail -3 gpurun_out/pytest_j.log
for g in 1 0; do
AVID_CUDA_GRAPH=$g timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2j_$g.json 2> gpurun_out/bench_r2j_$g.err; echo "bench graph=$g rc=$?"; grep -v "^$" gpurun_out/bench_r2j_$g.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2j_$g.json').read())
print('graph=$g value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'loss', d['last_loss'])
PY
done
