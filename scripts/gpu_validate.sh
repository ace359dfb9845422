# validation pass: smoke, gpu tests (with durations), bench bf16x3, launch list of one warm step
mkdir -p gpurun_out
T0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s)-T0 ))"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s)-T0 ))"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed" gpurun_out/pytest_gpu.log | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --math bf16x3 --no-cpu-baseline > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench rc=$? t=$(( $(date +%s)-T0 ))"; tail -2 gpurun_out/bench_bf16x3.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_bf16x3.json"))
print({k:d[k] for k in ("value","ms_per_step","dtype","gpu_launches","last_loss","clocks")}, d["e2e"]["value"])
for k,v in d["roofline"]["families"].items(): print(k, v)
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python bench.py --steps 1 --warmup 1 --math bf16x3 --skip-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu t=$(( $(date +%s)-T0 ))"; wc -l gpurun_out/launches_bf16x3.csv
