mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))|assert|Error|bf16 single" gpurun_out/pytest_gpu.log | head -60
for m in bf16x3 bf16; do
timeout 600 python bench.py --steps 10 --warmup 3 --math $m --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err; echo "bench $m rc=$?"; tail -2 gpurun_out/bench_$m.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$m.json"))
print({k:d[k] for k in ("value","ms_per_step","dtype","gpu_launches","last_loss")}, d["e2e"]["value"])
for k,v in d["roofline"]["families"].items(): print(k, v)
PY
done
