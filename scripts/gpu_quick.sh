# quick loop: tensor-core conv tests + tower tests + one bench line (no ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_towers_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error" gpurun_out/pytest_quick.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --math ${MATH:-bf16x3} --no-cpu-baseline --dump-launches gpurun_out/launch_table.txt > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","last_loss","clocks")}, d["e2e"]["value"])
for k,v in d["roofline"]["families"].items(): print(k, {a:round(b,3) for a,b in v.items()})
PY
