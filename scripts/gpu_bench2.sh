mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv,noheader
for i in 1 2 3; do
timeout 600 python bench.py --steps 20 --warmup 5 --math bf16x3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_rep$i.json 2> gpurun_out/bench_rep$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_rep$i.json"))
print({k:d[k] for k in ("value","ms_per_step","clocks")})
print({k:round(v["ms_per_step"],3) for k,v in d["roofline"]["families"].items()})
PY
done
