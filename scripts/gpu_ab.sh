mkdir -p gpurun_out
for v in twosweep; do

AVID_PROFILE_ALL=1 timeout 600 python bench.py --steps 10 --warmup 3 --math bf16x3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab_$v.json"))
print("$v", round(d["value"],1), {k:round(x["ms_per_step"],3) for k,x in d["roofline"]["families"].items() if k.startswith("filter")})
PY
done
unset AVID_FILTER_ELEMENTWISE
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --math bf16x3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_plain$i.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_plain$i.json')); print('plain', round(d['value'],1), d['ms_per_step'])"
done
