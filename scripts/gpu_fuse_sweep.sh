mkdir -p gpurun_out
for k in 2304 1152 576 192 0; do
AVID_FUSE_BN_BWD_MIN_K=$k timeout 600 python bench.py --steps 10 --warmup 3 --math bf16x3 --no-cpu-baseline --skip-e2e > gpurun_out/bench_fuse$k.json 2> gpurun_out/bench_fuse$k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_fuse$k.json"))
print($k, round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["families"].items() if "dgrad" in k}, d["last_loss"] if "last_loss" in d else "")
PY
done
