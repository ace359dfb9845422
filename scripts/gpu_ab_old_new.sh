# same-box A/B of the whole step: the tree at the start of this session (_ab_old/, commit c8bdb3e) vs the working tree, interleaved
mkdir -p gpurun_out
for rep in 1 2; do
  for side in old new; do
    if [ $side = old ]; then d=_ab_old; else d=.; fi
    (cd $d && timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | tail -1) > gpurun_out/ab_${side}_${rep}.json
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${side}_${rep}.json").read().strip().splitlines()[-1])
print("$side $rep", round(d["value"],1), "clips/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), d["clocks"])
PY
  done
done
