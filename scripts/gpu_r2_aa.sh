# warp-uniform elect.sync issue loops in every tcgen05 kernel: parity suite, layer probes, bench with conv_pair on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_aa.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_aa.log | head; tail -3 gpurun_out/pytest_aa.log
python scripts/probe_pair.py 2>&1 | grep -v "^conv_pair"
AVID_CONV_PAIR=0 python scripts/probe_conv.py 2>&1 | grep -E "^[a-z]|full \(stats\)|no epilogue work|no MMAs  "
for c in 1 0; do
AVID_CONV_PAIR=$c timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline --dump-launches gpurun_out/launches_aa_$c.txt > gpurun_out/bench_r2aa_$c.json 2> gpurun_out/bench_r2aa_$c.err; echo "bench pair=$c rc=$?"; grep -v "^$\|Warning\|warn\|run_backward" gpurun_out/bench_r2aa_$c.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2aa_$c.json').read())
print('pair=$c value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'loss', d['last_loss'], 'roof', d['roofline'].get('frac'))
PY
done
