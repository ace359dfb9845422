mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_ab.log | head; tail -3 gpurun_out/pytest_ab.log
for c in 1; do
AVID_CONV_PAIR=$c timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline --dump-launches gpurun_out/launches_ab_$c.txt > gpurun_out/bench_r2ab_$c.json 2> gpurun_out/bench_r2ab_$c.err; echo "bench pair=$c rc=$?"; grep -v "^$\|Warning\|warn\|run_backward" gpurun_out/bench_r2ab_$c.err | tail -3
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2ab_$c.json').read())
print('pair=$c value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'loss', d['last_loss'], 'roof', d['roofline'].get('frac'))
PY
done
AVID_PROFILE_ALL=1 timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-gpu-baseline --skip-e2e --dump-launches gpurun_out/launches_ab_all.txt > gpurun_out/bench_r2ab_all.json 2> gpurun_out/bench_r2ab_all.err; echo "profile-all rc=$?"
