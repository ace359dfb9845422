for c in 0 2 1; do
AVID_EW_CTAS_PER_SM=$c AVID_WGRAD_STREAM=1 timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2w_$c.json 2> gpurun_out/bench_r2w_$c.err; echo "bench ew_ctas=$c rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2w_$c.json').read())
print('ew_ctas=$c wgrad_stream=1 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2))
PY
done
AVID_EW_CTAS_PER_SM=2 AVID_WGRAD_STREAM=0 timeout 900 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2w_x.json 2> gpurun_out/bench_r2w_x.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2w_x.json').read()); print('ew_ctas=2 wgrad_stream=0 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2))"
