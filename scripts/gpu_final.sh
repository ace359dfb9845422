mkdir -p gpurun_out
T0=$(date +%s)
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$? t=$(( $(date +%s)-T0 ))"
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks')}); print('e2e',d['e2e']); print('cpu',d.get('cpu_baseline'))
r=d['roofline']; print({k:r[k] for k in ('kernel','bound','achieved','peak','frac','traffic','executed_frac','avg_launch_ms','algorithmic_flops_per_launch')}); print(r.get('nce'))"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300; echo "ref t=$(( $(date +%s)-T0 ))"
python bench.py --steps 10 --warmup 3 --math bf16 --no-cpu-baseline --skip-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bf16', round(d['value'],1), round(d['ms_per_step'],2), d['last_loss'] if 'last_loss' in d else '')"
