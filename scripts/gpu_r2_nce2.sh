# NCE gather kernel, round 2 second pass: parity tests, K sweep (cold and warm caches), one --set full capture at K = 1024
mkdir -p gpurun_out
python -m pytest tests/test_criterion_gpu.py tests/test_warm_start_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/bench_nce.py --out gpurun_out/r2_nce_sweep_v2.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r['bank_rows'],r['K'],'%.1f us'%(1e3*r['ms_median']),'%.3f'%r['frac_of_measured_hbm'])
    except Exception: print(l.rstrip())
"
echo warm; python scripts/bench_nce.py --banks 2000000 --flush none 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r['bank_rows'],r['K'],'%.1f us'%(1e3*r['ms_median']),'%.3f'%r['frac_of_measured_hbm'])
    except Exception: print(l.rstrip())
"
ncu --set full --clock-control none --import-source on -k regex:nce_gather_kernel -s 3 -c 1 -o gpurun_out/r2_nce_k1024_v2 -f python scripts/bench_nce.py --banks 2000000 --negatives 1024 --iters 3 --warmup 2 > gpurun_out/r2_ncu_nce_v2.log 2>&1
ls -la gpurun_out/r2_nce_k1024_v2.ncu-rep
