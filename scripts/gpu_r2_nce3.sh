# NCE gather kernel: parity tests + K sweep
mkdir -p gpurun_out
python -m pytest tests/test_criterion_gpu.py tests/test_warm_start_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/bench_nce.py --out gpurun_out/r2_nce_sweep_v3.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r['bank_rows'],r['K'],'%.1f us'%(1e3*r['ms_median']),'%.3f'%r['frac_of_measured_hbm'])
    except Exception: print(l.rstrip())
"
