mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_criterion_gpu.py tests/test_sharded_bank.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_nce.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|assert|Error|error" gpurun_out/pytest_nce.log | head -20
timeout 300 python scripts/bench_nce.py --banks 2000000 --out gpurun_out/nce_sweep.json 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['bank_rows'], d['K'], round(d['ms_median']*1e3,1), 'us', round(d['GB/s']), 'GB/s', round(d['frac_of_measured_hbm'],3))
"
