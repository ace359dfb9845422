# NCE gather kernel at K = 16384 / 4096: negatives per CTA forced through AVID_NCE_KC (tuning experiment)
fmt='import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r["bank_rows"],r["K"],"%.1f us"%(1e3*r["ms_median"]),"%.3f"%r["frac_of_measured_hbm"])
    except Exception: print(l.rstrip())
'
for kc in 0 592 912 1264 1648 2064 2736 4096; do
  echo "kc=$kc"; AVID_NCE_KC=$kc python scripts/bench_nce.py --banks 2000000 --negatives 16384 --iters 20 2>&1 | python -c "$fmt"
done
for kc in 0 176 352 528 688 1024; do
  echo "kc=$kc"; AVID_NCE_KC=$kc python scripts/bench_nce.py --banks 2000000 --negatives 4096 1024 --iters 20 2>&1 | python -c "$fmt"
done
