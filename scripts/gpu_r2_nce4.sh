# NCE gather kernel: how much of the K = 1024 time is the state the L2 flush leaves behind?
fmt='import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r["bank_rows"],r["K"],r["flush"],r["launches_per_event_pair"],"%.1f us"%(1e3*r["ms_median"]),"%.3f"%r["frac_of_measured_hbm"])
    except Exception: print(l.rstrip())
'
for mode in "--flush write" "--flush read" "--flush none" "--flush none --group 8" "--flush read --group 8"; do
  python scripts/bench_nce.py --banks 2000000 $mode 2>&1 | python -c "$fmt"
done
