#!/usr/bin/env python
"""Per-kernel summary of an `ncu --csv` launch list (one row per launch x metric):
    python scripts/ncu_summarise.py gpurun_out/launches.csv profiles/r1_bf16x3_kernels [--steps 2]
writes <out>.csv (kernel, launches, total_us, share, dram MB read/written per launch, time-weighted tensor-pipe %) and <out>.json."""
import collections
import csv
import json
import re
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 1
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    c_id, c_name, c_metric, c_unit, c_val = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= c_val:
            continue
        d = launches.setdefault(r[c_id], {"name": re.sub(r"\(.*", "", r[c_name]).replace("void ", "").strip()})
        v = float(r[c_val].replace(",", ""))
        unit = r[c_unit]
        m = r[c_metric]
        if m == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "s": v * 1e6, "second": v * 1e6}[unit]
        elif m.startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[m] = v
    agg = collections.OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tensor_us": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["us"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_us"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0
    total = sum(a["us"] for a in agg.values())
    table = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        table.append({"kernel": name, "launches_per_step": a["launches"] / steps, "ms_per_step": a["us"] / 1e3 / steps, "share": a["us"] / total,
                      "dram_read_MB_per_launch": a["rd"] / 1e6 / a["launches"], "dram_write_MB_per_launch": a["wr"] / 1e6 / a["launches"],
                      "dram_GBps": (a["rd"] + a["wr"]) / 1e9 / (a["us"] * 1e-6) if a["us"] else 0.0,
                      "tensor_pipe_active_pct": 100.0 * a["tensor_us"] / a["us"] if a["us"] else 0.0})
    with open(out + ".csv", "w") as f:
        f.write("# %s: %d launches, %.2f ms per step under ncu (serialised, cold-cache: compare shares)\n" % (src, len(launches), total / 1e3 / steps))
        w = csv.DictWriter(f, fieldnames=list(table[0].keys()))
        w.writeheader()
        for t in table:
            w.writerow({k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()})
    json.dump({"source": src, "steps": steps, "ms_per_step_under_ncu": total / 1e3 / steps, "kernels": table}, open(out + ".json", "w"), indent=1)
    for t in table[:16]:
        print("%-44s n=%6.1f %8.3f ms %5.1f%%  dram %7.1f/%7.1f MB  %6.0f GB/s  tensor %5.1f%%" % (
            t["kernel"][:44], t["launches_per_step"], t["ms_per_step"], 100 * t["share"], t["dram_read_MB_per_launch"], t["dram_write_MB_per_launch"],
            t["dram_GBps"], t["tensor_pipe_active_pct"]))


if __name__ == "__main__":
    main()
