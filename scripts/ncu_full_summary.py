#!/usr/bin/env python
"""Markdown summary of `ncu --set full` reports (read with `ncu -i <rep> --page raw --csv`): one table per captured launch.
    python scripts/ncu_full_summary.py profiles/r1_ncu_full_summary.md gpurun_out/r1_conv_tc.ncu-rep ..."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs/thread", "launch__registers_per_thread"),
    ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("L2->SM bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor-core unit busy % (incl. operand fetch)", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots busy %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["# ncu `--set full` captures (B200, `--clock-control none`, warm launches of the bf16x3 bench step unless noted)\n",
             "Read with `ncu -i <rep> --page raw --csv`; the .ncu-rep files stay in gpurun_out/ (scratch).  One table per captured launch.\n"]
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        lines.append("\n## %s\n" % rep.split("/")[-1])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            lines.append("\n`%s`\n\n| metric | value |\n|---|---|" % name[:140])
            for label, key in METRICS:
                col = [i for i, h in enumerate(hdr) if h == key or h.endswith("." + key) or h.endswith(key)]
                if not col:
                    continue
                i = col[0]
                lines.append("| %s (`%s`) | %s %s |" % (label, key, r[i], units[i]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main()
