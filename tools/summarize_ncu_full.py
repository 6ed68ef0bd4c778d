#!/usr/bin/env python
"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into a markdown table of
the metrics the roofline argument uses: duration, SM clock, tensor-pipe activity, DRAM bytes, L2 / L1 throughput.
Usage: summarize_ncu_full.py raw.csv [out.md] [--labels a,b,c,...]"""
import csv
import sys

src = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
labels = None
for a in sys.argv[2:]:
    if a.startswith("--labels="):
        labels = a.split("=", 1)[1].split(",")
rows = list(csv.reader(open(src)))
hdr, data = rows[0], rows[2:]
H = {h: i for i, h in enumerate(hdr)}
cols = [("grid", "Grid Size"), ("us", "gpu__time_duration.sum"), ("SM GHz", "sm__cycles_elapsed.avg.per_second"),
        ("tensor pipe % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("UTCHMMA bf16 ops % of peak", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
        ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
        ("dram %", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("xbar->L1 read %", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("smem KB", "launch__shared_mem_per_block_dynamic")]
lines = ["| # | kernel | " + " | ".join(c[0] for c in cols) + " |", "|---|---|" + "---|" * len(cols)]
for k, r in enumerate(data):
    name = r[H["Kernel Name"]].replace("void xv::", "").split("(")[0]
    if labels and k < len(labels):
        name += " " + labels[k]
    vals = []
    for _, key in cols:
        v = r[H[key]] if key in H else ""
        try:
            f = float(v.replace(",", ""))
            v = ("%.1f" % f) if abs(f) < 1000 else ("%.0f" % f)
        except ValueError:
            pass
        vals.append(v)
    lines.append("| %d | `%s` | %s |" % (k, name, " | ".join(vals)))
text = "\n".join(lines)
print(text)
if out:
    open(out, "w").write(text + "\n")
