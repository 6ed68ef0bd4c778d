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
# per-step DRAM traffic of the captured launches (bench.py reports it as roofline.traffic)
tot_r = sum(float(r[H["dram__bytes_read.sum"]]) for r in data)
tot_w = sum(float(r[H["dram__bytes_write.sum"]]) for r in data)
unit_r = rows[1][H["dram__bytes_read.sum"]]
lines.append("")
lines.append("DRAM traffic of these %d launches: read %.1f %s, write %.1f %s" % (len(data), tot_r, unit_r, tot_w,
                                                                                rows[1][H["dram__bytes_write.sum"]]))
text = "\n".join(lines)
print(text)
for a in sys.argv[2:]:
    if a.startswith("--traffic-json="):
        import json
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        json.dump({"launches": len(data), "dram_read_bytes": tot_r * scale.get(unit_r, 1e6),
                   "dram_write_bytes": tot_w * scale.get(rows[1][H["dram__bytes_write.sum"]], 1e6),
                   "source": src}, open(a.split("=", 1)[1], "w"), indent=1)
if out:
    open(out, "w").write(text + "\n")
