#!/usr/bin/env python
"""In-situ per-kernel durations of the CUDA-graphed config-2 step (CUPTI through torch.profiler): unlike the ncu launch
list these are warm-cache, back-to-back times, i.e. what each kernel really costs inside the step.
    python tools/step_kernel_times.py [steps] [out.md]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
out = sys.argv[2] if len(sys.argv) > 2 else None
tr = Trainer(ParamsPlain(**dict(bench.PD)), "/tmp/xv_profile_model")
tr.build("train", bench.D, bench.LOSS, bench.C)
x, y = bench.synthetic_batch(bench.B_PER_GPU, 100)
x, y = x.cuda(), y.cuda()
for i in range(6):
    tr.train_step(x, y, 0.01, i, fetch_loss=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        tr.train_step(x, y, 0.01, 10 + i, fetch_loss=False)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
seq = []
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        name = ev.name.split("(")[0][:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        seq.append((name, ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total))
tot = sum(v[1] for v in agg.values())
lines = ["| kernel | launches/step | us/step (in graph, CUPTI) | share |", "|---|---|---|---|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| `%s` | %.1f | %.1f | %.1f%% |" % (n, c / steps, t / steps, 100 * t / tot))
lines.append("| **total kernel time** | %.1f | %.1f | 100%% |" % (sum(v[0] for v in agg.values()) / steps, tot / steps))
text = "\n".join(lines)
print(text)
# per-launch sequence of the last step (GEMM launches in order)
n_per = len(seq) // steps
print("\nlast step sequence:")
for name, t in seq[-n_per:]:
    print("  %-60s %8.1f us" % (name[:60], t))
if out:
    open(out, "w").write(text + "\n")
