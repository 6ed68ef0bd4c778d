#!/usr/bin/env python
"""CUPTI per-kernel durations of the captured config-4 step (self-attention pooling, key net 1500-1500 tanh, H = 1, T = 200).
    python tools/step_kernel_times_c4.py [steps] [out.md]"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from torch.profiler import ProfilerActivity, profile

import config_bench as CB
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
out = sys.argv[2] if len(sys.argv) > 2 else None
H, T = 1, 200
pd = dict(CB.BASE, **CB.MARGIN, pooling_type="self_attention", amsoftmax_m=0.2, feature_norm=True, feature_scaling_factor=30,
          att_key_input="tdnn4_relu", att_key_num_nodes=[1500, 1500], att_key_network_type=3, att_value_input="tdnn5_relu",
          att_value_num_nodes=[], att_value_network_type=0, att_apply_nonlinear=False, att_use_scale=True, att_num_heads=H,
          att_split_key=(H > 1), att_penalty_term=(0.01 if H > 1 else 0.0))
tr = Trainer(ParamsPlain(**pd), "/tmp/xv_profile_c4")
tr.build("train", 30, "additive_margin_softmax", 7200)
x, y = CB.batch(128, T, 30, 7200)
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
        t = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        name = ev.name.split("(")[0][:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        seq.append((name, t))
tot = sum(v[1] for v in agg.values())
lines = ["| kernel | launches/step | us/step (in graph, CUPTI) | share |", "|---|---|---|---|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append("| `%s` | %.1f | %.1f | %.1f%% |" % (n, c / steps, t / steps, 100 * t / tot))
lines.append("| **total kernel time** | %.1f | %.1f | 100%% |" % (sum(v[0] for v in agg.values()) / steps, tot / steps))
print("\n".join(lines))
n_per = len(seq) // steps
print("\nlast step sequence:")
for name, t in seq[-n_per:]:
    print("  %-64s %8.1f us" % (name[:64], t))
if out:
    open(out, "w").write("\n".join(lines) + "\n")
