#!/usr/bin/env python
"""Run a few EAGER (un-graphed) config-2 training steps so ncu can attribute every launch.
Usage: profile_step.py [steps] [T]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pd = dict(bench.PD)
pd["cuda_graph"] = False
tr = Trainer(ParamsPlain(**pd), "/tmp/xv_profile_model")
tr.build("train", bench.D, bench.LOSS, bench.C)
x, y = bench.synthetic_batch(bench.B_PER_GPU, 100)
x, y = x.cuda(), y.cuda()
for i in range(steps):
    tr.train_step(x, y, 0.01, i, fetch_loss=False)
torch.cuda.synchronize()
print("profile_step: %d eager steps, %d launches" % (steps, tr.engine.launches))
