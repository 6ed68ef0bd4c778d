#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one B200 (the headline config 2 at T=200 is bench.py):
  C1  softmax, 1000 speakers, batch 64              C2  AAM at T = 300 / 400
  C3  A-softmax m=4, 23-dim, 4300 speakers, momentum   C4  self-attention pooling (+AM), T = 200 / 400, H = 1 / 4
  C5  extraction of ragged utterances U[25, 10000] frames (frames/s, utterances/s) through extract_embeddings
Device-resident inputs, CUDA-graphed step, CUDA events; algorithmic FLOPs per SURVEY 8(d).
    python tools/config_bench.py [--json out.json] [--steps 50]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer
from tf_kaldi_speaker_b200.runtime import set_engine, Engine

BASE = dict(seed=0, network_type="tdnn", last_layer_no_bn=False, last_layer_linear=True, feature_norm=False,
            pooling_type="statistics_pooling", embedding_node="tdnn6_dense", learning_rate=0.01, use_nesterov=False,
            clip_gradient=False, clip_gradient_norm=3, weight_l2_regularizer=1e-2, batchnorm_momentum=0.99)
MARGIN = {}
for pre in ("asoftmax", "amsoftmax", "arcsoftmax"):
    MARGIN.update({pre + "_lambda_min": 0, pre + "_lambda_base": 1000, pre + "_lambda_gamma": 1e-5, pre + "_lambda_power": 5})


def flops_fwd(t, d, c):
    return 2 * 512 * (5 * d * (t - 4) + 2560 * (t - 8) + (3584 + 512 + 1500) * (t - 14)) + 2 * (3000 * 512 + 512 * 512 + 512 * c)


def batch(b, t, d, c, seed=1):
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(b, 1, d, generator=g)
    s = 0.5 + torch.rand(b, 1, d, generator=g)
    return (m + s * torch.randn(b, t, d, generator=g)).cuda(), torch.randint(0, c, (b,), generator=g, dtype=torch.int32).cuda()


def train_case(name, pd, loss, B, T, D, C, steps, lr=0.01, extra_flops_fwd=0.0):
    set_engine(Engine())
    tr = Trainer(ParamsPlain(**pd), "/tmp/xv_cfg_bench")
    tr.build("train", D, loss, C)
    x, y = batch(B, T, D, C)
    for i in range(6):
        tr.train_step(x, y, lr, i, fetch_loss=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.train_step(x, y, lr, 10 + i, fetch_loss=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ff = flops_fwd(T, D, C) + extra_flops_fwd
    ftrain = 3 * ff - 2 * 512 * 5 * D * (T - 4)
    loss_v = tr.train_step(x, y, lr, 10 + steps, fetch_loss=True)["raw_loss"]
    r = {"case": name, "batch": B, "frames": T, "dim": D, "speakers": C, "ms_per_step": ms, "segments_per_s": B / ms * 1e3,
         "algorithmic_tflops": B * ftrain / ms / 1e9, "raw_loss": loss_v}
    print(json.dumps(r), flush=True)
    return r


def extraction_case(steps_utts=96):
    from tf_kaldi_speaker_b200.extract import extract_embeddings
    set_engine(Engine())
    pd = dict(BASE)
    tr = Trainer(ParamsPlain(**pd), "/tmp/xv_cfg_bench_x")
    tr.build("predict", 30)
    rng = np.random.RandomState(0)
    lens = rng.randint(25, 10001, size=steps_utts)
    utts = [("utt%04d" % i, rng.randn(int(t), 30).astype(np.float32)) for i, t in enumerate(lens)]
    extract_embeddings(tr, utts)                # warm-up pass: pinned staging buffers and workspaces of the final sizes
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = extract_embeddings(tr, utts)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    frames = int(lens.sum())
    r = {"case": "C5 extraction (host arrays in, embeddings out; wall clock incl. H2D/D2H)", "utterances": len(out),
         "frames": frames, "seconds": dt, "frames_per_s": frames / dt, "utterances_per_s": len(out) / dt,
         "algorithmic_tflops": frames * 8.5e6 / dt / 1e12}
    print(json.dumps(r), flush=True)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    rows = []
    pd = dict(BASE, loss_func="softmax", last_layer_linear=False)
    rows.append(train_case("C1 softmax B=64", pd, "softmax", 64, 200, 30, 1000, a.steps))
    for T in (200, 300, 400):
        pd = dict(BASE, **MARGIN, feature_norm=True, feature_scaling_factor=64, arcsoftmax_m=0.2)
        rows.append(train_case("C2 AAM T=%d" % T, pd, "additive_angular_margin_softmax", 128, T, 30, 7200, a.steps))
    pd = dict(BASE, **MARGIN)
    pd.update(asoftmax_m=4, asoftmax_lambda_min=10, optimizer="momentum", momentum=0.9)
    rows.append(train_case("C3 A-softmax m=4 momentum", pd, "asoftmax", 128, 200, 23, 4300, a.steps, lr=1e-3))
    for T, H in ((200, 1), (400, 1), (200, 4), (1000, 4)):
        pd = dict(BASE, **MARGIN, pooling_type="self_attention", amsoftmax_m=0.2, feature_norm=True, feature_scaling_factor=30,
                  att_key_input="tdnn4_relu", att_key_num_nodes=[1500, 1500], att_key_network_type=3,
                  att_value_input="tdnn5_relu", att_value_num_nodes=[], att_value_network_type=0,
                  att_apply_nonlinear=False, att_use_scale=True, att_num_heads=H, att_split_key=(H > 1),
                  att_penalty_term=(0.01 if H > 1 else 0.0))
        B = 128 if T <= 400 else 64
        extra = 2.0 * (T - 14) * (512 * 1500 + 1500 * 1500)
        rows.append(train_case("C4 attention H=%d T=%d" % (H, T), pd, "additive_margin_softmax", B, T, 30, 7200,
                               max(10, a.steps // 2), extra_flops_fwd=extra))
    rows.append(extraction_case())
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
