#!/usr/bin/env python
"""Isolated timing of the HBM-bound frame-level kernels at the config-2 shapes (B=128, T=200) through the C ABI.
Each kernel runs over a rotation of NSET buffer sets (> L2 in total) so that every launch streams from HBM; the
achieved GB/s is ALGORITHMIC bytes (tensor reads + writes) / time, against MEASURED_PEAKS.json:hbm_gbs.
    python tools/layers_bench.py [--reps 40] [--json out.json]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from tf_kaldi_speaker_b200 import _lib as L

NSET = 6


def timeit(fn, reps):
    for i in range(NSET):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        fn(i % NSET)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=36)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    lib = L.load()
    dev = "cuda"
    peak = 6552.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    B, T, valid = 128, 200, 186
    R = B * T
    st = L.stream_ptr
    nul = L.ptr(None)
    rows = []

    def report(name, us, nbytes):
        gbs = nbytes / us / 1e3
        rows.append({"kernel": name, "us": us, "algorithmic_MB": nbytes / 1e6, "GBps": gbs, "frac_of_hbm_peak": gbs / peak})
        print("%-34s %8.1f us  %7.1f MB  %7.0f GB/s  %.2f of measured HBM peak" % (name, us, nbytes / 1e6, gbs, gbs / peak),
              flush=True)

    for Cn in (512, 1536):
        g = torch.Generator(device=dev).manual_seed(Cn)
        ys = [torch.randn(R, Cn, generator=g, device=dev).to(torch.bfloat16) for _ in range(NSET)]
        das = [(torch.randn(R, Cn, generator=g, device=dev) * 1e-2).to(torch.bfloat16) for _ in range(NSET)]
        outs = [torch.empty(R, Cn, dtype=torch.bfloat16, device=dev) for _ in range(NSET)]
        scale = 1 + 0.1 * torch.randn(Cn, generator=g, device=dev)
        shift = 0.1 * torch.randn(Cn, generator=g, device=dev)
        mean = 0.1 * torch.randn(Cn, generator=g, device=dev)
        rstd = 1 + 0.1 * torch.rand(Cn, generator=g, device=dev)
        dg = torch.zeros(Cn, device=dev)
        db = torch.zeros(Cn, device=dev)
        nb = R * Cn * 2
        us = timeit(lambda i: L.check(lib.xv_bn_act_apply(L.ptr(ys[i]), L.ptr(outs[i]), L.ptr(scale), L.ptr(shift), nul, 1,
                                                          C.c_int64(R), Cn, C.c_int64(Cn), T, valid, nul, st())), args.reps)
        report("bn_act_apply C=%d" % Cn, us, 2 * nb)
        us = timeit(lambda i: L.check(lib.xv_bn_act_bwd_reduce(L.ptr(ys[i]), L.ptr(das[i]), L.ptr(scale), L.ptr(shift),
                                                               L.ptr(mean), L.ptr(rstd), nul, 1, C.c_int64(R), Cn,
                                                               C.c_int64(Cn), T, valid, nul, L.ptr(dg), L.ptr(db), nul, nul,
                                                               nul, 0, 0, st())), args.reps)
        report("bn_act_bwd_reduce C=%d" % Cn, us, 2 * nb)
        us = timeit(lambda i: L.check(lib.xv_bn_act_bwd_apply(L.ptr(ys[i]), L.ptr(das[i]), L.ptr(outs[i]), L.ptr(scale),
                                                              L.ptr(shift), L.ptr(mean), L.ptr(rstd), L.ptr(dg), L.ptr(db),
                                                              C.c_float(float(B * valid)), nul, 1, C.c_int64(R), Cn,
                                                              C.c_int64(Cn), T, valid, nul, nul, nul, 0, 0, st())), args.reps)
        report("bn_act_bwd_apply C=%d" % Cn, us, 3 * nb)
        if Cn == 1536:
            pooled = torch.randn(B, 2 * Cn, device=dev).abs() + 0.1
            dpooled = torch.randn(B, 2 * Cn, device=dev) * 1e-2
            out3 = torch.empty(B, 6 * Cn, dtype=torch.bfloat16, device=dev)
            sums = torch.empty(B, 4, Cn, device=dev)
            us = timeit(lambda i: L.check(lib.xv_stats_pool_fwd(L.ptr(ys[i]), L.ptr(pooled), L.ptr(out3), B, T, valid, nul,
                                                                1500, Cn, C.c_int64(Cn), L.ptr(scale), L.ptr(shift), nul, 1,
                                                                nul, nul, nul, st())), args.reps)
            report("stats_pool_fwd fused BN+ReLU", us, nb)
            us = timeit(lambda i: L.check(lib.xv_stats_pool_fwd(L.ptr(ys[i]), L.ptr(pooled), L.ptr(out3), B, T, valid, nul,
                                                                1500, Cn, C.c_int64(Cn), L.ptr(scale), L.ptr(shift), nul, 1,
                                                                L.ptr(mean), L.ptr(rstd), L.ptr(sums), st())), args.reps)
            report("stats_pool_fwd + bwd sums", us, nb)
            us = timeit(lambda i: L.check(lib.xv_bn_act_bwd_apply(L.ptr(ys[i]), nul, L.ptr(outs[i]), L.ptr(scale),
                                                                  L.ptr(shift), L.ptr(mean), L.ptr(rstd), L.ptr(dg),
                                                                  L.ptr(db), C.c_float(float(B * valid)), nul, 1,
                                                                  C.c_int64(R), Cn, C.c_int64(Cn), T, valid, nul,
                                                                  L.ptr(pooled), L.ptr(dpooled), Cn, 1500, st())), args.reps)
            report("bn_act_bwd_apply fused pool", us, 2 * nb)
            us = timeit(lambda i: L.check(lib.xv_pool_bn_bwd_reduce(L.ptr(pooled), L.ptr(dpooled), L.ptr(sums), B, valid,
                                                                    nul, 1500, Cn, L.ptr(dg), L.ptr(db), st())), args.reps)
            report("pool_bn_bwd_reduce", us, B * 8 * Cn * 4)
    # copy reference: what a plain bf16 copy kernel of the same size reaches here (torch)
    a = [torch.empty(R, 1536, dtype=torch.bfloat16, device=dev) for _ in range(NSET)]
    b = [torch.empty(R, 1536, dtype=torch.bfloat16, device=dev) for _ in range(NSET)]
    us = timeit(lambda i: b[i].copy_(a[i]), args.reps)
    report("torch copy_ 78.6 MB (reference)", us, 2 * R * 1536 * 2)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
