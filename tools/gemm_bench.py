#!/usr/bin/env python
"""Isolated timing of the tcgen05 implicit-GEMM on the contraction shapes of the config-2 step, next to cuBLAS
(torch.matmul, bf16) on a plain GEMM of the same M/N/K.  CUDA events around REPS back-to-back launches after a
warm-up; operands larger than they look because every launch re-reads them through L2.
    python tools/gemm_bench.py [--reps 20] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from tf_kaldi_speaker_b200 import _lib as L


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = "cuda"
    B, T = 128, 200
    R = B * T
    g = torch.Generator(device=dev).manual_seed(0)
    mk = lambda *shape: (torch.randn(*shape, generator=g, device=dev) * 0.1).to(torch.bfloat16)
    rows = []

    def run(name, fn, M, N, K):
        us = timeit(fn, args.reps)
        a, b = mk(M, K), mk(K, N)
        us_ref = timeit(lambda: torch.matmul(a, b), args.reps)
        fl = 2.0 * M * N * K
        rows.append({"case": name, "M": M, "N": N, "K": K, "us": us, "tflops": fl / us / 1e6, "cublas_us": us_ref,
                     "cublas_tflops": fl / us_ref / 1e6})
        print("%-22s M=%6d N=%5d K=%6d  ours %7.1f us %7.1f TF/s | cuBLAS %7.1f us %7.1f TF/s"
              % (name, M, N, K, us, fl / us / 1e6, us_ref, fl / us_ref / 1e6), flush=True)

    for name, k, cin, cout in (("tdnn1 fwd", 1, 192, 512), ("tdnn2 fwd", 5, 512, 512), ("tdnn3 fwd", 7, 512, 512),
                               ("tdnn4 fwd", 1, 512, 512), ("tdnn5 fwd", 1, 512, 1536)):
        x, w = mk(R, cin), mk(k * cin, cout)
        y = torch.empty(R, cout, dtype=torch.bfloat16, device=dev)
        st = torch.zeros(2, cout, device=dev)
        a_op = L.operand(x, False, div=(cin if k > 1 else 0), tap_rows=(1 if k > 1 else 0))
        # as in the training step: no bias (BN follows), BN statistics of the stored tensor from the epilogue
        run(name + " +stats", lambda: L.gemm(a_op, L.operand(w, True), R, cout, k * cin, y, epilogue=L.EPI_BF16,
                                             col_sum=st[0], col_sumsq=st[1], seg_len=T, seg_valid=T - 14),
            R, cout, k * cin)
        if k == 1:
            run(name + " plain", lambda: L.gemm(a_op, L.operand(w, True), R, cout, k * cin, y, epilogue=L.EPI_BF16),
                R, cout, k * cin)
        if name in ("tdnn1 fwd",):
            continue
        dy = mk(R, cout)
        dx = torch.empty(R, cin, dtype=torch.bfloat16, device=dev)
        run(name.replace("fwd", "dgrad"),
            lambda: L.gemm(L.operand(dy, False, div=(cout if k > 1 else 0), tap_rows=(-1 if k > 1 else 0)),
                           L.operand(w, False, div=(cout if k > 1 else 0), tap_rows=(cin if k > 1 else 0)), R, cin,
                           k * cout, dx, epilogue=L.EPI_BF16), R, cin, k * cout)
        if cin == 512:      # the same dgrad with the fused BN-backward reductions of the producer layer in its epilogue
            yprev = mk(R, cin)
            cst = [torch.rand(cin, device=dev) + 0.5 for _ in range(4)]
            acc2 = torch.zeros(2, cin, device=dev)
            run(name.replace("fwd", "dgrad+bnbwd"),
                lambda: L.gemm(L.operand(dy, False, div=(cout if k > 1 else 0), tap_rows=(-1 if k > 1 else 0)),
                               L.operand(w, False, div=(cout if k > 1 else 0), tap_rows=(cin if k > 1 else 0)), R, cin,
                               k * cout, dx, epilogue=L.EPI_BF16, col_sum=acc2[0], col_sumsq=acc2[1],
                               bn_bwd=(yprev, cst[0], cst[1], cst[2], cst[3], 0.0)), R, cin, k * cout)
        gw = torch.zeros(k * cin, cout, device=dev)
        tiles = ((k * cin + 255) // 256) * ((cout + 255) // 256)
        splits = max(1, min(32, 74 // tiles))
        run(name.replace("fwd", "wgrad") + " s%d" % splits,
            lambda: L.gemm(L.operand(x, True, div=(cin if k > 1 else 0), tap_rows=(1 if k > 1 else 0)),
                           L.operand(dy, True), k * cin, cout, R, gw, epilogue=L.EPI_F32, splits=splits), k * cin, cout, R)
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
