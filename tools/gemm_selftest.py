#!/usr/bin/env python
"""GPU self-test of the tcgen05 implicit-GEMM (xv_gemm_bf16) against a plain PyTorch fp32 reference.

Each case runs in its own subprocess (a trapped kernel poisons the CUDA context) with a timeout, so one
bad descriptor cannot hang or hide the other cases.  Used by tests/test_gemm_gpu.py and by hand:
    python tools/gemm_selftest.py            # all cases, summary table
    python tools/gemm_selftest.py --case conv_fwd
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _mk(shape, gen, scale=1.0):
    import torch
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def _err(out, ref):
    import torch
    out = out.float()
    ref = ref.float()
    denom = ref.abs().max().item() + 1e-20
    return (out - ref).abs().max().item() / denom


def case_plain(M, N, K, a_mn, b_mn, epi, splits=1):
    import torch
    from tf_kaldi_speaker_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(1)
    A = _mk((K, M) if a_mn else (M, K), g)
    B = _mk((K, N) if b_mn else (N, K), g)
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float().t() if b_mn else B.float()
    ref = Af @ Bf.t()
    ldc = (N + 7) // 8 * 8
    if epi == L.EPI_F32:
        out = torch.zeros(M, ldc, device="cuda", dtype=torch.float32)
    else:
        out = torch.zeros(M, ldc, device="cuda", dtype=torch.bfloat16)
    L.gemm(L.operand(A, a_mn), L.operand(B, b_mn), M, N, K, out, epilogue=epi, splits=splits)
    torch.cuda.synchronize()
    return {"err": _err(out[:, :N], ref), "pad_clean": float(out[:, N:].float().abs().max().item()) if ldc > N else 0.0}


def case_conv_fwd(B_=3, T=40, cin=128, cout=512, k=5, stats=True):
    """Forward temporal conv as implicit GEMM: A taps +1, B = kernel [k*cin, cout] MN-major, bias + column stats."""
    import torch
    from tf_kaldi_speaker_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(2)
    x = _mk((B_ * T, cin), g)
    w = _mk((k * cin, cout), g, 0.05)
    bias = torch.randn(cout, generator=g, device="cuda")
    R = B_ * T
    y = torch.zeros(R, cout, device="cuda", dtype=torch.bfloat16)
    cs = torch.zeros(cout, device="cuda")
    cq = torch.zeros(cout, device="cuda")
    valid = T - (k - 1)
    # with statistics the layer passes no bias (the stored pre-BN tensor is bias-free; the sums are those of the STORED
    # bf16 values over the valid rows)
    L.gemm(L.operand(x, False, div=cin, tap_rows=1), L.operand(w, True), R, cout, k * cin, y,
           epilogue=L.EPI_BF16, bias=None if stats else bias, col_sum=cs if stats else None,
           col_sumsq=cq if stats else None, seg_len=T, seg_valid=valid)
    torch.cuda.synchronize()
    xf = x.float().reshape(B_, T, cin)
    wf = w.float().reshape(k, cin, cout)
    ref = torch.zeros(B_, valid, cout, device="cuda")
    for j in range(k):
        ref += xf[:, j:j + valid] @ wf[j]
    yv = y.float().reshape(B_, T, cout)[:, :valid]
    res = {"err": _err(yv, ref if stats else ref + bias)}
    if stats:
        res["err_sum"] = _err(cs, yv.sum((0, 1)))
        res["err_sumsq"] = _err(cq, (yv ** 2).sum((0, 1)))
    return res


def case_dgrad(B_=3, T=40, cin=256, cout=512, k=5):
    """dX[q] = sum_j dY[q-j] W_j^T : A = dY K-major taps -1, B = kernel [k*cin, cout] K-major with tap rows."""
    import torch
    from tf_kaldi_speaker_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(3)
    R = B_ * T
    valid = T - (k - 1)
    dy = _mk((R, cout), g)
    dyv = dy.reshape(B_, T, cout).clone()
    dyv[:, valid:] = 0          # invalid rows of dY are forced to zero by the BN-backward kernel
    dy = dyv.reshape(R, cout).contiguous()
    w = _mk((k * cin, cout), g, 0.05)
    dx = torch.zeros(R, cin, device="cuda", dtype=torch.bfloat16)
    L.gemm(L.operand(dy, False, div=cout, tap_rows=-1), L.operand(w, False, div=cout, tap_rows=cin),
           R, cin, k * cout, dx, epilogue=L.EPI_BF16)
    torch.cuda.synchronize()
    wf = w.float().reshape(k, cin, cout)
    dyf = dy.float().reshape(B_, T, cout)
    ref = torch.zeros(B_, T, cin, device="cuda")
    for j in range(k):
        ref[:, j:] += dyf[:, :T - j] @ wf[j].t()
    return {"err": _err(dx.float().reshape(B_, T, cin), ref)}


def case_dgrad_bnbwd(B_=4, T=70, cin=512, cout=512, k=5, neg_slope=0.0):
    """dgrad whose epilogue also accumulates the BN-backward reductions of the producer layer:
    g = dX * act'(y*scale + shift); dbeta = sum g; dgamma = sum g * (y - mean) * rstd."""
    import torch
    from tf_kaldi_speaker_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    R = B_ * T
    valid = T - (k - 1)
    dy = _mk((R, cout), g)
    dyv = dy.reshape(B_, T, cout).clone()
    dyv[:, valid:] = 0
    dy = dyv.reshape(R, cout).contiguous()
    w = _mk((k * cin, cout), g, 0.05)
    y = _mk((R, cin), g)
    scale = 0.5 + torch.rand(cin, generator=g, device="cuda")
    scale[::7] *= -1           # negative gamma: the activation mask must follow the sign of z, not of y
    shift = 0.3 * torch.randn(cin, generator=g, device="cuda")
    mean = 0.2 * torch.randn(cin, generator=g, device="cuda")
    rstd = 0.5 + torch.rand(cin, generator=g, device="cuda")
    dbeta = torch.zeros(cin, device="cuda")
    dgamma = torch.zeros(cin, device="cuda")
    dx = torch.zeros(R, cin, device="cuda", dtype=torch.bfloat16)
    L.gemm(L.operand(dy, False, div=cout, tap_rows=-1), L.operand(w, False, div=cout, tap_rows=cin),
           R, cin, k * cout, dx, epilogue=L.EPI_BF16, col_sum=dbeta, col_sumsq=dgamma,
           bn_bwd=(y, scale, shift, mean, rstd, neg_slope))
    torch.cuda.synchronize()
    wf = w.float().reshape(k, cin, cout)
    dyf = dy.float().reshape(B_, T, cout)
    ref = torch.zeros(B_, T, cin, device="cuda")
    for j in range(k):
        ref[:, j:] += dyf[:, :T - j] @ wf[j].t()
    ref = ref.reshape(R, cin)
    z = y.float() * scale + shift
    # the reductions use the gradient as it is STORED (bf16), like the stand-alone xv_bn_act_bwd_reduce kernel
    gg = ref.to(torch.bfloat16).float() * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, neg_slope))
    return {"err": _err(dx.float(), ref), "err_dbeta": _err(dbeta, gg.sum(0)),
            "err_dgamma": _err(dgamma, (gg * (y.float() - mean) * rstd).sum(0))}


def case_wgrad(B_=4, T=64, cin=128, cout=512, k=5, splits=3):
    """dW_j = sum_r X[r+j]^T dY[r] : A = X MN-major (div=cin, tap +1), B = dY MN-major, split-K atomics."""
    import torch
    from tf_kaldi_speaker_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(4)
    R = B_ * T
    valid = T - (k - 1)
    x = _mk((R, cin), g)
    dy = _mk((R, cout), g)
    dyv = dy.reshape(B_, T, cout).clone()
    dyv[:, valid:] = 0
    dy = dyv.reshape(R, cout).contiguous()
    dw = torch.zeros(k * cin, cout, device="cuda", dtype=torch.float32)
    L.gemm(L.operand(x, True, div=cin, tap_rows=1), L.operand(dy, True), k * cin, cout, R, dw,
           epilogue=L.EPI_F32, splits=splits)
    torch.cuda.synchronize()
    xf = x.float().reshape(B_, T, cin)
    dyf = dy.float().reshape(B_, T, cout)
    ref = torch.zeros(k, cin, cout, device="cuda")
    for j in range(k):
        ref[j] = torch.einsum("btc,btn->cn", xf[:, j:j + valid], dyf[:, :valid])
    return {"err": _err(dw.reshape(k, cin, cout), ref)}


def _cases():
    from tf_kaldi_speaker_b200 import _lib as L
    return {
        "kk_f32": lambda: case_plain(256, 512, 256, False, False, L.EPI_F32),
        "kk_bf16": lambda: case_plain(256, 512, 256, False, False, L.EPI_BF16),
        "k_mn_f32": lambda: case_plain(256, 512, 256, False, True, L.EPI_F32),
        "mn_k_f32": lambda: case_plain(256, 512, 256, True, False, L.EPI_F32),
        "mn_mn_f32": lambda: case_plain(256, 512, 256, True, True, L.EPI_F32),
        "mn_mn_split": lambda: case_plain(384, 512, 2048, True, True, L.EPI_F32, splits=5),
        "edges_f32": lambda: case_plain(200, 1000, 200, False, False, L.EPI_F32),
        "edges_bf16_mn": lambda: case_plain(130, 1000, 192, False, True, L.EPI_BF16),
        "k1_tiny": lambda: case_plain(128, 256, 64, False, False, L.EPI_F32),
        "persistent_big": lambda: case_plain(25600, 512, 512, False, True, L.EPI_BF16),
        "conv_fwd": lambda: case_conv_fwd(),
        "conv_fwd_k7": lambda: case_conv_fwd(B_=5, T=100, cin=512, cout=512, k=7),
        "conv_fwd_bias_nostats": lambda: case_conv_fwd(stats=False),
        "conv_fwd_stats_edges": lambda: case_conv_fwd(B_=3, T=45, cin=128, cout=1000, k=5),
        "conv_fwd_stats_pairs": lambda: case_conv_fwd(B_=64, T=200, cin=128, cout=512, k=5),
        "dense_stats_pairs_1536": lambda: case_conv_fwd(B_=128, T=200, cin=512, cout=1536, k=1),
        "dgrad": lambda: case_dgrad(),
        "dgrad_bnbwd_relu": lambda: case_dgrad_bnbwd(),
        "dgrad_bnbwd_lrelu_dense": lambda: case_dgrad_bnbwd(B_=3, T=50, cin=512, cout=1536, k=1, neg_slope=0.2),
        "wgrad": lambda: case_wgrad(),
        "wgrad_nosplit": lambda: case_wgrad(splits=1),
    }


TOL = 2e-2   # bf16 outputs; fp32 outputs are far tighter (asserted separately below)


def run_case(name):
    res = _cases()[name]()
    res["case"] = name
    return res


def run_all(timeout=180):
    from tf_kaldi_speaker_b200 import _lib as L   # noqa: F401  (fail early if the .so is missing)
    names = list(_cases().keys())
    results = []
    for n in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", n], capture_output=True,
                               text=True, timeout=timeout)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                results.append(json.loads(line[-1]))
            else:
                results.append({"case": n, "error": (r.stderr or r.stdout)[-600:]})
        except subprocess.TimeoutExpired:
            results.append({"case": n, "error": "timeout"})
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    args = ap.parse_args()
    if args.case:
        print(json.dumps(run_case(args.case)))
    else:
        allr = run_all()
        bad = 0
        for r in allr:
            ok = "error" not in r and all(v < TOL for k, v in r.items() if k.startswith("err"))
            bad += 0 if ok else 1
            print(("PASS " if ok else "FAIL ") + json.dumps(r))
        print("gemm_selftest: %d/%d cases passed" % (len(allr) - bad, len(allr)))
        sys.exit(1 if bad else 0)
