"""tcgen05 implicit-GEMM (xv_gemm_bf16) against plain PyTorch fp32 on the same bf16 operands: operand majors, taps
(conv forward / dgrad / wgrad), split-K through the TMA reduce-add, edge tiles, BN-statistic and BN-backward
epilogues.  Each case runs in its own process (tools/gemm_selftest.py): a trapped kernel poisons its CUDA context."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


def test_gemm_selftest_cases():
    import gemm_selftest as G
    results = G.run_all(timeout=240)
    bad = []
    for r in results:
        if "error" in r:
            bad.append(r)
            continue
        for k, v in r.items():
            if k.startswith("err"):
                tol = G.TOL if ("bf16" in r["case"] or k == "err" and r["case"] in
                                ("conv_fwd", "conv_fwd_k7", "conv_fwd_bias_nostats", "conv_fwd_stats_edges",
                                 "conv_fwd_stats_pairs", "dense_stats_pairs_1536", "dgrad", "persistent_big",
                                 "dgrad_bnbwd_relu", "dgrad_bnbwd_lrelu_dense")) else 2e-3
                if not (v < tol):
                    bad.append((r["case"], k, v, tol))
    assert not bad, bad
