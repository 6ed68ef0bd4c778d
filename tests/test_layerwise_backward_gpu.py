"""Mask-agnostic check of the frame-level backward pass INSIDE a full training step.

End-to-end gradients of a bf16-activation pipeline differ from fp64 by 6-16 % (ReLU-mask flips of the units nearest
zero), which says nothing about kernel quality.  What is tight: given the tensors the step itself
STORED (pre-BN outputs y, activations a, upstream gradients da, bf16 kernels), every backward kernel must reproduce the
fp64 formulas of SURVEY Appendix A.1 to rounding -- the mask is then the pipeline's own, so no flip enters:

    g = da * 1[y*scale + shift > 0]        dbeta = sum g         dgamma = sum g * (y - mean) * rstd      (<= 1e-3, fp32)
    dy = scale * (g - dbeta/n - yhat*dgamma/n)                                                            (<= 1e-2, bf16)
    dW_j = sum_r a_prev[r+j]^T dy[r]                                                                     (<= 1e-3, fp32)
    dx[q] = sum_j dy[q-j] W_j^T                                                                          (<= 1e-2, bf16)

for tdnn2 / tdnn3 / tdnn4 (tdnn5's BN backward is fused with the pooling gradient and covered by
tests/test_kernels_gpu.py::test_stats_pool_forward_backward; its wgrad / dgrad are checked here from its stored dy).
Covers the dgrad-fused BN reductions (tdnn1, tdnn2, tdnn4), the stand-alone reduce kernel (tdnn3), split-K wgrads through
the TMA reduce-add and the implicit-GEMM tap addressing at a shape with several row tiles and ragged segment ends."""
import pytest
import torch

from tests.xv_testlib import base_params, head_params, make_batch

pytestmark = pytest.mark.gpu

AAM = "additive_angular_margin_softmax"


def _ws(eng, name):
    hits = [t for (n, shape, dt), t in ((k, v) for k, v in eng.ws.items() if len(k) == 3) if n == name]
    assert len(hits) == 1, (name, len(hits))
    return hits[0]


def _relfro(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


@pytest.mark.parametrize("B,T", [(16, 80), (40, 203)])
def test_frame_backward_on_stored_tensors(B, T):
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    D, C = 30, 400
    pd = base_params(**head_params(AAM))
    pd.update(feature_norm=True, feature_scaling_factor=64)
    tr = Trainer(ParamsPlain(**pd), "/tmp/xv_layerwise_model")
    tr.build("train", D, AAM, C)
    eng, st = tr.engine, tr.engine.store
    g = torch.Generator().manual_seed(9)
    vals = st.export_tf()
    for k in vals:       # non-trivial BN affine parameters (negative gammas included: the mask follows z, not y)
        if k.endswith("/gamma"):
            vals[k] = vals[k] * (0.3 + torch.rand(vals[k].shape, generator=g).numpy()) * \
                torch.where(torch.rand(vals[k].shape, generator=g) < 0.1, -1.0, 1.0).numpy()
        elif k.endswith("/beta"):
            vals[k] = vals[k] + 0.3 * torch.randn(vals[k].shape, generator=g).numpy()
    st.load_tf(vals)
    x, y = make_batch(B, T, D, C, seed=4)
    tr.forward_backward(x, y, 20000)
    torch.cuda.synchronize()
    R = B * T
    valid = {1: T - 4, 2: T - 8, 3: T - 14, 4: T - 14, 5: T - 14}
    taps = {2: 5, 3: 7, 4: 1, 5: 1}
    t_idx = torch.arange(R, device="cuda") % T
    worst = {}
    for Ln in (2, 3, 4, 5):
        name, kind = "tdnn%d" % Ln, ("conv" if Ln <= 3 else "dense")
        k = taps[Ln]
        vmask = (t_idx < valid[Ln]).double().unsqueeze(1)
        n = float(B * valid[Ln])
        a_prev = _ws(eng, "tdnn%d/a" % (Ln - 1)).double()                 # [R, 512], zero on its invalid rows
        dy = _ws(eng, name + "/dy").double()                              # [R, cout_pad]
        W = st.shadow_view("tdnn/%s_%s/kernel" % (name, kind)).double()   # [k*512, cout_pad] bf16 copy the GEMMs read
        cin = a_prev.shape[1]
        if Ln != 5:
            yv = _ws(eng, name + "/y").double()
            da = _ws(eng, name + "/a/grad").double()
            scale, shift = _ws(eng, name + "/scale").double(), _ws(eng, name + "/shift").double()
            mean, rstd = _ws(eng, name + "/save_mean").double(), _ws(eng, name + "/save_rstd").double()
            z = yv * scale + shift
            gg = da * (z > 0).double() * vmask
            yhat = (yv - mean) * rstd
            dbeta_ref, dgamma_ref = gg.sum(0), (gg * yhat).sum(0)
            dbeta = st.grad("tdnn/%s_bn/beta" % name).double()
            dgamma = st.grad("tdnn/%s_bn/gamma" % name).double()
            worst[name + " dbeta"] = _relfro(dbeta, dbeta_ref)
            worst[name + " dgamma"] = _relfro(dgamma, dgamma_ref)
            assert worst[name + " dbeta"] <= 1e-3 and worst[name + " dgamma"] <= 1e-3, worst
            dy_ref = scale * (gg - dbeta / n - yhat * dgamma / n) * vmask
            e = float((dy - dy_ref).abs().max() / dy_ref.abs().max())
            worst[name + " dy"] = e
            assert e <= 1e-2, worst
            assert float((dy * (1 - vmask)).abs().max()) == 0.0           # invalid rows of dY are exactly zero
        # wgrad from the STORED dy: dW[j*cin + c, n] = sum_r a_prev[r + j, c] dy[r, n]
        dW_ref = torch.zeros(k * cin, dy.shape[1], dtype=torch.float64, device="cuda")
        for j in range(k):
            dW_ref[j * cin:(j + 1) * cin] = a_prev[j:R].t() @ dy[:R - j]
        dW = st.grad("tdnn/%s_%s/kernel" % (name, kind)).double()
        worst[name + " dW"] = _relfro(dW, dW_ref)
        assert worst[name + " dW"] <= 1e-3, worst
        # dgrad from the stored dy and the bf16 kernel: dx[q, c] = sum_j dy[q - j, :] W_j[c, :]^T
        dx_ref = torch.zeros(R, cin, dtype=torch.float64, device="cuda")
        for j in range(k):
            dx_ref[j:] += dy[:R - j] @ W[j * cin:(j + 1) * cin].t()
        dx = _ws(eng, "tdnn%d/a/grad" % (Ln - 1)).double()
        pm = (t_idx < valid[Ln - 1]).double().unsqueeze(1)
        worst[name + " dx"] = _relfro(dx * pm, dx_ref * pm)
        assert worst[name + " dx"] <= 1e-2, worst
    print("layerwise backward, B=%d T=%d:" % (B, T), {k: "%.2e" % v for k, v in worst.items()})
