"""Per-kernel GPU parity through the C ABI (ctypes), each against a plain PyTorch / oracle evaluation of the SAME
inputs, so tolerances are tight (no accumulated bf16 pipeline noise):
  frame-level BN apply / backward (plain and fused-with-pooling), statistics pooling forward/backward (masked,
  golden vectors of multitask_v1/pooling.py:68-83), utterance-level BN, the fused margin heads on the reference's
  adversarial golden inputs (model/test_utils.py:157-318), and the optimizer kernels."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O

pytestmark = pytest.mark.gpu


def _lib():
    from tf_kaldi_speaker_b200 import _lib as L
    return L, L.load()


def _act(z, act, alpha):
    if act == 1:
        return torch.relu(z)
    if act == 2:
        return torch.nn.functional.leaky_relu(z, 0.2)
    if act == 3:
        return torch.where(z > 0, z, alpha * z)
    if act == 4:
        return torch.tanh(z)
    return z


def _valid_mask(B, T, valid, lengths, dev):
    t = torch.arange(T, device=dev).unsqueeze(0)
    if lengths is None:
        return (t < valid).expand(B, T)
    return t < lengths.unsqueeze(1)


@pytest.mark.parametrize("act", [1, 2, 3, 4])
@pytest.mark.parametrize("use_lengths", [False, True])
def test_bn_act_forward_backward(act, use_lengths):
    L, lib = _lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(act)
    B, T, Cn = 5, 37, 264
    R = B * T
    y = torch.randn(R, Cn, generator=g, device=dev).to(torch.bfloat16)
    da = (torch.randn(R, Cn, generator=g, device=dev) * 1e-2).to(torch.bfloat16)
    gamma = 1 + 0.3 * torch.randn(Cn, generator=g, device=dev)
    beta = 0.2 * torch.randn(Cn, generator=g, device=dev)
    alpha = 0.05 + 0.1 * torch.rand(Cn, generator=g, device=dev)
    valid = 30
    lengths = torch.tensor([30, 12, 37, 1, 25], dtype=torch.int32, device=dev) if use_lengths else None
    mask = _valid_mask(B, T, valid, lengths, dev).reshape(R, 1)
    # reference in fp64 on the same bf16 inputs
    yd, dad = y.double(), da.double()
    n = mask.sum().double()
    mean = (yd * mask).sum(0) / n
    var = (((yd - mean) ** 2) * mask).sum(0) / n
    rstd = torch.rsqrt(var + 1e-3)
    scale = (gamma.double() * rstd)
    shift = beta.double() - mean * scale
    z = yd * scale + shift
    a_ref = _act(z, act, alpha.double()) * mask
    zz = z.clone().requires_grad_(True)
    aa = _act(zz, act, alpha.double())
    (gz,) = torch.autograd.grad((aa * dad * mask).sum(), zz)
    gz = gz * mask
    yh = (yd - mean) * rstd
    dbeta_ref = gz.sum(0)
    dgamma_ref = (gz * yh).sum(0)
    dy_ref = scale * (gz - dbeta_ref / n - yh * dgamma_ref / n) * mask

    scale_f, shift_f = scale.float().contiguous(), shift.float().contiguous()
    mean_f, rstd_f = mean.float().contiguous(), rstd.float().contiguous()
    a = torch.full((R, Cn), 7.0, device=dev, dtype=torch.bfloat16)
    lp = L.ptr(lengths)
    L.check(lib.xv_bn_act_apply(L.ptr(y), L.ptr(a), L.ptr(scale_f), L.ptr(shift_f), L.ptr(alpha), act, C.c_int64(R), Cn,
                                C.c_int64(Cn), T, valid, lp, L.stream_ptr()))
    assert torch.allclose(a.double(), a_ref, rtol=1e-2, atol=1e-2)
    assert (a.double()[~mask.expand(R, Cn)] == 0).all()
    dgamma = torch.zeros(Cn, device=dev)
    dbeta = torch.zeros(Cn, device=dev)
    dalpha = torch.zeros(Cn, device=dev)
    nul = L.ptr(None)
    L.check(lib.xv_bn_act_bwd_reduce(L.ptr(y), L.ptr(da), L.ptr(scale_f), L.ptr(shift_f), L.ptr(mean_f), L.ptr(rstd_f),
                                     L.ptr(alpha), act, C.c_int64(R), Cn, C.c_int64(Cn), T, valid, lp, L.ptr(dgamma),
                                     L.ptr(dbeta), L.ptr(dalpha), nul, nul, 0, 0, L.stream_ptr()))
    assert torch.allclose(dbeta.double(), dbeta_ref, rtol=2e-4, atol=1e-5)
    assert torch.allclose(dgamma.double(), dgamma_ref, rtol=2e-4, atol=1e-5)
    if act == 3:
        dalpha_ref = (dad * torch.clamp(z, max=0) * mask).sum(0)
        assert torch.allclose(dalpha.double(), dalpha_ref, rtol=2e-4, atol=1e-5)
    dy = torch.full((R, Cn), 3.0, device=dev, dtype=torch.bfloat16)
    L.check(lib.xv_bn_act_bwd_apply(L.ptr(y), L.ptr(da), L.ptr(dy), L.ptr(scale_f), L.ptr(shift_f), L.ptr(mean_f),
                                    L.ptr(rstd_f), L.ptr(dgamma), L.ptr(dbeta), C.c_float(float(n)), L.ptr(alpha), act,
                                    C.c_int64(R), Cn, C.c_int64(Cn), T, valid, lp, nul, nul, 0, 0, L.stream_ptr()))
    err = (dy.double() - dy_ref).abs().max() / dy_ref.abs().max()
    assert err < 1e-2, err


@pytest.mark.parametrize("use_lengths", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_stats_pool_forward_backward(use_lengths, fused):
    L, lib = _lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(11)
    B, T, c_real, cpad = 6, 41, 300, 320
    R = B * T
    y = torch.randn(R, cpad, generator=g, device=dev).to(torch.bfloat16)
    y[:, c_real:] = 0
    y.view(B, T, cpad)[0] = 0          # an all-zero segment: variance floor path (multitask_v1/pooling.py:63)
    scale = (1 + 0.2 * torch.randn(cpad, generator=g, device=dev)).contiguous()
    shift = (0.3 * torch.randn(cpad, generator=g, device=dev)).contiguous()
    valid = 33
    lengths = torch.tensor([33, 5, 41, 1, 20, 33], dtype=torch.int32, device=dev) if use_lengths else None
    ln = lengths.long() if use_lengths else torch.full((B,), valid, device=dev)
    if fused:
        x_in = torch.relu(y.double() * scale.double() + shift.double())
    else:
        x_in = y.double()
    xin = x_in.view(B, T, cpad)[:, :, :c_real].cpu().clone().requires_grad_(True)
    ref = O.statistics_pooling(xin, ln.cpu())
    out = torch.zeros(B, 2 * cpad, device=dev)
    out3 = torch.zeros(B, 6 * cpad, device=dev, dtype=torch.bfloat16)
    nul = L.ptr(None)
    # saved BN statistics of the producing layer (arbitrary but consistent: yhat = (y - mean) * rstd)
    mean = (0.1 * torch.randn(cpad, generator=g, device=dev)).contiguous()
    rstd = (1 + 0.3 * torch.rand(cpad, generator=g, device=dev)).contiguous()
    sums = torch.full((B, 4, cpad), float("nan"), device=dev)
    if fused:
        L.check(lib.xv_stats_pool_fwd(L.ptr(y), L.ptr(out), L.ptr(out3), B, T, valid, L.ptr(lengths), c_real, cpad,
                                      C.c_int64(cpad), L.ptr(scale), L.ptr(shift), nul, 1, L.ptr(mean), L.ptr(rstd),
                                      L.ptr(sums), L.stream_ptr()))
    else:
        L.check(lib.xv_stats_pool_fwd(L.ptr(y), L.ptr(out), L.ptr(out3), B, T, valid, L.ptr(lengths), c_real, cpad,
                                      C.c_int64(cpad), nul, nul, nul, 0, nul, nul, nul, L.stream_ptr()))
    got = torch.cat([out[:, :c_real], out[:, cpad:cpad + c_real]], 1).double().cpu()
    assert torch.allclose(got, ref.detach(), rtol=2e-5, atol=2e-6), (got - ref).abs().max()
    assert (out[:, c_real:cpad] == 0).all() and (out[:, cpad + c_real:] == 0).all()
    rec = out3[:, :2 * cpad].float() + out3[:, 4 * cpad:].float()      # hi + lo reconstructs fp32 to ~2^-16
    assert torch.allclose(rec, out, rtol=1e-4, atol=1e-6)
    # backward
    gp = torch.randn(B, 2 * cpad, generator=g, device=dev) * 1e-2
    gref_in = torch.cat([gp[:, :c_real], gp[:, cpad:cpad + c_real]], 1).double().cpu()
    (gx,) = torch.autograd.grad((ref * gref_in).sum(), xin)
    if not fused:
        dx = torch.full((R, cpad), 5.0, device=dev, dtype=torch.bfloat16)
        L.check(lib.xv_stats_pool_bwd(L.ptr(y), L.ptr(out), L.ptr(gp), L.ptr(dx), B, T, valid, L.ptr(lengths), c_real,
                                      cpad, C.c_int64(cpad), L.stream_ptr()))
        got = dx.view(B, T, cpad)[:, :, :c_real].double().cpu()
        err = (got - gx).abs().max() / gx.abs().max()
        assert err < 1e-2, err
        assert (dx.view(B, T, cpad)[:, :, c_real:] == 0).all()
    else:
        # fused: the BN backward of the producing layer evaluates the pooling gradient on the fly
        dgamma = torch.zeros(cpad, device=dev)
        dbeta = torch.zeros(cpad, device=dev)
        L.check(lib.xv_bn_act_bwd_reduce(L.ptr(y), nul, L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(rstd), nul, 1,
                                         C.c_int64(R), cpad, C.c_int64(cpad), T, valid, L.ptr(lengths), L.ptr(dgamma),
                                         L.ptr(dbeta), nul, L.ptr(out), L.ptr(gp), cpad, c_real, L.stream_ptr()))
        z = (y.double() * scale.double() + shift.double()).view(B, T, cpad)[:, :, :c_real].cpu()
        gz = gx * (z > 0)
        dbeta_ref = gz.sum((0, 1))
        assert torch.allclose(dbeta[:c_real].double().cpu(), dbeta_ref, rtol=1e-3, atol=1e-6), (dbeta[:c_real].double().cpu() - dbeta_ref).abs().max()
        yh = ((y.double() - mean.double()) * rstd.double()).view(B, T, cpad)[:, :, :c_real].cpu()
        assert torch.allclose(dgamma[:c_real].double().cpu(), (gz * yh).sum((0, 1)), rtol=1e-3, atol=1e-6)
        # the same reductions from the per-(segment, channel) sums emitted by the forward kernel
        dgamma2 = torch.zeros(cpad, device=dev)
        dbeta2 = torch.zeros(cpad, device=dev)
        L.check(lib.xv_pool_bn_bwd_reduce(L.ptr(out), L.ptr(gp), L.ptr(sums), B, valid, L.ptr(lengths), c_real, cpad,
                                          L.ptr(dgamma2), L.ptr(dbeta2), L.stream_ptr()))
        assert torch.allclose(dbeta2[:c_real].double().cpu(), dbeta_ref, rtol=1e-3, atol=2e-6), (dbeta2[:c_real].double().cpu() - dbeta_ref).abs().max()
        assert torch.allclose(dgamma2[:c_real].double().cpu(), (gz * yh).sum((0, 1)), rtol=1e-3, atol=2e-6)
        assert float(dbeta2[c_real:].abs().max()) == 0.0 and float(dgamma2[c_real:].abs().max()) == 0.0
        dy = torch.zeros(R, cpad, device=dev, dtype=torch.bfloat16)
        L.check(lib.xv_bn_act_bwd_apply(L.ptr(y), nul, L.ptr(dy), L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(rstd),
                                        L.ptr(dgamma), L.ptr(dbeta), C.c_float(float(ln.sum())), nul, 1, C.c_int64(R), cpad,
                                        C.c_int64(cpad), T, valid, L.ptr(lengths), L.ptr(out), L.ptr(gp), cpad, c_real,
                                        L.stream_ptr()))
        n = float(ln.sum())
        dy_ref = scale.double().cpu()[:c_real] * (gz - dbeta_ref / n - yh * (gz * yh).sum((0, 1)) / n)
        tmask = (torch.arange(T).unsqueeze(0) < ln.cpu().unsqueeze(1)).unsqueeze(2)
        dy_ref = dy_ref * tmask
        got = dy.view(B, T, cpad)[:, :, :c_real].double().cpu()
        assert (got - dy_ref).abs().max() / dy_ref.abs().max() < 1e-2


def test_stats_pool_golden(golden_dir):
    """multitask_v1/pooling.py:68-83 golden vectors through the CUDA kernel (inputs rounded to bf16 storage)."""
    L, lib = _lib()
    gd = np.load(os.path.join(golden_dir, "statpool.npz"))
    x = torch.from_numpy(gd["x"]).cuda()
    B, T, Cn = x.shape
    xb = x.to(torch.bfloat16).reshape(B * T, Cn).contiguous()
    ln = torch.from_numpy(gd["length"].astype(np.int32)).cuda()
    out = torch.zeros(B, 2 * Cn, device="cuda")
    nul = L.ptr(None)
    L.check(lib.xv_stats_pool_fwd(L.ptr(xb), L.ptr(out), nul, B, T, T, L.ptr(ln), Cn, Cn, C.c_int64(Cn), nul, nul, nul, 0,
                                  nul, nul, nul, L.stream_ptr()))
    # exact reference on the bf16-rounded inputs, and the reference's own golden output within bf16 input rounding
    ref_b = O.statistics_pooling(xb.double().view(B, T, Cn).cpu(), ln.long().cpu())
    assert torch.allclose(out.double().cpu(), ref_b, rtol=2e-5, atol=2e-6)
    assert np.allclose(out.cpu().numpy(), gd["out"], rtol=1e-2, atol=2e-3)


@pytest.mark.parametrize("mode,act", [(1, 1), (1, 3), (2, 1), (0, 2), (1, 0)])
def test_bn_rows(mode, act):
    L, lib = _lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    B, Cn = 37, 200
    y = torch.randn(B, Cn, generator=g, device=dev)
    da = torch.randn(B, Cn, generator=g, device=dev)
    gamma = 1 + 0.3 * torch.randn(Cn, generator=g, device=dev)
    beta = 0.2 * torch.randn(Cn, generator=g, device=dev)
    alpha = 0.05 + 0.1 * torch.rand(Cn, generator=g, device=dev)
    mm = 0.1 * torch.randn(Cn, generator=g, device=dev)
    mv = 0.5 + torch.rand(Cn, generator=g, device=dev)
    mm0, mv0 = mm.clone(), mv.clone()
    yd = y.double().requires_grad_(True)
    if mode == 1:
        mean = yd.mean(0)
        var = ((yd - mean) ** 2).mean(0)
    elif mode == 2:
        mean, var = mm.double(), mv.double()
    if mode == 0:
        z = yd
    else:
        z = (yd - mean) * torch.rsqrt(var + 1e-3) * gamma.double() + beta.double()
    a_ref = _act(z, act, alpha.double())
    (dy_ref,) = torch.autograd.grad((a_ref * da.double()).sum(), yd)
    a = torch.zeros(B, Cn, device=dev)
    a3 = torch.zeros(B, 3 * Cn, device=dev, dtype=torch.bfloat16)
    bn_out = torch.zeros(B, Cn, device=dev)
    smean, srstd = torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev)
    L.check(lib.xv_bn_rows_fwd(L.ptr(y), B, Cn, mode, L.ptr(gamma), L.ptr(beta), L.ptr(mm), L.ptr(mv), C.c_float(0.9),
                               C.c_float(1e-3), L.ptr(alpha), act, L.ptr(bn_out), L.ptr(a), L.ptr(a3), 3, L.ptr(smean),
                               L.ptr(srstd), L.stream_ptr()))
    assert torch.allclose(a.double(), a_ref.detach(), rtol=1e-4, atol=1e-5)
    rec = a3[:, :Cn].float() + a3[:, 2 * Cn:].float()
    assert torch.allclose(rec, a, rtol=1e-4, atol=1e-6)
    if mode == 1:
        assert torch.allclose(mm.double(), mm0.double() * 0.9 + mean.detach() * 0.1, rtol=1e-4, atol=1e-6)
        assert torch.allclose(mv.double(), mv0.double() * 0.9 + var.detach() * 0.1, rtol=1e-4, atol=1e-6)
    dy = torch.zeros(B, Cn, device=dev)
    dyb = torch.zeros(B, Cn, device=dev, dtype=torch.bfloat16)
    dgamma, dbeta, dalpha, dbias = (torch.zeros(Cn, device=dev) for _ in range(4))
    L.check(lib.xv_bn_rows_bwd(L.ptr(y), L.ptr(da), B, Cn, mode, L.ptr(gamma), L.ptr(beta), L.ptr(smean), L.ptr(srstd),
                               L.ptr(alpha), act, L.ptr(dy), L.ptr(dyb), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dalpha),
                               L.ptr(dbias), L.stream_ptr()))
    assert torch.allclose(dy.double(), dy_ref, rtol=2e-3, atol=2e-5), (dy.double() - dy_ref).abs().max()
    assert torch.allclose(dbias.double(), dy_ref.sum(0), rtol=1e-3, atol=1e-4)


def _head_engine(E, Cn, loss_type, w):
    from tf_kaldi_speaker_b200.runtime import Engine, VarSpec, set_engine, UttAct, _pad_to
    eng = set_engine(Engine())
    cpad = _pad_to(Cn, 8)
    eng.store.declare(VarSpec("softmax/output/kernel", (E, Cn), (E, cpad)))
    if loss_type == "softmax":
        eng.store.declare(VarSpec("softmax/output/bias", (Cn,), (cpad,)))
    eng.store.finalize()
    eng.store.load_tf({"softmax/output/kernel": w})
    return eng, UttAct


def test_margin_heads_on_reference_golden(golden_dir):
    """The reference's adversarial rows (theta~0, theta~pi, tiny / x10 norm; model/tdnn.py:271-277) through the fused
    head; expected losses come from model/test_utils.py:157-318 (tests/golden/heads.npz).  Gradients against the
    fp64 oracle's autograd."""
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model import loss as LS
    from tf_kaldi_speaker_b200.runtime import ScaledUtt
    gd = np.load(os.path.join(golden_dir, "heads.npz"))
    labels = torch.from_numpy(gd["labels"].astype(np.int32)).cuda()
    w = gd["w"]
    E, Cn = w.shape
    P = {"softmax/output/kernel": torch.from_numpy(w.astype(np.float64))}
    fns = {"asoftmax": LS.asoftmax, "additive_margin_softmax": LS.additive_margin_softmax,
           "additive_angular_margin_softmax": LS.additive_angular_margin_softmax}
    for i in range(len(gd["case_loss"])):
        lt = str(gd["case_loss_type"][i])
        m = float(gd["case_m"][i])
        fnorm = bool(gd["case_feature_norm"][i])
        factor = float(gd["case_scaling_factor"][i])
        pd = dict(weight_l2_regularizer=1e-5, global_step=int(gd["case_global_step"][i]), feature_norm=fnorm,
                  feature_scaling_factor=factor)
        for pre in ("asoftmax", "amsoftmax", "arcsoftmax"):
            pd.update({pre + "_lambda_min": 10, pre + "_lambda_base": 1000, pre + "_lambda_gamma": 1, pre + "_lambda_power": 4})
        pd["asoftmax_m"], pd["amsoftmax_m"], pd["arcsoftmax_m"] = (int(m) if lt == "asoftmax" else 4), m, m
        emb = gd["emb_x10"] if lt == "asoftmax" else gd["emb"]
        eng, UttAct = _head_engine(E, Cn, lt, w)
        params = ParamsPlain(**pd)
        u = UttAct(torch.from_numpy(emb).cuda().contiguous(), None, "u")
        u.needs_grad = True
        feats = ScaledUtt(u, factor) if fnorm else u
        eng.begin_step(True)
        loss, _ = fns[lt](feats, labels, Cn, params, is_training=True)
        eng.backward()
        torch.cuda.synchronize()
        got = float(loss.item())
        want = float(gd["case_loss"][i])
        assert abs(got - want) <= 2e-4 * abs(want) + 1e-5, (lt, m, fnorm, factor, got, want)
        # gradients vs fp64 oracle autograd
        po = O.ParamsPlain(**pd)
        x = torch.from_numpy(emb.astype(np.float64)).requires_grad_(True)
        Wt = P["softmax/output/kernel"].clone().requires_grad_(True)
        xin = O.l2_scaling(x, factor) if fnorm else x
        lo, _ = O.loss_network(lt, xin, labels.cpu(), {"softmax/output/kernel": Wt}, po)
        gx, gw = torch.autograd.grad(lo, [x, Wt])
        gxe = u.grad.double().cpu()
        gwe = torch.from_numpy(eng.store.export_tf(grads=True)["softmax/output/kernel"]).double()
        ex = (gxe - gx).norm() / gx.norm()
        ew = (gwe - gw).norm() / gw.norm()
        assert ex < 2e-2 and ew < 2e-2, (lt, m, fnorm, factor, float(ex), float(ew))


def test_softmax_head_with_bias_and_ragged_classes():
    """Plain softmax head (loss.py:29-35) with a class count that is neither a multiple of 8 nor of the N tile."""
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model import loss as LS
    g = torch.Generator().manual_seed(9)
    B, E, Cn = 70, 512, 1003
    w = (torch.rand(E, Cn, generator=g) - 0.5) * 0.2
    b = torch.randn(Cn, generator=g) * 0.1
    x = torch.randn(B, E, generator=g)
    y = torch.randint(0, Cn, (B,), generator=g, dtype=torch.int32)
    eng, UttAct = _head_engine(E, Cn, "softmax", w.numpy())
    eng.store.load_tf({"softmax/output/bias": b.numpy()})
    params = ParamsPlain(weight_l2_regularizer=1e-4, global_step=0, debug_logits=True)
    u = UttAct(x.cuda().contiguous(), None, "u")
    u.needs_grad = True
    eng.begin_step(True)
    loss, ep = LS.softmax(u, y.cuda(), Cn, params, is_training=True)
    eng.backward()
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True)
    lo, logits = O.softmax_head(xd, y, {"softmax/output/kernel": wd, "softmax/output/bias": bd})
    gx, gw, gb = torch.autograd.grad(lo, [xd, wd, bd])
    assert abs(float(loss.item()) - lo.item()) <= 1e-4 * abs(lo.item())
    assert torch.allclose(ep["logits"].double().cpu(), logits.detach(), rtol=1e-3, atol=2e-3)
    ge = eng.store.export_tf(grads=True)
    assert (u.grad.double().cpu() - gx).norm() / gx.norm() < 2e-2
    assert (torch.from_numpy(ge["softmax/output/kernel"]).double() - gw).norm() / gw.norm() < 2e-2
    assert (torch.from_numpy(ge["softmax/output/bias"]).double() - gb).norm() / gb.norm() < 2e-2


@pytest.mark.parametrize("opt", ["sgd", "momentum", "nesterov", "adam"])
@pytest.mark.parametrize("clip", [False, True])
def test_optimizer_kernels(opt, clip):
    from tf_kaldi_speaker_b200 import _lib as L
    from tf_kaldi_speaker_b200.runtime import Engine, VarSpec, set_engine
    eng = set_engine(Engine())
    st = eng.store
    shapes = {"a/kernel": (70, 33), "a/bias": (33,), "b/kernel": (512, 512), "b/gamma": (512,)}
    st.declare(VarSpec("a/kernel", shapes["a/kernel"], (72, 40), l2=1e-2, shadow="plain"))
    st.declare(VarSpec("a/bias", shapes["a/bias"], (40,)))
    st.declare(VarSpec("b/kernel", shapes["b/kernel"], (512, 512), l2=5e-3, shadow="split"))
    st.declare(VarSpec("b/gamma", shapes["b/gamma"], (512,)))
    st.finalize()
    g = torch.Generator().manual_seed(2)
    P = {k: torch.randn(v, generator=g, dtype=torch.float64) for k, v in shapes.items()}
    st.load_tf({k: v.numpy() for k, v in P.items()})
    code = {"sgd": L.OPT_SGD, "momentum": L.OPT_MOMENTUM, "nesterov": L.OPT_NESTEROV, "adam": L.OPT_ADAM}[opt]
    po = O.ParamsPlain(optimizer="momentum" if opt == "nesterov" else opt, momentum=0.9, use_nesterov=(opt == "nesterov"))
    state = {}
    l2 = {"a/kernel": 1e-2, "a/bias": 0.0, "b/kernel": 5e-3, "b/gamma": 0.0}
    for step in range(3):
        G = {k: torch.randn(v, generator=g, dtype=torch.float64) for k, v in shapes.items()}
        for k in shapes:
            st.grad(k).zero_()
            st.grad(k)[tuple(slice(0, d) for d in shapes[k])] = G[k].float().cuda()
        full = {k: G[k] + l2[k] * P[k] for k in shapes}
        if clip:
            gn = math.sqrt(sum(float((v ** 2).sum()) for v in full.values()))
            sc = 3.0 / max(gn, 3.0)
            full = {k: v * sc for k, v in full.items()}
        P, state = O.apply_optimizer(P, full, state, po, 0.05)
        eng.scalars.zero_()
        eng.set_hyper(0.05, 0.9, float(step + 1), 3.0 if clip else 0.0)
        eng.optimizer_step(code, clip=clip)
        got = st.export_tf()
        for k in shapes:
            assert np.allclose(got[k], P[k].numpy(), rtol=2e-5, atol=2e-6), (opt, clip, step, k)
    # bf16 shadows track the fp32 masters (plain and [hi; lo; hi] split)
    wa = st.view("a/kernel")
    assert torch.equal(st.shadow_view("a/kernel"), wa.to(torch.bfloat16))
    wb = st.view("b/kernel")
    sv = st.shadow_view("b/kernel")
    hi = wb.to(torch.bfloat16)
    assert torch.equal(sv[:512], hi) and torch.equal(sv[1024:], hi)
    assert torch.equal(sv[512:1024], (wb - hi.float()).to(torch.bfloat16))


@pytest.mark.parametrize("B,T,D", [(128, 200, 30), (3, 37, 23), (2, 5, 32), (1, 9, 1)])
def test_pack_input_bit_exact(B, T, D):
    """tdnn1's im2col operand (model/tdnn.py:35-44 input side): out[b*T+t, j*32+c] = bf16(x[b, t+j, c]), zero elsewhere."""
    from tf_kaldi_speaker_b200.runtime import Engine
    eng = Engine()
    g = torch.Generator().manual_seed(B + T + D)
    x = torch.randn(B, T, D, generator=g).cuda()
    fa = eng.pack_input(x)
    torch.cuda.synchronize()
    k, dpad = 5, 32
    want = torch.zeros(B, T, 192, dtype=torch.bfloat16, device="cuda")
    for j in range(k):
        if T - j > 0:
            want[:, :T - j, j * dpad:j * dpad + D] = x[:, j:, :].to(torch.bfloat16)
    assert torch.equal(fa.data.view(B, T, 192), want)
