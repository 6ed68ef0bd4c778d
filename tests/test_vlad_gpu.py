"""GPU parity of the NetVLAD / GhostVLAD pooling kernels (csrc/xv_vlad.cu) through the C ABI:
  * forward against the committed golden vectors generated from the reference's NumPy known-answer code
    (model/test_utils.py:421-436 compute_ghost_vlad -> tests/golden/vlad.npz);
  * forward + backward against fp64 autograd of the oracle's ghost_vlad on the SAME bf16 inputs (ghost clusters, final
    normalisation on / off, ragged lengths, more clusters than one accumulator pass, accumulate mode);
  * one full training step with pooling_type = "ghost_vlad" against the fp64 oracle, and batched variable-length
    extraction == one call per utterance."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O
from tests.xv_testlib import base_params, head_params, make_batch, rel_fro, min_cosine

pytestmark = pytest.mark.gpu


def _pad(n, m):
    return (n + m - 1) // m * m


def _run_kernels(logits, value, centers, K, G, final, lengths=None, dout=None):
    """logits f32 [B,T,K+G], value f32 [B,T,dv] (bf16-representable), centers f32 [K+G, dv] -> dict of torch results."""
    from tf_kaldi_speaker_b200 import _lib as L
    lib = L.load()
    dev = "cuda"
    B, T, KG = logits.shape
    dv = value.shape[2]
    ldl, cpad = _pad(KG, 64), _pad(dv, 64)
    lg = torch.zeros(B * T, ldl, dtype=torch.bfloat16, device=dev)
    vd = torch.zeros(B * T, cpad, dtype=torch.bfloat16, device=dev)
    lg[:, :KG] = logits.reshape(B * T, KG).to(dev).to(torch.bfloat16)
    lg[:, KG:] = 7.0          # padded logit columns must be ignored
    vd[:, :dv] = value.reshape(B * T, dv).to(dev).to(torch.bfloat16)
    cen = torch.zeros(KG, cpad, device=dev)
    cen[:, :dv] = centers.to(dev).float()
    ln = None if lengths is None else lengths.to(dev).to(torch.int32)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    post = torch.full((B, T, KG), float("nan"), device=dev)
    L.check(lib.xv_vlad_post_fwd(L.ptr(lg), L.ptr(post), B, T, T, L.ptr(ln), KG, ldl, s))
    res = torch.full((B, K, cpad), float("nan"), device=dev)
    mass = torch.full((B, K), float("nan"), device=dev)
    sumsq = torch.full((B, K), float("nan"), device=dev)
    out = torch.full((B, K * cpad), float("nan"), device=dev)
    out3 = torch.empty(B, 3 * K * cpad, dtype=torch.bfloat16, device=dev)
    L.check(lib.xv_vlad_pool_fwd(L.ptr(vd), L.ptr(post), L.ptr(cen), L.ptr(res), L.ptr(mass), L.ptr(sumsq), L.ptr(out),
                                 L.ptr(out3), B, T, T, L.ptr(ln), K, KG, dv, cpad, C.c_int64(cpad), cpad, int(final), s))
    o = out.view(B, K, cpad)
    r = {"post": post.clone(), "out": o[:, :, :dv].reshape(B, K * dv).clone(), "out_pad": o[:, :, dv:].clone(), "out3": out3,
         "out_full": out}
    if dout is not None:
        dp = torch.zeros(B, K, cpad, device=dev)
        dp[:, :, :dv] = dout.reshape(B, K, dv).to(dev).float()
        gres = torch.empty(B, K, cpad, device=dev)
        gc = torch.empty(B, K, device=dev)
        dl = torch.full((B * T, ldl), float("nan"), dtype=torch.bfloat16, device=dev)
        dvd = torch.full((B * T, cpad), float("nan"), dtype=torch.bfloat16, device=dev)
        dcen = torch.zeros(KG, cpad, device=dev)
        args = lambda dvd_, dcen_, acc: (L.ptr(vd), L.ptr(post), L.ptr(cen), L.ptr(mass), L.ptr(sumsq), L.ptr(out),
                                         L.ptr(dp.view(B, K * cpad)), L.ptr(gres), L.ptr(gc), L.ptr(dl), L.ptr(dvd_),
                                         L.ptr(dcen_), B, T, T, L.ptr(ln), K, KG, dv, cpad, C.c_int64(cpad), ldl, cpad,
                                         int(final), acc, s)
        L.check(lib.xv_vlad_pool_bwd(*args(dvd, dcen, 0)))
        dvd2, dcen2 = dvd.clone(), dcen.clone()
        L.check(lib.xv_vlad_pool_bwd(*args(dvd2, dcen2, 1)))      # accumulate: a second pass doubles both gradients
        torch.cuda.synchronize()
        r.update(dlogits=dl.float().reshape(B, T, ldl)[:, :, :KG], dlogits_pad=dl.float().reshape(B, T, ldl)[:, :, KG:],
                 dvalue=dvd.float().reshape(B, T, cpad)[:, :, :dv], dvalue_pad=dvd.float().reshape(B, T, cpad)[:, :, dv:],
                 dcenters=dcen[:, :dv], dcenters_pad=dcen[:, dv:], dvalue2=dvd2.float().reshape(B, T, cpad)[:, :, :dv],
                 dcenters2=dcen2[:, :dv])
    torch.cuda.synchronize()
    return r


def _oracle(logits, value, centers, K, G, final, lengths=None):
    KG = K + G
    p = O.ParamsPlain(vlad_value_input="v", vlad_key_input="k", vlad_key_num_nodes=[], vlad_value_num_nodes=[],
                      vlad_num_centers=K, vlad_num_ghosts=G, vlad_final_l2_norm=bool(final), batchnorm_momentum=0.99)
    P = {"tdnn/vlad/vlad_weight_affine/kernel": torch.eye(KG, dtype=torch.float64),
         "tdnn/vlad/vlad_weight_affine/bias": torch.zeros(KG, dtype=torch.float64), "tdnn/vlad/vlad_centers": centers}
    ep = {"v": value, "k": logits}
    out = O.ghost_vlad(ep, P, p, True, None, lengths)
    return out, ep["vlad_weights"]


def test_vlad_forward_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "vlad.npz"))
    for tag in ("k8_g2", "k4_g0_final", "k5_g1_final"):
        K, G, final = [int(v) for v in g[tag + "/cfg"]]
        # the kernels read bf16 activations: round the fixture inputs and evaluate the oracle on the rounded values too
        value = torch.from_numpy(g[tag + "/value"]).to(torch.bfloat16).float()
        key = torch.from_numpy(g[tag + "/key"]).to(torch.bfloat16).float()
        centers = torch.from_numpy(g[tag + "/centers"])
        r = _run_kernels(key, value, centers, K, G, final)
        out_o, post_o = _oracle(key.double(), value.double(), centers.double(), K, G, final)
        assert torch.allclose(r["post"].cpu().double(), post_o, rtol=2e-5, atol=1e-7), tag
        assert torch.allclose(r["out"].cpu().double(), out_o, rtol=1e-4, atol=2e-6), tag
        # and against the reference's own NumPy output (differs only by the bf16 rounding of the inputs)
        ref = torch.from_numpy(g[tag + "/out"])
        assert rel_fro(r["out"].cpu(), ref) <= 1e-2, (tag, rel_fro(r["out"].cpu(), ref))
        # the bf16 split copy reproduces the fp32 row to ~2^-16
        B = value.shape[0]
        W = r["out_full"].shape[1]
        o3 = r["out3"].float().view(B, 3, W)
        assert torch.equal(o3[:, 0], o3[:, 1])
        assert float((o3[:, 0] + o3[:, 2] - r["out_full"]).abs().max()) <= 2e-5


CASES = [
    # B, T, dv, K, G, final, ragged
    (4, 37, 1500, 8, 2, False, False),       # tdnn5-wide value, one accumulator pass, ghosts
    (3, 50, 200, 10, 0, True, True),         # two accumulator passes, NetVLAD (no ghosts), final normalisation, ragged
    (2, 300, 512, 40, 24, True, True),       # 64 clusters in total: two logits per lane, long segments
    (5, 21, 72, 3, 1, False, True),          # value dim not a multiple of 64
]


@pytest.mark.parametrize("B,T,dv,K,G,final,ragged", CASES)
def test_vlad_forward_backward(B, T, dv, K, G, final, ragged):
    g = torch.Generator().manual_seed(B * 100 + K)
    KG = K + G
    value = (torch.relu(torch.randn(B, T, dv, generator=g)) + 0.3 * torch.randn(B, 1, dv, generator=g)).to(torch.bfloat16).float()
    logits = (2.0 * torch.randn(B, T, KG, generator=g)).to(torch.bfloat16).float()
    centers = 0.5 * torch.randn(KG, dv, generator=g)
    lengths = None
    if ragged:
        lengths = torch.randint(5, T + 1, (B,), generator=g)
        lengths[0] = T
    dout = torch.randn(B, K * dv, generator=g)
    r = _run_kernels(logits, value, centers, K, G, final, lengths, dout)

    l64 = logits.double().requires_grad_(True)
    v64 = value.double().requires_grad_(True)
    c64 = centers.double().requires_grad_(True)
    out_o, post_o = _oracle(l64, v64, c64, K, G, final, lengths)
    gl, gv, gcn = torch.autograd.grad((out_o * dout.double()).sum(), [l64, v64, c64])

    assert torch.allclose(r["post"].cpu().double(), post_o.detach(), rtol=1e-4, atol=1e-7)
    assert rel_fro(r["out"].cpu(), out_o.detach()) <= 1e-5, rel_fro(r["out"].cpu(), out_o.detach())
    assert float(r["out_pad"].abs().max()) == 0.0 if r["out_pad"].numel() else True
    # frame-level gradients are emitted as bf16, the centres' in fp32
    assert rel_fro(r["dvalue"].cpu(), gv) <= 4e-3, rel_fro(r["dvalue"].cpu(), gv)
    assert rel_fro(r["dlogits"].cpu(), gl) <= 4e-3, rel_fro(r["dlogits"].cpu(), gl)
    assert rel_fro(r["dcenters"].cpu(), gcn) <= 1e-4, rel_fro(r["dcenters"].cpu(), gcn)
    assert float(gcn[K:].abs().max() if G else 0.0) == 0.0 and float(r["dcenters"][K:].abs().max() if G else 0.0) == 0.0
    for name in ("dvalue_pad", "dlogits_pad", "dcenters_pad"):
        assert float(r[name].abs().max()) == 0.0 if r[name].numel() else True, name
    assert rel_fro(r["dvalue2"].cpu(), 2 * gv) <= 8e-3
    assert rel_fro(r["dcenters2"].cpu(), 2 * gcn) <= 2e-4
    if lengths is not None:     # frames beyond the length: zero posterior, zero gradient
        for b in range(B):
            n = int(lengths[b])
            if n < T:
                assert float(r["post"][b, n:].abs().max()) == 0.0
                assert float(r["dvalue"][b, n:].abs().max()) == 0.0
                assert float(r["dlogits"][b, n:].abs().max()) == 0.0


def test_vlad_rejects_too_many_clusters():
    from tf_kaldi_speaker_b200 import _lib as L
    lib = L.load()
    x = torch.zeros(8, 128, dtype=torch.bfloat16, device="cuda")
    post = torch.zeros(1, 8, 65, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    with pytest.raises(ValueError):
        L.check(lib.xv_vlad_post_fwd(L.ptr(x), L.ptr(post), 1, 8, 8, L.ptr(None), 65, 128, s))


VLAD_CASES = [
    ("ghostvlad_k8_g2", dict(vlad_num_centers=8, vlad_num_ghosts=2, vlad_key_input="tdnn5_relu", vlad_key_num_nodes=[],
                             vlad_value_input="tdnn5_relu", vlad_value_num_nodes=[], vlad_final_l2_norm=False)),
    ("netvlad_k6_nets_final", dict(vlad_num_centers=6, vlad_num_ghosts=0, vlad_key_input="tdnn4_relu",
                                   vlad_key_num_nodes=[96], vlad_value_input="tdnn5_relu", vlad_value_num_nodes=[120],
                                   vlad_final_l2_norm=True)),
]


@pytest.mark.parametrize("name,vl", VLAD_CASES, ids=[c[0] for c in VLAD_CASES])
def test_vlad_train_step(name, vl):
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    loss_type = "additive_margin_softmax"
    # B = 64: the utterance-level batch-norms amplify bf16 rounding noise on tiny batches; at B = 32 the
    # second case sat at 1.09e-3 against the 1e-3 gate
    B, T, D, Cn = 64, 60, 30, 200
    pd = base_params(**head_params(loss_type))
    pd.update(vl)
    pd.update(pooling_type="ghost_vlad", feature_norm=True, feature_scaling_factor=30, num_nodes_pooling_layer=200)
    x, y = make_batch(B, T, D, Cn, seed=2)
    po = O.ParamsPlain(**dict(pd))
    P = O.init_params(D, po, Cn, loss_type, seed=4)
    gen = torch.Generator().manual_seed(6)
    for k in P:
        if k.endswith("/gamma"):
            P[k] = P[k] + 0.2 * torch.randn(P[k].shape, generator=gen, dtype=torch.float64)
        elif k.endswith("/beta") or k.endswith("/bias"):
            P[k] = P[k] + 0.1 * torch.randn(P[k].shape, generator=gen, dtype=torch.float64)
    P["tdnn/vlad/vlad_weight_affine/kernel"] = P["tdnn/vlad/vlad_weight_affine/kernel"] * 3.0    # visibly non-uniform posteriors
    gstep, lr = 300000, 0.01
    loss_o, total_o, _, newP_o, _, ep_o = O.train_step(P, {}, x.double(), y, po, loss_type, lr, gstep)
    grads_o = ep_o["__raw_grads"]

    tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_test_model_vlad")
    tr.build("train", D, loss_type, Cn)
    st = tr.engine.store
    assert set(st.specs.keys()) == set(P.keys()), set(st.specs.keys()) ^ set(P.keys())
    st.load_tf({k: v.numpy() for k, v in P.items()})
    res = tr.train_step(x, y, lr, gstep, fetch_loss=True)
    torch.cuda.synchronize()
    loss_rel = abs(res["raw_loss"] - loss_o.item()) / abs(loss_o.item())
    total_rel = abs(res["loss"] - total_o.item()) / abs(total_o.item())
    emb = tr.endpoints["tdnn6_dense"].dense().cpu().numpy()
    cos = min_cosine(emb, ep_o["tdnn6_dense"].detach().numpy())
    pool = tr.endpoints["pooling"].dense().cpu().double()
    pool_err = rel_fro(pool, ep_o["pooling"].detach())
    w_err = float((tr.endpoints["vlad_weights"].cpu().double() - ep_o["vlad_weights"].detach()).abs().max())
    print(name, "loss_rel %.2e total_rel %.2e emb_cos %.6f pooling rel %.2e posteriors max abs err %.2e"
          % (loss_rel, total_rel, cos, pool_err, w_err))
    assert loss_rel <= 1e-3 and total_rel <= 1e-3
    assert cos >= 0.999
    assert pool_err <= 2e-2
    assert w_err <= 8e-2      # bf16 logits of magnitude ~10 carry an absolute error ~3e-2, which peaked posteriors pass on
                              # (3.1e-2 .. 3.9e-2 over runs; the kernel-level test pins the posteriors to 1e-4 on equal inputs)
    ge = st.export_tf(grads=True)
    s = float(pd["weight_l2_regularizer"])
    worst = {}
    for n, go in grads_o.items():
        gv = ge[n].astype(np.float64)
        if O.l2_regularised(n):
            gv = gv + s * P[n].numpy()
        if np.linalg.norm(go.numpy()) < 1e-9:
            assert float(np.abs(gv).max()) <= 1e-3, n
            continue
        worst[n] = rel_fro(gv, go.numpy())
        c = float(np.dot(gv.ravel(), go.numpy().ravel()) / (np.linalg.norm(gv) * np.linalg.norm(go.numpy()) + 1e-300))
        # the bias of the cluster-assignment layer: every frame's dlogits sum to zero (softmax Jacobian), so this gradient is
        # what survives the cancellation of ~B*T bf16-rounded terms -- 0.28 .. 0.31 run to run
        gate = 0.40 if n.endswith("vlad_weight_affine/bias") else 0.30
        assert worst[n] <= gate and c >= 0.95, (n, worst[n], c)
    print("  worst grads vs fp64:", sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    newv = st.export_tf()
    for n in ("tdnn/vlad/vlad_centers", "tdnn/vlad/vlad_weight_affine/kernel", "tdnn/tdnn4_dense/kernel"):
        assert rel_fro(newv[n], newP_o[n].numpy()) <= 5e-2, n
    # second and third call: CUDA-graph capture + replay of the same step shape
    for i in range(2):
        res2 = tr.train_step(x, y, lr, gstep + 1 + i, fetch_loss=True)
        assert np.isfinite(res2["loss"])


def test_vlad_predict_ragged_matches_single():
    """Batched variable-length extraction with GhostVLAD pooling == one call per utterance (masked posteriors)."""
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    pd = base_params()
    pd.update(VLAD_CASES[0][1])
    pd.update(pooling_type="ghost_vlad", num_nodes_pooling_layer=200)
    D = 24
    tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_test_model_vlad2")
    tr.build("predict", D)
    g = torch.Generator().manual_seed(9)
    lens = [90, 40, 64]
    feats = np.zeros((3, 90, D), dtype=np.float32)
    singles = []
    for i, n in enumerate(lens):
        f = torch.randn(n, D, generator=g).numpy()
        feats[i, :n] = f
        singles.append(tr.predict(f))
    batched = tr.predict_batch_padded(feats, lens)
    for i in range(3):
        c = float(np.dot(batched[i], singles[i]) / (np.linalg.norm(batched[i]) * np.linalg.norm(singles[i])))
        assert c >= 0.9999, (i, c)
        assert np.allclose(batched[i], singles[i], rtol=2e-3, atol=2e-3), i
