"""Pin the oracle (oracle/xvector_oracle.py) against the committed golden vectors that
tests/golden/make_golden.py produced from the REFERENCE's NumPy known-answer functions
(model/test_utils.py:157-376, model/multitask_v1/pooling.py:68-83).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O


def _params(**kw):
    p = O.ParamsPlain(weight_l2_regularizer=1e-5)
    for pre in ("asoftmax", "amsoftmax", "arcsoftmax"):
        p.dict[pre + "_lambda_min"] = 10
        p.dict[pre + "_lambda_base"] = 1000
        p.dict[pre + "_lambda_gamma"] = 1
        p.dict[pre + "_lambda_power"] = 4
    p.dict.update(kw)
    return p


def test_heads_match_reference_numpy(golden_dir):
    g = np.load(os.path.join(golden_dir, "heads.npz"))
    labels = torch.from_numpy(g["labels"].astype(np.int64))
    P = {"softmax/output/kernel": torch.from_numpy(g["w"].astype(np.float64))}
    n = len(g["case_loss"])
    assert n == 27
    for i in range(n):
        lt = str(g["case_loss_type"][i])
        m = float(g["case_m"][i])
        p = _params(feature_norm=bool(g["case_feature_norm"][i]),
                    feature_scaling_factor=float(g["case_scaling_factor"][i]),
                    global_step=int(g["case_global_step"][i]))
        emb = g["emb_x10"] if lt == "asoftmax" else g["emb"]
        x = torch.from_numpy(emb.astype(np.float64)).requires_grad_(True)
        if lt == "asoftmax":
            p.asoftmax_m = int(m)
        elif lt == "additive_margin_softmax":
            p.amsoftmax_m = m
        else:
            p.arcsoftmax_m = m
        xin = O.l2_scaling(x, p.feature_scaling_factor) if p.feature_norm else x
        loss, _ = O.loss_network(lt, xin, labels, P, p)
        (gx,) = torch.autograd.grad(loss, x)
        assert not torch.isnan(gx).any(), "Gradient should not be nan"      # model/tdnn.py:282,313,342
        # same tolerance the reference's own self-tests use (np.allclose defaults)
        assert np.allclose(loss.item(), g["case_loss"][i], rtol=1e-5, atol=1e-8), (lt, m, loss.item(), g["case_loss"][i])


def test_masked_statistics_pooling(golden_dir):
    g = np.load(os.path.join(golden_dir, "statpool.npz"))
    x = torch.from_numpy(g["x"].astype(np.float64))
    ln = torch.from_numpy(g["length"].astype(np.int64))
    out = O.statistics_pooling(x, ln).numpy()
    assert np.allclose(out, g["out"], rtol=1e-5, atol=1e-7)
    # unmasked == masked with full lengths
    full = torch.full((x.shape[0],), x.shape[1], dtype=torch.int64)
    assert np.allclose(O.statistics_pooling(x).numpy(), O.statistics_pooling(x, full).numpy(), rtol=1e-12, atol=1e-12)


def test_self_attention(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    for tag in ("h4_split", "h1_shared", "h2_shared_noscale"):
        heads, split, scale = [int(v) for v in g[tag + "/cfg"]]
        value = torch.from_numpy(g[tag + "/value"].astype(np.float64))
        key = torch.from_numpy(g[tag + "/key"].astype(np.float64))
        p = O.ParamsPlain(att_value_input="v", att_key_input="k", att_key_num_nodes=[], att_value_num_nodes=[],
                          att_key_network_type=0, att_value_network_type=0, att_num_heads=heads,
                          att_split_key=bool(split), att_use_scale=bool(scale), att_penalty_term=0.5,
                          batchnorm_momentum=0.99)
        P = {"tdnn/attention/query": torch.from_numpy(g[tag + "/query"].astype(np.float64))}
        att, pen = O.self_attention({"v": value, "k": key}, P, p, True, None)
        dv = value.shape[2]
        # the reference numpy adds 1e-12 inside the sqrt instead of flooring (test_utils.py:369): rtol 1e-3 as pooling.py:474
        assert np.allclose(att.numpy()[:, :dv], g[tag + "/att"][:, :dv], rtol=1e-6, atol=1e-9)
        assert np.allclose(att.numpy()[:, dv:], g[tag + "/att"][:, dv:], rtol=1e-3, atol=2e-6)
        assert np.allclose(pen.item(), float(g[tag + "/penalty"]), rtol=1e-8)


def test_ghost_vlad(golden_dir):
    """GhostVLAD pooling (model/pooling.py:195-277) against model/test_utils.py:421-436 compute_ghost_vlad."""
    g = np.load(os.path.join(golden_dir, "vlad.npz"))
    for tag in ("k8_g2", "k4_g0_final", "k5_g1_final"):
        k, gh, final = [int(v) for v in g[tag + "/cfg"]]
        value = torch.from_numpy(g[tag + "/value"].astype(np.float64))
        key = torch.from_numpy(g[tag + "/key"].astype(np.float64))
        p = O.ParamsPlain(vlad_value_input="v", vlad_key_input="k", vlad_key_num_nodes=[], vlad_value_num_nodes=[],
                          vlad_num_centers=k, vlad_num_ghosts=gh, vlad_final_l2_norm=bool(final), batchnorm_momentum=0.99)
        # the cluster logits ARE the key here: identity weight affine
        P = {"tdnn/vlad/vlad_weight_affine/kernel": torch.eye(k + gh, dtype=torch.float64),
             "tdnn/vlad/vlad_weight_affine/bias": torch.zeros(k + gh, dtype=torch.float64),
             "tdnn/vlad/vlad_centers": torch.from_numpy(g[tag + "/centers"].astype(np.float64))}
        out = O.ghost_vlad({"v": value, "k": key}, P, p, True, None)
        assert out.shape == (value.shape[0], k * value.shape[2])
        assert np.allclose(out.numpy(), g[tag + "/out"], rtol=1e-9, atol=1e-12)
        # full lengths == unmasked; a shorter length == the truncated utterance
        full = torch.full((value.shape[0],), value.shape[1], dtype=torch.int64)
        assert np.allclose(O.ghost_vlad({"v": value, "k": key}, P, p, True, None, full).numpy(), out.numpy(), rtol=1e-12)
        short = O.ghost_vlad({"v": value, "k": key}, P, p, True, None, full - 5)
        trunc = O.ghost_vlad({"v": value[:, :-5], "k": key[:, :-5]}, P, p, True, None)
        assert np.allclose(short.numpy(), trunc.numpy(), rtol=1e-12)


def test_metric_losses_match_reference_numpy(golden_dir):
    """Semi-hard / angular triplet losses and the GE2E validation loss (model/loss.py:358-705) against the reference's
    NumPy known answers (model/test_utils.py:118-154, 488-650, 21-86) on the data of model/tdnn.py:355-445."""
    g = np.load(os.path.join(golden_dir, "triplet.npz"))
    lab = torch.from_numpy(g["labels"].astype(np.int64))
    x = O.l2_scaling(torch.from_numpy(g["semihard/emb"].astype(np.float64)), 1.0)   # "L2 normalization should be applied before"
    for sq, m, want in g["semihard/cases"]:
        p = O.ParamsPlain(margin=float(m), triplet_loss_squared=bool(sq))
        assert np.allclose(float(O.semihard_triplet_loss(x, lab, p)), want, rtol=1e-9)
    x = torch.from_numpy(g["angular/emb"].astype(np.float64))
    names = ["asoftmax", "additive_margin_softmax", "additive_angular_margin_softmax"]
    for code, ti, m, want in g["angular/cases"]:
        p = O.ParamsPlain(margin=float(m), triplet_type=("all", "hard")[int(ti)], loss_type=names[int(code)])
        # arc form: the sqrt floor (1e-12, finite gradients at |cos| = 1) moves the duplicated pair by 1e-6 * sin(m)
        assert np.allclose(float(O.angular_triplet_loss(x, lab, p)), want, rtol=1e-7 if int(code) == 2 else 1e-9)
    p = O.ParamsPlain(num_valid_speakers_per_batch=int(g["num_speakers"]), num_valid_segments_per_speaker=int(g["num_segments"]))
    assert np.allclose(float(O.e2e_valid_loss(x, lab, p)), float(g["e2e/loss"]), rtol=1e-9)


def test_generalized_triplet_matches_reference_numpy(golden_dir):
    """generalized_angular_triplet_loss (model/loss.py:708-901) against model/test_utils.py:653-852
    compute_generalized_triplet_loss: triplet / centre / between-class parts for learnable and moving-average centres,
    top-n = 1, 0, k, and the updated centres themselves."""
    g = np.load(os.path.join(golden_dir, "gtriplet.npz"))
    w = torch.from_numpy(g["w"].astype(np.float64))
    for ci, (avg, topn, m, tm) in enumerate(g["cases"]):
        lab = torch.from_numpy(g["case%d/labels" % ci].astype(np.int64))
        x = torch.from_numpy(g["emb"][:len(lab)].astype(np.float64))
        p = O.ParamsPlain(triplet_center="average" if avg else "learnable", triplet_center_momentum=0.9, loss_compute="raw",
                          triplet_topn=int(topn), margin=float(m), target_margin=float(tm), triplet_loss_weight=1.0,
                          center_loss_weight=0.5, between_loss_weight=0.25, l2_loss_weight=0.0)
        updates = {}
        loss, parts = O.generalized_angular_triplet_loss(x, lab, {"softmax/output/kernel": w.clone()}, p, True, updates)
        want = g["case%d/parts" % ci]
        got = [float(parts[k]) for k in ("triplet_loss", "center_loss", "between_loss")]
        assert np.allclose(got, want, rtol=1e-9), (ci, got, want)
        assert np.allclose(float(loss), want[0] + 0.5 * want[1] + 0.25 * want[2], rtol=1e-9)
        assert np.allclose(parts["average_centers"].numpy(), g["case%d/w_update" % ci], rtol=1e-12, atol=1e-14)
        assert ("softmax/output/kernel" in updates) == bool(avg)
        # validation graph: averaged centres are NOT updated when is_training is False
        _, pv = O.generalized_angular_triplet_loss(x, lab, {"softmax/output/kernel": w.clone()}, p, False, {})
        assert torch.equal(pv["average_centers"], w)
    p = O.ParamsPlain(triplet_center="learnable", loss_compute="softplus", triplet_topn=1, margin=0.1, target_margin=0.0,
                      triplet_loss_weight=1.0, center_loss_weight=0.0, between_loss_weight=0.0, l2_loss_weight=0.0)
    with pytest.raises(NotImplementedError):          # loss.py:826
        O.generalized_angular_triplet_loss(x, lab, {"softmax/output/kernel": w}, p, True, {})


def test_aux_losses_match_reference_numpy(golden_dir):
    """Ring loss and MHE (model/loss.py:985-1037) against model/test_utils.py:855-884 (compute_ring_loss, compute_mhe)."""
    g = np.load(os.path.join(golden_dir, "aux.npz"))
    labels = torch.from_numpy(g["labels"].astype(np.int64))
    x = torch.from_numpy(g["emb"].astype(np.float64))
    P = {"softmax/output/kernel": torch.from_numpy(g["w"].astype(np.float64))}
    for r, lam, want in g["ring_cases"]:
        P["softmax_ringloss/r"] = torch.tensor(float(r), dtype=torch.float64)
        p = O.ParamsPlain(aux_loss_func=["ring_loss"], ring_loss_lambda=float(lam))
        assert np.allclose(float(O.aux_loss(x, labels, P, p)), want, rtol=1e-9)
    for lam, want in g["mhe_cases"]:
        p = O.ParamsPlain(aux_loss_func=["mhe_loss"], mhe_lambda=float(lam))
        # the NumPy known answer omits the 1e-6 the TF graph adds to the mean (test_utils.py:884 vs loss.py:1029)
        assert np.allclose(float(O.aux_loss(x, labels, P, p)), want, rtol=1e-5)
