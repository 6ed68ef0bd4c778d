#!/usr/bin/env python
"""Generate the committed golden vectors under tests/golden/ from the REFERENCE's own NumPy
known-answer functions (run in the build container, where /root/reference exists):

  heads.npz      model/test_utils.py:157-318 (compute_asoftmax / compute_amsoftmax / compute_arcsoftmax)
                 on the adversarial rows of model/tdnn.py:271-277 (theta~0, theta~pi, tiny norm, x10 norm),
                 m grids and feature_norm on/off as in model/tdnn.py:254-343.
  statpool.npz   the inline NumPy check ``compute_stat_pooling`` of model/multitask_v1/pooling.py:68-80, exec'd from the
                 reference source text (it lives under ``__main__``).
  aux.npz        model/test_utils.py:855-884 compute_ring_loss / compute_mhe (auxiliary losses of model/loss.py:985-1037).
  triplet.npz    model/test_utils.py:118-154 compute_triplet_loss, :488-650 *_angular_triplet_loss, :21-86 compute_ge2e_loss
                 (semi-hard / angular triplet losses and the softmax GE2E validation loss of model/loss.py:358-705).
  gtriplet.npz   model/test_utils.py:653-852 compute_generalized_triplet_loss (triplet / centre / between-class parts of
                 model/loss.py:708-901 for learnable and moving-average class centres, top-n = 1, 0, k).
  vlad.npz       model/test_utils.py:421-436 compute_ghost_vlad (NetVLAD / GhostVLAD pooling, model/pooling.py:195-277).
  attention.npz  model/test_utils.py:321-376 compute_self_attention, exec'd from the reference source with
                 the two py2 integer divisions (``value_dim/n_heads``, ``key_dim/n_heads``) turned into ``//``.

Usage: python tests/golden/make_golden.py   (writes next to this file; seeded, deterministic)
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("XV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


class ParamsPlain(object):      # misc/utils.py:44-61 stand-in (the original imports TensorFlow)
    @property
    def dict(self):
        return self.__dict__


def load_test_utils():
    sys.path.insert(0, REF)
    from model import test_utils          # numpy + six only
    return test_utils


def xavier(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def make_heads(tu):
    rng = np.random.RandomState(20240)
    n, d, c = 100, 512, 10
    labels = rng.randint(0, c, size=(n,)).astype(np.int32)
    w = xavier(rng, d, c)
    emb = rng.rand(n, d).astype(np.float32)
    emb[0] = w[:, labels[0]] + 1e-5          # theta ~ 0
    emb[1] = -1 * w[:, labels[1]] + 1e-5     # theta ~ pi
    emb[2] = 1e-4 * emb[2]                   # tiny norm
    emb_x10 = emb.copy()
    emb_x10[3] = 10 * emb_x10[3]             # large norm (asoftmax test only, tdnn.py:277)
    out = {"labels": labels, "w": w, "emb": emb, "emb_x10": emb_x10}
    cases = []
    p = ParamsPlain()
    for pre in ("asoftmax", "amsoftmax", "arcsoftmax"):
        p.dict[pre + "_lambda_min"] = 10
        p.dict[pre + "_lambda_base"] = 1000
        p.dict[pre + "_lambda_gamma"] = 1
        p.dict[pre + "_lambda_power"] = 4
    # the reference tests use feature_scaling_factor 0.1; add 2.0 as an extra scaled case
    for scaling, factor in ((True, 0.1), (False, 0.1), (True, 2.0)):
        p.dict["feature_norm"] = scaling
        p.dict["feature_scaling_factor"] = factor
        for m in (1, 2, 4):
            p.dict["global_step"] = 1
            p.dict["asoftmax_m"] = m
            v = tu.compute_asoftmax(emb_x10.astype(np.float64), labels, p, w.astype(np.float64))
            cases.append(("asoftmax", scaling, factor, float(m), 1, float(v)))
        for m in (0, 0.1, 0.5):
            p.dict["global_step"] = 1000
            p.dict["amsoftmax_m"] = m
            v = tu.compute_amsoftmax(emb.astype(np.float64), labels, p, w.astype(np.float64))
            cases.append(("additive_margin_softmax", scaling, factor, float(m), 1000, float(v)))
            p.dict["arcsoftmax_m"] = m
            v = tu.compute_arcsoftmax(emb.astype(np.float64), labels, p, w.astype(np.float64))
            cases.append(("additive_angular_margin_softmax", scaling, factor, float(m), 1000, float(v)))
    out["case_loss_type"] = np.array([c_[0] for c_ in cases])
    out["case_feature_norm"] = np.array([c_[1] for c_ in cases])
    out["case_scaling_factor"] = np.array([c_[2] for c_ in cases], dtype=np.float64)
    out["case_m"] = np.array([c_[3] for c_ in cases], dtype=np.float64)
    out["case_global_step"] = np.array([c_[4] for c_ in cases], dtype=np.int64)
    out["case_loss"] = np.array([c_[5] for c_ in cases], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "heads.npz"), **out)
    print("heads.npz: %d cases" % len(cases))


def make_statpool():
    # model/multitask_v1/pooling.py:68-80: the inline ``compute_stat_pooling`` of the reference's self-test, exec'd from
    # the reference SOURCE TEXT (it lives under ``__main__`` next to TensorFlow calls, so it cannot be imported);
    # inputs as at :62-64 (uniform features, row 0 all zeros, random lengths).
    src = open(os.path.join(REF, "model", "multitask_v1", "pooling.py")).read()
    m = re.search(r"^( *)def compute_stat_pooling\(features, length\):\n(?:\1 +.*\n|\s*\n)+", src, re.M)
    indent = len(m.group(1))
    body = "\n".join(line[indent:] for line in m.group(0).splitlines())
    ns = {"np": np, "range": range}
    exec(body, ns)
    fn = ns["compute_stat_pooling"]
    rng = np.random.RandomState(20241)
    n, l, d = 12, 96, 40
    x = rng.rand(n, l, d).astype(np.float32)
    x[0] = 0
    length = rng.randint(10, l + 1, size=(n,)).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "statpool.npz"), x=x, length=length, out=fn(x, length))
    print("statpool.npz (reference text exec'd: %d lines)" % len(body.splitlines()))


def make_aux(tu):
    # model/test_utils.py:855-884 compute_ring_loss / compute_mhe, the known answers of model/loss.py:1054-1087
    rng = np.random.RandomState(20243)
    n, d, c = 64, 128, 37
    labels = rng.randint(0, c, size=(n,)).astype(np.int32)
    w = xavier(rng, d, c)
    emb = (rng.randn(n, d) * (0.5 + rng.rand(n, 1) * 3)).astype(np.float32)
    p = ParamsPlain()
    out = {"labels": labels, "w": w, "emb": emb}
    ring, mhe = [], []
    for r, lam in ((0.1, 0.01), (20.0, 0.01), (5.0, 1.0)):
        p.dict["ring_loss_lambda"] = lam
        ring.append((r, lam, float(tu.compute_ring_loss(emb.astype(np.float64), p, r))))
    for lam in (0.01, 0.1, 1.0):
        p.dict["mhe_lambda"] = lam
        mhe.append((lam, float(tu.compute_mhe(labels, p, w.astype(np.float64).copy()))))
    out["ring_cases"] = np.array(ring, dtype=np.float64)
    out["mhe_cases"] = np.array(mhe, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "aux.npz"), **out)
    print("aux.npz: %d ring + %d mhe cases" % (len(ring), len(mhe)))


def make_attention(tu):
    src = open(os.path.join(REF, "model", "test_utils.py")).read()
    m = re.search(r"def compute_self_attention\(.*?\n(?=def )", src, re.S)
    body = m.group(0).replace("value_dim/n_heads", "value_dim//n_heads").replace("key_dim/n_heads", "key_dim//n_heads")
    ns = {"np": np, "softmax": tu.softmax, "range": range}
    exec(body, ns)
    fn = ns["compute_self_attention"]
    rng = np.random.RandomState(20242)
    out = {}
    # model/pooling.py:430-476 (commented-out self-test): rows *1e-8, =0, *100, =100
    for tag, (heads, split, scale, dk, dv) in {"h4_split": (4, True, True, 24, 32),
                                               "h1_shared": (1, False, True, 24, 32),
                                               "h2_shared_noscale": (2, False, False, 24, 32)}.items():
        b, l = 6, 20
        value = rng.rand(b, l, dv).astype(np.float32)
        key = rng.rand(b, l, dk).astype(np.float32)
        value[0] *= 1e-8
        value[1] = 0
        value[2] *= 100
        value[3] = 100
        query = (rng.randn(heads, dk // heads if split else dk) * 0.1).astype(np.float32)
        p = ParamsPlain()
        p.dict.update(att_split_key=split, att_use_scale=scale, att_penalty_term=0.5)
        att, pen = fn(value.astype(np.float64), key.astype(np.float64), query.astype(np.float64), p)
        out[tag + "/value"], out[tag + "/key"], out[tag + "/query"] = value, key, query
        out[tag + "/att"], out[tag + "/penalty"] = att, np.float64(pen)
        out[tag + "/cfg"] = np.array([heads, int(split), int(scale)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "attention.npz"), **out)
    print("attention.npz")


def make_vlad(tu):
    # model/test_utils.py:421-436 compute_ghost_vlad, the known answer of model/pooling.py:195-277
    rng = np.random.RandomState(20244)
    out = {}
    for tag, (k, g, final, b, l, d) in {"k8_g2": (8, 2, False, 5, 23, 40), "k4_g0_final": (4, 0, True, 4, 17, 24),
                                        "k5_g1_final": (5, 1, True, 6, 30, 72)}.items():
        value = (rng.randn(b, l, d) * (0.5 + rng.rand(b, 1, 1) * 2) + rng.randn(b, 1, d)).astype(np.float32)
        key = (rng.randn(b, l, k + g) * 2).astype(np.float32)
        centers = xavier(rng, k + g, d)
        p = ParamsPlain()
        p.dict.update(vlad_num_centers=k, vlad_num_ghosts=g, vlad_final_l2_norm=final)
        o = tu.compute_ghost_vlad(value.astype(np.float64), key.astype(np.float64), centers.astype(np.float64), p)
        out[tag + "/value"], out[tag + "/key"], out[tag + "/centers"], out[tag + "/out"] = value, key, centers, o
        out[tag + "/cfg"] = np.array([k, g, int(final)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "vlad.npz"), **out)
    print("vlad.npz")


def make_triplet(tu):
    # model/test_utils.py:118-154 compute_triplet_loss, :488-650 {asoftmax,amsoftmax,arcsoftmax}_angular_triplet_loss,
    # :21-86 compute_ge2e_loss -- the known answers of model/loss.py:358-705, on the data of model/tdnn.py:355-445
    # (speaker-ordered labels, one duplicated and one negated embedding)
    rng = np.random.RandomState(20245)
    n_spk, n_seg, dim = 6, 4, 16
    n = n_spk * n_seg
    labels = np.repeat(np.arange(n_spk), n_seg).astype(np.int32)
    out = {"labels": labels, "num_speakers": np.int64(n_spk), "num_segments": np.int64(n_seg)}
    emb = rng.rand(n, dim).astype(np.float32)
    emb[-1, :] = emb[-2, :]                                   # tdnn.py:359
    out["semihard/emb"] = emb
    semi = []
    for squared in (True, False):
        for margin in (0.2, 0.5):
            semi.append((float(squared), margin, float(tu.compute_triplet_loss(emb.astype(np.float64).copy(), labels, margin, squared))))
    out["semihard/cases"] = np.array(semi, dtype=np.float64)
    emb2 = rng.rand(n, dim).astype(np.float32) - 0.3
    emb2[1, :] = emb2[0, :]                                   # tdnn.py:377-378
    emb2[2, :] = -emb2[0, :]
    out["angular/emb"] = emb2
    ang = []
    kinds = {"asoftmax": (0, tu.asoftmax_angular_triplet_loss), "additive_margin_softmax": (1, tu.amsoftmax_angular_triplet_loss),
             "additive_angular_margin_softmax": (2, tu.arcsoftmax_angular_triplet_loss)}
    for ti, ttype in enumerate(("all", "hard")):
        for name, margins in (("asoftmax", (1, 2, 4)), ("additive_margin_softmax", (0.1, 0.35)),
                              ("additive_angular_margin_softmax", (0.3,))):
            code, fn = kinds[name]
            for m in margins:
                ang.append((code, ti, float(m), float(fn(emb2.astype(np.float64).copy(), labels, m, ttype))))
    out["angular/cases"] = np.array(ang, dtype=np.float64)
    out["e2e/loss"] = np.float64(tu.compute_ge2e_loss(emb2.astype(np.float64).copy(), labels, 20, 0, "softmax"))
    np.savez_compressed(os.path.join(HERE, "triplet.npz"), **out)
    print("triplet.npz: %d semihard + %d angular cases + e2e" % (len(semi), len(ang)))


def make_gtriplet(tu):
    # model/test_utils.py:653-852 compute_generalized_triplet_loss, the known answer of model/loss.py:708-901
    rng = np.random.RandomState(20246)
    n, dim, c = 24, 16, 11
    emb = (rng.randn(n, dim) * (0.5 + rng.rand(n, 1))).astype(np.float32)
    w = xavier(rng, dim, c)
    out = {"emb": emb, "w": w}
    cases = []
    for ci, (center, topn, margin, tmargin, repeat) in enumerate([("learnable", 1, 0.3, 0.2, True), ("learnable", 0, 0.1, 0.5, True),
                                                                  ("learnable", 3, 0.3, 0.0, True), ("average", 1, 0.3, 0.2, False),
                                                                  ("average", 4, 0.2, 0.1, False)]):
        # the NumPy code updates the averaged centres one sample at a time, the graph all at once (scatter_nd): they agree
        # only when no label repeats, so the "average" cases use distinct labels
        labels = (rng.randint(0, c, size=(n,)) if repeat else rng.permutation(c)[:min(n, c)]).astype(np.int32)
        e = emb[:len(labels)]
        p = ParamsPlain()
        p.dict.update(triplet_center=center, triplet_center_momentum=0.9, loss_compute="raw", triplet_topn=topn, margin=margin,
                      target_margin=tmargin, center_loss_weight=0.5, between_loss_weight=0.25, l2_loss_weight=0.0)
        loss, w_upd = tu.compute_generalized_triplet_loss(e.astype(np.float64).copy(), w.astype(np.float64).copy(), labels, p, c)
        out["case%d/labels" % ci] = labels
        out["case%d/w_update" % ci] = np.asarray(w_upd)
        out["case%d/parts" % ci] = np.array([float(loss["triplet_loss"]), float(loss["center_loss"]), float(loss["between_loss"])])
        cases.append((float(center == "average"), topn, margin, tmargin))
    out["cases"] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "gtriplet.npz"), **out)
    print("gtriplet.npz: %d cases" % len(cases))


if __name__ == "__main__":
    tu = load_test_utils()
    make_gtriplet(tu)
    make_triplet(tu)
    make_vlad(tu)
    make_heads(tu)
    make_statpool()
    make_attention(tu)
    make_aux(tu)
