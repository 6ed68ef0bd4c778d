"""CPU tests of the host-side logic that mirrors the reference: the parameter schema (TF variable names, padded
internal layouts and their round trip), Params loading of the shipped nnet_conf JSON keys, the margin-annealing
schedule (loss.py:144-147), work-per-segment formulas and the oracle's optimizers."""
import json
import math
import os

import numpy as np
import torch

from oracle import xvector_oracle as O
from tf_kaldi_speaker_b200.misc.utils import Params, ParamsPlain
from tf_kaldi_speaker_b200.model.loss import margin_lambda, margin_schedule
from tf_kaldi_speaker_b200.runtime import VarSpec, ParamStore, _pad_to


def test_params_json_roundtrip(tmp_path):
    # keys of egs/voxceleb/v1/nnet_conf/tdnn_arcsoftmax_m0.20_linear_bn_1e-2.json (strings for numerics are legal there)
    cfg = {"seed": 0, "network_type": "tdnn", "last_layer_linear": True, "loss_func": "additive_angular_margin_softmax",
           "arcsoftmax_m": 0.20, "arcsoftmax_lambda_min": "0", "arcsoftmax_lambda_base": 1000,
           "arcsoftmax_lambda_gamma": 0.00001, "arcsoftmax_lambda_power": 5, "pooling_type": "statistics_pooling",
           "embedding_node": "tdnn6_dense", "weight_l2_regularizer": 1e-2, "batchnorm_momentum": 0.99}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(cfg))
    params = Params(str(p))
    assert params.arcsoftmax_m == 0.2 and params.dict["pooling_type"] == "statistics_pooling"
    params.dict["num_nodes_pooling_layer"] = 1500           # operators inject defaults through .dict (tdnn.py:111-113)
    assert params.num_nodes_pooling_layer == 1500
    params.save(str(tmp_path / "out.json"))
    assert json.load(open(tmp_path / "out.json"))["num_nodes_pooling_layer"] == 1500


def test_margin_schedule_matches_reference_formula():
    # lambda = 1000, 620.9, 31.25, 10 (clamped) at global_step 0, 1e4, 1e5, 1e6 for (min 10, base 1000, 1e-5, 5)
    for step, want in ((0, 1000.0), (10 ** 4, 620.921), (10 ** 5, 31.25), (10 ** 6, 10.0)):
        lam, fa, fs = margin_lambda(10, 1000, 1e-5, 5, step)
        assert abs(lam - want) / want < 1e-4
        assert abs(fa - 1 / (1 + lam)) < 1e-15 and abs(fa + fs - 1) < 1e-15
    p = ParamsPlain(asoftmax_m=4, asoftmax_lambda_min="10", asoftmax_lambda_base=1000, asoftmax_lambda_gamma=1e-5,
                    asoftmax_lambda_power=5)
    fa, fs = margin_schedule("asoftmax", p, 10 ** 5)
    assert abs(fa - 1 / 32.25) < 1e-12
    p.asoftmax_m = 1
    assert margin_schedule("asoftmax", p, 5) == (0.0, 1.0)       # m = 1: plain xent, no lambda (loss.py:110-115)


def test_varspec_padded_layout_roundtrip():
    rng = np.random.RandomState(0)
    dim, dpad = 30, 32
    rm = (np.arange(5)[:, None] * dpad + np.arange(dim)[None, :]).reshape(-1)
    s = VarSpec("tdnn/tdnn1_conv/kernel", (1, 5, dim, 512), (192, 512), row_map=rm)
    k = rng.randn(1, 5, dim, 512).astype(np.float32)
    ki = s.to_internal(k)
    assert ki.shape == (192, 512)
    assert np.array_equal(ki[2 * dpad + 7], k[0, 2, 7]) and not ki[dim:dpad].any() and not ki[160:].any()
    assert np.array_equal(s.to_tf(ki), k)
    P, Pp = 1500, _pad_to(1500, 64)
    rows = np.concatenate([np.arange(P), Pp + np.arange(P)])
    s6 = VarSpec("tdnn/tdnn6_dense/kernel", (2 * P, 512), (2 * Pp, 512), row_map=rows)
    w = rng.randn(2 * P, 512).astype(np.float32)
    wi = s6.to_internal(w)
    assert np.array_equal(wi[Pp + 3], w[P + 3]) and not wi[P:Pp].any()
    assert np.array_equal(s6.to_tf(wi), w)
    g = VarSpec("tdnn/tdnn5_bn/moving_variance", (P,), (Pp,), trainable=False, pad_value=1.0)
    gi = g.to_internal(np.full(P, 0.5, dtype=np.float32))
    assert gi[:P].max() == 0.5 and (gi[P:] == 1.0).all()


def test_flops_per_segment_match_baseline_md():
    assert abs(O.flops_fwd(200, 30, 7200) / 1e9 - 1.6102) < 5e-4
    assert abs(O.flops_train(200, 30, 7200) / 1e9 - 4.8006) < 5e-4
    assert abs(O.flops_train(400, 30, 7200) / 1e9 - 9.8731) < 5e-4
    assert abs(O.flops_train(200, 23, 4300) / 1e9 - 4.7776) < 5e-4


def test_oracle_optimizers_follow_tf_semantics():
    P = {"w/kernel": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    g = {"w/kernel": torch.tensor([0.5, 0.25], dtype=torch.float64)}
    p1, _ = O.apply_optimizer(P, g, {}, O.ParamsPlain(optimizer="sgd"), 0.1)
    assert torch.allclose(p1["w/kernel"], torch.tensor([0.95, -2.025], dtype=torch.float64))
    pm = O.ParamsPlain(optimizer="momentum", momentum=0.9, use_nesterov=False)
    p2, st = O.apply_optimizer(P, g, {}, pm, 0.1)
    p3, st = O.apply_optimizer(p2, g, st, pm, 0.1)          # accum = 0.9*g + g
    assert torch.allclose(p3["w/kernel"], P["w/kernel"] - 0.1 * g["w/kernel"] - 0.1 * 1.9 * g["w/kernel"])
    pa = O.ParamsPlain(optimizer="adam")
    p4, _ = O.apply_optimizer(P, g, {}, pa, 0.1)            # first Adam step moves by ~lr * sign(g)
    assert torch.allclose(p4["w/kernel"], P["w/kernel"] - 0.1 * torch.sign(g["w/kernel"]), atol=1e-6)


def test_extraction_chunking_rule():
    """extract.py:69-87: n = ceil((T - cs)/(cs/2)) + 1 chunks, hop cs/2, length-weighted mean of the chunk embeddings."""
    po = O.ParamsPlain(weight_l2_regularizer=1e-2, batchnorm_momentum=0.99, pooling_type="statistics_pooling",
                       embedding_node="tdnn6_dense", num_nodes_pooling_layer=64)
    P = O.init_params(8, po, seed=1)
    x = torch.randn(130, 8, dtype=torch.float64)
    e = O.extract_embedding(x, P, po, chunk_size=50, min_chunk_size=25)
    starts, lens = [0, 25, 50, 75, 100], [50, 50, 50, 50, 30]
    embs = torch.stack([O.predict(x[s:s + l], P, po) for s, l in zip(starts, lens)])
    ln = torch.tensor(lens, dtype=torch.float64).unsqueeze(1)
    assert torch.allclose(e, (embs * ln).sum(0) / ln.sum(), rtol=1e-10)
    assert O.extract_embedding(x[:20], P, po, chunk_size=50, min_chunk_size=25) is None


def test_extraction_packers_lay_out_batches_exactly():
    """extract._pack_ragged (padding-free layout: chunks back to back + (starts, lengths)) and extract._pack (zero-padded
    [N, tmax, D] + lengths), including the multi-threaded copy path and slot reuse with a smaller batch."""
    from tf_kaldi_speaker_b200 import extract as X
    rng = np.random.RandomState(3)
    lens = [25, 300, 7000, 26, 513, 9000, 100, 4000, 64, 2500, 1200, 33, 800, 5000]
    jobs = [(i, 0, rng.randn(n, 30).astype(np.float32)) for i, n in enumerate(lens)]
    slot = {"buf": None, "event": None}
    old_threads = X._COPY_THREADS
    try:
        for threads, group in ((1, list(range(len(jobs)))), (4, list(range(len(jobs)))), (4, [5, 2, 9])):
            X._COPY_THREADS, X._copy_executor = threads, None
            flat, (starts, lengths) = X._pack_ragged(slot, group, jobs, 30)
            assert flat.shape == (sum(lens[j] for j in group), 30)
            assert list(lengths) == [lens[j] for j in group]
            assert list(starts) == list(np.concatenate([[0], np.cumsum([lens[j] for j in group])[:-1]]))
            for k, j in enumerate(group):
                assert np.array_equal(flat.numpy()[starts[k]:starts[k] + lengths[k]], jobs[j][2])
    finally:
        X._COPY_THREADS, X._copy_executor = old_threads, None
    group = [0, 3, 11, 8]
    batch, lengths = X._pack({"buf": None, "event": None}, group, jobs, 256, 30)
    assert batch.shape == (4, 256, 30) and list(lengths) == [25, 26, 33, 64]
    for r, j in enumerate(group):
        assert np.array_equal(batch.numpy()[r, :lens[j]], jobs[j][2])
        assert not batch.numpy()[r, lens[j]:].any()
