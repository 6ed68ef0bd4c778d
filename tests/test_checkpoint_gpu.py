"""Trainer.save / load (model/trainer.py:142-166): variables AND optimizer slots survive a round trip in both on-disk
formats -- the package's .npz and tf.train.Saver's tensor bundle with TF slot names (<var>/Momentum, <var>/Adam,
<var>/Adam_1, beta1_power) -- old checkpoints are pruned to keep_checkpoint_max, and training resumed from the file takes
the same next step as training that never stopped."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.xv_testlib import base_params, head_params, make_batch

pytestmark = pytest.mark.gpu

AAM = "additive_angular_margin_softmax"


@pytest.mark.parametrize("fmt,opt", [("npz", "momentum"), ("tf", "momentum"), ("tf", "adam"), ("npz", "sgd")])
def test_save_load_roundtrip_with_optimizer_slots(tmp_path, fmt, opt):
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    B, T, D, C = 16, 60, 30, 120
    pd = base_params(**head_params(AAM))
    pd.update(feature_norm=True, feature_scaling_factor=64, optimizer=opt, checkpoint_format=fmt, keep_checkpoint_max=2,
              cuda_graph=False)
    if opt == "momentum":
        pd["momentum"] = 0.9
    model = str(tmp_path / "exp")
    tr = Trainer(ParamsPlain(**dict(pd)), model)
    tr.build("train", D, AAM, C)
    for i in range(3):
        x, y = make_batch(B, T, D, C, seed=40 + i)
        tr.train_step(x, y, 0.01, i)
        tr.save(i + 1)
    nnet = os.path.join(model, "nnet")
    kept = sorted(glob.glob(os.path.join(nnet, "model-*.npz" if fmt == "npz" else "model-*.index")))
    assert [os.path.basename(p).split(".")[0] for p in kept] == ["model-2", "model-3"]          # keep_checkpoint_max = 2
    if fmt == "tf":
        assert sorted(glob.glob(os.path.join(nnet, "model-*.data-*"))) == [p[:-6] + ".data-00000-of-00001" for p in kept]
        assert not glob.glob(os.path.join(nnet, ".tmp-*"))
    st = tr.engine.store
    want = st.export_tf()
    want_s1 = st.export_tf(which="state1")
    want_s2 = st.export_tf(which="state2")

    tr2 = Trainer(ParamsPlain(**dict(pd)), model)
    tr2.build("train", D, AAM, C)
    assert tr2.load() == 3 and tr2.global_step == 3
    got = tr2.engine.store.export_tf()
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    if opt != "sgd":
        got_s1 = tr2.engine.store.export_tf(which="state1")
        assert want_s1 and all(np.array_equal(got_s1[k], want_s1[k]) for k in want_s1)
        assert any(np.abs(v).max() > 0 for v in want_s1.values())
    if opt == "adam":
        got_s2 = tr2.engine.store.export_tf(which="state2")
        assert all(np.array_equal(got_s2[k], want_s2[k]) for k in want_s2) and tr2.adam_t == tr.adam_t == 3
    # the resumed trainer and the uninterrupted one take the same step 4
    x, y = make_batch(B, T, D, C, seed=50)
    ra = tr.train_step(x, y, 0.01, 3, fetch_loss=True)
    rb = tr2.train_step(x, y, 0.01, 3, fetch_loss=True)
    torch.cuda.synchronize()
    assert abs(ra["raw_loss"] - rb["raw_loss"]) <= 5e-4 * abs(ra["raw_loss"])
    pa, pb = st.export_tf(), tr2.engine.store.export_tf()
    num = sum(float(np.linalg.norm(pa[k].astype(np.float64) - pb[k])) ** 2 for k in pa)
    den = sum(float(np.linalg.norm(pa[k].astype(np.float64) - want[k])) ** 2 for k in pa)
    assert (num / den) ** 0.5 <= 0.25, (num / den) ** 0.5          # same update up to the run-to-run rounding floor (16 segments)
