"""The reference's entry points on a synthetic Kaldi directory of compressed archives: the
training driver with the command line of egs/voxceleb/v1/nnet/lib/train.py:13-23 -- epochs of Trainer.train(data_dir,
spklist, lr) / Trainer.valid(data_dir, spklist, ...), variable segment lengths, checkpoints, the learning_rate / valid_loss
/ feature_dim files, ``-c`` continuation -- and the command line of nnet/lib/extract.py:11-96 that wrap/extract_wrapper.sh
invokes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O
from tests.xv_testlib import make_kaldi_dir

pytestmark = pytest.mark.gpu

CONFIG = dict(seed=0, network_type="tdnn", last_layer_no_bn=False, last_layer_linear=True, feature_norm=True,
              feature_scaling_factor=64, loss_func="additive_angular_margin_softmax", arcsoftmax_m=0.2,
              arcsoftmax_lambda_min=0, arcsoftmax_lambda_base=1000, arcsoftmax_lambda_gamma=1e-5, arcsoftmax_lambda_power=5,
              pooling_type="statistics_pooling", embedding_node="tdnn6_dense", weight_l2_regularizer=1e-2,
              batchnorm_momentum=0.99, optimizer="momentum", momentum=0.9, use_nesterov=False, clip_gradient=False,
              clip_gradient_norm=3, learning_rate=0.01, num_epochs=2, num_steps_per_epoch=8, show_training_progress=4,
              save_summary_steps=100, save_checkpoints_steps=5, keep_checkpoint_max=3, valid_max_iterations=3,
              reduce_lr_epochs=2, num_parallel_datasets=2, max_queue_size=4, num_speakers_per_batch=6,
              num_segments_per_speaker=2, min_segment_len=60, max_segment_len=90, batch_type="softmax", data_seed=1)


def test_train_driver_and_extract_cli(tmp_path):
    from tf_kaldi_speaker_b200.dataset.kaldi_io import read_vec_flt_ark, write_mat
    from tf_kaldi_speaker_b200.misc.utils import Params
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    from tf_kaldi_speaker_b200.nnet import train as train_driver
    from tf_kaldi_speaker_b200 import extract as extract_cli

    data, spklist, _ = make_kaldi_dir(tmp_path / "train", num_speakers=8, utts_per_speaker=4, dim=30, min_frames=150,
                                      max_frames=300, seed=1)
    vdata, vspk, _ = make_kaldi_dir(tmp_path / "valid", num_speakers=8, utts_per_speaker=4, dim=30, min_frames=150,
                                    max_frames=300, seed=2)
    cfg = str(tmp_path / "config.json")
    with open(cfg, "w") as f:
        json.dump(CONFIG, f)
    model = str(tmp_path / "exp")
    assert train_driver.main(["--config", cfg, data, spklist, vdata, vspk, model]) == 0
    nnet = os.path.join(model, "nnet")
    assert open(os.path.join(nnet, "feature_dim")).read().strip() == "30"
    assert 'model-16' in open(os.path.join(nnet, "checkpoint")).read()
    assert os.path.isfile(os.path.join(nnet, "model-16.npz"))
    assert len(open(os.path.join(nnet, "learning_rate")).read().strip().splitlines()) == 3      # epoch 0 twice + epoch 1
    vl = [l.split() for l in open(os.path.join(nnet, "valid_loss")).read().strip().splitlines()]
    assert len(vl) == 2 and all(np.isfinite(float(x[1])) and 0.0 <= float(x[2]) <= 1.0 for x in vl)
    z = np.load(os.path.join(nnet, "model-16.npz"))
    assert any(k.startswith("__slot:") and k.endswith("/Momentum") for k in z.files)           # optimizer slots saved
    # continue training for one more epoch from the checkpoint (-c): resumes at step 16
    c2 = dict(CONFIG)
    c2["num_epochs"] = 3
    with open(os.path.join(nnet, "config.json"), "w") as f:
        json.dump(c2, f)
    assert train_driver.main(["-c", "--config", cfg, data, spklist, vdata, vspk, model]) == 0
    assert 'model-24' in open(os.path.join(nnet, "checkpoint")).read()

    # ---- extraction command line on a float ark: short (skipped), normal and > chunk-size utterances
    rng = np.random.RandomState(5)
    ark = str(tmp_path / "feats.ark")
    utts = [("tooshort", 20), ("u1", 130), ("u2", 77), ("long", 260), ("u3", 200)]
    feats = {}
    with open(ark, "wb") as f:
        for key, n in utts:
            m = (rng.randn(1, 30) + rng.randn(n, 30)).astype(np.float32)
            feats[key] = m
            write_mat(f, m, key=key)
    out_ark = str(tmp_path / "xvector.ark")
    assert extract_cli.main(["-g", "0", "-s", "100", "-m", "25", model, "ark:" + ark, out_ark]) == 0
    got = dict(read_vec_flt_ark(out_ark))
    assert list(got.keys()) == ["u1", "u2", "long", "u3"]          # input order, the short utterance skipped
    # oracle on the trained parameters: chunk + length-weighted average rule of extract.py:65-94
    params = Params(os.path.join(nnet, "config.json"))
    tr = Trainer(params, model)
    tr.build("predict", dim=30)
    tr.load()
    P = {k: torch.from_numpy(v).double() for k, v in tr.engine.store.export_tf().items()}
    po = O.ParamsPlain(**dict(params.dict))
    for key in got:
        ref = O.extract_embedding(torch.from_numpy(feats[key]).double(), P, po, chunk_size=100, min_chunk_size=25).numpy()
        cos = float(np.dot(got[key], ref) / (np.linalg.norm(got[key]) * np.linalg.norm(ref)))
        assert cos >= 0.999, (key, cos)
        assert got[key].dtype == np.float32 and got[key].shape == (512,)


def test_train_driver_angular_triplet_end2end_valid(tmp_path):
    """The same epoch loop with a metric-learning loss: loss_func = angular_triplet_loss on speakers x segments batches,
    validated with the softmax GE2E loss on ``batch_type = end2end`` batches (model/trainer.py:272-275, 669-672)."""
    from tf_kaldi_speaker_b200.nnet import train as train_driver
    data, spklist, _ = make_kaldi_dir(tmp_path / "train", num_speakers=8, utts_per_speaker=4, dim=30, min_frames=150,
                                      max_frames=300, seed=3)
    vdata, vspk, _ = make_kaldi_dir(tmp_path / "valid", num_speakers=8, utts_per_speaker=4, dim=30, min_frames=150,
                                    max_frames=300, seed=4)
    c = dict(CONFIG)
    for k in [k for k in c if k.startswith("arcsoftmax")]:
        del c[k]
    c.update(loss_func="angular_triplet_loss", loss_type="additive_margin_softmax", triplet_type="all", margin=0.2,
             feature_norm=False, batch_type="end2end", num_valid_speakers_per_batch=4, num_valid_segments_per_speaker=3,
             num_epochs=1, num_steps_per_epoch=6, embedding_node="tdnn7_dense")
    c.pop("feature_scaling_factor")
    cfg = str(tmp_path / "config.json")
    with open(cfg, "w") as f:
        json.dump(c, f)
    model = str(tmp_path / "exp")
    assert train_driver.main(["--config", cfg, data, spklist, vdata, vspk, model]) == 0
    nnet = os.path.join(model, "nnet")
    assert 'model-6' in open(os.path.join(nnet, "checkpoint")).read()
    z = np.load(os.path.join(nnet, "model-6.npz"))
    assert not any(k.startswith("softmax/") for k in z.files)            # no speaker matrix with a metric-learning loss
    vl = [l.split() for l in open(os.path.join(nnet, "valid_loss")).read().strip().splitlines()]
    # GE2E cross entropy over 4 speakers: finite, positive, below the uniform-posterior value by a margin after training
    assert len(vl) == 1 and 0.0 < float(vl[0][1]) < 2.0 * np.log(4) and 0.0 <= float(vl[0][2]) <= 1.0
