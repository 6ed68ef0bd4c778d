"""xv_cm_decode on the GPU against the REFERENCE reader's golden vectors (tests/golden/cm_golden.npz, made by
dataset/kaldi_io.py:767-868 of the reference) and against the oracle on random bytes.  Bit-exact."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import kaldi_cm_oracle as CM

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    gd = np.load(os.path.join(golden_dir, "cm_golden.npz"))
    ark = open(os.path.join(golden_dir, "cm_golden.ark"), "rb").read()
    offs = dict(zip([str(k) for k in gd["__offsets_keys"]], [int(o) for o in gd["__offsets"]]))
    return gd, ark, offs


def test_decode_matches_reference_reader(golden_dir):
    from tf_kaldi_speaker_b200.dataset import kaldi_io as K
    from tf_kaldi_speaker_b200.dataset.feeder import CompressedSegmentBatch
    gd, ark, offs = _golden(golden_dir)
    n = 0
    for name in gd.files:
        if name.startswith("__"):
            continue
        parts = name.split("/")
        fd = io.BytesIO(ark)
        fd.seek(offs[parts[1]])
        raw = K.read_compressed_raw(fd) if parts[0] == "full" else K.read_compressed_raw(fd, int(parts[2]), int(parts[3]))
        batch = CompressedSegmentBatch(1, raw.data.shape[1], raw.cols)
        batch.set(0, raw)
        got = batch.to_device().cpu().numpy()[0]
        want = gd[name]
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, np.abs(got - want).max())
        n += 1
    assert n >= 16


@pytest.mark.parametrize("B,T,D", [(128, 200, 30), (7, 333, 23), (3, 31, 80), (2, 1, 1), (1, 10000, 30)])
def test_decode_random_batches_bit_exact(B, T, D):
    """Batches of segments cut from different matrices (own min / range / percentiles each), ragged tile edges, the
    extraction-size single utterance; every byte value occurs."""
    from tf_kaldi_speaker_b200.dataset.feeder import CompressedSegmentBatch
    rng = np.random.RandomState(B * 1000 + T + D)
    batch = CompressedSegmentBatch(B, T, D)
    batch.data[:] = rng.randint(0, 256, size=(B, D, T)).astype(np.uint8)
    batch.headers[:] = np.sort(rng.randint(0, 65536, size=(B, D, 4)), axis=-1).astype(np.uint16)
    batch.glob[:, 0] = rng.randn(B).astype(np.float32) * 20
    batch.glob[:, 1] = np.abs(rng.randn(B)).astype(np.float32) * 50 + 1e-3
    out = torch.full((B, T, D), float("nan"), device="cuda")
    got = batch.decode_into(out).cpu().numpy()
    for b in range(B):
        want = CM.decode(batch.glob[b, 0], batch.glob[b, 1], batch.headers[b], batch.data[b])
        assert np.array_equal(got[b].view(np.uint32), want.view(np.uint32)), (b, np.abs(got[b] - want).max())


def test_train_step_accepts_compressed_batch(golden_dir):
    """Trainer.train_step fed with a CompressedSegmentBatch == the same step fed with the host-decoded float32 batch."""
    from tf_kaldi_speaker_b200.dataset import kaldi_io as K
    from tf_kaldi_speaker_b200.dataset.feeder import CompressedSegmentBatch
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    from tests.xv_testlib import base_params
    gd, ark, offs = _golden(golden_dir)
    B, T, D, C = 8, 60, 30, 50
    batch = CompressedSegmentBatch(B, T, D)
    dense = np.zeros((B, T, D), dtype=np.float32)
    for i in range(B):
        fd = io.BytesIO(ark)
        fd.seek(offs["spk1-utt1"])
        batch.set(i, K.read_compressed_raw(fd, 25 * i, T))
        dense[i] = gd["full/spk1-utt1"][25 * i:25 * i + T]
    y = np.arange(B, dtype=np.int32) % C
    losses = []
    for feats in (batch, dense):
        pd = base_params(cuda_graph=False)
        tr = Trainer(ParamsPlain(**pd), "/tmp/xv_cm_test_model")
        tr.build("train", D, "softmax", C)
        r = tr.train_step(feats, y, 0.01, 0, fetch_loss=True)
        losses.append(r["raw_loss"])
    assert abs(losses[0] - losses[1]) <= 2e-3 * abs(losses[1]), losses      # same inputs; BN-statistic atomics reorder
