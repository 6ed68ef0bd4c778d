"""Extraction (config 5) parity: batched, length-masked CUDA extraction == one-utterance-at-a-time oracle
(extract.py:65-94 semantics: skip short, chunk + length-weighted average for long, optional L2 normalisation)."""
import io

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O
from tests.xv_testlib import base_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("normalize", [False, True])
def test_batched_extraction_matches_oracle(normalize):
    from tf_kaldi_speaker_b200.dataset import kaldi_io
    from tf_kaldi_speaker_b200.extract import extract_embeddings
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    D = 30
    pd = base_params()
    po = O.ParamsPlain(**dict(pd))
    P = O.init_params(D, po, seed=4)
    g = torch.Generator().manual_seed(8)
    for k in P:      # non-trivial inference statistics and affine parameters
        if k.endswith("moving_mean") or k.endswith("/beta") or k.endswith("/bias"):
            P[k] = 0.1 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
        elif k.endswith("moving_variance"):
            P[k] = 0.5 + torch.rand(P[k].shape, generator=g, dtype=torch.float64)
        elif k.endswith("/gamma"):
            P[k] = 1 + 0.2 * torch.randn(P[k].shape, generator=g, dtype=torch.float64)
    lens = [24, 25, 26, 40, 133, 256, 257, 300, 301, 515, 700, 90]      # 24 is skipped; > 300 are chunked
    utts = []
    for i, t in enumerate(lens):
        m = torch.randn(1, D, generator=g)
        utts.append(("utt%02d" % i, (m + torch.randn(t, D, generator=g)).numpy().astype(np.float32)))
    tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_extract_test")
    tr.build("predict", D)
    tr.engine.store.load_tf({k: v.numpy() for k, v in P.items()})
    ark = io.BytesIO()
    for k, f in utts:
        kaldi_io.write_mat(ark, f, key=k)
    ark.seek(0)
    out = io.BytesIO()
    res = extract_embeddings(tr, kaldi_io.read_mat_ark(ark), out, chunk_size=300, min_chunk_size=25,
                             normalize=normalize, max_batch_frames=2048)
    assert [k for k, _ in res] == [k for k, f in utts if f.shape[0] >= 25]
    back = dict(kaldi_io.read_vec_flt_ark(io.BytesIO(out.getvalue())))
    for k, e in res:
        f = dict(utts)[k]
        ref = O.extract_embedding(torch.from_numpy(f).double(), P, po, chunk_size=300, min_chunk_size=25,
                                  normalize=normalize).numpy()
        cos = float(np.dot(e, ref) / (np.linalg.norm(e) * np.linalg.norm(ref)))
        assert cos >= 0.999, (k, f.shape[0], cos)
        assert np.linalg.norm(e - ref) / np.linalg.norm(ref) < 2e-2, (k, f.shape[0])
        assert np.array_equal(back[k], e)


def test_predict_matches_reference_shapes():
    """Trainer.predict: [T, D] -> [E] and [N, T, D] -> [N, E] (trainer.py:708-726); ragged == one at a time."""
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    D = 23
    pd = base_params()
    tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_predict_test")
    tr.build("predict", D)
    rng = np.random.RandomState(1)
    x = rng.randn(4, 120, D).astype(np.float32)
    e1 = tr.predict(x[0])
    eN = tr.predict(x)
    assert e1.shape == (512,) and eN.shape == (4, 512)
    assert np.allclose(e1, eN[0], rtol=1e-3, atol=1e-3)
    lens = np.array([120, 60, 33, 100], dtype=np.int32)
    er = tr.predict_batch_padded(x, lens)
    for i in range(4):
        ei = tr.predict(x[i, :lens[i]])
        assert np.allclose(er[i], ei, rtol=2e-3, atol=2e-3), i


def test_ragged_layout_equals_one_at_a_time():
    """Trainer.predict_ragged: utterances concatenated in one flat row space (no padding) give the embeddings of
    one predict() call per utterance -- the frames straddling two utterances never reach a pooled row -- and the same as
    the padded, length-masked batch."""
    from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
    from tf_kaldi_speaker_b200.model.trainer import Trainer
    D = 30
    pd = base_params()
    tr = Trainer(ParamsPlain(**dict(pd)), "/tmp/xv_ragged_test")
    tr.build("predict", D)
    rng = np.random.RandomState(3)
    lens = np.array([25, 300, 15 + 1, 77, 1000, 26, 513], dtype=np.int32)
    feats = [(10.0 * rng.randn(1, D) + rng.randn(int(t), D) * (1 + 5 * rng.rand())).astype(np.float32) for t in lens]
    starts = np.zeros_like(lens)
    starts[1:] = np.cumsum(lens)[:-1]
    er = tr.predict_ragged(np.concatenate(feats, 0), starts, lens)
    assert er.shape == (len(lens), 512)
    tmax = int(lens.max())
    padded = np.zeros((len(lens), tmax, D), dtype=np.float32)
    for i, f in enumerate(feats):
        padded[i, :f.shape[0]] = f
    ep = tr.predict_batch_padded(padded, lens)
    for i, f in enumerate(feats):
        ei = tr.predict(f)
        assert np.allclose(er[i], ei, rtol=2e-3, atol=2e-3), (i, int(lens[i]))
        assert np.allclose(er[i], ep[i], rtol=2e-3, atol=2e-3), (i, int(lens[i]))
    # order independence: the neighbours of an utterance do not matter
    perm = np.array([4, 0, 6, 2, 5, 1, 3])
    st2 = np.zeros_like(lens)
    st2[1:] = np.cumsum(lens[perm])[:-1]
    e2 = tr.predict_ragged(np.concatenate([feats[i] for i in perm], 0), st2, lens[perm])
    assert np.allclose(e2, er[perm], rtol=2e-3, atol=2e-3)
