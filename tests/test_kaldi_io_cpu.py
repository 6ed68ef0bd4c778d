"""Host-side ark I/O on either side of the extraction path (reference dataset/kaldi_io.py:624-655, 683-740) and the
chunking rule of extract.py:69-80."""
import io
import struct

import numpy as np

from tf_kaldi_speaker_b200.dataset import kaldi_io
from tf_kaldi_speaker_b200.extract import split_chunks


def test_write_vec_flt_bytes_match_kaldi_layout():
    v = np.arange(5, dtype=np.float32) * 0.5
    buf = io.BytesIO()
    kaldi_io.write_vec_flt(buf, v, key="utt-1")
    raw = buf.getvalue()
    assert raw == b"utt-1 " + b"\0B" + b"FV " + b"\x04" + struct.pack("<I", 5) + v.tobytes()
    (k, r), = list(kaldi_io.read_vec_flt_ark(io.BytesIO(raw)))
    assert k == "utt-1" and np.array_equal(r, v)


def test_mat_ark_roundtrip_and_ragged():
    rng = np.random.RandomState(0)
    mats = [("a", rng.randn(3, 4).astype(np.float32)), ("bb", rng.randn(1, 4).astype(np.float64)),
            ("c", np.zeros((0, 4), dtype=np.float32))]
    buf = io.BytesIO()
    for k, m in mats:
        kaldi_io.write_mat(buf, m, key=k)
    buf.seek(0)
    got = list(kaldi_io.read_mat_ark(buf))
    assert [k for k, _ in got] == ["a", "bb", "c"]
    for (k, m), (_, g) in zip(mats, got):
        assert g.shape == m.shape and np.array_equal(g, m)
    assert list(kaldi_io.read_mat_ark(io.BytesIO(b""))) == []


def test_split_chunks_follows_extract_py():
    assert split_chunks(100, 10000) == [(0, 100)]
    assert split_chunks(10000, 10000) == [(0, 10000)]
    assert split_chunks(10001, 10000) == [(0, 10000), (5000, 5001)]
    assert split_chunks(130, 50) == [(0, 50), (25, 50), (50, 50), (75, 50), (100, 30)]
    # every frame is covered and only the last chunk is shorter
    for t, cs in ((25001, 10000), (777, 100)):
        ch = split_chunks(t, cs)
        assert ch[-1][0] + ch[-1][1] == t and all(l == cs for _, l in ch[:-1])
