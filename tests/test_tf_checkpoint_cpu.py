"""TF tensor-bundle (V2 checkpoint) reader / writer restated without TensorFlow (misc/tf_checkpoint.py): the format the
reference's tf.train.Saver writes (model/trainer.py:142-166).  Round trips, the crc32c known answer, the table layout
(footer magic, block trailers) and the Trainer-facing variable names.  No TF-written file exists in this image."""
import os
import struct

import numpy as np
import pytest

from tf_kaldi_speaker_b200.misc import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283            # the standard CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA               # 32 zero bytes (RFC 3720 B.4)
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    for c in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert T.unmask_crc(T.mask_crc(c)) == c
    assert T.mask_crc(T.crc32c(b"foo")) != T.crc32c(b"foo")


def _variables(rng, many=False):
    v = {"tdnn/tdnn1_conv/kernel": rng.randn(1, 5, 30, 512).astype(np.float32),
         "tdnn/tdnn1_conv/bias": rng.randn(512).astype(np.float32),
         "tdnn/tdnn1_bn/moving_variance": np.abs(rng.randn(512)).astype(np.float32),
         "softmax/output/kernel": rng.randn(512, 37).astype(np.float32),
         "global_step": np.array(123456, dtype=np.int64),
         "tdnn/tdnn1_conv/kernel/Momentum": rng.randn(1, 5, 30, 512).astype(np.float32)}
    if many:        # more than one data block in the index, shared key prefixes across restart points
        for i in range(300):
            v["aux/layer_%03d/w" % i] = rng.randn(3, 2).astype(np.float64)
    return v


@pytest.mark.parametrize("many", [False, True])
def test_roundtrip(tmp_path, many):
    rng = np.random.RandomState(3)
    v = _variables(rng, many)
    prefix = str(tmp_path / "model-1000")
    T.write_tf_checkpoint(prefix, v)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    meta = T.list_variables(prefix)
    assert sorted(meta) == sorted(v)
    assert meta["tdnn/tdnn1_conv/kernel"][:2] == (np.dtype(np.float32), (1, 5, 30, 512))
    assert meta["global_step"][:2] == (np.dtype(np.int64), ())
    got = T.read_tf_checkpoint(prefix, verify_crc=True)
    for k in v:
        assert got[k].dtype == v[k].dtype and got[k].shape == v[k].shape and np.array_equal(got[k], v[k]), k
    only = T.read_tf_checkpoint(prefix, names={"global_step"})
    assert list(only) == ["global_step"] and int(only["global_step"]) == 123456


def test_table_layout_and_corruption(tmp_path):
    rng = np.random.RandomState(4)
    prefix = str(tmp_path / "model-7")
    T.write_tf_checkpoint(prefix, _variables(rng))
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xdb4775248b80fb57 and len(idx) > 48
    # first data block: type byte 0 and a valid masked crc over (block + type) right after it
    pos = len(idx) - 48
    _, pos = T._get_varint(idx, pos); msz, pos = T._get_varint(idx, pos)
    ioff, pos = T._get_varint(idx, pos); isz, pos = T._get_varint(idx, pos)
    first = T._read_block(idx, ioff, isz)[0][1]
    boff, p = T._get_varint(first, 0); bsz, p = T._get_varint(first, p)
    assert boff == 0 and idx[bsz] == 0
    assert struct.unpack_from("<I", idx, bsz + 1)[0] == T.mask_crc(T.crc32c(idx[:bsz + 1]))
    # the first key of a bundle is the empty string (header)
    assert T._read_block(idx, boff, bsz)[0][0] == b""
    # flipped data byte -> crc mismatch when verification is on
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(IOError):
        T.read_tf_checkpoint(prefix, verify_crc=True)
    with pytest.raises(ValueError):
        open(prefix + ".index", "wb").write(idx[:-1] + b"\x00")
        T.list_variables(prefix)
