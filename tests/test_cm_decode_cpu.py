"""Kaldi compressed-matrix ('CM ') decode: the oracle restatement and the product's raw reader against golden vectors
that the REFERENCE's own reader produced (tests/golden/make_golden_cm.py ran dataset/kaldi_io.py:767-868 on
tests/golden/cm_golden.ark).  Bit-exact: this is byte/float work with a fixed operation order."""
import io
import os

import numpy as np

from oracle import kaldi_cm_oracle as CM


def _entries(golden_dir):
    gd = np.load(os.path.join(golden_dir, "cm_golden.npz"))
    ark = open(os.path.join(golden_dir, "cm_golden.ark"), "rb").read()
    offs = dict(zip([str(k) for k in gd["__offsets_keys"]], [int(o) for o in gd["__offsets"]]))
    return gd, ark, offs


def test_oracle_reproduces_reference_reader_bit_exactly(golden_dir):
    gd, ark, offs = _entries(golden_dir)
    n_full = n_sub = 0
    for name in gd.files:
        if name.startswith("__"):
            continue
        parts = name.split("/")
        fd = io.BytesIO(ark)
        fd.seek(offs[parts[1]])
        assert fd.read(3) == b"CM "
        if parts[0] == "full":
            got = CM.read_compressed_mat(fd)
            n_full += 1
        else:
            got = CM.read_compressed_mat(fd, int(parts[2]), int(parts[3]))
            n_sub += 1
        want = gd[name]
        assert got.dtype == np.float32 and got.shape == want.shape, (name, got.shape, want.shape)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, np.abs(got - want).max())
    assert n_full == 7 and n_sub >= 9


def test_legacy_float64_percentiles_within_rounding(golden_dir):
    """NumPy 1.x evaluated the uint16 -> float map in float64 (value-based casting) and rounded once; the float32
    evaluation (NumPy 2, Kaldi's C++) rounds three times: a few ulp of the header magnitude apart at most."""
    gd, ark, offs = _entries(golden_dir)
    fd = io.BytesIO(ark)
    fd.seek(offs["spk1-utt1"] + 3)
    gmin, grange, rows, hdr, data = CM.read_compressed_raw(fd)
    a = CM.percentiles_to_float(hdr, gmin, grange)
    b = CM.percentiles_to_float(hdr, gmin, grange, percentile_f64=True)
    ulp = float(np.spacing(np.float32(max(abs(float(gmin)), abs(float(grange))))))
    assert np.all(np.abs(a.astype(np.float64) - b.astype(np.float64)) <= 4 * ulp)


def test_product_raw_reader_matches_reference_bytes(golden_dir):
    """dataset.kaldi_io.read_compressed_raw / read_cm_ark return exactly the header fields and the byte crop that the
    reference reads (kaldi_io.py:800-806, 850-866) -- no decoding on the host."""
    from tf_kaldi_speaker_b200.dataset import kaldi_io as K
    gd, ark, offs = _entries(golden_dir)
    for key, (start, length) in (("spk1-utt1", (57, 200)), ("allbytes", (188, 68)), ("scalar", (0, 1))):
        fd = io.BytesIO(ark)
        fd.seek(offs[key])
        raw = K.read_compressed_raw(fd, start, length)
        fd2 = io.BytesIO(ark)
        fd2.seek(offs[key] + 3)
        gmin, grange, rows, hdr, data = CM.read_compressed_raw(fd2, start, length)
        assert raw.rows == rows and raw.cols == hdr.shape[0] and raw.start == start
        assert np.float32(raw.globmin) == gmin and np.float32(raw.globrange) == grange
        assert np.array_equal(raw.headers, hdr) and np.array_equal(raw.data, data)
        assert raw.data.dtype == np.uint8 and raw.data.shape == (hdr.shape[0], length)
    keys = [k for k, _ in K.read_cm_ark(io.BytesIO(ark))]
    assert keys == ["spk1-utt1", "spk2-utt7", "oneframe", "scalar", "allbytes", "tinyrange", "hugerange"]


def test_segment_batch_packing(golden_dir):
    from tf_kaldi_speaker_b200.dataset import kaldi_io as K
    from tf_kaldi_speaker_b200.dataset.feeder import CompressedSegmentBatch
    gd, ark, offs = _entries(golden_dir)
    batch = CompressedSegmentBatch(3, 100, 30, pin=False)
    for i, start in enumerate((0, 57, 200)):
        fd = io.BytesIO(ark)
        fd.seek(offs["spk1-utt1"])
        batch.set(i, K.read_compressed_raw(fd, start, 100))
    assert batch.data.shape == (3, 30, 100) and batch.headers.shape == (3, 30, 4) and batch.glob.shape == (3, 2)
    assert batch.h2d_bytes == 3 * 30 * 100 + 3 * 30 * 8 + 3 * 8
    want = gd["full/spk1-utt1"]
    for i, start in enumerate((0, 57, 200)):
        got = CM.decode(batch.glob[i, 0], batch.glob[i, 1], batch.headers[i], batch.data[i])
        assert np.array_equal(got, want[start:start + 100])
    try:
        fd = io.BytesIO(ark)
        fd.seek(offs["spk2-utt7"])
        batch.set(0, K.read_compressed_raw(fd, 0, 100))          # 23-dim features into a 30-dim batch
        assert False
    except ValueError:
        pass


def test_scp_offset_reader_reads_segments(golden_dir, tmp_path):
    """CompressedFeatureReader.read_segment = the reference's FeatureReader.read_segment (kaldi_io.py:112-149) minus the
    dequantisation: 'utt path:offset' entries, file descriptors kept open, row ranges cut with the same seek pattern."""
    from tf_kaldi_speaker_b200.dataset import kaldi_io as K
    gd, ark, offs = _entries(golden_dir)
    path = tmp_path / "feats.ark"
    path.write_bytes(ark)
    rd = K.CompressedFeatureReader()
    try:
        for key, (start, length) in (("spk1-utt1", (100, 200)), ("spk2-utt7", (3, 100)), ("oneframe", (0, 1))):
            entry = "%s %s:%d" % (key, path, offs[key] - 2)          # scp offsets point at the '\0B' marker
            raw = rd.read_segment(entry, length=length, start=start)
            got = CM.decode(raw.globmin, raw.globrange, raw.headers, raw.data)
            assert np.array_equal(got, gd["sub/%s/%d/%d" % (key, start, length)])
            whole = rd.read_segment(entry)
            assert whole.data.shape[1] == whole.rows and np.array_equal(
                CM.decode(whole.globmin, whole.globrange, whole.headers, whole.data), gd["full/" + key])
        assert len(rd.fd) == 1                                         # one descriptor for the archive, reused
    finally:
        rd.close()


def test_column_sharded_varspec_slices_full_checkpoints():
    """Class-sharded head variables: a full-shape checkpoint array is cut to the shard's columns on load, a shard-shaped
    one is taken as is, and the padded internal layout round-trips (runtime.VarSpec)."""
    from tf_kaldi_speaker_b200.runtime import VarSpec
    full = np.arange(6 * 21, dtype=np.float32).reshape(6, 21)
    spec = VarSpec("softmax/output/kernel", (6, 5), (6, 8), full_shape=(6, 21), col_range=(16, 21))
    local = full[:, 16:21]
    inner = spec.to_internal(local)
    assert inner.shape == (6, 8) and np.array_equal(inner[:, :5], local) and not inner[:, 5:].any()
    assert np.array_equal(spec.to_tf(inner), local)
    assert spec.full_shape == (6, 21) and spec.col_range == (16, 21)
