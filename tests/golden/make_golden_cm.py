#!/usr/bin/env python
"""Golden vectors of the Kaldi compressed-matrix ('CM ') decode, produced by the REFERENCE's own reader.

Runs in the build container only (needs /root/reference): imports dataset/kaldi_io.py (NumPy + six), reads
tests/golden/cm_golden.ark -- written here by a small format-1 compressor that follows Kaldi's
compressed-matrix.cc (test input generator, not product code) -- with the reference's
``_read_compressed_mat`` (kaldi_io.py:767-811) and ``_read_compressed_submat`` (kaldi_io.py:814-868), and commits
the float32 results in tests/golden/cm_golden.npz.  The oracle restatement (oracle/kaldi_cm_oracle.py) and the CUDA
kernel (xv_cm_decode) must reproduce them bit-exactly.

Usage: python tests/golden/make_golden_cm.py   (seeded, deterministic)
"""
import io
import os
import struct
import sys

import numpy as np

REF = os.environ.get("XV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def compress(mat):
    """float32 [rows, cols] -> (global header bytes, uint16 [cols, 4], uint8 [cols, rows]) in Kaldi format 1."""
    mat = np.asarray(mat, dtype=np.float32)
    rows, cols = mat.shape
    gmin, gmax = float(mat.min()), float(mat.max())
    grange = max(gmax - gmin, 1e-5)
    hdr = np.zeros((cols, 4), dtype=np.uint16)
    data = np.zeros((cols, rows), dtype=np.uint8)
    for c in range(cols):
        col = np.sort(mat[:, c])
        q = [col[0], col[rows // 4], col[(3 * rows) // 4], col[-1]]
        u = [int(min(65535, max(0, (v - gmin) / grange * 65535 + 0.499))) for v in q]
        u[1] = min(max(u[1], u[0] + 1), 65533)
        u[2] = min(max(u[2], u[1] + 1), 65534)
        u[3] = max(u[3], u[2] + 1)
        hdr[c] = u
        p = [gmin + grange * 1.52590218966964e-05 * v for v in u]
        x = mat[:, c].astype(np.float64)
        lo = np.clip((x - p[0]) / (p[1] - p[0]) * 64 + 0.5, 0, 64)
        mid = np.clip(64 + (x - p[1]) / (p[2] - p[1]) * 128 + 0.5, 64, 192)
        hi = np.clip(192 + (x - p[2]) / (p[3] - p[2]) * 63 + 0.5, 192, 255)
        data[c] = np.where(x < p[1], lo, np.where(x < p[2], mid, hi)).astype(np.uint8)
    return struct.pack("<ffii", gmin, grange, rows, cols), hdr, data


def entry(key, ghdr, hdr, data):
    return key.encode() + b" " + b"\0B" + b"CM " + ghdr + hdr.tobytes() + data.tobytes()


def main():
    rng = np.random.RandomState(20181017)
    items = []
    # 1. MFCC-like speech features (30-dim, 300 frames), 2. SRE-like 23-dim, 3. one frame, 4. one element
    for key, rows, cols in (("spk1-utt1", 300, 30), ("spk2-utt7", 127, 23), ("oneframe", 1, 5), ("scalar", 1, 1)):
        m = (rng.randn(rows, cols) * (1.0 + np.arange(cols)) + rng.randn(1, cols) * 3).astype(np.float32)
        items.append((key,) + compress(m))
    # 5. every byte value 0..255 in every column, adversarial headers: equal percentiles (zero-width pieces),
    #    full uint16 range, negative minimum, tiny and huge ranges
    cols = 6
    hdr = np.array([[0, 16384, 49152, 65535], [100, 100, 100, 100], [0, 0, 65535, 65535], [65535, 40000, 20000, 0],
                    [1, 2, 3, 4], [12345, 23456, 34567, 45678]], dtype=np.uint16)
    data = np.tile(np.arange(256, dtype=np.uint8), (cols, 1))
    items.append(("allbytes", struct.pack("<ffii", -37.25, 101.5, 256, cols), hdr, data))
    items.append(("tinyrange", struct.pack("<ffii", 3.0e-3, 1.0e-6, 256, cols), hdr, data))
    items.append(("hugerange", struct.pack("<ffii", -1.0e6, 3.0e6, 256, cols), hdr, data))
    ark = b"".join(entry(*it) for it in items)
    with open(os.path.join(HERE, "cm_golden.ark"), "wb") as f:
        f.write(ark)

    sys.path.insert(0, REF)
    from dataset import kaldi_io as K          # the reference reader
    out = {}
    offsets = {}
    fd = io.BytesIO(ark)
    while True:
        key = K.read_key(fd)
        if not key:
            break
        assert fd.read(2) == b"\0B"
        offsets[key] = fd.tell()
        fmt = fd.read(3).decode()
        out["full/" + key] = K._read_compressed_mat(fd, fmt).astype(np.float32)
    crops = {"spk1-utt1": [(0, 200), (57, 200), (100, 200), (299, 1)], "spk2-utt7": [(0, 127), (3, 100)],
             "allbytes": [(60, 10), (188, 68)], "oneframe": [(0, 1)]}
    for key, lst in crops.items():
        for start, length in lst:
            fd.seek(offsets[key])
            fmt = fd.read(3).decode()
            out["sub/%s/%d/%d" % (key, start, length)] = K._read_compressed_submat(fd, fmt, start, length).astype(np.float32)
    out["__offsets_keys"] = np.array(sorted(offsets))
    out["__offsets"] = np.array([offsets[k] for k in sorted(offsets)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "cm_golden.npz"), **out)
    print("wrote cm_golden.ark (%d bytes), cm_golden.npz (%d arrays), numpy %s" % (len(ark), len(out), np.__version__))


if __name__ == "__main__":
    main()
