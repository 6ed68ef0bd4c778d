"""CPU restatement of the reference's Kaldi compressed-matrix ('CM ', format 1) reader.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/dataset/kaldi_io.py:
  :767-811  _read_compressed_mat      whole matrix
  :814-868  _read_compressed_submat   rows [start, start+length) of every column, seeking past the rest
  :775-777 / :823-824                 16-byte global header (min f32, range f32, rows i32, cols i32; '.format' not stored)
                                      and 8-byte per-column header of four uint16 percentiles (0, 25, 75, 100)
  :780-781                            uint16 -> float:  min + range * 1.52590218966964e-05 * value
  :784-797                            uint8  -> float:  three linear pieces, [0,64] / (64,192] / (192,255]
  :811, :868                          data is column-major; the result is transposed to [rows, cols]

Arithmetic: every operand is a NumPy float32 scalar / uint8 array, so under NumPy >= 2 (NEP 50; the NumPy of this image)
each operation rounds to float32 in the order written -- the same order Kaldi's own C++ uses
(compressed-matrix.h Uint16ToFloat / CharToFloat).  Under the NumPy 1.x the reference was developed with, the
percentile expression is evaluated in float64 and rounded once; the two differ by at most 1 ulp of the percentile.
``percentile_f64=True`` selects that legacy variant.

Parity pinning: PINNED.  This is the one part of the path whose reference code runs in this image (NumPy + six only):
tests/golden/make_golden_cm.py imports /root/reference/dataset/kaldi_io.py, reads tests/golden/cm_golden.ark with
its _read_compressed_mat / _read_compressed_submat and commits the outputs (tests/golden/cm_golden.npz); this
restatement and the CUDA kernel xv_cm_decode must both reproduce them BIT-EXACTLY (tests/test_cm_decode_*.py).
"""
import struct

import numpy as np

U16_STEP = 1.52590218966964e-05


def read_header(fd):
    """After the '\\0B' binary marker and the 'CM ' token: (min, range, rows, cols) -- kaldi_io.py:800, 850."""
    globmin, globrange, rows, cols = struct.unpack("<ffii", fd.read(16))
    return np.float32(globmin), np.float32(globrange), int(rows), int(cols)


def percentiles_to_float(col_headers_u16, globmin, globrange, percentile_f64=False):
    """uint16 [cols, 4] -> float32 [cols, 4] (kaldi_io.py:780-781)."""
    h = np.asarray(col_headers_u16, dtype=np.uint16)
    if percentile_f64:
        return (np.float64(globmin) + np.float64(globrange) * U16_STEP * h.astype(np.float64)).astype(np.float32)
    step = np.float32(globrange) * np.float32(U16_STEP)              # range * 1.5259e-05   (float32)
    return (np.float32(globmin) + step * h.astype(np.float32)).astype(np.float32)


def bytes_to_float(data_u8, pf):
    """data uint8 [cols, n], pf float32 [cols, 4] -> float32 [cols, n] (kaldi_io.py:784-797)."""
    data = np.asarray(data_u8, dtype=np.uint8)
    v = data.astype(np.float32)
    p0, p25, p75, p100 = (pf[:, i:i + 1].astype(np.float32) for i in range(4))
    lo = p0 + (p25 - p0) / np.float32(64.) * v
    mid = p25 + (p75 - p25) / np.float32(128.) * (v - np.float32(64.))
    hi = p75 + (p100 - p75) / np.float32(63.) * (v - np.float32(192.))
    return np.where(data <= 64, lo, np.where(data <= 192, mid, hi)).astype(np.float32)


def read_compressed_raw(fd, start=None, length=None):
    """The bytes the reference reads, undecoded: (min, range, rows, headers u16 [cols,4], data u8 [cols, length]).
    start/length select a row range as _read_compressed_submat does (kaldi_io.py:852-866)."""
    globmin, globrange, rows, cols = read_header(fd)
    if start is None:
        start, length = 0, rows
    assert rows >= start + length, "The number of frames is not enough for length %d" % length
    headers = np.frombuffer(fd.read(cols * 8), dtype=np.uint16, count=cols * 4).reshape(cols, 4)
    body = np.frombuffer(fd.read(cols * rows), dtype=np.uint8, count=cols * rows).reshape(cols, rows)
    return globmin, globrange, rows, headers.copy(), body[:, start:start + length].copy()


def decode(globmin, globrange, headers, data, percentile_f64=False):
    """-> float32 [length, cols] (the reference's ``mat.T``)."""
    pf = percentiles_to_float(headers, globmin, globrange, percentile_f64)
    return np.ascontiguousarray(bytes_to_float(data, pf).T)


def read_compressed_mat(fd, start=None, length=None, percentile_f64=False):
    globmin, globrange, _, headers, data = read_compressed_raw(fd, start, length)
    return decode(globmin, globrange, headers, data, percentile_f64)
