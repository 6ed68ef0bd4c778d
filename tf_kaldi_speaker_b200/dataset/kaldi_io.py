"""Kaldi ark I/O used on either side of the extraction path (reference dataset/kaldi_io.py, a modified copy of
Vesely's kaldi_io): reading float feature matrices from an ark stream (`read_mat_ark`, kaldi_io.py:683-740) and
writing binary float vectors (`write_vec_flt`, kaldi_io.py:624-655) byte-for-byte as Kaldi's `copy-vector` expects:
``key SPACE \\0 B F V SPACE \\x04 <uint32 dim> <dim x float32>``.  Host-side code; no device work here.

Training archives are Kaldi *compressed* matrices ('CM ', format 1; the reference's segment reader accepts nothing
else, kaldi_io.py:743-749).  ``read_compressed_raw`` / ``read_cm_ark`` / ``CompressedFeatureReader`` read exactly the
bytes the reference reads (kaldi_io.py:814-868: 16-byte global header, 8-byte per-column uint16 percentiles, the uint8
row range of every column) but do NOT dequantise them: the uint8 crop goes to the GPU as it is (4x fewer PCIe bytes
than float32 features) and ``xv_cm_decode`` dequantises + transposes it there (dataset/feeder.py).
"""
import struct
import subprocess
import sys

import numpy as np


class UnknownMatrixHeader(Exception):
    pass


class CompressedRaw(object):
    """Undecoded rows [start, start+length) of a Kaldi 'CM ' matrix: headers uint16 [cols, 4] (percentiles 0/25/75/100),
    data uint8 [cols, length] (column-major, as stored), globmin / globrange float32, rows = frames of the whole matrix."""
    __slots__ = ("globmin", "globrange", "rows", "cols", "start", "headers", "data")

    def __init__(self, globmin, globrange, rows, cols, start, headers, data):
        self.globmin, self.globrange, self.rows, self.cols = globmin, globrange, rows, cols
        self.start, self.headers, self.data = start, headers, data


class UnsupportedDataType(Exception):
    pass


def open_or_fd(file, mode="rb"):
    """Open a file, an ``ark:``-prefixed path, ``-`` (stdin/stdout) or a ``cmd |`` input pipe; pass through open fds."""
    if not isinstance(file, str):
        return file
    if file.startswith("ark:") or file.startswith("scp:"):
        file = file.split(":", 1)[1]
    file = file.strip()
    if file.endswith("|"):
        return subprocess.Popen(file[:-1], shell=True, stdout=subprocess.PIPE).stdout
    if file.startswith("|"):
        return subprocess.Popen(file[1:], shell=True, stdin=subprocess.PIPE).stdin
    if file == "-":
        return sys.stdin.buffer if "r" in mode else sys.stdout.buffer
    return open(file, mode)


def read_key(fd):
    """Read the utterance key (up to the first space); '' at end of stream (kaldi_io.py read_key)."""
    key = b""
    while True:
        ch = fd.read(1)
        if ch == b"":
            break
        if ch == b" ":
            break
        key += ch
    key = key.decode("latin1").strip()
    return key if key != "" else None


def _read_mat_binary(fd):
    header = fd.read(3).decode()
    if header.startswith("CM"):
        raise UnknownMatrixHeader("compressed matrices ('%s') are not produced by the extraction feature pipe "
                                  "(apply-cmvn-sliding | select-voiced-frames); decompress with copy-feats first" % header)
    if header == "FM ":
        dt, size = np.float32, 4
    elif header == "DM ":
        dt, size = np.float64, 8
    else:
        raise UnknownMatrixHeader("The header contained '%s'" % header)
    raw = fd.read(10)
    s1, rows, s2, cols = struct.unpack("<bibi", raw)
    assert s1 == 4 and s2 == 4
    buf = fd.read(rows * cols * size)
    return np.frombuffer(buf, dtype=dt).reshape(rows, cols)


def _read_mat_ascii(fd):
    rows = []
    while True:
        line = fd.readline().decode()
        if len(line) == 0:
            raise ValueError("unexpected end of an ascii matrix")
        if len(line.strip()) == 0:
            continue
        arr = line.strip().split()
        if arr[-1] != "]":
            rows.append(np.array(arr, dtype="float32"))
        else:
            rows.append(np.array(arr[:-1], dtype="float32"))
            return np.vstack(rows)


def read_mat(fd):
    binary = fd.read(2).decode()
    if binary == "\0B":
        return _read_mat_binary(fd)
    assert binary == " ["
    return _read_mat_ascii(fd)


def read_mat_ark(file_or_fd):
    """generator(key, mat) over an ark file / stream (kaldi_io.py:683-704)."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            yield key, read_mat(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_mat(file_or_fd, m, key=""):
    """Binary float matrix (the inverse of read_mat; used to build test arks)."""
    fd = open_or_fd(file_or_fd, mode="wb")
    try:
        if key != "":
            fd.write((key + " ").encode("latin1"))
        fd.write(b"\0B")
        if m.dtype == np.float32:
            fd.write(b"FM ")
        elif m.dtype == np.float64:
            fd.write(b"DM ")
        else:
            raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % m.dtype)
        fd.write(b"\x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]))
        fd.write(np.ascontiguousarray(m).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


def write_vec_flt(file_or_fd, v, key=""):
    """Binary float vector, byte-identical to kaldi_io.py:624-655."""
    fd = open_or_fd(file_or_fd, mode="wb")
    try:
        if key != "":
            fd.write((key + " ").encode("latin1"))
        fd.write(b"\0B")
        if v.dtype == np.float32:
            fd.write(b"FV ")
        elif v.dtype == np.float64:
            fd.write(b"DV ")
        else:
            raise UnsupportedDataType("'%s', please use 'float32' or 'float64'" % v.dtype)
        fd.write(b"\x04")
        fd.write(struct.pack("<I", v.shape[0]))
        fd.write(np.ascontiguousarray(v).tobytes())
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_vec_flt_ark(file_or_fd):
    """generator(key, vec) over a binary float-vector ark (test helper / downstream check)."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            assert fd.read(2) == b"\0B"
            header = fd.read(3).decode()
            dt = np.float32 if header == "FV " else np.float64
            assert fd.read(1) == b"\x04"
            (dim,) = struct.unpack("<I", fd.read(4))
            yield key, np.frombuffer(fd.read(dim * np.dtype(dt).itemsize), dtype=dt)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


def read_compressed_raw(fd, start=None, length=None):
    """fd positioned at the matrix token (after the ``\\0B`` marker).  Reads what _read_submat_binary /
    _read_compressed_submat read (kaldi_io.py:743-749, 814-868) -- 'CM ' only, the per-column seek pattern included --
    and returns the bytes undecoded (CompressedRaw)."""
    header = fd.read(3).decode()
    if not header.startswith("CM"):
        raise ValueError("The features should be in the compressed format.")              # kaldi_io.py:749
    if header != "CM ":
        raise UnknownMatrixHeader("The formats CM2, CM3 are not supported (kaldi_io.py:819)")
    globmin, globrange, rows, cols = struct.unpack("<ffii", fd.read(16))
    if start is None:
        start, length = 0, rows
    assert rows >= (start + length), "The number of frames is not enough for length %d" % length
    headers = np.frombuffer(fd.read(cols * 8), dtype=np.uint16, count=cols * 4).reshape(cols, 4).copy()
    data = np.empty((cols, length), dtype=np.uint8)
    col_left = 0
    for i in range(cols):
        fd.seek(col_left + start, 1)
        data[i] = np.frombuffer(fd.read(length), dtype=np.uint8, count=length)
        col_left = rows - (start + length)
    fd.seek(col_left, 1)
    return CompressedRaw(np.float32(globmin), np.float32(globrange), int(rows), int(cols), int(start), headers, data)


def read_cm_ark(file_or_fd):
    """Generator of (key, CompressedRaw) over an ark of compressed matrices (seekable stream)."""
    fd = open_or_fd(file_or_fd)
    try:
        key = read_key(fd)
        while key:
            if fd.read(2) != b"\0B":
                raise UnknownMatrixHeader("compressed matrices are binary")
            yield key, read_compressed_raw(fd)
            key = read_key(fd)
    finally:
        if fd is not file_or_fd:
            fd.close()


class CompressedFeatureReader(object):
    """The reference's FeatureReader.read_segment (kaldi_io.py:112-149) without the host-side dequantisation: keeps the
    archive file descriptors open, seeks to ``path:offset`` of a feats.scp entry and returns the raw crop."""

    def __init__(self):
        self.fd = {}

    def close(self):
        for f in self.fd.values():
            f.close()
        self.fd = {}

    def read_segment(self, scp_entry, length=None, start=None):
        """scp_entry: 'utt path:offset' or 'path:offset' -> CompressedRaw of rows [start, start+length) (all rows if None)."""
        loc = scp_entry.strip().split(" ")[-1]
        filename, offset = loc.rsplit(":", 1)
        if filename not in self.fd:
            self.fd[filename] = open(filename, "rb")
        fd = self.fd[filename]
        fd.seek(int(offset))
        if fd.read(2) != b"\0B":
            raise IOError("Cannot read features from %s" % scp_entry)
        if length is None:
            return read_compressed_raw(fd)
        return read_compressed_raw(fd, 0 if start is None else start, length)
