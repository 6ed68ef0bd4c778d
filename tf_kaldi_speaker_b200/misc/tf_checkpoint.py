"""TensorFlow-free reader / writer of TF "tensor bundle" (V2) checkpoints: ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``.

The reference saves and restores its models with ``tf.train.Saver`` (model/trainer.py:142-166: ``saver.save(sess,
os.path.join(self.model, "model"), global_step=step)`` and ``saver.restore``), i.e. in this format, with the variable
names of SURVEY Appendix B as keys.  TensorFlow cannot be installed in this image, so the format is restated here from
its published description (tensorflow/core/util/tensor_bundle/tensor_bundle.h, tensorflow/core/lib/io/format.h -- the
leveldb table format):

  index file   = data blocks, meta-index block, index block, 48-byte footer
     footer    = BlockHandle(meta-index), BlockHandle(index), zero padding to 40 bytes, magic 0xdb4775248b80fb57 (LE)
     handle    = varint64 offset, varint64 size            (size excludes the 5-byte block trailer)
     block     = entries, uint32 restart offsets[n], uint32 n;  trailer = 1 byte compression type (0 none, 1 snappy)
                 + uint32 masked crc32c(block + type)
     entry     = varint32 shared key bytes, varint32 unshared key bytes, varint32 value bytes, key delta, value
     key ""    -> BundleHeaderProto {1: num_shards, 2: endianness (0 little), 3: VersionDef {1: producer}}
     key name  -> BundleEntryProto  {1: dtype, 2: TensorShapeProto {2: Dim {1: size}}, 3: shard_id, 4: offset, 5: size,
                                     6: fixed32 masked crc32c of the tensor bytes}
  data file    = the raw little-endian tensor bytes at those offsets

Status: round-trip tested here (tests/test_tf_checkpoint_cpu.py, including the crc32c known answer); NOT yet validated
against a file written by TensorFlow itself -- none is available in this image ("parity unpinned" for the file format).
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
MASK_DELTA = 0xa282ead8
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}

_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    return _CRC_TABLE


def _crc32c_bytewise(data, crc=0):
    t = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in bytes(data):
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _gf2_apply(cols, x):
    """y = M x over GF(2) for a 32 x 32 matrix given by its 32 uint32 columns; x may be an array."""
    y = np.zeros_like(x)
    for j in range(32):
        y ^= np.where((x >> np.uint32(j)) & np.uint32(1), cols[j], np.uint32(0)).astype(np.uint32)
    return y


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), as crc32c::Extend.  Large buffers (a 30 MB speaker matrix) are cut into 2^k equal chunks
    whose raw CRC registers advance in lock-step as NumPy vectors, then merged pairwise with the GF(2) "append n zero
    bytes" operator (the crc32_combine construction): seconds of pure-Python byte loop become milliseconds."""
    data = bytes(data)
    n = len(data)
    if n < (1 << 16):
        return _crc32c_bytewise(data, crc)
    table = np.array(_crc_table(), dtype=np.uint32)
    levels = 12
    N = 1 << levels
    L = -(-n // N)
    buf = np.zeros(N * L, dtype=np.uint8)
    buf[N * L - n:] = np.frombuffer(data, dtype=np.uint8)       # leading zeros leave a zero-initialised register at zero
    chunks = buf.reshape(N, L)
    reg = np.zeros(N, dtype=np.uint32)
    for i in range(L):
        reg = table[(reg ^ chunks[:, i]) & np.uint32(0xFF)] ^ (reg >> np.uint32(8))
    # operator "append one zero byte" as matrix columns, raised to the power L by repeated squaring
    one = np.uint32(1)
    cols = np.array([table[(one << np.uint32(j)) & np.uint32(0xFF)] ^ ((one << np.uint32(j)) >> np.uint32(8)) for j in range(32)],
                    dtype=np.uint32)

    def power(c, e):
        result = np.array([one << np.uint32(j) for j in range(32)], dtype=np.uint32)      # identity
        base = c.copy()
        while e:
            if e & 1:
                result = _gf2_apply(base, result)
            base = _gf2_apply(base, base)
            e >>= 1
        return result
    shift = power(cols, L)
    for _ in range(levels):
        reg = _gf2_apply(shift, reg[0::2]) ^ reg[1::2]
        shift = _gf2_apply(shift, shift)
    init = np.array([(crc ^ 0xFFFFFFFF) & 0xFFFFFFFF], dtype=np.uint32)
    head = _gf2_apply(power(cols, n), init)                    # the initial register pushed through n bytes
    return int(reg[0] ^ head[0]) ^ 0xFFFFFFFF


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / minimal protobuf -----------------------------------------------------------------------------------
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_proto(buf):
    """-> {field number: [values]}; varint fields as int, length-delimited as bytes, fixed32/64 as int."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype_code, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    msg = _field(1, 0, _put_varint(dtype_code)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", mask_crc(crc)))
    return msg


# ---- table blocks ------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size):
    kind = buf[offset + size]
    if kind != 0:
        raise NotImplementedError("compressed index blocks (type %d) are not supported" % kind)
    block = buf[offset:offset + size]
    n_restarts = struct.unpack_from("<I", block, size - 4)[0]
    end = size - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _build_block(items, restart_interval=16):
    body, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    body += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(body)


def _with_trailer(block):
    return block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00")))


# ---- public API ----------------------------------------------------------------------------------------------------------
def _read_index(prefix):
    buf = open(prefix + ".index", "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s.index is not a TensorFlow tensor-bundle index (bad magic)" % prefix)
    pos = len(buf) - 48
    _, pos = _get_varint(buf, pos)              # meta-index handle (unused)
    _, pos = _get_varint(buf, pos)
    ioff, pos = _get_varint(buf, pos)
    isz, pos = _get_varint(buf, pos)
    out, num_shards = {}, 1
    for _, handle in _read_block(buf, ioff, isz):
        boff, p = _get_varint(handle, 0)
        bsz, p = _get_varint(handle, p)
        for key, value in _read_block(buf, boff, bsz):
            if key == b"":
                hdr = _parse_proto(value)
                num_shards = hdr.get(1, [1])[0]
                if hdr.get(2, [0])[0] != 0:
                    raise NotImplementedError("big-endian bundles are not supported")
                continue
            e = _parse_proto(value)
            if 7 in e:
                raise NotImplementedError("%s: partitioned (sliced) variables are not supported" % key.decode())
            code = e.get(1, [0])[0]
            if code not in DTYPES:       # e.g. DT_STRING bookkeeping entries: only an error if somebody asks for them
                out[key.decode()] = (None, code, e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0], None)
                continue
            shape = []
            if 2 in e:
                for d in _parse_proto(e[2][0]).get(2, []):
                    sz = _parse_proto(d).get(1, [0])[0]
                    shape.append(sz - (1 << 64) if sz >= (1 << 63) else sz)
            out[key.decode()] = (np.dtype(DTYPES[code]), tuple(shape), e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0],
                                 unmask_crc(e.get(6, [0])[0]) if 6 in e else None)
    return out, num_shards


def list_variables(prefix):
    """-> {name: (dtype, shape, shard_id, offset, size, crc32c)} of ``<prefix>.index``."""
    return _read_index(prefix)[0]


def read_tf_checkpoint(prefix, names=None, verify_crc=False):
    """-> {variable name: ndarray} (all variables, or ``names``).  ``prefix`` as in ``saver.restore(sess, prefix)``."""
    meta, num = _read_index(prefix)
    files = {}
    out = {}
    try:
        for name, (dt, shape, shard, offset, size, crc) in meta.items():
            if names is not None and name not in names:
                continue
            if dt is None:
                if names is not None:
                    raise NotImplementedError("%s: dtype code %d" % (name, shape))
                continue
            if shard not in files:
                files[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, num), "rb")
            f = files[shard]
            f.seek(offset)
            raw = f.read(size)
            if len(raw) != size:
                raise IOError("%s: truncated data file" % name)
            if verify_crc and crc is not None and crc32c(raw) != crc:
                raise IOError("%s: crc32c mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(shape).astype(dt, copy=True)
    finally:
        for f in files.values():
            f.close()
    return out


def write_tf_checkpoint(prefix, variables, checksums=True):
    """Write ``{name: ndarray}`` as a single-shard bundle.  ``checksums=False`` stores a zero crc32c per tensor (the
    pure-Python crc runs at a few MB/s); TensorFlow verifies these on restore, so keep it on for files TF must read."""
    names = sorted(variables, key=lambda n: n.encode())
    data_path = "%s.data-00000-of-00001" % prefix
    items, offset = [], 0
    with open(data_path, "wb") as f:
        for n in names:
            a = np.asarray(variables[n], order="C")          # (ascontiguousarray would turn scalars into [1])
            if a.dtype not in DTYPE_CODES:
                raise NotImplementedError("%s: dtype %s" % (n, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((n.encode(), _entry_proto(DTYPE_CODES[a.dtype], a.shape, offset, len(raw),
                                                   crc32c(raw) if checksums else 0)))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))   # num_shards, version
    items = [(b"", header)] + items
    out = bytearray()
    index_entries = []
    BLOCK = 64          # entries per data block
    for i in range(0, len(items), BLOCK):
        chunk = items[i:i + BLOCK]
        block = _build_block(chunk)
        index_entries.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(block))))
        out += _with_trailer(block)
    meta = _build_block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out += _with_trailer(meta)
    index = _build_block(index_entries, restart_interval=1)
    index_handle = _put_varint(len(out)) + _put_varint(len(index))
    out += _with_trailer(index)
    footer = meta_handle + index_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    return prefix
