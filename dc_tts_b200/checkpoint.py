"""TensorFlow checkpoint (tensor bundle, "V2" format) reader and writer -- no TensorFlow needed.

Replaces what `tf.train.latest_checkpoint` + `tf.train.Saver(var_list).restore(sess, path)` do for
the synthesis path (/root/reference/synthesize.py:31-41; written by `sv.saver.save` at
/root/reference/train.py:152): name -> ndarray for every variable of a checkpoint, which
`Engine.load_params` then packs for the kernels.  SURVEY.md 8(f) rank 2.

The format lives in TensorFlow (absent, version un-pinned ">= 1.3", reference README.md:7), so this is a
restatement of its published on-disk layout (tensorflow/core/util/tensor_bundle, lib/io/table = the LevelDB
table format, lib/hash/crc32c):

  <prefix>.index                  an immutable sorted string table:
      data blocks   [entries][restart offsets: uint32 x n][n: uint32]  + 1 byte compression type (0 none,
                    1 snappy) + 4 bytes masked CRC-32C of block+type.  Entry = varint32 shared-key-bytes,
                    varint32 unshared-key-bytes, varint32 value-bytes, key suffix, value.
      metaindex block, index block (separator key -> BlockHandle{varint64 offset, varint64 size}),
      48-byte footer: metaindex handle, index handle, zero padding, magic 0xdb4775248b80fb57 (LE).
      key ""   -> BundleHeaderProto  {1: num_shards, 2: endianness (0 little), 3: VersionDef{1: producer}}
      key name -> BundleEntryProto   {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id,
                                      4: offset, 5: size, 6: fixed32 masked crc32c, 7: slices (partitioned)}
  <prefix>.data-SSSSS-of-NNNNN    the tensors' raw little-endian bytes at [offset, offset + size).
  checkpoint                      text proto: model_checkpoint_path: "<prefix>" (relative to its directory).

PARITY UNPINNED: no TensorFlow and no checkpoint file exist offline.  The reader is held by (i) the
published constants above (table magic, CRC-32C test vectors of RFC 3720, the CRC mask), (ii) a round trip
through the writer below, which emits prefix-compressed multi-block tables exactly as the layout
prescribes (tests/test_checkpoint.py).  Both directions verify every checksum.
"""
import ctypes as C
import os
import re
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------ checksums
def _crc_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_TABLE = _crc_table()
_native = None


def _native_crc():
    """dctts_crc32c from the C-ABI library (slicing-by-8); the pure-Python loop is kept for small inputs and
    for hosts where the library has not been built."""
    global _native
    if _native is None:
        try:
            from ._lib import load
            _native = load().dctts_crc32c
        except Exception:
            _native = False
    return _native


def crc32c(data, crc=0):
    data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data
    fn = _native_crc() if len(data) > 4096 else None
    if fn:
        buf = np.frombuffer(data, np.uint8)
        return int(fn(C.c_uint32(crc), C.c_void_p(buf.ctypes.data), C.c_int64(buf.size)))
    c = crc ^ 0xffffffff
    for b in bytes(data):
        c = _TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def mask_crc(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xffffffff


def unmask_crc(m):
    rot = (m - _MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


# ------------------------------------------------------------------------------------ varints / protobuf
def _get_varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """Yields (field number, wire type, value) of one protobuf message; value is int or bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for f, _, v in _proto_fields(buf):
        if f == 2:                                   # Dim
            size = 0
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for f, _, v in _proto_fields(buf):
        if f == 1: e["dtype"] = v
        elif f == 2: e["shape"] = _parse_shape(v)
        elif f == 3: e["shard_id"] = v
        elif f == 4: e["offset"] = v
        elif f == 5: e["size"] = v
        elif f == 6: e["crc32c"] = v
        elif f == 7: e["slices"] += 1
    return e


def _parse_header(buf):
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    for f, _, v in _proto_fields(buf):
        if f == 1: h["num_shards"] = v
        elif f == 2: h["endianness"] = v
        elif f == 3:
            for f2, _, v2 in _proto_fields(v):
                if f2 == 1:
                    h["producer"] = v2
    return h


# ------------------------------------------------------------------------------------ snappy (index blocks may use it)
def _snappy_decompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:                                # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little"); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little"); pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                          # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------ table reader
def _read_block(data, offset, size):
    raw = data[offset:offset + size]
    if len(raw) != size or offset + size + 5 > len(data):
        raise ValueError("table block out of range")
    ctype = data[offset + size]
    stored = struct.unpack_from("<I", data, offset + size + 1)[0]
    if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
        raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return bytes(raw)
    if ctype == 1:
        return _snappy_decompress(bytes(raw))
    raise ValueError("unknown block compression type %d" % ctype)


def _block_entries(block):
    """(key, value) pairs of one table block, in order."""
    if len(block) < 4:
        raise ValueError("table block too short")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise ValueError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key):
            raise ValueError("corrupt key prefix")
        key = key[:shared] + block[pos:pos + unshared]; pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path):
    """All (key, value) pairs of a LevelDB-format table file (the `.index` of a bundle)."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError("%s: too short for a table footer" % path)
    footer = data[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("%s: not a tensor-bundle index (bad magic)" % path)
    pos = 0
    _mo, pos = _get_varint(footer, pos); _ms, pos = _get_varint(footer, pos)
    io, pos = _get_varint(footer, pos); isz, pos = _get_varint(footer, pos)
    out = []
    for _sep, handle in _block_entries(_read_block(data, io, isz)):
        off, p2 = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p2)
        out.extend(_block_entries(_read_block(data, off, size)))
    return out


# ------------------------------------------------------------------------------------ public reader
def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the prefix named by `<dir>/checkpoint`, or None."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.isfile(state):
        return None
    with open(state) as f:
        m = re.search(r'^\s*model_checkpoint_path:\s*"((?:[^"\\]|\\.)*)"', f.read(), re.M)
    if not m:
        return None
    p = m.group(1).encode().decode("unicode_escape")
    if not os.path.isabs(p):
        p = os.path.join(checkpoint_dir, p)
    return p if os.path.isfile(p + ".index") else None


def list_variables(prefix):
    """[(name, shape, dtype)] like tf.train.list_variables."""
    out = []
    for k, v in read_table(prefix + ".index"):
        if k:
            e = _parse_entry(v)
            out.append((k.decode(), e["shape"], np.dtype(_DTYPES[e["dtype"]]) if e["dtype"] in _DTYPES else None))
    return out


def load_checkpoint(prefix, names=None, verify=True):
    """name -> ndarray for the variables of checkpoint `prefix` (all of them, or `names`).
    Raises on bad checksums, unsupported dtypes, partitioned (sliced) variables or big-endian bundles."""
    entries = {}
    header = None
    for k, v in read_table(prefix + ".index"):
        if k == b"":
            header = _parse_header(v)
        else:
            entries[k.decode()] = _parse_entry(v)
    if header is None:
        raise ValueError("%s.index: no bundle header" % prefix)
    if header["endianness"] != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    want = list(entries) if names is None else list(names)
    missing = [n for n in want if n not in entries]
    if missing:
        raise KeyError("not in checkpoint %s: %s" % (prefix, ", ".join(missing[:5]) + (" ..." if len(missing) > 5 else "")))
    shards = {}
    out = {}
    for name in want:
        e = entries[name]
        if e["slices"]:
            raise ValueError("%s is a partitioned variable (tensor slices): not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise ValueError("%s: unsupported dtype id %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"]), np.uint8, "r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if raw.size != e["size"] or n * dt.itemsize != e["size"]:
            raise ValueError("%s: %d bytes on disk for shape %s %s" % (name, e["size"], e["shape"], dt))
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(np.ascontiguousarray(raw).tobytes()):
            raise ValueError("%s: data checksum mismatch" % name)
        out[name] = np.frombuffer(np.ascontiguousarray(raw).tobytes(), dt).reshape(e["shape"]).copy()
    return out


class Saver:
    """The part of tf.train.Saver the synthesis script uses: restore(sess, save_path) with an optional
    var_list of names or scope prefixes ("Text2Mel", "SSRN", "gs" -- synthesize.py:32,37-38)."""

    def __init__(self, var_list=None):
        self.var_list = var_list

    def restore(self, sess_or_engine, save_path):
        from .arch import param_shapes
        if save_path is None:
            raise ValueError("Can't load save_path when it is None.")           # TF's own message
        shapes = param_shapes()
        avail = {n for n, _, _ in list_variables(save_path)}
        scopes = self.var_list
        names = [n for n in shapes if n in avail and (scopes is None or any(n == s or n.startswith(s.rstrip("/") + "/") for s in scopes))]
        tensors = load_checkpoint(save_path, names)
        engine = getattr(sess_or_engine, "engine", sess_or_engine)
        if hasattr(engine, "stage_params"):
            engine.stage_params(tensors)
        return tensors


# ------------------------------------------------------------------------------------ writer
def _proto_varint_field(field, v):
    return _put_varint(field << 3) + _put_varint(v & ((1 << 64) - 1))


def _proto_bytes_field(field, b):
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def _encode_entry(dtype_id, shape, shard_id, offset, size, crc):
    dims = b"".join(_proto_bytes_field(2, _proto_varint_field(1, int(d))) for d in shape)
    out = _proto_varint_field(1, dtype_id) + _proto_bytes_field(2, dims)
    if shard_id: out += _proto_varint_field(3, shard_id)
    if offset: out += _proto_varint_field(4, offset)
    out += _proto_varint_field(5, size)
    out += _put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc))
    return out


class _BlockBuilder:
    def __init__(self, restart_interval):
        self.ri = restart_interval
        self.reset()

    def reset(self):
        self.buf = bytearray(); self.restarts = [0]; self.count = 0; self.last = b""

    def add(self, key, value):
        shared = 0
        if self.count < self.ri:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf)); self.count = 0
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key; self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(path, items, block_size=4096, restart_interval=16):
    """Writes sorted (key, value) byte pairs as a LevelDB-format table (no compression, like BundleWriter)."""
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block); out.append(0)
        out.extend(struct.pack("<I", mask_crc(crc32c(bytes(block) + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index = _BlockBuilder(1)
    data = _BlockBuilder(restart_interval)
    last_key = None
    for key, value in items:
        if last_key is not None and key <= last_key:
            raise ValueError("table keys must be strictly increasing")
        data.add(key, value); last_key = key
        if data.size() >= block_size:
            index.add(last_key, emit(data.finish())); data.reset()
    if data.buf:
        index.add(last_key, emit(data.finish()))
    meta_handle = emit(_BlockBuilder(1).finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


def save_checkpoint(prefix, tensors, update_state=True, block_size=4096):
    """Writes {name: ndarray} as a one-shard tensor bundle `prefix`.{index,data-00000-of-00001} and, like
    tf.train.Saver.save, points `<dir>/checkpoint` at it."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    items = [(b"", _proto_varint_field(1, 1) + _proto_bytes_field(3, _proto_varint_field(1, 1)))]   # 1 shard, little endian, producer 1
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])                 # (ascontiguousarray would turn a scalar into shape (1,))
            if not a.flags.c_contiguous:
                a = a.copy(order="C")
            if a.dtype not in _DTYPE_IDS:
                raise ValueError("%s: dtype %s cannot be stored" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode(), _encode_entry(_DTYPE_IDS[a.dtype], a.shape, 0, offset, len(raw), crc32c(raw))))
            offset += len(raw)
    write_table(prefix + ".index", items, block_size=block_size)
    if update_state:
        base = os.path.basename(prefix)
        with open(os.path.join(d or ".", "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix
