"""TF tensor-bundle checkpoint reader (dc_tts_b200/checkpoint.py).
No TensorFlow and no checkpoint exist offline: the reader is pinned by the format's published constants
(RFC 3720 CRC-32C vectors, table magic) and by round trips through the writer."""
import os
import struct

import numpy as np
import pytest

from dc_tts_b200 import checkpoint as ck
from dc_tts_b200.arch import param_shapes
from dc_tts_b200.params import init_params


def _py_crc(data):
    c = 0xffffffff
    for b in data:
        c = ck._TABLE[(c ^ b) & 0xff] ^ (c >> 8)
    return c ^ 0xffffffff


def test_crc32c_known_answers():
    # RFC 3720 B.4 and the usual check value
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert ck.crc32c(b"6789", ck.crc32c(b"12345")) == 0xE3069283        # continuation


def test_native_crc32c_matches_python():
    from dc_tts_b200._lib import load
    import ctypes as C
    fn = load().dctts_crc32c
    rng = np.random.default_rng(0)
    buf = rng.integers(0, 256, 70001, dtype=np.uint8)
    for start, n in [(0, 0), (0, 1), (1, 7), (3, 8), (5, 4097), (0, 70001), (7, 65536)]:
        view = buf[start:start + n]
        got = fn(C.c_uint32(0), C.c_void_p(view.ctypes.data), C.c_int64(n))
        assert got == _py_crc(view.tobytes()), (start, n)
    assert fn(C.c_uint32(0), C.c_char_p(b"123456789"), 9) == 0xE3069283
    assert ck.crc32c(buf.tobytes()) == _py_crc(buf.tobytes())           # the > 4096-byte route of the module


def test_crc_mask_roundtrip():
    for c in (0, 1, 0xE3069283, 0xffffffff, 0x12345678):
        m = ck.mask_crc(c)
        assert 0 <= m <= 0xffffffff and ck.unmask_crc(m) == c
    assert ck.mask_crc(0) == 0xa282ead8
    assert ck.mask_crc(ck.mask_crc(0xE3069283)) != 0xE3069283


def test_varint_and_snappy():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 + 5):
        b = ck._put_varint(v)
        assert ck._get_varint(b, 0) == (v, len(b))
    # literal "abc" then an overlapping copy (offset 3, length 9)
    assert ck._snappy_decompress(bytes([12, 0x08]) + b"abc" + bytes([0x15, 3])) == b"abcabcabcabc"
    # 2-byte-offset copy
    assert ck._snappy_decompress(bytes([8, 0x0c]) + b"wxyz" + bytes([(3 << 2) | 2, 4, 0])) == b"wxyzwxyz"
    with pytest.raises(ValueError):
        ck._snappy_decompress(bytes([5, 0x08]) + b"abc")


def test_table_roundtrip_multi_block(tmp_path):
    rng = np.random.default_rng(1)
    items = [(b"", b"header")]
    for i in range(700):
        items.append((("Text2Mel/TextEnc/HC_%03d/conv1d/kernel%s" % (i // 2, "" if i % 2 == 0 else "/Adam")).encode(),
                      rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8).tobytes()))
    items = sorted(set(items), key=lambda kv: kv[0])
    path = str(tmp_path / "t.index")
    ck.write_table(path, items, block_size=512, restart_interval=4)
    assert ck.read_table(path) == items
    raw = open(path, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) < sum(len(k) + len(v) + 8 for k, v in items)
    # a flipped byte in a data block is caught by the block checksum; a bad magic by the footer check
    bad = bytearray(raw); bad[10] ^= 0x40
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        ck.read_table(path)
    bad = bytearray(raw); bad[-1] ^= 1
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        ck.read_table(path)
    with pytest.raises(ValueError, match="increasing"):
        ck.write_table(path, [(b"b", b""), (b"a", b"")])


def _subset(params, prefixes):
    return {k: v for k, v in params.items() if any(k.startswith(p) for p in prefixes)}


def test_bundle_roundtrip_and_errors(tmp_path):
    P = init_params(3, "perturbed")
    sub = _subset(P, ["Text2Mel/TextEnc/embed_1", "Text2Mel/TextEnc/C_2", "Text2Mel/AudioDec/HC_3", "SSRN/D_4"])
    extra = {"gs": np.array(123456, np.int32), "Text2Mel/TextEnc/C_2/conv1d/kernel/Adam": np.zeros((1, 128, 512), np.float32),
             "beta1_power": np.array(0.5, np.float32), "stats/f64": np.arange(6, dtype=np.float64).reshape(2, 3),
             "stats/i64": np.array([-1, 2 ** 40], np.int64), "stats/flag": np.array([True, False]),
             "stats/empty": np.zeros((0, 4), np.float32)}
    d = tmp_path / "LJ01-1"
    prefix = ck.save_checkpoint(str(d / "model_gs_123k"), {**sub, **extra}, block_size=300)
    assert ck.latest_checkpoint(str(d)) == prefix
    assert ck.latest_checkpoint(str(tmp_path)) is None
    listed = {n: (s, t) for n, s, t in ck.list_variables(prefix)}
    assert listed["gs"] == ((), np.dtype(np.int32)) and listed["stats/f64"] == ((2, 3), np.dtype(np.float64))
    got = ck.load_checkpoint(prefix)
    assert set(got) == set(sub) | set(extra)
    for k, v in {**sub, **extra}.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(got[k], v)
    only = ck.load_checkpoint(prefix, names=["gs", "SSRN/D_4/conv1d_transpose/kernel"] if "SSRN/D_4/conv1d_transpose/kernel" in sub else ["gs"])
    assert "gs" in only and len(only) <= 2
    with pytest.raises(KeyError):
        ck.load_checkpoint(prefix, names=["no/such/variable"])
    # corrupt one tensor byte: the per-tensor CRC catches it (and only when verification is on)
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read()); raw[len(raw) // 2] ^= 0x01
    open(data, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        ck.load_checkpoint(prefix)
    ck.load_checkpoint(prefix, verify=False)


def test_saver_selects_scopes(tmp_path):
    P = init_params(0)
    shapes = param_shapes()
    t2m = {k: v for k, v in P.items() if k.startswith("Text2Mel/TextEnc/C_")}
    ssrn = {k: v for k, v in P.items() if k.startswith("SSRN/C_1")}
    prefix = ck.save_checkpoint(str(tmp_path / "both" / "model"), {**t2m, **ssrn, "gs": np.array(7, np.int32),
                                                                    "Text2Mel/TextEnc/C_1/conv1d/kernel/Adam_1": np.ones((1, 128, 512), np.float32)})

    class Sink:
        def __init__(self): self.got = {}
        def stage_params(self, d): self.got.update(d)

    s = Sink()
    ck.Saver(var_list=["Text2Mel"]).restore(s, prefix)
    assert set(s.got) == set(t2m) and all(k in shapes for k in s.got)       # no Adam slots, no gs, no SSRN
    s2 = Sink()
    ck.Saver(var_list=["SSRN", "gs"]).restore(s2, prefix)
    assert set(s2.got) == set(ssrn)
    with pytest.raises(ValueError):
        ck.Saver().restore(s, None)


@pytest.mark.gpu
def test_engine_restore_from_checkpoints(tmp_path):
    """synthesize.py:31-41 end to end: two checkpoint directories -> Engine.restore -> same SSRN output as the
    engine loaded from the in-memory dict."""
    import torch
    from dc_tts_b200.engine import Engine
    P = init_params(5, "perturbed")
    rng = np.random.default_rng(0)
    slots = {k + "/Adam": rng.standard_normal(v.shape).astype(np.float32) for k, v in list(P.items())[:5]}
    ck.save_checkpoint(str(tmp_path / "LJ01-1" / "model_gs_5k"), {**{k: v for k, v in P.items() if k.startswith("Text2Mel/")}, **slots,
                                                                  "gs": np.array(5000, np.int32)})
    ck.save_checkpoint(str(tmp_path / "LJ01-2" / "model_gs_9k"), {**{k: v for k, v in P.items() if k.startswith("SSRN/")},
                                                                  "gs": np.array(9000, np.int32)})
    a, b = Engine(0), Engine(0)
    n = a.restore(str(tmp_path / "LJ01-1"), str(tmp_path / "LJ01-2"))
    assert n == b.load_params(P)
    Y = torch.from_numpy(rng.uniform(0, 1, (2, 16, 80)).astype(np.float32)).cuda()
    za, zb = a.ssrn(Y)[1], b.ssrn(Y)[1]
    assert torch.equal(za, zb)
    L = np.zeros((1, 180), np.int32); L[0, :5] = [3, 4, 5, 6, 1]
    assert torch.equal(a.text2mel_generate(L, steps=3)[0], b.text2mel_generate(L, steps=3)[0])
