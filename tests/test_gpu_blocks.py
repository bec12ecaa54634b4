"""GPU parity tests, block and network level: CUDA path (through the C-ABI) vs the oracle
on identical seeded inputs.  Tolerances are absolute, in float32: 2e-4 per block (pure
re-association noise is ~1e-5), 1e-3 per network (the north-star tolerance)."""
import zlib

import numpy as np
import pytest
import torch

from dc_tts_b200 import arch
from dc_tts_b200.hyperparams import Hyperparams as hp
from dc_tts_b200.params import synthetic_text
from oracle import ref_torch as rt

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("path")]
BLOCK_TOL = 2e-4
NET_TOL = 1e-3


def _rand(shape, seed, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(np.float32)


def _layer(net, scope):
    for l in arch.NETWORKS[net]():
        if l.scope == scope:
            return l
    raise KeyError(scope)


def test_embed(engine, params):
    ids = np.random.default_rng(0).integers(0, 32, (3, hp.max_N)).astype(np.int32)
    ids[:, -20:] = 0
    out = engine.embed("Text2Mel/TextEnc/embed_1", ids).cpu().numpy()
    ref = rt.embed(params, torch.from_numpy(ids), "Text2Mel/TextEnc/embed_1").numpy()
    assert np.array_equal(out, ref)                 # a gather: bit exact
    assert (out[ids == 0] == 0).all()


@pytest.mark.parametrize("C,scope", [(80, "Text2Mel/AudioDec/C_11/normalize"), (256, "Text2Mel/AudioEnc/HC_4/H1"),
                                     (512, "SSRN/D_4/normalize"), (1024, "SSRN/HC_11/H2"), (1025, "SSRN/C_16/normalize")])
def test_normalize(engine, params, C, scope):
    x = _rand((37, C), C, -3, 3)
    x[5] = 0.0                                      # zero-variance row: eps 1e-12 path (quirk Q4)
    x[6] = 2.5
    out = engine.normalize(scope, x).cpu().numpy()
    ref = rt.normalize(torch.from_numpy(x), torch.from_numpy(params[scope + "/gamma"]),
                       torch.from_numpy(params[scope + "/beta"])).numpy()
    assert np.abs(out - ref).max() < 1e-4


CONV_CASES = [
    # net, scope, B, L   (rate / padding / act from the layer table)
    ("Text2Mel/TextEnc", "C_2", 2, 180), ("Text2Mel/TextEnc", "C_3", 1, 33),
    ("Text2Mel/AudioEnc", "C_1", 3, 210), ("Text2Mel/AudioDec", "C_1", 2, 85),
    ("Text2Mel/AudioDec", "C_11", 2, 210), ("SSRN", "C_1", 2, 64), ("SSRN", "C_10", 1, 700),
    ("SSRN", "C_13", 1, 70), ("SSRN", "C_15", 1, 300), ("SSRN", "C_16", 2, 17), ("SSRN", "C_14", 1, 1),
]


@pytest.mark.parametrize("net,scope,B,L", CONV_CASES)
def test_conv1d_block(engine, params, net, scope, B, L):
    l = _layer(net, scope)
    x = _rand((B, L, l.cin), zlib.crc32(scope.encode()) % 1000)
    full = net + "/" + scope
    out = engine.conv1d(full, x, l.cout, l.rate, l.pad == "CAUSAL", 1 if l.act == "relu" else 0).cpu().numpy()
    ref = rt.conv1d(params, torch.from_numpy(x), full, l.rate, l.pad, l.act).numpy()
    assert out.shape == ref.shape == (B, L, l.cout)
    assert np.abs(out - ref).max() < BLOCK_TOL


HC_CASES = [
    ("Text2Mel/TextEnc", "HC_4", 2, 180), ("Text2Mel/TextEnc", "HC_7", 1, 180), ("Text2Mel/TextEnc", "HC_15", 2, 50),
    ("Text2Mel/AudioEnc", "HC_4", 2, 210), ("Text2Mel/AudioEnc", "HC_7", 1, 210), ("Text2Mel/AudioEnc", "HC_13", 3, 20),
    ("Text2Mel/AudioDec", "HC_5", 1, 210), ("Text2Mel/AudioDec", "HC_5", 32, 85), ("SSRN", "HC_3", 1, 210),
    ("SSRN", "HC_9", 1, 840), ("SSRN", "HC_11", 1, 520), ("SSRN", "HC_12", 2, 9), ("SSRN", "HC_2", 1, 1),
]


@pytest.mark.parametrize("net,scope,B,L", HC_CASES)
def test_hc_block(engine, params, net, scope, B, L):
    l = _layer(net, scope)
    x = _rand((B, L, l.cin), zlib.crc32(scope.encode()) % 1000 + L)
    full = net + "/" + scope
    out = engine.hc(full, x, l.rate, l.pad == "CAUSAL").cpu().numpy()
    ref = rt.hc(params, torch.from_numpy(x), full, l.rate, l.pad).numpy()
    assert np.abs(out - ref).max() < BLOCK_TOL


def test_hc_both_paddings_any_rate(engine, params):
    """the op-level API takes rate/padding from the call, as the reference signature does"""
    x = _rand((2, 100, 256), 77)
    full = "Text2Mel/AudioEnc/HC_5"
    for rate, causal in [(1, True), (5, False), (27, False), (60, True)]:
        out = engine.hc(full, x, rate, causal).cpu().numpy()
        ref = rt.hc(params, torch.from_numpy(x), full, rate, "CAUSAL" if causal else "SAME").numpy()
        assert np.abs(out - ref).max() < BLOCK_TOL, (rate, causal)


@pytest.mark.parametrize("B,L", [(1, 210), (2, 420), (3, 7), (1, 1)])
def test_conv1d_transpose_block(engine, params, B, L):
    x = _rand((B, L, hp.c), L)
    for scope in ("SSRN/D_4", "SSRN/D_7"):
        out = engine.conv1d_transpose(scope, x).cpu().numpy()
        ref = rt.conv1d_transpose(params, torch.from_numpy(x), scope).numpy()
        assert out.shape == (B, 2 * L, hp.c)
        assert np.abs(out - ref).max() < BLOCK_TOL


@pytest.mark.parametrize("monotonic", [False, True])
def test_attention(engine, monotonic):
    B = 4
    Q, K, V = _rand((B, hp.max_T, hp.d), 1), _rand((B, hp.max_N, hp.d), 2), _rand((B, hp.max_N, hp.d), 3)
    pma = np.array([0, 57, 178, 179], np.int32)
    R, A, M = engine.attention(Q, K, V, monotonic, pma)
    Rr, Ar, Mr = rt.Attention(torch.from_numpy(Q), torch.from_numpy(K), torch.from_numpy(V), monotonic, pma)
    assert np.abs(R.cpu().numpy() - Rr.numpy()).max() < 1e-4
    assert np.abs(A.cpu().numpy() - Ar.numpy()).max() < 1e-5
    assert M.dtype == torch.int64 and np.array_equal(M.cpu().numpy(), Mr.numpy())
    if monotonic:
        A = A.cpu().numpy()
        for b, p in enumerate(pma):
            live = np.zeros(hp.max_N, bool); live[p:p + 3] = True
            assert (A[b][~live] == 0).all()          # exact zeros outside the window


def test_textenc(engine, params):
    L = synthetic_text(3, 70, seed=5)
    L[2] = 0; L[2, :5] = [3, 4, 5, 6, 1]            # a very short sentence: mostly padding (quirk Q2)
    K, V = engine.textenc(L)
    Kr, Vr = rt.TextEnc(params, L)
    assert np.abs(K.cpu().numpy() - Kr.numpy()).max() < NET_TOL
    assert np.abs(V.cpu().numpy() - Vr.numpy()).max() < NET_TOL


@pytest.mark.parametrize("B,T", [(2, 210), (1, 37)])
def test_audioenc_audiodec(engine, params, B, T):
    S = _rand((B, T, hp.n_mels), 9, 0, 1)
    Q = engine.audioenc(S)
    Qr = rt.AudioEnc(params, torch.from_numpy(S))
    assert np.abs(Q.cpu().numpy() - Qr.numpy()).max() < NET_TOL
    R = _rand((B, T, 2 * hp.d), 10)
    logits, Y = engine.audiodec(R)
    lr, Yr = rt.AudioDec(params, torch.from_numpy(R))
    assert np.abs(logits.cpu().numpy() - lr.numpy()).max() < NET_TOL
    assert np.abs(Y.cpu().numpy() - Yr.numpy()).max() < NET_TOL


@pytest.mark.parametrize("B,T", [(1, 12), (2, 210)])
def test_ssrn(engine, params, B, T):
    Y = _rand((B, T, hp.n_mels), 12, 0, 1)
    logits, Z = engine.ssrn(Y)
    lr, Zr = rt.SSRN(params, torch.from_numpy(Y))
    assert Z.shape == (B, 4 * T, 1 + hp.n_fft // 2)
    assert np.abs(Z.cpu().numpy() - Zr.numpy()).max() < NET_TOL
    assert np.abs(logits.cpu().numpy() - lr.numpy()).max() < 5e-3      # logits are O(10)


def test_ssrn_golden(engine):
    from conftest import golden
    g = golden("ssrn_T12.npz")
    Y = np.random.default_rng(12).uniform(0, 1, (1, 12, hp.n_mels)).astype(np.float32)
    _, Z = engine.ssrn(Y, want_logits=False)
    assert np.abs(Z.cpu().numpy() - g["Z"]).max() < NET_TOL


def test_unfused_composition_matches_library_tables(engine, params):
    """networks.py composed block by block (arch.py tables) == the library's own chains."""
    from dc_tts_b200 import networks
    from dc_tts_b200.modules import variable_scope
    L = synthetic_text(1, 40, seed=8)
    with variable_scope("Text2Mel"), variable_scope("TextEnc"):
        K1, V1 = networks.TextEnc(L, training=False, fused=True)
        K2, V2 = networks.TextEnc(L, training=False, fused=False)
    assert torch.equal(K1, K2) and torch.equal(V1, V2)
    Y = _rand((1, 6, hp.n_mels), 3, 0, 1)
    with variable_scope("SSRN"):
        _, Z1 = networks.SSRN(Y, training=False, fused=True)
        _, Z2 = networks.SSRN(Y, training=False, fused=False)
    assert (Z1 - Z2).abs().max().item() < 1e-6


def test_error_paths(engine):
    from dc_tts_b200.engine import DcttsError
    x = _rand((1, 4, 256), 0)
    with pytest.raises(DcttsError):
        engine.hc("Text2Mel/AudioEnc/NOPE_1", x)
    with pytest.raises(DcttsError):
        engine.hc("Text2Mel/AudioEnc/C_1", x)            # wrong block kind
    with pytest.raises(DcttsError):
        engine.textenc(np.zeros((1, 50), np.int32))      # N must be max_N
