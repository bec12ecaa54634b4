"""CPU tests: the oracle against itself (two independent restatements), against the
committed golden vectors, and structural facts of the reference it must reproduce."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from dc_tts_b200 import arch
from dc_tts_b200.data_load import load_data
from dc_tts_b200.hyperparams import Hyperparams as hp
from dc_tts_b200.params import init_params, num_params, synthetic_text
from oracle import ref_numpy as rn
from oracle import ref_torch as rt


def test_param_counts_match_survey():
    # Text2Mel 23,970,288 params, SSRN 28,410,383
    assert num_params("Text2Mel") == 23970288
    assert num_params("SSRN") == 28410383
    assert len(arch.textenc_layers()) == 14 and len(arch.audioenc_layers()) == 13
    assert len(arch.audiodec_layers()) == 11 and len(arch.ssrn_layers()) == 16


def test_initialisers_follow_reference():
    P = init_params(0)
    k = P["SSRN/HC_11/conv1d/kernel"]                      # [3,1024,2048], fan_in 3072
    assert abs(k.std() - np.sqrt(2.6 / 3072) * 0.8796) < 2e-4     # truncated-normal std factor
    assert np.abs(k).max() <= 2 * np.sqrt(2.6 / 3072) + 1e-7
    assert P["SSRN/HC_11/conv1d/bias"].max() == 0 and P["SSRN/HC_11/H1/gamma"].min() == 1
    assert np.abs(P["Text2Mel/TextEnc/embed_1/lookup_table"]).max() <= 0.2 + 1e-7


@pytest.mark.parametrize("k,rate,pad", [(1, 1, "SAME"), (3, 1, "SAME"), (3, 9, "SAME"), (3, 27, "CAUSAL"), (3, 3, "CAUSAL")])
def test_conv_two_restatements(k, rate, pad):
    rng = np.random.default_rng(k * 100 + rate)
    x = rng.standard_normal((2, 40, 24)).astype(np.float32)
    W = rng.standard_normal((k, 24, 16)).astype(np.float32)
    b = rng.standard_normal(16).astype(np.float32)
    a = rt._conv(torch.from_numpy(x), torch.from_numpy(W), torch.from_numpy(b), rate, pad).numpy()
    c = rn._conv(x, W, b, rate, pad)
    assert np.abs(a - c).max() < 1e-4
    if pad == "CAUSAL":        # output at t must not depend on inputs after t
        x2 = x.copy(); x2[:, 20:] += 1.0
        c2 = rn._conv(x2, W, b, rate, pad)
        assert np.array_equal(c[:, :20], c2[:, :20])


def test_deconv_two_restatements_and_tap_layout():
    P = {"D/conv2d_transpose/kernel": np.random.default_rng(0).standard_normal((1, 3, 8, 8)).astype(np.float32),
         "D/conv2d_transpose/bias": np.random.default_rng(1).standard_normal(8).astype(np.float32),
         "D/normalize/gamma": np.ones(8, np.float32), "D/normalize/beta": np.zeros(8, np.float32)}
    x = np.random.default_rng(2).standard_normal((2, 9, 8)).astype(np.float32)
    a = rt.conv1d_transpose(P, torch.from_numpy(x), "D").numpy()
    c = rn.conv1d_transpose(P, x, "D")
    assert a.shape == (2, 18, 8) and np.abs(a - c).max() < 1e-5
    # out[2t] = W0 x[t] + W2 x[t-1], out[2t+1] = W1 x[t] (before LN)
    W = P["D/conv2d_transpose/kernel"][0]
    pre = np.zeros((2, 18, 8), np.float32)
    for t in range(9):
        pre[:, 2 * t] = x[:, t] @ W[0].T + (x[:, t - 1] @ W[2].T if t > 0 else 0)
        pre[:, 2 * t + 1] = x[:, t] @ W[1].T
    pre += P["D/conv2d_transpose/bias"]
    assert np.abs(rn.normalize(pre, 1.0, 0.0) - c).max() < 1e-5


def test_attention_window_semantics():
    rng = np.random.default_rng(5)
    Q = rng.standard_normal((3, hp.max_T, hp.d)).astype(np.float32)
    K = rng.standard_normal((3, hp.max_N, hp.d)).astype(np.float32)
    V = rng.standard_normal((3, hp.max_N, hp.d)).astype(np.float32)
    pma = np.array([0, 100, 179])
    R, A, M = rt.Attention(torch.from_numpy(Q), torch.from_numpy(K), torch.from_numpy(V), True, pma)
    R2, A2, M2 = rn.Attention(Q, K, V, True, pma)
    A = A.numpy()
    for b, p in enumerate(pma):      # keys p <= n < p+3 only (networks.py:141-147), exact zeros elsewhere
        live = np.zeros(hp.max_N, bool); live[p:p + 3] = True
        assert (A[b][~live] == 0).all() and np.allclose(A[b][live].sum(0), 1, atol=1e-6)
    assert np.abs(A - A2).max() < 1e-6 and np.array_equal(M.numpy(), M2)
    assert np.abs(R.numpy() - R2).max() < 1e-5


def test_full_graph_two_restatements(params):
    L = synthetic_text(2, 60, 0)
    mels = np.random.default_rng(1).uniform(0, 1, (2, hp.max_T, hp.n_mels)).astype(np.float32)
    pma = np.array([5, 178])
    a = rt.text2mel_forward(params, L, mels, pma)
    b = rn.text2mel_forward(params, L, mels, pma)
    for k in ("K", "Q", "R", "Y", "alignments"):
        assert np.abs(a[k].numpy() - b[k]).max() < 1e-4, k
    assert np.array_equal(a["max_attentions"].numpy(), b["max_attentions"])


def test_ssrn_fp32_vs_fp64(params):
    Y = np.random.default_rng(12).uniform(0, 1, (1, 12, hp.n_mels)).astype(np.float32)
    _, z = rt.SSRN(params, torch.from_numpy(Y))
    P64 = {k: v.astype(np.float64) for k, v in params.items() if k.startswith("SSRN")}
    _, z64 = rn.SSRN(P64, Y.astype(np.float64))
    assert z.shape == (1, 48, 1 + hp.n_fft // 2)
    assert np.abs(z.numpy() - z64).max() < 1e-4
    g = golden("ssrn_T12.npz")
    assert np.abs(z.numpy() - g["Z"]).max() < 2e-5


def test_golden_t2m_forward(params):
    g = golden("t2m_forward.npz")
    L = synthetic_text(1, 60, seed=3)
    mels = np.random.default_rng(11).uniform(0, 1, (1, hp.max_T, hp.n_mels)).astype(np.float32)
    o = rn.text2mel_forward(params, L, mels, np.array([7]))          # the OTHER restatement
    assert np.abs(o["Y"] - g["Y"]).max() < 1e-4
    assert np.array_equal(o["max_attentions"], g["max_attentions"])
    assert np.abs(o["alignments"][:, 7:10, :] - g["align_win"]).max() < 1e-5


def test_q1_history_is_remasked(params):
    """Quirk Q1: Y[:, j] depends on the window applied to EARLIER rows, so
    a decoder that freezes old rows' attention is not the reference."""
    L = synthetic_text(1, 60, seed=3)
    mels = np.random.default_rng(11).uniform(0, 1, (1, hp.max_T, hp.n_mels)).astype(np.float32)
    a = rt.text2mel_forward(params, L, mels, np.array([7]))["Y"][:, 100]
    b = rt.text2mel_forward(params, L, mels, np.array([8]))["Y"][:, 100]
    assert (a - b).abs().max() > 1e-3
    # ... but only through the 85-frame AudioDec receptive field (App. A)
    R7 = rt.text2mel_forward(params, L, mels, np.array([7]))["R"]
    R7m = R7.clone(); R7m[:, :100 - 84] = 0
    y1 = rt.AudioDec(params, R7)[1][:, 100]
    y2 = rt.AudioDec(params, R7m)[1][:, 100]
    assert torch.equal(y1, y2)


def test_golden_synthesize_prefix(params):
    """First AR steps of synthesize.py:45-54 under the LITERAL schedule reproduce the golden."""
    g = golden("synth_harvard1.npz")
    L = load_data("synthesize", os.path.join(ROOT, "harvard_sentences.txt"))[:1]
    assert np.array_equal(L, g["L"])
    r = rt.synthesize(params, L, steps=6, literal=True, record=True)
    assert np.abs(r["Y"].numpy()[:, :6] - g["Y"][:, :6]).max() < 1e-5
    assert np.array_equal(r["p_hist"].numpy(), g["p_hist"][:, :6])
