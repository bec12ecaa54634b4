"""GPU parity tests, graph level: one sess.run (train.py:48-68) and the autoregressive
loop (synthesize.py:45-57) against the oracle's literal full-recompute schedule."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from dc_tts_b200.data_load import load_data
from dc_tts_b200.hyperparams import Hyperparams as hp
from dc_tts_b200.params import synthetic_text
from oracle import ref_torch as rt

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("path")]
TOL = 1e-3          # north star: max-abs 1e-3 on mel and linear magnitudes


def test_full_graph_pass(engine, params):
    L = synthetic_text(2, 60, seed=0)
    mels = np.random.default_rng(1).uniform(0, 1, (2, hp.max_T, hp.n_mels)).astype(np.float32)
    pma = np.array([5, 178], np.int32)
    Y, M, A = engine.text2mel_forward(L, mels, pma)
    o = rt.text2mel_forward(params, L, mels, pma)
    assert np.abs(Y.cpu().numpy() - o["Y"].numpy()).max() < TOL
    assert np.array_equal(M.cpu().numpy(), o["max_attentions"].numpy())
    assert np.abs(A.cpu().numpy() - o["alignments"].numpy()).max() < 1e-4


def test_full_graph_golden(engine):
    g = golden("t2m_forward.npz")
    L = synthetic_text(1, 60, seed=3)
    mels = np.random.default_rng(11).uniform(0, 1, (1, hp.max_T, hp.n_mels)).astype(np.float32)
    Y, M, A = engine.text2mel_forward(L, mels, np.array([7], np.int32))
    assert np.abs(Y.cpu().numpy() - g["Y"]).max() < TOL
    assert np.array_equal(M.cpu().numpy(), g["max_attentions"])
    assert np.abs(A.cpu().numpy()[:, 7:10, :] - g["align_win"]).max() < 1e-4


def test_generate_matches_golden_harvard(engine):
    """config 1: Harvard sentence #1, all 210 steps, free running, vs the oracle's literal
    schedule.  The window trajectory (argmax feedback) must be identical; golden margins
    (>= 5e-3 in probability) are far above float32 noise."""
    g = golden("synth_harvard1.npz")
    Y, P, M, A = engine.text2mel_generate(g["L"], want_final_attention=True)
    assert g["margin"].min() > 1e-3
    assert np.array_equal(P.cpu().numpy(), g["p_hist"])
    assert np.abs(Y.cpu().numpy() - g["Y"]).max() < TOL
    _, Z = engine.ssrn(Y, want_logits=False)
    assert np.abs(Z.cpu().numpy()[:, ::8, ::8] - g["Z_sub"]).max() < TOL
    # the final sess.run's attention outputs: every row under the last step's window
    p_last = int(g["p_hist"][0, -1])
    A = A.cpu().numpy()
    assert (A[0, :p_last] == 0).all() and (A[0, p_last + 3:] == 0).all()
    assert np.allclose(A[0].sum(0), 1, atol=1e-5)
    assert ((M.cpu().numpy() >= p_last) & (M.cpu().numpy() < p_last + 3)).all()


def test_generate_batch_vs_oracle(engine, params):
    """B=3 synthetic sentences of different lengths; 60 free-running steps vs the oracle,
    then teacher-forced comparison of every step's row so a near-tie cannot hide an error."""
    L = np.concatenate([synthetic_text(1, n, seed=20 + n) for n in (30, 100, 170)])
    steps = 60
    r = rt.synthesize(params, L, steps=steps, literal=False, record=True)
    Y, P, _, _ = engine.text2mel_generate(L, steps=steps)
    Yo, Po = r["Y"].numpy(), r["p_hist"].numpy()
    ok = r["margin_hist"].numpy().min(1) > 1e-4            # rows whose argmax feedback is well separated
    assert ok.any()
    assert np.array_equal(P.cpu().numpy()[ok, :steps], Po[ok])
    assert np.abs(Y.cpu().numpy()[ok] - Yo[ok]).max() < TOL
    # teacher forced: feed the oracle's own Y prefix and window at step j, compare row j
    for j in (0, 1, 17, 59):
        mels = Yo.copy(); mels[:, j:] = 0
        Yg, Mg, _ = engine.text2mel_forward(L, mels, Po[:, j].astype(np.int32), want_alignments=False)
        assert np.abs(Yg.cpu().numpy()[:, j] - Yo[:, j]).max() < TOL


def test_generate_equals_stepwise_api(engine):
    """The CUDA-graph loop and the step-wise sess.run loop (full recompute each step, as
    synthesize.py does) are the same computation."""
    L = synthetic_text(2, 80, seed=42)
    steps = 30
    Yg, Pg, _, _ = engine.text2mel_generate(L, steps=steps)
    Y = torch.zeros((2, hp.max_T, hp.n_mels), device=engine.device)
    pma = torch.zeros((2,), dtype=torch.int32, device=engine.device)
    for j in range(steps):
        _Y, _M, _ = engine.text2mel_forward(L, Y, pma, want_alignments=False)
        Y[:, j] = _Y[:, j]
        assert torch.equal(pma, Pg[:, j])
        pma = _M[:, j].to(torch.int32)
    assert (Y - Yg).abs().max().item() < 1e-5


def test_generate_is_deterministic_and_batch_independent(engine):
    L = synthetic_text(4, 90, seed=7)
    Y1, P1, _, _ = engine.text2mel_generate(L, steps=40)
    Y2, P2, _, _ = engine.text2mel_generate(L, steps=40)
    assert torch.equal(Y1, Y2) and torch.equal(P1, P2)
    # utterances never interact: permuting the batch permutes the result, bit for bit
    perm = [2, 0, 3, 1]
    Yp, Pp, _, _ = engine.text2mel_generate(L[perm], steps=40)
    assert torch.equal(Yp, Y1[perm]) and torch.equal(Pp, P1[perm])
    # a different batch size may select a different GEMM tiling (different summation
    # order), so across batch sizes the guarantee is the parity tolerance, not bits
    Ys, Ps, _, _ = engine.text2mel_generate(L[2:3], steps=40)
    assert torch.equal(Ps[0], P1[2]) and (Ys[0] - Y1[2]).abs().max().item() < 1e-4


def test_synthesize_host_end_to_end(engine):
    g = golden("synth_harvard1.npz")
    L = np.repeat(g["L"], 2, 0)
    Yh, Zh = engine.synthesize_host(L)
    assert Yh.shape == (2, hp.max_T, hp.n_mels) and Zh.shape == (2, hp.max_T * hp.r, 1 + hp.n_fft // 2)
    assert np.abs(Yh.numpy()[0] - g["Y"][0]).max() < TOL
    assert np.abs(Zh.numpy()[:, ::8, ::8] - g["Z_sub"]).max() < TOL
    assert torch.equal(Yh[0], Yh[1])
    assert engine.launch_count() > 0


def test_graph_session_api(engine):
    """The reference-facing objects: Graph(mode='synthesize') + Session.run with feeds
    (synthesize.py:26,48-57), fused and block-by-block evaluation."""
    from dc_tts_b200.train import Graph, Session
    g = golden("synth_harvard1.npz")
    L = g["L"]
    for fused in (True, False):
        gr = Graph(mode="synthesize", fused=fused)
        with Session() as sess:
            Y = np.zeros((1, hp.max_T, hp.n_mels), np.float32)
            pma = np.zeros((1,), np.int32)
            for j in range(3):
                _gs, _Y, _M, _A = sess.run([gr.global_step, gr.Y, gr.max_attentions, gr.alignments],
                                           {gr.L: L, gr.mels: Y, gr.prev_max_attentions: pma})
                Y[:, j, :] = _Y[:, j, :]
                pma = _M[:, j]
            assert _A.shape == (1, hp.max_N, hp.max_T) and _M.dtype == np.int64
            assert np.abs(Y[:, :3] - g["Y"][:, :3]).max() < TOL
            Z = sess.run(gr.Z, {gr.Y: g["Y"]})
            assert np.abs(Z[:, ::8, ::8] - g["Z_sub"]).max() < TOL
    with pytest.raises(ValueError):
        Graph(mode="train")                  # the training graph needs its input pipeline (tests/test_trainer.py, test_train.py)


def test_synthesize_script(engine, tmp_path, monkeypatch):
    from dc_tts_b200 import synthesize as syn
    monkeypatch.chdir(tmp_path)
    sent = os.path.join(ROOT, "harvard_sentences.txt")
    g = golden("synth_harvard1.npz")
    Y, Z = syn.synthesize(sentences=sent, fast=True, write=True)
    assert Y.shape[0] == 20 and os.path.exists(tmp_path / "samples" / "20.wav")       # synthesize.py:60-64
    assert np.abs(Y[0] - g["Y"][0]).max() < TOL
    from scipy.io.wavfile import read as read_wav
    sr, wav = read_wav(tmp_path / "samples" / "1.wav")
    assert sr == hp.sr and wav.dtype == np.float32 and 0 < len(wav) <= hp.hop_length * (hp.max_T * hp.r - 1)
    assert np.isfinite(wav).all()


def test_synthesize_without_checkpoint_raises(engine, tmp_path, monkeypatch):
    """a missing logdir-1 / logdir-2 checkpoint must fail like the reference's Saver.restore(sess, None)
    (synthesize.py:33,39), never fall back silently to random weights."""
    from dc_tts_b200 import synthesize as syn
    from dc_tts_b200.engine import Engine, set_engine
    monkeypatch.chdir(tmp_path)
    fresh = Engine(0)                                   # no parameters committed
    set_engine(fresh)
    try:
        with pytest.raises(FileNotFoundError):
            syn.synthesize(sentences=os.path.join(ROOT, "harvard_sentences.txt"), write=False)
        Y, Z = syn.synthesize(sentences=os.path.join(ROOT, "harvard_sentences.txt"), write=False, allow_random_init=True, seed=3)
        assert Y.shape[0] == 20 and np.isfinite(Z).all()
    finally:
        set_engine(engine)
        fresh.close()
