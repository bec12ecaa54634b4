"""Vocoder (reference utils.py:67-114): CPU tests pin the oracle's librosa restatement by its
invariants; the GPU test compares the CUDA Griffin-Lim with the oracle."""
import numpy as np
import pytest

from dc_tts_b200.hyperparams import Hyperparams as hp
from oracle import ref_vocoder as rv


def test_stft_istft_roundtrip_and_shapes():
    x = np.random.default_rng(0).standard_normal(hp.hop_length * 39).astype(np.float32)
    S = rv.stft(x)
    assert S.shape == (1 + hp.n_fft // 2, 40) and S.dtype == np.complex64
    y = rv.istft(S)
    assert len(y) == hp.hop_length * 39
    assert np.abs(y - x).max() < 5e-6                    # perfect reconstruction (window sum-square normalisation)
    # linearity and the Hermitian extension: a pure bin comes back as a windowed cosine of that frequency
    k = 100
    S1 = np.zeros((1025, 9), np.complex64); S1[k, 4] = 1.0
    y1 = rv.istft(S1)
    n = np.arange(len(y1))
    peak = 4 * hp.hop_length
    assert np.argmax(np.abs(y1)) in range(peak - 20, peak + 20)
    assert abs(np.abs(np.fft.rfft(y1 * 1.0)).argmax() * hp.n_fft / len(y1) - k) < 2


def test_window_and_sumsquare():
    w = rv.hann_padded()
    lpad = (hp.n_fft - hp.win_length) // 2
    assert w.shape == (2048,) and w[:lpad].max() == 0 and w[lpad + hp.win_length:].max() == 0
    assert abs(w[lpad + hp.win_length // 2] - 1.0) < 1e-6 and w[lpad] == 0.0      # periodic Hann
    wss = rv.window_sumsquare(40)
    mid = wss[4000:8000]
    assert mid.min() > 1.4 and mid.max() < 1.6           # 1102/275 ~ 4 overlapping hann^2 -> ~1.5


def test_trim_indices():
    y = np.zeros(50000, np.float32)
    y[12000:30000] = np.random.default_rng(1).standard_normal(18000).astype(np.float32) * 0.1
    s, e = rv.trim_indices(y)
    assert s % 512 == 0 and 10000 <= s <= 12000 and 30000 <= e <= 32000
    assert rv.trim_indices(np.zeros(5000, np.float32)) == (0, 5000) or rv.trim_indices(np.zeros(5000, np.float32)) == (0, 0)


def test_griffin_lim_float32_vs_float64():
    mag = np.random.default_rng(2).uniform(0.2, 0.9, (60, 1025)).astype(np.float32)
    w32, _, f32 = rv.spectrogram2wav(mag, n_iter=8)
    w64, _, f64 = rv.spectrogram2wav(mag, n_iter=8, dtype=np.float64)
    assert len(f32) == hp.hop_length * 59
    assert np.abs(f32 - f64).max() < 1e-4 * max(1.0, np.abs(f64).max())


@pytest.mark.gpu
@pytest.mark.parametrize("T,n_iter", [(60, 0), (60, 5), (840, 3)])
def test_gpu_vocoder_vs_oracle(engine, T, n_iter):
    mag = np.random.default_rng(T).uniform(0.1, 0.95, (2, T, 1025)).astype(np.float32)
    mag[1, T // 2:] *= 0.05                                  # a quiet second half: exercises trim
    wav, trim = engine.spectrogram2wav(mag, n_iter=n_iter)
    wav = wav.cpu().numpy()
    for b in range(2):
        w, se, full = rv.spectrogram2wav(mag[b], n_iter=n_iter)
        scale = np.abs(full).max()
        assert wav.shape[1] == len(full)
        assert np.abs(wav[b] - full).max() < 2e-3 * scale, (b, np.abs(wav[b] - full).max(), scale)
        assert abs(int(trim[b, 0]) - se[0]) <= 512 and abs(int(trim[b, 1]) - se[1]) <= 512


@pytest.mark.gpu
def test_gpu_vocoder_full_default_iterations(engine):
    """hp.n_iter = 50 on one utterance-sized spectrogram: Griffin-Lim keeps the target magnitudes"""
    from dc_tts_b200 import utils
    mag = np.random.default_rng(7).uniform(0.3, 0.8, (210, 1025)).astype(np.float32)
    wav = utils.spectrogram2wav(mag)
    assert wav.dtype == np.float32 and 0 < len(wav) <= hp.hop_length * 209
    w, se, full = rv.spectrogram2wav(mag)                    # oracle, float32, 50 iterations
    assert abs(len(wav) - len(w)) <= 1024
    n = min(len(wav), len(w))
    # 50 projections amplify float32 rounding differently on the two sides: compare energies, not samples
    assert abs(np.sqrt(np.mean(wav[:n] ** 2)) / np.sqrt(np.mean(w[:n] ** 2)) - 1) < 0.05


def test_stft_istft_match_torch_librosa_compatible_implementations():
    """torch.stft / torch.istft follow librosa's conventions by design (centred reflect padding, window zero-padded to n_fft,
    window-sum-square normalisation, n_fft/2 trimmed): an independent implementation of what oracle/ref_vocoder.py restates."""
    import torch
    from dc_tts_b200.hyperparams import Hyperparams as hp
    from oracle import ref_vocoder as rv
    y = np.random.default_rng(0).standard_normal(8000).astype(np.float32)
    win = torch.hann_window(hp.win_length, periodic=True)
    S = rv.stft(y)
    St = torch.stft(torch.from_numpy(y), hp.n_fft, hp.hop_length, hp.win_length, window=win, center=True, pad_mode="reflect",
                    return_complex=True).numpy()
    assert S.shape == St.shape and np.abs(S - St).max() < 1e-6 * np.abs(St).max()
    yi = rv.istft(S)
    yt = torch.istft(torch.from_numpy(St), hp.n_fft, hp.hop_length, hp.win_length, window=win, center=True).numpy()
    assert yi.shape == yt.shape and np.abs(yi - yt).max() < 5e-6
