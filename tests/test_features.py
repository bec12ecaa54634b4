"""Feature extraction (get_spectrograms / load_spectrograms, reference utils.py:20-65,147-162): oracle invariants on
the CPU, CUDA vs oracle on the GPU.  librosa is absent offline: PARITY UNPINNED (oracle/ref_features.py)."""
import numpy as np
import pytest

from dc_tts_b200.hyperparams import Hyperparams as hp
from oracle import ref_features as rf
from oracle import ref_vocoder as rv


def _speechlike(seed, seconds=2.0, lead=3000, tail=5000):
    rng = np.random.default_rng(seed)
    n = int(hp.sr * seconds)
    t = np.arange(n) / hp.sr
    y = 0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.05 * rng.standard_normal(n)
    y[:lead] *= 1e-5
    y[n - tail:] *= 1e-5
    return y.astype(np.float32)


def test_mel_basis_properties():
    W = rf.mel_basis()
    assert W.shape == (hp.n_mels, 1 + hp.n_fft // 2) and W.min() >= 0
    df = hp.sr / 2.0 / (hp.n_fft // 2)
    np.testing.assert_allclose(W.sum(1) * df, 1.0, atol=0.02)            # area normalisation (norm=1)
    for row in W:                                                         # each filter is one contiguous triangle
        nz = np.flatnonzero(row)
        assert nz.size >= 2 and np.all(np.diff(nz) == 1)
        pk = row.argmax()
        assert np.all(np.diff(row[nz[0]:pk + 1]) >= 0) and np.all(np.diff(row[pk:nz[-1] + 1]) <= 0)
    centres = (W * np.arange(W.shape[1])).sum(1) / W.sum(1)
    assert np.all(np.diff(centres) > 0)
    # Slaney scale: linear below 1 kHz (equal spacing), geometric above
    f = rf.mel_to_hz(np.linspace(rf.hz_to_mel(0.0), rf.hz_to_mel(hp.sr / 2.0), hp.n_mels + 2))
    lin = f[f < 1000]
    np.testing.assert_allclose(np.diff(lin), np.diff(lin)[0], rtol=1e-9)
    np.testing.assert_allclose(rf.hz_to_mel(rf.mel_to_hz(np.array([3.0, 15.0, 40.0]))), [3.0, 15.0, 40.0], rtol=1e-12)
    assert abs(rf.hz_to_mel(1000.0) - 15.0) < 1e-12


def test_oracle_shapes_ranges_and_reduction():
    y = _speechlike(0)
    mel, mag = rf.get_spectrograms(y)
    s, e = rv.trim_indices(y)
    assert 0 < s < e < len(y)
    T = 1 + (e - s) // hp.hop_length
    assert mel.shape == (T, hp.n_mels) and mag.shape == (T, 1 + hp.n_fft // 2)
    assert mel.dtype == np.float32 and mel.min() >= 1e-8 and mel.max() <= 1 and mag.max() <= 1
    m2, g2 = rf.load_spectrograms(y)
    assert g2.shape[0] % hp.r == 0 and m2.shape[0] == g2.shape[0] // hp.r
    np.testing.assert_array_equal(m2, np.pad(mel, [[0, g2.shape[0] - T], [0, 0]])[::hp.r])
    # a pure tone lands in the right linear bin
    t = np.arange(hp.sr) / hp.sr
    tone = (1e-3 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)      # quiet enough not to clip at 1.0
    _, g = rf.get_spectrograms(tone)
    assert abs(int(g[g.shape[0] // 2].argmax()) - round(1000.0 / (hp.sr / hp.n_fft))) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed,seconds", [(0, 2.0), (1, 0.7), (2, 5.3)])
def test_gpu_features_vs_oracle(engine, seed, seconds):
    y = _speechlike(seed, seconds)
    mel_o, mag_o = rf.get_spectrograms(y)
    mel, mag, trim = engine.get_spectrograms(y)
    assert trim == rv.trim_indices(y)
    mel, mag = mel.cpu().numpy(), mag.cpu().numpy()
    assert mel.shape == mel_o.shape and mag.shape == mag_o.shape
    # normalised dB scale: 1e-4 is 0.01 dB; bins at the float32 FFT noise floor are compared in amplitude instead
    lin = lambda z: 10.0 ** ((z * hp.max_db - hp.max_db + hp.ref_db) / 20.0)
    peak = lin(mag_o).max()
    np.testing.assert_allclose(lin(mag), lin(mag_o), atol=2e-6 * peak, rtol=2e-3)
    np.testing.assert_allclose(lin(mel), lin(mel_o), atol=2e-6 * lin(mel_o).max(), rtol=2e-3)
    loud = mag_o > 0.35                                                       # > -45 dB re full scale
    assert np.abs(mag - mag_o)[loud].max() < 1e-4
    assert np.abs(mel - mel_o)[mel_o > 0.35].max() < 1e-4


@pytest.mark.gpu
def test_gpu_load_spectrograms_from_wav_file(engine, tmp_path):
    from scipy.io import wavfile
    from dc_tts_b200 import utils
    from dc_tts_b200.engine import set_engine
    set_engine(engine)
    y = _speechlike(3, 1.5)
    pcm = np.round(y * 32767).astype(np.int16)
    path = str(tmp_path / "LJ001-0001.wav")
    wavfile.write(path, hp.sr, pcm)
    fname, mel, mag = utils.load_spectrograms(path)
    m_o, g_o = rf.load_spectrograms(pcm.astype(np.float32) / 32768.0)
    assert fname == "LJ001-0001.wav" and mel.shape == m_o.shape and mag.shape == g_o.shape
    assert np.abs(mag - g_o)[g_o > 0.35].max() < 1e-4 and np.abs(mel - m_o)[m_o > 0.35].max() < 1e-4
    with pytest.raises(ValueError):
        wavfile.write(path, 16000, pcm)
        utils.load_spectrograms(path)


def test_mel_basis_matches_transformers_slaney_filter_bank():
    """transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") documents itself as librosa.filters.mel's
    equivalent: an independent implementation of what oracle/ref_features.mel_basis restates."""
    au = pytest.importorskip("transformers.audio_utils")
    M = au.mel_filter_bank(num_frequency_bins=1 + hp.n_fft // 2, num_mel_filters=hp.n_mels, min_frequency=0.0,
                           max_frequency=hp.sr / 2, sampling_rate=hp.sr, norm="slaney", mel_scale="slaney")
    assert np.abs(M.T - rf.mel_basis()).max() < 1e-12
