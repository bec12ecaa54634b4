"""ORACLE (test infrastructure, not product) -- numpy restatement of the reference's vocoder
`spectrogram2wav` (/root/reference/utils.py:67-114): de-normalise, amplitude ** hp.power,
Griffin-Lim (50 x istft -> stft -> keep phase), de-pre-emphasis, trim.

The signal-processing primitives live in a third-party dependency that is absent here and not
pinned by the reference (`librosa`, README.md:7; the code's call signatures --
`librosa.stft(X_t, hp.n_fft, hp.hop_length, win_length=...)` with positional n_fft/hop -- are
those of librosa 0.5/0.6).  This file restates the published librosa 0.6 algorithms:
  core.stft      center=True -> reflect-pad n_fft//2, periodic Hann of win_length zero-padded
                 (centred) to n_fft, frames of n_fft every hop, rfft;
  core.istft     per frame irfft x the same padded window, overlap-add, divide by the summed
                 squared window where it exceeds tiny(), drop n_fft//2 samples at both ends;
  effects.trim   top_db=60, frame_length=2048, hop_length=512, RMS per centred (reflect-padded)
                 frame, 10*log10 relative to the maximum, first/last frame above -60 dB.
PARITY: librosa itself and vectors of it do not exist offline.  The restatement is pinned by (i) torch.stft / torch.istft,
which implement librosa's conventions by design (agreement 2e-7 relative / 1e-6 absolute, tests/test_vocoder.py), (ii) the
reference's own utils.spectrogram2wav executed with these primitives plugged in for librosa (tests/test_reference_shim.py:
the composition), (iii) self-consistency (istft(stft(x)) == x).  `effects.trim` has no independent implementation here.
dtype: the reference runs in float32 / complex64 (Z is float32, librosa's default dtypes);
`dtype=np.float64` gives the double-precision variant used to bound float32 noise.
"""
import numpy as np
import scipy.fft
import scipy.signal

from dc_tts_b200.hyperparams import Hyperparams as hp


def hann_padded(n_fft=None, win_length=None, dtype=np.float32):
    """scipy.signal.get_window('hann', win_length, fftbins=True), util.pad_center'ed to n_fft."""
    n_fft = n_fft or hp.n_fft
    win_length = win_length or hp.win_length
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, np.float64)
    out[lpad:lpad + win_length] = w
    return out.astype(dtype)


def stft(y, n_fft=None, hop=None, win_length=None):
    """librosa.core.stft(y, n_fft, hop, win_length=win_length) -> (1 + n_fft/2, frames)."""
    n_fft = n_fft or hp.n_fft
    hop = hop or hp.hop_length
    cdt = np.complex64 if y.dtype == np.float32 else np.complex128
    w = hann_padded(n_fft, win_length, y.dtype)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * w[:, None]
    return scipy.fft.fft(frames, axis=0)[:1 + n_fft // 2].astype(cdt)


def window_sumsquare(n_frames, n_fft=None, hop=None, win_length=None, dtype=np.float32):
    """librosa.filters.window_sumsquare('hann', n_frames, hop, win_length, n_fft)."""
    n_fft = n_fft or hp.n_fft
    hop = hop or hp.hop_length
    n = n_fft + hop * (n_frames - 1)
    x = np.zeros(n, dtype)
    win_sq = hann_padded(n_fft, win_length, np.float64) ** 2
    for i in range(n_frames):
        s = i * hop
        x[s:min(n, s + n_fft)] += win_sq[:max(0, min(n_fft, n - s))].astype(dtype)
    return x


def istft(S, hop=None, win_length=None):
    """librosa.core.istft(S, hop, win_length=win_length, window='hann') (center=True)."""
    hop = hop or hp.hop_length
    n_fft = 2 * (S.shape[0] - 1)
    n_frames = S.shape[1]
    rdt = np.float32 if S.dtype in (np.float32, np.complex64) else np.float64
    w = hann_padded(n_fft, win_length, rdt)
    y = np.zeros(n_fft + hop * (n_frames - 1), rdt)
    full = np.concatenate((S, np.conj(S[-2:0:-1])), 0)
    frames = scipy.fft.ifft(full, axis=0).real.astype(rdt) * w[:, None]
    for i in range(n_frames):
        y[i * hop:i * hop + n_fft] += frames[:, i]
    wss = window_sumsquare(n_frames, n_fft, hop, win_length, rdt)
    nz = wss > np.finfo(rdt).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:-(n_fft // 2)]


def griffin_lim(spectrogram, n_iter=None):
    """utils.py:96-107."""
    n_iter = hp.n_iter if n_iter is None else n_iter
    X_best = spectrogram.copy()
    for _ in range(n_iter):
        X_t = istft(X_best)
        est = stft(X_t)
        phase = est / np.maximum(1e-8, np.abs(est))
        X_best = spectrogram * phase
    return np.real(istft(X_best))


def trim_indices(y, top_db=60, frame_length=2048, hop_length=512):
    """librosa.effects.trim(y)[1]: (start, end) sample indices."""
    yp = np.pad(y, frame_length // 2, mode="reflect")
    n_frames = 1 + (len(yp) - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    mse = np.mean(np.abs(yp[idx]) ** 2, axis=0)
    db = 10.0 * np.log10(np.maximum(1e-10, mse)) - 10.0 * np.log10(np.maximum(1e-10, mse.max()))
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return 0, 0
    return int(nz[0] * hop_length), min(len(y), int((nz[-1] + 1) * hop_length))


def denormalise(mag):
    """utils.py:78-85 on a (T, 1+n_fft/2) float32 magnitude: -> amplitude ** hp.power, (F, T)."""
    m = mag.T
    m = (np.clip(m, 0, 1) * hp.max_db) - hp.max_db + hp.ref_db
    m = np.power(10.0, m * 0.05)
    return m ** hp.power


def spectrogram2wav(mag, n_iter=None, dtype=np.float32, trim=True):
    """utils.py:67-94.  Returns (wav float32, (start, end), untrimmed wav)."""
    S = denormalise(np.asarray(mag, np.float32)).astype(dtype)
    wav = griffin_lim(S, n_iter)
    wav = scipy.signal.lfilter([1], [1, -hp.preemphasis], wav)
    se = trim_indices(wav) if trim else (0, len(wav))
    return wav[se[0]:se[1]].astype(np.float32), se, wav.astype(np.float32)
