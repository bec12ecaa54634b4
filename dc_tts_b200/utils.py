"""Vocoder half of the reference's utils.py (/root/reference/utils.py:67-114) with the same names:
`spectrogram2wav(mag)`, `griffin_lim(spectrogram)`, `invert_spectrogram(spectrogram)`.

The reference runs librosa's stft / istft on the CPU, 50 + 51 times per utterance; here the whole
Griffin-Lim loop runs on the GPU (csrc/kernels_vocoder.cu: one CTA per STFT frame, a 2048-point FFT
in shared memory) behind `dctts_spectrogram2wav`.  This is the first "next" row of SURVEY.md 8(f), not part
of the Text2Mel + SSRN hot path.  Feature extraction (`get_spectrograms`, `load_spectrograms`,
utils.py:20-65,147-162) runs on the GPU too (`dctts_get_spectrograms`: trim, pre-emphasis, STFT, mel
filterbank, dB, normalisation in one kernel per utterance); `librosa.load` is replaced by scipy's WAV reader
for files that already have hp.sr (LJ Speech does) -- resampling, plotting and the training helpers of the
reference's utils.py stay out of scope.
"""
import os

import numpy as np

from .engine import get_engine
from .hyperparams import Hyperparams as hp


def spectrogram2wav(mag):
    """utils.py:67-94.  mag: (T, 1+n_fft//2) normalised magnitudes -> trimmed float32 wav (numpy)."""
    wav, trim = get_engine().spectrogram2wav(np.asarray(mag, np.float32)[None])
    s, e = int(trim[0, 0]), int(trim[0, 1])
    return wav[0, s:e].cpu().numpy().astype(np.float32)


def spectrograms2wavs(mags):
    """Batched form: (B, T, F) -> list of trimmed wavs (one device pass for the whole batch)."""
    wav, trim = get_engine().spectrogram2wav(mags)
    wav = wav.cpu().numpy()
    return [wav[b, int(trim[b, 0]):int(trim[b, 1])].astype(np.float32) for b in range(wav.shape[0])]


def griffin_lim(spectrogram):
    """utils.py:96-107.  spectrogram: (1+n_fft//2, t) amplitude (already ** hp.power) -> waveform."""
    S = np.asarray(spectrogram, np.float32).T
    # undo the de-normalisation the device entry point applies: S = (10 ^ ((z*max_db - max_db + ref_db)/20)) ^ power
    z = (20.0 * np.log10(np.maximum(S, 1e-30) ** (1.0 / hp.power)) + hp.max_db - hp.ref_db) / hp.max_db
    if z.min() < 0 or z.max() > 1:
        raise ValueError("griffin_lim: amplitude outside the range spectrogram2wav can produce")
    e = get_engine()
    wav, _ = e.spectrogram2wav(z[None].astype(np.float32))
    # spectrogram2wav also de-pre-emphasises; Griffin-Lim alone does not: invert y[n] = x[n] + c y[n-1]
    y = wav[0].cpu().numpy().astype(np.float64)
    x = y.copy()
    x[1:] -= hp.preemphasis * y[:-1]
    return x.astype(np.float32)


def invert_spectrogram(spectrogram):
    """utils.py:109-114: one inverse STFT (librosa.istft conventions) = Griffin-Lim with zero iterations."""
    S = np.asarray(spectrogram)
    if np.iscomplexobj(S):
        raise NotImplementedError("invert_spectrogram: only zero-phase (real) input is exposed; the complex "
                                  "iterations run inside dctts_spectrogram2wav")
    z = (20.0 * np.log10(np.maximum(S.T, 1e-30) ** (1.0 / hp.power)) + hp.max_db - hp.ref_db) / hp.max_db
    e = get_engine()
    wav, _ = e.spectrogram2wav(np.clip(z, 0, 1)[None].astype(np.float32), n_iter=0)
    y = wav[0].cpu().numpy().astype(np.float64)
    x = y.copy()
    x[1:] -= hp.preemphasis * y[:-1]
    return x.astype(np.float32)


def _load_wav(fpath):
    """What `librosa.load(fpath, sr=hp.sr)` returns for a mono PCM / float WAV that already has hp.sr."""
    from scipy.io import wavfile
    sr, y = wavfile.read(fpath)
    if sr != hp.sr:
        raise ValueError("%s: sample rate %d != hp.sr %d (resampling is not implemented)" % (fpath, sr, hp.sr))
    if y.ndim > 1:
        y = y.mean(axis=1)
    if y.dtype == np.int16:
        y = y.astype(np.float32) / 32768.0
    elif y.dtype == np.int32:
        y = y.astype(np.float32) / 2147483648.0
    elif y.dtype == np.uint8:
        y = (y.astype(np.float32) - 128.0) / 128.0
    return np.ascontiguousarray(y, np.float32)


def get_spectrograms(fpath):
    """utils.py:20-65.  `fpath`: a WAV file path, or the already loaded waveform (1-D float array at hp.sr).
    Returns normalised mel (T, n_mels) and linear magnitude (T, 1+n_fft/2), float32 numpy."""
    y = _load_wav(fpath) if isinstance(fpath, (str, bytes, os.PathLike)) else np.asarray(fpath, np.float32)
    mel, mag, _ = get_engine().get_spectrograms(y)
    return mel.cpu().numpy(), mag.cpu().numpy()


def load_spectrograms(fpath):
    """utils.py:147-162: pads T to a multiple of hp.r and keeps every r-th mel frame."""
    fname = os.path.basename(fpath) if isinstance(fpath, (str, bytes, os.PathLike)) else None
    mel, mag = get_spectrograms(fpath)
    t = mel.shape[0]
    num_paddings = hp.r - (t % hp.r) if t % hp.r != 0 else 0
    mel = np.pad(mel, [[0, num_paddings], [0, 0]], mode="constant")
    mag = np.pad(mag, [[0, num_paddings], [0, 0]], mode="constant")
    mel = mel[::hp.r, :]
    return fname, mel, mag


def guided_attention(g=0.2):
    """utils.py:134-140: W[n, t] = 1 - exp(-(t/max_T - n/max_N)^2 / (2 g^2)), shape (max_N, max_T) float32 (the device
    training step builds the same table itself: dctts_train_init)."""
    n = np.arange(hp.max_N, dtype=np.float64)[:, None] / float(hp.max_N)
    t = np.arange(hp.max_T, dtype=np.float64)[None, :] / float(hp.max_T)
    return (1.0 - np.exp(-(t - n) ** 2 / (2.0 * g * g))).astype(np.float32)


def learning_rate_decay(init_lr, global_step, warmup_steps=4000.):
    """utils.py:142-145, the Noam scheme: step = global_step + 1; lr * warmup^0.5 * min(step * warmup^-1.5, step^-0.5)."""
    step = float(global_step + 1)
    return float(init_lr) * warmup_steps ** 0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)
