"""Time the Griffin-Lim vocoder (dctts_spectrogram2wav) on synthetic spectrograms.
   python tools/profile_vocoder.py [B ...]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from dc_tts_b200.engine import Engine
from dc_tts_b200.hyperparams import Hyperparams as hp

eng = Engine(device=0)
Bs = [int(a) for a in sys.argv[1:]] or [1, 32]
T, F = hp.max_T * hp.r, hp.n_fft // 2 + 1
for B in Bs:
    g = torch.Generator(device="cuda").manual_seed(0)
    Z = torch.rand(B, T, F, device="cuda", generator=g) * 0.6 + 0.2
    for _ in range(2):
        eng.spectrogram2wav(Z)
    torch.cuda.synchronize()
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        wav, _ = eng.spectrogram2wav(Z)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("vocoder B=%d: %.2f ms, %.0fx real time (%d iterations)" % (B, dt * 1e3, B * wav.shape[1] / hp.sr / dt, hp.n_iter))
