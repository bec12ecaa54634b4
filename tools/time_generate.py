"""Time Text2Mel generation (TextEnc + 210 CUDA-graph steps) for a list of batch sizes.
   python tools/time_generate.py 1 32"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dc_tts_b200.engine import Engine
from dc_tts_b200.params import init_params, synthetic_text

e = Engine(0)
e.load_params(init_params(0, "perturbed"))
for B in [int(x) for x in sys.argv[1:]] or [1, 32]:
    L = synthetic_text(B, 100, seed=0)
    for _ in range(2):
        e.text2mel_generate(L)
    torch.cuda.synchronize()
    n = 5
    t0 = time.perf_counter()
    for _ in range(n):
        Y = e.text2mel_generate(L)[0]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("generate B=%d: %.2f ms (%.1f us/step)  checksum %.6f" % (B, dt * 1e3, dt * 1e6 / 210, float(Y.double().sum())))
