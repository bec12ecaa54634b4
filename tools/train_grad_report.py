"""Per-tensor gradient deviation of the CUDA training steps from the autograd oracle (max |g - g_ref| / max |g_ref|)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dc_tts_b200.engine import Engine
from dc_tts_b200.hyperparams import Hyperparams as hp
from dc_tts_b200.params import init_params
from oracle import ref_train as rtr

P = init_params(0, "perturbed")
eng = Engine(0)
eng.load_params(P)
for (B, T, rate, seed) in [(2, 16, 0.0, 0), (2, 12, 0.05, 9), (2, 16, 0.05, 1), (1, 8, 0.0, 0)]:
    mels = np.random.default_rng(3).uniform(0, 1, (B, T, hp.n_mels)).astype(np.float32)
    mags = np.random.default_rng(4).uniform(0, 1, (B, 4 * T, 1025)).astype(np.float32)
    _, _, info = rtr.train_step_ssrn(P, mels, mags, global_step=3999, seed=seed, rate=rate)
    eng.train_init_ssrn(B, T, rate)
    out = eng.train_step_ssrn(mels, mags, global_step=3999, seed=seed, apply=False)
    errs = []
    for n, ref in info["grads"].items():
        g = eng.train_tensor(n, "grad")
        errs.append((float(np.abs(np.clip(g, -1, 1) - ref).max() / max(np.abs(ref).max(), 1e-8)), n))
    errs.sort(reverse=True)
    print("SSRN B=%d T=%d rate=%.2f: loss %.6f vs %.6f; worst:" % (B, T, rate, out["loss"], info["loss"]),
          ["%s %.1e" % (n.replace("SSRN/", ""), e) for e, n in errs[:6]], "median %.1e" % np.median([e for e, _ in errs]), flush=True)
