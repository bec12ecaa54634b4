"""Generates the committed golden fixtures from the ORACLE (oracle/ref_torch.py, fp32).

The reference itself cannot run here (TF1 absent) and ships no vectors, so these pin
the oracle in time rather than the reference ("parity unpinned").  Inputs
are re-derivable from seeds (dc_tts_b200.params.init_params / synthetic_text), outputs
are stored.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dc_tts_b200.data_load import load_data                      # noqa: E402
from dc_tts_b200.hyperparams import Hyperparams as hp            # noqa: E402
from dc_tts_b200.params import init_params, synthetic_text       # noqa: E402
from oracle import ref_torch as rt                               # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(os.cpu_count())
P = init_params(0, "perturbed")

# 1. one full-graph pass (train.py:48-68) on seeded inputs
L = synthetic_text(1, 60, seed=3)
mels = np.random.default_rng(11).uniform(0, 1, (1, hp.max_T, hp.n_mels)).astype(np.float32)
pma = np.array([7], np.int32)
o = rt.text2mel_forward(P, L, mels, pma)
np.savez_compressed(os.path.join(OUT, "t2m_forward.npz"),
                    Y=o["Y"].numpy(), max_attentions=o["max_attentions"].numpy(),
                    Q_sub=o["Q"].numpy()[:, ::10, :16], K_sub=o["K"].numpy()[:, ::10, :16],
                    R_sub=o["R"].numpy()[:, ::10, ::16],
                    align_win=o["alignments"].numpy()[:, 7:10, :])

# 2. SSRN on a short mel (12 frames -> 48 x 1025)
Ys = np.random.default_rng(12).uniform(0, 1, (1, 12, hp.n_mels)).astype(np.float32)
zl, z = rt.SSRN(P, torch.from_numpy(Ys))
np.savez_compressed(os.path.join(OUT, "ssrn_T12.npz"), Z=z.numpy(), Z_logits_sub=zl.numpy()[:, :, ::8])

# 3. the synthesize loop (synthesize.py:45-57) on Harvard sentence #1, literal schedule
Lh = load_data("synthesize", os.path.join(ROOT, "harvard_sentences.txt"))[:1]
with torch.no_grad():
    r = rt.synthesize(P, Lh, literal=False, record=True)
np.savez_compressed(os.path.join(OUT, "synth_harvard1.npz"), L=Lh, Y=r["Y"].numpy(),
                    p_hist=r["p_hist"].numpy().astype(np.int32), margin=r["margin_hist"].numpy(),
                    Z_sub=r["Z"].numpy()[:, ::8, ::8])
print("golden fixtures written to", OUT)
