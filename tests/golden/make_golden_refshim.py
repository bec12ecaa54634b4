"""Generates golden fixtures from the REFERENCE'S OWN CODE (/root/reference/{modules,networks,train}.py), executed
under the TensorFlow API stand-in of tf_shim.py -- same seeded inputs as make_golden.py, so the two fixture sets
are directly comparable.  Needs /root/reference (this container only); run from the repo root:
    python tests/golden/make_golden_refshim.py          (~10 min: 210 full-graph passes on the CPU)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import tf_shim                                                   # noqa: E402
from dc_tts_b200.data_load import load_data                      # noqa: E402
from dc_tts_b200.hyperparams import Hyperparams as hp            # noqa: E402
from dc_tts_b200.params import init_params, synthetic_text       # noqa: E402

P = init_params(0, "perturbed")
store = tf_shim.Store(P)
tf_shim.install(store)

# 1. one full-graph pass (train.py:48-68) on seeded inputs
L = synthetic_text(1, 60, seed=3)
mels = np.random.default_rng(11).uniform(0, 1, (1, hp.max_T, hp.n_mels)).astype(np.float32)
pma = np.array([7], np.int32)
o = tf_shim.run_graph(L, mels, pma, fetch=("Y", "max_attentions", "alignments", "Q", "K", "R"))
np.savez_compressed(os.path.join(HERE, "refshim_t2m_forward.npz"), Y=o["Y"], max_attentions=o["max_attentions"],
                    Q_sub=o["Q"][:, ::10, :16], K_sub=o["K"][:, ::10, :16], R_sub=o["R"][:, ::10, ::16],
                    align_win=o["alignments"][:, 7:10, :])

# 2. SSRN on a short mel (12 frames -> 48 x 1025)
Ys = np.random.default_rng(12).uniform(0, 1, (1, 12, hp.n_mels)).astype(np.float32)
zl, z = tf_shim.run_ssrn(Ys)
np.savez_compressed(os.path.join(HERE, "refshim_ssrn_T12.npz"), Z=z, Z_logits_sub=zl[:, :, ::8])
missing = sorted(set(P) - store.requested)
assert not missing, "variables the reference graph never asked for: %s" % missing[:5]

# 3. the synthesize loop (synthesize.py:45-57) on Harvard sentence #1: 210 full-graph passes, then SSRN
Lh = load_data("synthesize", os.path.join(ROOT, "harvard_sentences.txt"))[:1]
t0 = time.time()
r = tf_shim.synthesize(Lh)
np.savez_compressed(os.path.join(HERE, "refshim_synth_harvard1.npz"), L=Lh, Y=r["Y"], p_hist=r["p_hist"],
                    max_attentions=r["max_attentions"], Z_sub=r["Z"][:, ::8, ::8])
print("reference-under-shim fixtures written to %s (%.0f s for the loop)" % (HERE, time.time() - t0))
