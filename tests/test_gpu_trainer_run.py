"""The trainer loop END TO END on the GPU (it had only been driven with a fake engine): 1,001 real
optimiser steps of each trainer (train.py:137-160) on a small synthetic data set written as prepo.py would write it
(mels/*.npy, mags/*.npy), the model_gs_001k bundle it writes restored into a fresh engine (synthesize.py:31-41) and into a
resumed run (tf.train.Supervisor semantics)."""
import os

import numpy as np
import pytest
import torch

from dc_tts_b200 import trainer
from dc_tts_b200.hyperparams import Hyperparams as hp

pytestmark = pytest.mark.gpu


def _write_dataset(root, n=24, seed=0):
    rng = np.random.default_rng(seed)
    d = root / "LJSpeech-1.0"
    (d / "wavs").mkdir(parents=True)
    (root / "mels").mkdir(); (root / "mags").mkdir()
    lines = []
    F = 1 + hp.n_fft // 2
    for i in range(n):
        nchar = int(rng.integers(20, 90))
        text = "".join(rng.choice(list("abcdefghijklmnopqrstuvwxyz '"), nchar))
        lines.append("LJ%03d|raw|%s" % (i, text))
        T = int(rng.integers(60, hp.max_T))
        # smooth positive "spectrograms" in [0, 1], low rank so that a few hundred steps visibly fit them
        base = 0.5 + 0.4 * np.sin(np.linspace(0, 3 + i % 5, T))[:, None] * np.cos(np.linspace(0, 2, hp.n_mels))[None, :]
        mel = np.clip(base + 0.02 * rng.standard_normal((T, hp.n_mels)), 0, 1).astype(np.float32)
        mag = np.clip(np.repeat(base[:, :1], hp.r, 0) * np.linspace(1, 0.2, F)[None, :] + 0.02 * rng.standard_normal((T * hp.r, F)), 0, 1).astype(np.float32)
        np.save(root / "mels" / ("LJ%03d.npy" % i), mel); np.save(root / "mags" / ("LJ%03d.npy" % i), mag)
    (d / "transcript.csv").write_text("\n".join(lines) + "\n", encoding="utf-8")
    return str(d)


@pytest.mark.parametrize("num", [1, 2])
def test_trainer_1001_steps_checkpoint_and_resume(tmp_path, num):
    from dc_tts_b200.engine import Engine
    from dc_tts_b200.params import init_params, synthetic_text
    d = _write_dataset(tmp_path)
    fpaths, lens, texts = trainer.load_train_data(d)
    loader = lambda p: trainer._load_spectrograms_npy(p, str(tmp_path / "mels"), str(tmp_path / "mags"))
    B = 4
    P = init_params(1)
    eng = Engine(0)
    eng.load_params(P)
    logdir = str(tmp_path / ("logdir/LJ01-%d" % num))
    losses = []
    gs = trainer.train(num, eng, trainer.fixed_size_batches(fpaths, texts, B=B, seed=0, loader=loader), num_iterations=1000,
                       logdir=logdir, save_every=1000, log=lambda s: losses.append(s))
    assert gs == 1001
    from dc_tts_b200.checkpoint import latest_checkpoint, load_checkpoint
    ck = latest_checkpoint(logdir)
    assert ck is not None and ck.endswith("model_gs_001k")
    meta = load_checkpoint(ck, ["gs/global_step"])
    assert int(meta["gs/global_step"]) == 1000
    # the loss logged at step 1000 is far below the loss of the untrained network
    fresh0 = Engine(0); fresh0.load_params(P)
    L0, m0, g0, _ = next(trainer.fixed_size_batches(fpaths, texts, B=B, seed=0, loader=loader))
    if num == 1:
        fresh0.train_init(B, 0.0); first = fresh0.train_step(L0, m0, global_step=0, seed=0, apply=False)["loss"]
        eng2 = eng; last = eng2.train_step(L0, m0, global_step=1001, seed=0, apply=False)["loss"]
    else:
        fresh0.train_init_ssrn(B, hp.max_T, 0.0); first = fresh0.train_step_ssrn(m0, g0, global_step=0, seed=0, apply=False)["loss"]
        last = eng.train_step_ssrn(m0, g0, global_step=1001, seed=0, apply=False)["loss"]
    assert np.isfinite(last) and last < 0.9 * first, (first, last)
    fresh0.close()
    # resume: a second call continues at 1000 and stops after one more step
    more = []
    gs2 = trainer.train(num, eng, trainer.fixed_size_batches(fpaths, texts, B=B, seed=1, loader=loader), num_iterations=1000,
                        logdir=logdir, save_every=1000, log=lambda s: more.append(s))
    assert gs2 == 1001 and any("resumed" in s for s in more)
    # the bundle restores into a fresh engine through the synthesis path's reader (the other network from a plain bundle)
    from dc_tts_b200 import checkpoint as ckpt
    other = "SSRN" if num == 1 else "Text2Mel"
    odir = str(tmp_path / ("logdir/LJ01-%d" % (3 - num)))
    ckpt.save_checkpoint(odir + "/model_gs_000k", {k: v for k, v in P.items() if k.startswith(other + "/")})
    syn = Engine(0)
    syn.restore(logdir if num == 1 else odir, odir if num == 1 else logdir)
    Y, _, _, _ = syn.text2mel_generate(synthetic_text(2, 50, seed=3), steps=30)
    _, Z = syn.ssrn(Y, want_logits=False)
    assert torch.isfinite(Y).all() and torch.isfinite(Z).all()
    name = "Text2Mel/AudioDec/C_11/conv1d/bias" if num == 1 else "SSRN/C_16/conv1d/bias"
    assert np.abs(eng.train_tensor(name, "param") - P[name]).max() > 1e-4          # the trained weights moved ...
    syn.close(); eng.close()
