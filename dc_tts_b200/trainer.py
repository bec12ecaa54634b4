"""The trainer loop around `Engine.train_step` / `train_step_ssrn` -- what `python train.py 1|2` does in the reference
(/root/reference/train.py:137-160: one optimiser step per batch, a checkpoint `model_gs_{NNN}k` every 1000 steps in
`hp.logdir + "-" + num`, stop after hp.num_iterations) on top of the LJ transcript parser of data_load.py:41-56 and the
pre-computed `mels/*.npy`, `mags/*.npy` of prepo.py (data_load.py:104-112).

Deliberate difference: the reference feeds length-bucketed, dynamically padded batches from a TF queue
(data_load.py:120-129); the CUDA training step takes FIXED-size batches (B, max_N) / (B, max_T, n_mels) /
(B, 4 max_T, F) as BASELINE config 5 specifies, so every utterance is zero-padded to hp.max_N / hp.max_T and utterances
longer than that are skipped.  The reference's losses already average over the padding inside a bucket (train.py:85-88
have no mask); padding to the global maximum changes that weighting, not the definition.
"""
import codecs
import os

import numpy as np

from .data_load import load_vocab, text_normalize
from .hyperparams import Hyperparams as hp


def load_train_data(data_dir=None):
    """data_load.py:41-56 (LJ Speech): transcript.csv -> (wav paths, text lengths, int32 id arrays ending in E)."""
    data_dir = data_dir or hp.data
    char2idx, _ = load_vocab()
    fpaths, text_lengths, texts = [], [], []
    for line in codecs.open(os.path.join(data_dir, "transcript.csv"), "r", "utf-8").readlines():
        fname, _, text = line.strip().split("|")
        ids = np.array([char2idx[ch] for ch in text_normalize(text) + "E"], np.int32)
        fpaths.append(os.path.join(data_dir, "wavs", fname + ".wav"))
        text_lengths.append(len(ids))
        texts.append(ids)
    return fpaths, text_lengths, texts


def _load_spectrograms_npy(fpath, mels_dir="mels", mags_dir="mags"):
    """data_load.py:105-109: what prepo.py wrote for this wav."""
    fname = os.path.basename(fpath)
    return fname, np.load(os.path.join(mels_dir, fname.replace("wav", "npy"))), np.load(os.path.join(mags_dir, fname.replace("wav", "npy")))


def fixed_size_batches(fpaths, texts, B=None, seed=0, loader=_load_spectrograms_npy, epochs=None):
    """Shuffled batches of exactly B utterances, zero-padded to (B, max_N), (B, max_T, n_mels), (B, 4 max_T, F).
    Utterances that do not fit are skipped; an incomplete last batch of an epoch is dropped (num_batch = len // B,
    data_load.py:97)."""
    B = B or hp.B
    F = 1 + hp.n_fft // 2
    rng = np.random.default_rng(seed)
    epoch = 0
    while epochs is None or epoch < epochs:
        order = rng.permutation(len(fpaths))
        L = np.zeros((B, hp.max_N), np.int32)
        mels = np.zeros((B, hp.max_T, hp.n_mels), np.float32)
        mags = np.zeros((B, hp.max_T * hp.r, F), np.float32)
        names, n = [], 0
        for i in order:
            text = texts[i]
            if len(text) > hp.max_N:
                continue
            fname, mel, mag = loader(fpaths[i])
            if mel.shape[0] > hp.max_T or mag.shape[0] > hp.max_T * hp.r:
                continue
            L[n, :len(text)] = text
            mels[n, :mel.shape[0]] = mel
            mags[n, :mag.shape[0]] = mag
            names.append(fname)
            n += 1
            if n == B:
                yield L, mels, mags, names
                L = np.zeros_like(L); mels = np.zeros_like(mels); mags = np.zeros_like(mags)
                names, n = [], 0
        epoch += 1


def checkpoint_name(logdir, gs):
    """train.py:152."""
    return os.path.join(logdir, "model_gs_{}".format(str(gs // 1000).zfill(3) + "k"))


def train(num, engine, batches, num_iterations=None, logdir=None, global_step=0, save_every=1000, log=print):
    """train.py:137-160 for num = 1 (Text2Mel) or 2 (SSRN).  `batches` yields (L, mels, mags, names); `engine` is an
    `Engine` with parameters loaded.  Returns the final global step."""
    if num not in (1, 2):
        raise ValueError("num: 1 for Text2Mel, 2 for SSRN (train.py:139)")
    num_iterations = hp.num_iterations if num_iterations is None else num_iterations
    logdir = logdir or (hp.logdir + "-" + str(num))
    os.makedirs(logdir, exist_ok=True)
    gs = int(global_step)
    initialised = False
    for L, mels, mags, _names in batches:
        if not initialised:
            if num == 1:
                engine.train_init(len(L))
            else:
                engine.train_init_ssrn(len(L), mels.shape[1])
            initialised = True
        if num == 1:
            losses = engine.train_step(L, mels, global_step=gs, seed=gs)
        else:
            losses = engine.train_step_ssrn(mels, mags, global_step=gs, seed=gs)
        gs += 1                                   # apply_gradients(..., global_step=...) increments (train.py:131)
        if gs % save_every == 0:                  # train.py:151-152
            engine.save_checkpoint(checkpoint_name(logdir, gs), gs, "Text2Mel" if num == 1 else "SSRN")
            log("step %d  %s" % (gs, "  ".join("%s %.4f" % kv for kv in sorted(losses.items()))))
        if gs > num_iterations:                   # train.py:160
            break
    return gs
