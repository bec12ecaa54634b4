"""ORACLE (test infrastructure, not product) -- torch-autograd restatement of ONE Text2Mel training step of the
reference (/root/reference/train.py:43-68 graph in mode="train", num=1; losses :83-99; optimiser :122-132;
learning-rate schedule utils.py:141-145; guided-attention weights utils.py:134-140), BASELINE config 5.

What the reference does per step: forward with dropout (rate hp.dropout_rate after every block, modules.py:139,
195,245) and FULL softmax attention (no window), loss = mean|Y - mels| + mean BCE(Y_logits, mels) +
sum|A * W_guided| / (B N T); gradients of all 23.97 M Text2Mel variables, clipped elementwise to [-1, 1];
Adam (TF defaults beta1 0.9, beta2 0.999, eps 1e-8, bias correction folded into the step size) with the Noam
learning rate.

Dropout: TF's random stream cannot be reproduced, so oracle and CUDA path share a stateless hash mask
(`dropout_keep`): element i of the output of the l-th block (blocks counted in graph order: TextEnc 0..13,
AudioEnc 14..26, AudioDec 27..37) is kept iff mix32(i, l, seed) >= rate * 2^32; kept values are scaled by
1 / (1 - rate) like tf.layers.dropout.  tests/test_train.py runs the reference's own train graph under the TF
stand-in with this mask plugged into `tf.layers.dropout` and compares the three losses.
PARITY: wiring and loss definitions pinned by the reference's code under the stand-in; op numerics unpinned.
"""
import numpy as np
import torch

from dc_tts_b200 import arch
from dc_tts_b200.hyperparams import Hyperparams as hp

from . import ref_torch as rt


def mix32(idx, layer, seed):
    """uint32 hash of (element index, block index, seed) -- same arithmetic as csrc/kernels_train.cu."""
    x = (np.asarray(idx, np.uint64) * np.uint64(0x9E3779B1)) & np.uint64(0xffffffff)
    x ^= np.uint64((int(layer) * 0x85EBCA77 + int(seed)) & 0xffffffff)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xffffffff)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xffffffff)
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def dropout_keep(shape, layer, seed, rate):
    """float32 multiplier tensor: 0 where dropped, 1/(1-rate) where kept."""
    n = int(np.prod(shape))
    if rate <= 0:
        return np.ones(shape, np.float32)
    thresh = np.uint32(min(int(rate * 4294967296.0), 0xffffffff))
    keep = mix32(np.arange(n, dtype=np.uint64), layer, seed) >= thresh
    return (keep.astype(np.float32) * np.float32(1.0 / (1.0 - rate))).reshape(shape)


def guided_attention(g=0.2):
    """utils.py:134-140 -> (max_N, max_T) float32."""
    n = np.arange(hp.max_N, dtype=np.float64)[:, None] / float(hp.max_N)
    t = np.arange(hp.max_T, dtype=np.float64)[None, :] / float(hp.max_T)
    return (1 - np.exp(-(t - n) ** 2 / (2 * g * g))).astype(np.float32)


def learning_rate(global_step, init_lr=None, warmup_steps=4000.0):
    """utils.py:141-145 (Noam): global_step is the value BEFORE this step's increment."""
    init_lr = hp.lr if init_lr is None else init_lr
    step = float(global_step + 1)
    return init_lr * warmup_steps ** 0.5 * min(step * warmup_steps ** -1.5, step ** -0.5)


def _chain(P, x, net, layers, counter, seed, rate):
    for l in layers:
        scope = "%s/%s" % (net, l.scope)
        if l.kind == "C":
            x = rt.conv1d(P, x, scope, l.rate, l.pad, l.act)
        elif l.kind == "HC":
            x = rt.hc(P, x, scope, l.rate, l.pad)
        else:
            x = rt.conv1d_transpose(P, x, scope)
        if rate > 0:
            x = x * torch.from_numpy(dropout_keep(tuple(x.shape), counter[0], seed, rate))
        counter[0] += 1
    return x


def forward(P, L, mels, seed=0, rate=None):
    """train.py:48-68 + :83-99 in training mode.  P: name -> tensor (requires_grad for gradients)."""
    rate = hp.dropout_rate if rate is None else rate
    mels = torch.as_tensor(mels, dtype=torch.float32)
    S = torch.cat((torch.zeros_like(mels[:, :1, :]), mels[:, :-1, :]), 1)
    c = [0]
    x = rt.embed(P, torch.as_tensor(L), "Text2Mel/TextEnc/embed_1").to(torch.float32)
    x = _chain(P, x, "Text2Mel/TextEnc", arch.textenc_layers(), c, seed, rate)
    K, V = torch.chunk(x, 2, dim=-1)
    Q = _chain(P, S, "Text2Mel/AudioEnc", arch.audioenc_layers(), c, seed, rate)
    R, alignments, _ = rt.Attention(Q, K, V, False, None)
    logits = _chain(P, R, "Text2Mel/AudioDec", arch.audiodec_layers(), c, seed, rate)
    Y = torch.sigmoid(logits)
    loss_mels = (Y - mels).abs().mean()
    loss_bd1 = torch.nn.functional.binary_cross_entropy_with_logits(logits, mels)
    gts = torch.from_numpy(guided_attention())
    A = alignments[:, :hp.max_N, :hp.max_T]          # fixed-size batches: the -1 padding of train.py:91 is empty
    loss_att = (A * gts).abs().sum() / float(A.numel())
    return dict(loss=loss_mels + loss_bd1 + loss_att, loss_mels=loss_mels, loss_bd1=loss_bd1, loss_att=loss_att,
                Y=Y, logits=logits, alignments=alignments, Q=Q, K=K, V=V, R=R)


def text2mel_names():
    return [n for n in arch.param_shapes() if n.startswith("Text2Mel/")]


def train_step(P, L, mels, state=None, global_step=0, seed=0, rate=None, lr=None,
               beta1=0.9, beta2=0.999, eps=1e-8):
    """One optimiser step (train.py:122-132).  Returns (new params dict (numpy), state, info) where info holds
    the losses and the clipped gradients."""
    names = text2mel_names()
    T = {n: torch.tensor(np.asarray(P[n], np.float32), requires_grad=True) for n in names}
    out = forward(T, L, mels, seed, rate)
    out["loss"].backward()
    lr_now = learning_rate(global_step, lr)
    t = global_step + 1
    lr_t = lr_now * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    state = state or {n: (np.zeros(T[n].shape, np.float32), np.zeros(T[n].shape, np.float32)) for n in names}
    newP, grads, newstate = dict(P), {}, {}
    for n in names:
        g = T[n].grad.numpy() if T[n].grad is not None else np.zeros(T[n].shape, np.float32)
        g = np.clip(g, -1.0, 1.0).astype(np.float32)
        m, v = state[n]
        m = (beta1 * m + (1 - beta1) * g).astype(np.float32)
        v = (beta2 * v + (1 - beta2) * g * g).astype(np.float32)
        newP[n] = (np.asarray(P[n], np.float32) - np.float32(lr_t) * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
        grads[n] = g
        newstate[n] = (m, v)
    info = {k: float(out[k].detach()) for k in ("loss", "loss_mels", "loss_bd1", "loss_att")}
    info["grads"] = grads
    info["lr"] = lr_now
    return newP, newstate, info


# ------------------------------------------------------------------------------------------ SSRN (num = 2)
def ssrn_names():
    return [n for n in arch.param_shapes() if n.startswith("SSRN/")]


def forward_ssrn(P, mels, mags, seed=0, rate=None):
    """train.py:69-72 + :100-108 in training mode: SSRN on the GROUND-TRUTH mels, L1 + binary divergence on the
    linear magnitudes.  Blocks are counted 0..15 for the dropout mask (the num=2 graph holds nothing else)."""
    rate = hp.dropout_rate if rate is None else rate
    mels = torch.as_tensor(mels, dtype=torch.float32)
    mags = torch.as_tensor(mags, dtype=torch.float32)
    logits = _chain(P, mels, "SSRN", arch.ssrn_layers(), [0], seed, rate)
    Z = torch.sigmoid(logits)
    loss_mags = (Z - mags).abs().mean()
    loss_bd2 = torch.nn.functional.binary_cross_entropy_with_logits(logits, mags)
    return dict(loss=loss_mags + loss_bd2, loss_mags=loss_mags, loss_bd2=loss_bd2, Z=Z, logits=logits)


def _adam(P, names, T, state, global_step, lr, beta1, beta2, eps):
    lr_now = learning_rate(global_step, lr)
    t = global_step + 1
    lr_t = lr_now * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    state = state or {n: (np.zeros(T[n].shape, np.float32), np.zeros(T[n].shape, np.float32)) for n in names}
    newP, grads, newstate = dict(P), {}, {}
    for n in names:
        g = T[n].grad.numpy() if T[n].grad is not None else np.zeros(T[n].shape, np.float32)
        g = np.clip(g, -1.0, 1.0).astype(np.float32)
        m, v = state[n]
        m = (beta1 * m + (1 - beta1) * g).astype(np.float32)
        v = (beta2 * v + (1 - beta2) * g * g).astype(np.float32)
        newP[n] = (np.asarray(P[n], np.float32) - np.float32(lr_t) * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
        grads[n] = g
        newstate[n] = (m, v)
    return newP, newstate, grads, lr_now


def train_step_ssrn(P, mels, mags, state=None, global_step=0, seed=0, rate=None, lr=None, beta1=0.9, beta2=0.999, eps=1e-8):
    """One SSRN optimiser step (train.py num=2: losses :100-108, optimiser :122-132)."""
    names = ssrn_names()
    T = {n: torch.tensor(np.asarray(P[n], np.float32), requires_grad=True) for n in names}
    out = forward_ssrn(T, mels, mags, seed, rate)
    out["loss"].backward()
    newP, newstate, grads, lr_now = _adam(P, names, T, state, global_step, lr, beta1, beta2, eps)
    info = {k: float(out[k].detach()) for k in ("loss", "loss_mags", "loss_bd2")}
    info["grads"] = grads
    info["lr"] = lr_now
    return newP, newstate, info
