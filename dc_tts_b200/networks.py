"""The five networks with the reference's signatures (/root/reference/networks.py).

`TextEnc(L, training)` :14, `AudioEnc(S, training)` :73,
`Attention(Q, K, V, mononotic_attention, prev_max_attentions)` :126 (the misspelt keyword
is part of the API), `AudioDec(R, training)` :157, `SSRN(Y, training)` :214.

Two execution modes, selected by `fused`:
  fused=True  (default) one C-ABI call per network (dctts_textenc, ...);
  fused=False           block-by-block through modules.py, following arch.py's tables --
                        the same composition the reference builds, used to cross-check
                        the library's internal layer tables.
Both must be called under the same variable scopes the reference uses
(train.py:49-76), e.g. `with variable_scope("Text2Mel"), variable_scope("TextEnc")`.
"""
import torch

from . import arch
from .engine import get_engine
from .hyperparams import Hyperparams as hp
from .modules import conv1d, conv1d_transpose, embed, hc, nn


def _chain(tensor, layers, training):
    for l in layers:
        if l.kind == "C":
            tensor = conv1d(tensor, filters=l.cout, size=l.size, rate=l.rate, padding=l.pad,
                            dropout_rate=hp.dropout_rate, activation_fn=(nn.relu if l.act == "relu" else None),
                            training=training, scope=l.scope)
        elif l.kind == "HC":
            tensor = hc(tensor, size=l.size, rate=l.rate, padding=l.pad, dropout_rate=hp.dropout_rate,
                        activation_fn=None, training=training, scope=l.scope)
        else:
            tensor = conv1d_transpose(tensor, scope=l.scope, dropout_rate=hp.dropout_rate, training=training)
    return tensor


def _inference_only(training):
    if training:
        raise NotImplementedError("only training=False (synthesis) is on the hot path (SURVEY.md 8)")


def TextEnc(L, training=True, fused=True):
    """L (B,N) int32 -> K, V (B,N,d).  networks.py:14-71."""
    _inference_only(training)
    if fused:
        return get_engine().textenc(L)
    tensor = embed(L, vocab_size=len(hp.vocab), num_units=hp.e, scope="embed_1")
    tensor = _chain(tensor, arch.textenc_layers(), training)
    K, V = torch.split(tensor, hp.d, dim=-1)
    return K.contiguous(), V.contiguous()


def AudioEnc(S, training=True, fused=True):
    """S (B,T/r,n_mels) -> Q (B,T/r,d).  networks.py:73-124."""
    _inference_only(training)
    if fused:
        return get_engine().audioenc(S)
    return _chain(get_engine()._f32(S), arch.audioenc_layers(), training)


def Attention(Q, K, V, mononotic_attention=False, prev_max_attentions=None):
    """-> R (B,T/r,2d), alignments (B,N,T/r), max_attentions (B,T/r).  networks.py:126-155."""
    return get_engine().attention(Q, K, V, mononotic_attention, prev_max_attentions)


def AudioDec(R, training=True, fused=True):
    """R (B,T/r,2d) -> logits, Y (B,T/r,n_mels).  networks.py:157-212."""
    _inference_only(training)
    if fused:
        return get_engine().audiodec(R)
    logits = _chain(get_engine()._f32(R), arch.audiodec_layers(), training)
    return logits, torch.sigmoid(logits)


def SSRN(Y, training=True, fused=True):
    """Y (B,T/r,n_mels) -> logits, Z (B,T,1+n_fft/2).  networks.py:214-292."""
    _inference_only(training)
    if fused:
        return get_engine().ssrn(Y)
    logits = _chain(get_engine()._f32(Y), arch.ssrn_layers(), training)
    return logits, torch.sigmoid(logits)
