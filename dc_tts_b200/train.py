"""`Graph(mode="synthesize")` -- the object synthesize.py talks to
(/root/reference/train.py:21-80, synthesize branch :43-46, :48-68, :74-80).

The reference builds a symbolic TF graph and evaluates it with `sess.run(fetches,
feed_dict)`.  Here the attributes (`L, mels, prev_max_attentions, S, K, V, Q, R,
alignments, max_attentions, Y_logits, Y, Z_logits, Z, global_step`) are lightweight
symbols and `Session.run` evaluates them eagerly on the GPU, so that the reference's
loop (synthesize.py:47-57) runs unmodified in structure.  `Graph.generate` is the fast
path: the whole loop on the device, replayed from a CUDA graph.

The training branch is not a symbolic graph here: losses, backward pass and optimiser (train.py:82-135) are
`Engine.train_step` / `train_step_ssrn`, and the loop of train.py:137-160 is `dc_tts_b200/trainer.py: train`;
`Graph(mode="train")` raises NotImplementedError and says so.
"""
import numpy as np
import torch

from .data_load import load_vocab
from .engine import get_engine
from .hyperparams import Hyperparams as hp
from .modules import variable_scope
from .networks import Attention, AudioDec, AudioEnc, SSRN, TextEnc


class Symbol:
    """Placeholder / fetchable node of the synthesize graph."""

    def __init__(self, graph, name):
        self.graph, self.name = graph, name

    def __repr__(self):
        return "<dc_tts_b200 graph tensor %s>" % self.name


_TEXT2MEL = ("S", "K", "V", "Q", "R", "alignments", "max_attentions", "Y_logits", "Y")
_FUSED_OK = {"Y", "max_attentions", "alignments", "global_step"}


class Graph:
    def __init__(self, num=1, mode="train", engine=None, fused=True):
        if mode != "synthesize":
            raise NotImplementedError("Graph(mode='train'): use dc_tts_b200.trainer.train (Engine.train_step / train_step_ssrn)")
        self.char2idx, self.idx2char = load_vocab()
        self.engine = engine or get_engine()
        self.fused = fused
        self.global_step_value = 0          # `gs/global_step` (train.py:79-80); no checkpoint offline
        for name in ("L", "mels", "prev_max_attentions") + _TEXT2MEL + ("Z_logits", "Z", "global_step"):
            setattr(self, name, Symbol(self, name))

    # ---------------------------------------------------------------- evaluation
    def _text2mel(self, L, mels, pma, want):
        """train.py:48-68 evaluated block by block through networks.py."""
        e = self.engine
        mels = e._f32(mels)
        vals = {}
        with variable_scope("Text2Mel"):
            vals["S"] = torch.cat((torch.zeros_like(mels[:, :1, :]), mels[:, :-1, :]), 1)     # train.py:51
            with variable_scope("TextEnc"):
                vals["K"], vals["V"] = TextEnc(L, training=False, fused=self.fused)
            with variable_scope("AudioEnc"):
                vals["Q"] = AudioEnc(vals["S"], training=False, fused=self.fused)
            with variable_scope("Attention"):
                vals["R"], vals["alignments"], vals["max_attentions"] = Attention(
                    vals["Q"], vals["K"], vals["V"], mononotic_attention=True, prev_max_attentions=pma)
            with variable_scope("AudioDec"):
                vals["Y_logits"], vals["Y"] = AudioDec(vals["R"], training=False, fused=self.fused)
        return vals

    def run(self, fetches, feed_dict=None, as_numpy=True):
        """`sess.run` equivalent.  Feeding `self.Y` cuts Text2Mel out of the evaluation,
        exactly as feeding g.Y does in the reference (synthesize.py:57)."""
        single = isinstance(fetches, Symbol)
        names = [fetches.name] if single else [f.name for f in fetches]
        feed = {k.name: v for k, v in (feed_dict or {}).items()}
        vals = dict(global_step=np.int64(self.global_step_value))
        need_t2m = any(n in _TEXT2MEL for n in names if n not in feed) or \
            (any(n in ("Z", "Z_logits") for n in names) and "Y" not in feed)
        if need_t2m:
            for k in ("L", "mels", "prev_max_attentions"):
                if k not in feed:
                    raise ValueError("placeholder %s must be fed" % k)
            if self.fused and set(names) <= _FUSED_OK:
                Y, M, A = self.engine.text2mel_forward(feed["L"], feed["mels"], feed["prev_max_attentions"],
                                                       want_alignments="alignments" in names)
                vals.update(Y=Y, max_attentions=M, alignments=A)
            else:
                vals.update(self._text2mel(feed["L"], feed["mels"], feed["prev_max_attentions"], names))
        vals.update(feed)
        if any(n in ("Z", "Z_logits") for n in names):
            with variable_scope("SSRN"):                                                      # train.py:74-77
                vals["Z_logits"], vals["Z"] = SSRN(vals["Y"], training=False, fused=self.fused)
        out = []
        for n in names:
            v = vals[n]
            if as_numpy and isinstance(v, torch.Tensor):
                v = v.cpu().numpy()
            out.append(v)
        return out[0] if single else out

    # ---------------------------------------------------------------- fast path
    def generate(self, L, steps=0):
        """synthesize.py:45-54 entirely on the device: returns the mel tensor Y
        (B, max_T, n_mels) as a CUDA tensor plus the prev_max_attentions history."""
        Y, P, _, _ = self.engine.text2mel_generate(L, steps)
        return Y, P


class Session:
    """Minimal stand-in for tf.Session used as `with Session() as sess: sess.run(...)`."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def run(self, fetches, feed_dict=None):
        g = (fetches if isinstance(fetches, Symbol) else fetches[0]).graph
        return g.run(fetches, feed_dict)
