"""ORACLE (test infrastructure, not product) -- torch-CPU restatement of the reference
DC-TTS synthesis path.

PARITY UNPINNED: the reference (/root/reference) ships no golden vectors, no tests and
no checkpoint, and its TF1 code cannot run here (TensorFlow absent, `tf.contrib` needs
TF 1.x / Python <= 3.7).  This file is therefore a literal restatement of the
reference's algorithm, with the TF op semantics of SURVEY.md App. B, cross-checked
against the independent numpy/fp64 restatement in `ref_numpy.py` (tests/test_oracle.py).
PINNED since: the reference's own modules.py / networks.py / train.py executed under the TF API stand-in
tests/golden/tf_shim.py reproduce this file to <= 2.4e-6 with the same window trajectory and variable schema
(tests/test_reference_shim.py, fixtures tests/golden/refshim_*.npz) -- that pins the wiring; the numerics of
the TF ops themselves remain restated, not observed.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import it; the product path never does.

Each function cites the reference lines it follows.  Layout is channels-last
(B, time, C) as in the reference.  `P` is the name -> array parameter dictionary
(dc_tts_b200.params), names per SURVEY.md App. C.
"""
import numpy as np
import torch
import torch.nn.functional as F

from dc_tts_b200.hyperparams import Hyperparams as hp

LN_EPS = 1e-12                       # tf.contrib.layers.layer_norm variance_epsilon
MASK_VALUE = float(-2 ** 32 + 1)     # networks.py:146 (python precedence: -(2**32)+1)


def _t(P, name, dtype):
    v = P[name]
    return v.to(dtype) if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)).to(dtype)


# --------------------------------------------------------------------------- modules.py
def embed(P, inputs, scope):
    """modules.py:13-42: lookup with row 0 forced to zeros (zero_pad=True)."""
    table = _t(P, scope + "/lookup_table", torch.float32)
    table = torch.cat((torch.zeros_like(table[:1]), table[1:]), 0)
    return table[inputs.long()]


def normalize(x, gamma, beta):
    """modules.py:45-64 -> tf.contrib.layers.layer_norm(begin_norm_axis=-1):
    biased variance, eps 1e-12, gamma/beta over the last axis."""
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, LN_EPS)


def _conv(x, W, b, rate, padding):
    """tf.layers.conv1d (modules.py:134,187): cross-correlation, kernel [k,Cin,Cout].
    CAUSAL = left pad (k-1)*rate then VALID (modules.py:121-125); SAME with stride 1
    pads (k-1)*rate in total, floor half on the left."""
    k = W.shape[0]
    tot = (k - 1) * rate
    left = tot if padding.lower() == "causal" else tot // 2
    xp = F.pad(x.transpose(1, 2), (left, tot - left))
    y = F.conv1d(xp, W.permute(2, 1, 0).contiguous(), b, dilation=rate)
    return y.transpose(1, 2)


def conv1d(P, x, scope, rate=1, padding="SAME", activation_fn=None):
    """modules.py:91-141 at training=False (dropout is the identity, :139)."""
    dt = x.dtype
    y = _conv(x, _t(P, scope + "/conv1d/kernel", dt), _t(P, scope + "/conv1d/bias", dt), rate, padding)
    y = normalize(y, _t(P, scope + "/normalize/gamma", dt), _t(P, scope + "/normalize/beta", dt))
    if activation_fn == "relu":
        y = torch.relu(y)
    return y


def hc(P, x, scope, rate=1, padding="SAME"):
    """modules.py:143-197: conv to 2C, split, LN each half, sigmoid gate on the first
    half, highway mix with the *unpadded* input (:193).  Info branch is linear."""
    dt = x.dtype
    y = _conv(x, _t(P, scope + "/conv1d/kernel", dt), _t(P, scope + "/conv1d/bias", dt), rate, padding)
    H1, H2 = torch.chunk(y, 2, dim=-1)
    H1 = normalize(H1, _t(P, scope + "/H1/gamma", dt), _t(P, scope + "/H1/beta", dt))
    H2 = normalize(H2, _t(P, scope + "/H2/gamma", dt), _t(P, scope + "/H2/beta", dt))
    H1 = torch.sigmoid(H1)
    return H1 * H2 + (1. - H1) * x


def conv1d_transpose(P, x, scope):
    """modules.py:199-247: conv2d_transpose(kernel (1,3), strides (1,2), 'same') on the
    expanded tensor, then LN, no activation.  Kernel variable [1,3,Cout,Cin];
    out[2t] = W0 x[t] + W2 x[t-1], out[2t+1] = W1 x[t]   which is
    torch conv_transpose1d(stride 2, padding 0) truncated to 2L outputs."""
    dt = x.dtype
    W = _t(P, scope + "/conv2d_transpose/kernel", dt)[0]          # (3, Cout, Cin)
    b = _t(P, scope + "/conv2d_transpose/bias", dt)
    L = x.shape[1]
    y = F.conv_transpose1d(x.transpose(1, 2), W.permute(2, 1, 0).contiguous(), b, stride=2)
    y = y[..., :2 * L].transpose(1, 2)
    return normalize(y, _t(P, scope + "/normalize/gamma", dt), _t(P, scope + "/normalize/beta", dt))


# --------------------------------------------------------------------------- networks.py
def _run_chain(P, x, net, layers):
    for l in layers:
        scope = "%s/%s" % (net, l.scope)
        if l.kind == "C":
            x = conv1d(P, x, scope, l.rate, l.pad, l.act)
        elif l.kind == "HC":
            x = hc(P, x, scope, l.rate, l.pad)
        else:
            x = conv1d_transpose(P, x, scope)
    return x


def TextEnc(P, L, dtype=torch.float32):
    """networks.py:14-71."""
    from dc_tts_b200.arch import textenc_layers
    x = embed(P, torch.as_tensor(L), "Text2Mel/TextEnc/embed_1").to(dtype)
    x = _run_chain(P, x, "Text2Mel/TextEnc", textenc_layers())
    K, V = torch.chunk(x, 2, dim=-1)
    return K, V


def AudioEnc(P, S):
    """networks.py:73-124."""
    from dc_tts_b200.arch import audioenc_layers
    return _run_chain(P, S, "Text2Mel/AudioEnc", audioenc_layers())


def Attention(Q, K, V, mononotic_attention=False, prev_max_attentions=None):
    """networks.py:126-155.  The window mask (:141-147) keeps keys p <= n < p+3 and is
    tiled over every query row."""
    A = torch.matmul(Q, K.transpose(1, 2)) * (1.0 / np.sqrt(float(hp.d)))
    if mononotic_attention:
        p = torch.as_tensor(prev_max_attentions).long()
        n = torch.arange(hp.max_N)
        key_masks = n[None, :] < p[:, None]                                   # sequence_mask(p, N)
        rev = (n[None, :] < (hp.max_N - hp.attention_win_size - p)[:, None]).flip(1)
        masks = (key_masks | rev)[:, None, :].expand(-1, A.shape[1], -1)
        A = torch.where(masks, torch.full_like(A, MASK_VALUE), A)
    A = torch.softmax(A, dim=-1)
    max_attentions = torch.argmax(A, dim=-1)
    R = torch.matmul(A, V)
    R = torch.cat((R, Q), -1)
    alignments = A.transpose(1, 2)
    return R, alignments, max_attentions


def AudioDec(P, R):
    """networks.py:157-212: returns (logits, sigmoid(logits))."""
    from dc_tts_b200.arch import audiodec_layers
    logits = _run_chain(P, R, "Text2Mel/AudioDec", audiodec_layers())
    return logits, torch.sigmoid(logits)


def SSRN(P, Y):
    """networks.py:214-292: returns (logits, sigmoid(logits))."""
    from dc_tts_b200.arch import ssrn_layers
    logits = _run_chain(P, Y, "SSRN", ssrn_layers())
    return logits, torch.sigmoid(logits)


# --------------------------------------------------------------------------- train.py Graph (synthesize)
def text2mel_forward(P, L, mels, prev_max_attentions, KV=None):
    """One `sess.run` of the synthesize graph (train.py:48-68): shift (:51), TextEnc,
    AudioEnc, Attention(monotonic), AudioDec over ALL rows."""
    mels = torch.as_tensor(mels)
    S = torch.cat((torch.zeros_like(mels[:, :1, :]), mels[:, :-1, :]), 1)
    K, V = KV if KV is not None else TextEnc(P, L, mels.dtype)
    Q = AudioEnc(P, S)
    R, alignments, max_attentions = Attention(Q, K, V, True, prev_max_attentions)
    Y_logits, Y = AudioDec(P, R)
    return dict(K=K, V=V, Q=Q, R=R, alignments=alignments, max_attentions=max_attentions,
                Y_logits=Y_logits, Y=Y, S=S)


# --------------------------------------------------------------------------- synthesize.py loop
@torch.no_grad()
def synthesize(P, L, steps=None, literal=True, dtype=torch.float32, record=False):
    """synthesize.py:45-57: zero Y, zero prev_max_attentions, `max_T` full-graph
    passes keeping row j of Y and the argmax of row j, then one SSRN pass.

    literal=True recomputes TextEnc in every pass exactly as the reference's
    `sess.run` does (used for the CPU baseline timing); literal=False evaluates
    TextEnc once -- same arithmetic on the same inputs, hence identical results.
    record=True additionally returns per-step prev_max_attentions and the top-2
    in-window score margin (used by parity tests to skip near-ties)."""
    L = np.asarray(L)
    B = L.shape[0]
    steps = hp.max_T if steps is None else steps
    Y = torch.zeros((B, hp.max_T, hp.n_mels), dtype=dtype)
    pma = torch.zeros((B,), dtype=torch.int64)
    KV = None
    hist_p, hist_margin = [], []
    for j in range(steps):
        out = text2mel_forward(P, L, Y, pma, KV)
        if not literal and KV is None:
            KV = (out["K"], out["V"])
        if record:
            hist_p.append(pma.clone())
            a = out["alignments"][:, :, j]                       # (B, N) probs of row j
            top2 = torch.topk(a, 2, dim=-1).values
            hist_margin.append((top2[:, 0] - top2[:, 1]).clone())
        Y[:, j, :] = out["Y"][:, j, :]
        pma = out["max_attentions"][:, j]
    Z_logits, Z = SSRN(P, Y)
    res = dict(Y=Y, Z=Z, Z_logits=Z_logits, prev_max_attentions=pma,
               alignments=out["alignments"], max_attentions=out["max_attentions"])
    if record:
        res["p_hist"] = torch.stack(hist_p, 1)           # (B, steps)
        res["margin_hist"] = torch.stack(hist_margin, 1)  # (B, steps)
    return res
