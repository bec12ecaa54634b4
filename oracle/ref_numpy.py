"""ORACLE (test infrastructure, not product) -- independent numpy restatement.

Second, deliberately different statement of the same reference algorithm as
`ref_torch.py`: convolutions are explicit per-tap matrix products with hand-built
index arithmetic (no library conv), layer-norm follows tf.nn.batch_normalization's
`x*inv + (beta - mean*inv)` form, the attention mask is built by per-row loops, and
everything can run in float64 to quantify float32 noise.  PARITY UNPINNED (see
ref_torch.py).  Each function cites the reference lines it follows.
"""
import numpy as np

from dc_tts_b200.hyperparams import Hyperparams as hp

LN_EPS = 1e-12
MASK_VALUE = -2 ** 32 + 1            # networks.py:146


def _p(P, name, dt):
    return np.asarray(P[name]).astype(dt)


def embed(P, ids, scope, dt=np.float32):
    """modules.py:13-42."""
    table = _p(P, scope + "/lookup_table", dt).copy()
    table[0, :] = 0                      # modules.py:36-38
    return table[np.asarray(ids)]


def normalize(x, gamma, beta):
    """modules.py:45-64; tf.nn.moments (biased variance) + batch_normalization form."""
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    inv = gamma / np.sqrt(var + x.dtype.type(LN_EPS))
    return x * inv + (beta - mean * inv)


def _conv(x, W, b, rate, padding):
    """y[t] = b + sum_j W[j]^T x[t + j*rate - left], zero outside [0, L) (modules.py:121-134)."""
    B, L, _ = x.shape
    k = W.shape[0]
    left = (k - 1) * rate if padding.lower() == "causal" else ((k - 1) * rate) // 2
    y = np.zeros((B, L, W.shape[2]), x.dtype) + b
    for j in range(k):
        off = j * rate - left
        lo, hi = max(0, -off), min(L, L - off)
        if hi > lo:
            y[:, lo:hi, :] += x[:, lo + off:hi + off, :] @ W[j]
    return y


def conv1d(P, x, scope, rate=1, padding="SAME", activation_fn=None):
    """modules.py:91-141 (training=False)."""
    dt = x.dtype
    y = _conv(x, _p(P, scope + "/conv1d/kernel", dt), _p(P, scope + "/conv1d/bias", dt), rate, padding)
    y = normalize(y, _p(P, scope + "/normalize/gamma", dt), _p(P, scope + "/normalize/beta", dt))
    return np.maximum(y, 0) if activation_fn == "relu" else y


def hc(P, x, scope, rate=1, padding="SAME"):
    """modules.py:143-197."""
    dt = x.dtype
    y = _conv(x, _p(P, scope + "/conv1d/kernel", dt), _p(P, scope + "/conv1d/bias", dt), rate, padding)
    C = y.shape[-1] // 2
    H1 = normalize(y[..., :C], _p(P, scope + "/H1/gamma", dt), _p(P, scope + "/H1/beta", dt))
    H2 = normalize(y[..., C:], _p(P, scope + "/H2/gamma", dt), _p(P, scope + "/H2/beta", dt))
    H1 = 1.0 / (1.0 + np.exp(-H1))
    return (H1 * H2 + (1 - H1) * x).astype(dt)


def conv1d_transpose(P, x, scope):
    """modules.py:199-247; scatter form of the stride-2 transposed conv:
    input row t adds W[j] x[t] into output row 2t + j, outputs beyond 2L dropped."""
    dt = x.dtype
    W = _p(P, scope + "/conv2d_transpose/kernel", dt)[0]        # (3, Cout, Cin)
    b = _p(P, scope + "/conv2d_transpose/bias", dt)
    B, L, _ = x.shape
    y = np.zeros((B, 2 * L + 1, W.shape[1]), dt)
    for j in range(3):
        y[:, j:j + 2 * L:2, :] += x @ W[j].T
    y = y[:, :2 * L, :] + b
    return normalize(y, _p(P, scope + "/normalize/gamma", dt), _p(P, scope + "/normalize/beta", dt))


def run_chain(P, x, net, layers):
    for l in layers:
        scope = "%s/%s" % (net, l.scope)
        if l.kind == "C":
            x = conv1d(P, x, scope, l.rate, l.pad, l.act)
        elif l.kind == "HC":
            x = hc(P, x, scope, l.rate, l.pad)
        else:
            x = conv1d_transpose(P, x, scope)
    return x


def TextEnc(P, L, dt=np.float32):
    """networks.py:14-71."""
    from dc_tts_b200.arch import textenc_layers
    x = run_chain(P, embed(P, L, "Text2Mel/TextEnc/embed_1", dt), "Text2Mel/TextEnc", textenc_layers())
    d = x.shape[-1] // 2
    return x[..., :d], x[..., d:]


def AudioEnc(P, S):
    """networks.py:73-124."""
    from dc_tts_b200.arch import audioenc_layers
    return run_chain(P, S, "Text2Mel/AudioEnc", audioenc_layers())


def Attention(Q, K, V, mononotic_attention=False, prev_max_attentions=None):
    """networks.py:126-155."""
    A = np.einsum("btd,bnd->btn", Q, K) * Q.dtype.type(1.0 / np.sqrt(float(hp.d)))
    if mononotic_attention:
        for b, p in enumerate(np.asarray(prev_max_attentions)):
            for n in range(hp.max_N):
                if n < p or n >= p + hp.attention_win_size:
                    A[b, :, n] = MASK_VALUE
    A = A - A.max(-1, keepdims=True)
    E = np.exp(A)
    A = E / E.sum(-1, keepdims=True)
    max_attentions = A.argmax(-1)
    R = np.concatenate((np.einsum("btn,bnd->btd", A, V), Q), -1)
    return R, A.transpose(0, 2, 1), max_attentions


def AudioDec(P, R):
    """networks.py:157-212."""
    from dc_tts_b200.arch import audiodec_layers
    logits = run_chain(P, R, "Text2Mel/AudioDec", audiodec_layers())
    return logits, 1.0 / (1.0 + np.exp(-logits))


def SSRN(P, Y):
    """networks.py:214-292."""
    from dc_tts_b200.arch import ssrn_layers
    logits = run_chain(P, Y, "SSRN", ssrn_layers())
    return logits, 1.0 / (1.0 + np.exp(-logits))


def text2mel_forward(P, L, mels, prev_max_attentions):
    """train.py:48-68, synthesize branch."""
    S = np.concatenate((np.zeros_like(mels[:, :1]), mels[:, :-1]), 1)
    K, V = TextEnc(P, L, mels.dtype)
    Q = AudioEnc(P, S)
    R, alignments, max_attentions = Attention(Q, K, V, True, prev_max_attentions)
    logits, Y = AudioDec(P, R)
    return dict(K=K, V=V, Q=Q, R=R, alignments=alignments, max_attentions=max_attentions,
                Y_logits=logits, Y=Y)
