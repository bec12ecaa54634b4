"""The number format of the tensor-core kernels, emulated in numpy: an fp32 operand x is carried as two fp16 planes
hi = fp16(s x), lo = fp16(s x - hi) with a power-of-two scale s, and a product is hi*hi' + hi*lo' + lo*hi' accumulated in
fp32 (csrc/kernels_tc.cu, csrc/kernels_gemm_tc.cu).  These tests pin the claims the design rests
on: the scheme is fp32-grade (2^-22 per operand), single-pass fp16 / bf16 are not, and the per-tensor scale of the training
GEMMs keeps 1e-7-sized gradients in fp16's normal range."""
import numpy as np


def _pow2_scale(x, top=13):
    m = float(np.abs(x).max())
    return 1.0 if m == 0 else 2.0 ** (top - int(np.floor(np.log2(m))))


def _planes(x, s):
    hi = (x * s).astype(np.float16)
    lo = (x * s - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def _gemm3(a, b):
    """(M,K) x (N,K)^T with the three-product scheme and per-tensor scales, fp32 accumulation."""
    sa, sb = _pow2_scale(a), _pow2_scale(b)
    ah, al = _planes(a, np.float32(sa))
    bh, bl = _planes(b, np.float32(sb))
    acc = ah @ bh.T + ah @ bl.T + al @ bh.T
    return acc * np.float32(1.0 / (sa * sb))


def _bf16(x):
    u = x.astype(np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def test_three_product_scheme_is_fp32_grade_and_single_pass_is_not():
    rng = np.random.default_rng(0)
    a = rng.normal(0, 1, (128, 768)).astype(np.float32)
    b = rng.normal(0, 0.05, (256, 768)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    scale = np.abs(ref).max()
    err3 = np.abs(_gemm3(a, b) - ref).max() / scale
    err_fp32 = np.abs(a @ b.T - ref).max() / scale
    err_fp16 = np.abs(a.astype(np.float16).astype(np.float32) @ b.astype(np.float16).astype(np.float32).T - ref).max() / scale
    err_bf16 = np.abs(_bf16(a) @ _bf16(b).T - ref).max() / scale
    assert err3 < 2e-6 and err3 < 8 * max(err_fp32, 1e-7)          # within a small factor of fp32 FMA arithmetic
    assert err_fp16 > 20 * err3 and err_bf16 > 100 * err3           # why one pass misses the 1e-3 budget after 25 blocks


def test_per_tensor_scale_keeps_tiny_gradients_in_range():
    rng = np.random.default_rng(1)
    dy = (rng.normal(0, 1, (420, 512)) * 3e-7).astype(np.float32)   # gradient-sized values: unscaled fp16 would flush them
    x = rng.normal(0, 1, (420, 256)).astype(np.float32)
    ref = x.astype(np.float64).T @ dy.astype(np.float64)            # the weight-gradient GEMM: X^T dY
    got = _gemm3(np.ascontiguousarray(x.T), np.ascontiguousarray(dy.T))
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6
    naive = np.ascontiguousarray(x.T).astype(np.float16).astype(np.float32) @ dy.astype(np.float16).astype(np.float32)
    assert np.abs(naive - ref).max() / np.abs(ref).max() > 1e-2     # unscaled fp16: subnormal / flushed


def test_elements_far_below_the_tensor_maximum_lose_bits_gracefully():
    """One scale per tensor: an element 2^-20 below the maximum still has fp16's 11 bits (the hi plane is normal down to
    2^-27 of the maximum), so its contribution is wrong by at most 2^-11 of ITSELF -- negligible against the tensor's max-norm,
    which is what the gradient-parity criterion (2e-3 of the max-norm) measures."""
    a = np.zeros((1, 16), np.float32); b = np.ones((1, 16), np.float32)
    a[0, 0] = 1.0; a[0, 1] = 2.0 ** -20 * 1.2345
    got = _gemm3(a, b)[0, 0]
    ref = float(a.astype(np.float64).sum())
    assert abs(got - ref) <= 2.0 ** -11 * a[0, 1] + 2.0 ** -22
