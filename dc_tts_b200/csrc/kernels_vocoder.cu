// kernels_vocoder.cu -- Griffin-Lim vocoder (`spectrogram2wav`, reference utils.py:67-114) on the GPU.
// This is the first "next" row of SURVEY.md 8(f): the step right after the synthesis path, serial
// per-utterance CPU work in the reference (librosa), 51 inverse + 50 forward STFTs per utterance.
//
// One CTA per STFT frame; the 2048-point real transform is a 1024-point complex Stockham radix-4 FFT
// (first and last pass in registers, three through 16 KB of shared memory), the
// Hann window (1102 non-zero taps, centred in the 2048 frame) and librosa's conventions
// (center=True reflect padding, division by the summed squared window, n_fft/2 trimmed at both
// ends) are applied on the fly, so per Griffin-Lim iteration only the (B, T, 1025) complex
// spectrum, the windowed frames and the waveform touch HBM:
//   voc_istft_kernel   spectrum row -> Hermitian extension -> IFFT -> x window -> frame buffer
//   voc_ola_kernel     overlap-add of the <= 5 frames covering a sample, / window sum-square
//   voc_stft_phase_kernel  reflect-padded frame x window -> FFT -> X = S * est / max(1e-8, |est|)
// plus de-normalisation (power law), the de-pre-emphasis IIR (float64 like scipy.signal.lfilter)
// and the frame energies librosa.effects.trim thresholds.
#include "kernels.cuh"

#include <cmath>
#include <vector>

namespace dctts {

constexpr int VC_N = 2048;
constexpr int VC_THREADS = 256;

constexpr int VC_H = VC_N / 2;                      // complex FFT length (real-input packing)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// 1024-point complex FFT, Stockham autosort, radix 4, 256 threads, one butterfly per thread per pass.
// Entry: v[k] = x[tid + 256 k]; exit: v[k] = X[tid + 256 k] (natural order, no scaling).  The first
// pass reads and the last pass writes registers only; the three passes between go through the two
// 8 KB ping-pong buffers (one __syncthreads each).  tw[k] = exp(-2 pi i k / 2048), k < 2048.
template <bool INV>
__device__ __forceinline__ void fft1024(float2 (&v)[4], float2* s0, float2* s1, const float2* __restrict__ tw) {
    const int tid = threadIdx.x;
    float2* buf = s0;
#pragma unroll
    for (int ls = 0; ls <= 8; ls += 2) {
        const int str = 1 << ls, q = tid & (str - 1), p = tid >> ls;
        const float2 a = v[0], b = v[1], c = v[2], d = v[3];
        const float2 apc = make_float2(a.x + c.x, a.y + c.y), amc = make_float2(a.x - c.x, a.y - c.y);
        const float2 bpd = make_float2(b.x + d.x, b.y + d.y), bmd = make_float2(b.x - d.x, b.y - d.y);
        const float2 jb = INV ? make_float2(bmd.y, -bmd.x) : make_float2(-bmd.y, bmd.x);      // (+-i)(b - d), sign folded: y1 = amc - jb
        float2 y0 = make_float2(apc.x + bpd.x, apc.y + bpd.y);
        float2 y1 = make_float2(amc.x - jb.x, amc.y - jb.y);
        float2 y2 = make_float2(apc.x - bpd.x, apc.y - bpd.y);
        float2 y3 = make_float2(amc.x + jb.x, amc.y + jb.y);
        if (ls == 8) { v[0] = y0; v[1] = y1; v[2] = y2; v[3] = y3; break; }                   // n = 4: p = 0, unit twiddles
        const int i1 = 2 * (tid - q);
        float2 w1 = tw[i1], w2 = tw[2 * i1], w3 = tw[3 * i1];
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        y1 = cmul(y1, w1); y2 = cmul(y2, w2); y3 = cmul(y3, w3);
        float2* o = buf + q + str * 4 * p;
        if (ls == 0) {
            reinterpret_cast<float4*>(o)[0] = make_float4(y0.x, y0.y, y1.x, y1.y);
            reinterpret_cast<float4*>(o)[1] = make_float4(y2.x, y2.y, y3.x, y3.y);
        } else { o[0] = y0; o[str] = y1; o[2 * str] = y2; o[3 * str] = y3; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = buf[tid + 256 * k];
        buf = (buf == s0) ? s1 : s0;
    }
}

__global__ void voc_twiddle_kernel(float2* tw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < VC_N) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)VC_N, &s, &c);
        tw[k] = make_float2((float)c, (float)s);
    }
}

// utils.py:78-85: amplitude target S = (10 ^ ((clip(z,0,1)*max_db - max_db + ref_db) * 0.05)) ^ power; X <- S (zero phase)
__global__ void voc_prepare_kernel(const float* __restrict__ mag, float* __restrict__ S, float2* __restrict__ X, long long n,
                                   float max_db, float ref_db, float power) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = fminf(fmaxf(mag[i], 0.f), 1.f) * max_db - max_db + ref_db;
    float v = powf(powf(10.0f, m * 0.05f), power);
    S[i] = v;
    X[i] = make_float2(v, 0.f);
}

// librosa.core.istft, one frame: grid (T, B).  fr: (B, T, win) windowed time-domain frames.
// The 2048-point Hermitian inverse is a 1024-point complex one: with E/O the spectra of the even/odd
// samples, X[k] = E[k] + W^k O[k], X[k+1024] = conj(X[1024-k]) = E[k] - W^k O[k]; z = IFFT(E + i O)
// carries x[2n] in its real and x[2n+1] in its imaginary part.
__global__ void __launch_bounds__(VC_THREADS) voc_istft_kernel(const float2* __restrict__ X, float* __restrict__ fr,
                                                               const float2* __restrict__ tw, const float* __restrict__ window,
                                                               int T, int F, int win, int lpad) {
    __shared__ __align__(16) float2 s0[VC_H];
    __shared__ __align__(16) float2 s1[VC_H];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float2* x = X + ((size_t)b * T + t) * F;
    float2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int kk = tid + 256 * k;
        float2 xk = x[kk], xc = x[VC_H - kk];
        if (kk == 0) { xk.y = 0.f; xc.y = 0.f; }        // ifft(...).real drops the imaginary parts of bins 0 and n_fft/2
        xc.y = -xc.y;
        const float2 e = make_float2(xk.x + xc.x, xk.y + xc.y), d = make_float2(xk.x - xc.x, xk.y - xc.y);
        float2 w = tw[kk]; w.y = -w.y;
        const float2 o = cmul(d, w);
        v[k] = make_float2(e.x - o.y, e.y + o.x);        // E + i O (the halves are folded into the 1/2048 below)
    }
    fft1024<true>(v, s0, s1, tw);
    float* o = fr + ((size_t)b * T + t) * win;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int m = 2 * (tid + 256 * k) - lpad;        // frame sample 2n -> position in the window
        if (m >= 0 && m < win) o[m] = v[k].x * (1.0f / VC_N) * window[m];
        if (m + 1 >= 0 && m + 1 < win) o[m + 1] = v[k].y * (1.0f / VC_N) * window[m + 1];
    }
}

// overlap-add + window sum-square normalisation + centre trim: y (B, Ly), Ly = hop*(T-1)
__global__ void voc_ola_kernel(const float* __restrict__ fr, const float* __restrict__ wss, float* __restrict__ y,
                               int T, int win, int lpad, int hop, int Ly, float tiny) {
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (sidx >= Ly) return;
    const int u = sidx + VC_N / 2;                       // index in the un-trimmed signal
    // frames t with lpad <= u - hop*t < lpad + win, in ascending order like librosa's loop
    int t_hi = (u - lpad) / hop;
    int t_lo = (u - lpad - win) / hop + 1;
    if (u - lpad - win < 0) t_lo = 0;
    t_hi = min(t_hi, T - 1);
    float acc = 0.f;
    for (int t = max(t_lo, 0); t <= t_hi; ++t) acc += fr[((size_t)b * T + t) * win + (u - hop * t - lpad)];
    const float w = wss[u];
    y[(size_t)b * Ly + sidx] = (w > tiny) ? acc / w : acc;
}

// librosa.core.stft of the current estimate, one frame, fused with the Griffin-Lim phase update
// (utils.py:101-104): X = S * est / max(1e-8, |est|).  grid (T, B).  Real input packed as
// z[n] = x[2n] + i x[2n+1]; est[k] = (Z[k] + conj Z[1024-k]) / 2 - i W^k (Z[k] - conj Z[1024-k]) / 2.
__global__ void __launch_bounds__(VC_THREADS) voc_stft_phase_kernel(const float* __restrict__ y, const float* __restrict__ S,
                                                                    float2* __restrict__ X, const float2* __restrict__ tw,
                                                                    const float* __restrict__ window, int T, int F, int win,
                                                                    int lpad, int hop, int Ly) {
    __shared__ __align__(16) float2 s0[VC_H];
    __shared__ __align__(16) float2 s1[VC_H];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* yb = y + (size_t)b * Ly;
    auto sample = [&](int n) -> float {
        const int m = n - lpad;
        if (m < 0 || m >= win) return 0.f;
        int u = t * hop + n - VC_N / 2;                   // np.pad(y, n_fft//2, mode='reflect')
        if (u < 0) u = -u;
        if (u >= Ly) u = 2 * (Ly - 1) - u;
        return yb[u] * window[m];
    };
    float2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int n = 2 * (tid + 256 * k); v[k] = make_float2(sample(n), sample(n + 1)); }
    fft1024<false>(v, s0, s1, tw);
#pragma unroll
    for (int k = 0; k < 4; ++k) s0[tid + 256 * k] = v[k];          // s0 was last read before the final barrier of the FFT
    __syncthreads();
    const float* Sb = S + ((size_t)b * T + t) * F;
    float2* x = X + ((size_t)b * T + t) * F;
    auto emit = [&](int kk, float2 e) {
        const float mag = fmaxf(1e-8f, sqrtf(e.x * e.x + e.y * e.y));
        const float a = Sb[kk];
        x[kk] = make_float2(a * (e.x / mag), a * (e.y / mag));
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int kk = tid + 256 * k;
        const float2 zk = v[k];
        float2 zc = s0[(VC_H - kk) & (VC_H - 1)]; zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = cmul(make_float2(d.y, -d.x), tw[kk]);     // -i d W^k
        emit(kk, make_float2(e.x + o.x, e.y + o.y));
        if (kk == 0) emit(VC_H, make_float2(zk.x - zk.y, 0.f));    // bin n_fft/2: E[0] - O[0]
    }
}

// scipy.signal.lfilter([1], [1, -c], wav): y[n] = x[n] + c*y[n-1], evaluated in float64 like scipy.
// The recurrence is linear, so it is cut into chunks of DE_LC samples: (1) every chunk's end state from a zero
// start, (2) per utterance the true state entering each chunk, carry[j] = end[j-1] + c^DE_LC * carry[j-1]
// (449 steps instead of 230 000), (3) every chunk again, seeded with its carry, writing float32.  Given the
// carry, step (3) is the reference's own sequence of float64 operations; the carries differ from the serial
// evaluation by float64 rounding only.  (One thread per utterance walking all samples took 5-13 ms.)
constexpr int DE_LC = 512;

__global__ void voc_deemph_local_kernel(const float* __restrict__ y, double* __restrict__ ends, int Ly, int nch, double c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= nch) return;
    const float* p = y + (size_t)b * Ly + (size_t)j * DE_LC;
    const int n = min(DE_LC, Ly - j * DE_LC);
    double acc = 0.0;
    for (int i = 0; i < n; ++i) acc = (double)p[i] + c * acc;
    ends[(size_t)b * nch + j] = acc;
}

__global__ void voc_deemph_carry_kernel(double* __restrict__ ends, int nch, int B, double c) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double cl = 1.0;
    for (int i = 0; i < DE_LC; ++i) cl *= c;
    double* e = ends + (size_t)b * nch;
    double carry = 0.0;
    for (int j = 0; j < nch; ++j) { const double local = e[j]; e[j] = carry; carry = local + cl * carry; }
}

__global__ void voc_deemph_apply_kernel(float* __restrict__ y, const double* __restrict__ carry, int Ly, int nch, double c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= nch) return;
    float* p = y + (size_t)b * Ly + (size_t)j * DE_LC;
    const int n = min(DE_LC, Ly - j * DE_LC);
    double acc = carry[(size_t)b * nch + j];
    for (int i = 0; i < n; ++i) { acc = (double)p[i] + c * acc; p[i] = (float)acc; }
}

// librosa.feature.rmse(y, 2048, 512)**2 per centred frame (reflect padding): mse (B, nfr)
__global__ void __launch_bounds__(256) voc_frame_mse_kernel(const float* __restrict__ y, float* __restrict__ mse, int Ly, int nfr,
                                                           int flen, int fhop) {
    __shared__ float red[8];
    const int f = blockIdx.x, b = blockIdx.y;
    const float* yb = y + (size_t)b * Ly;
    float acc = 0.f;
    for (int n = threadIdx.x; n < flen; n += 256) {
        int u = f * fhop + n - flen / 2;
        if (u < 0) u = -u;
        if (u >= Ly) u = 2 * (Ly - 1) - u;
        const float v = yb[u];
        acc = fmaf(v, v, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        mse[(size_t)b * nfr + f] = t / (float)flen;
    }
}

// ---------------------------------------------------------------------------------------- features
// get_spectrograms (reference utils.py:20-65) for one trimmed utterance, one CTA per STFT frame: pre-emphasis
// (float32 multiply then subtract, like numpy), reflect-padded Hann frame, the same real-packed FFT, |X|,
// the mel filterbank (each mel bin is a contiguous run of FFT bins), 20 log10, normalisation.  grid (T).
__global__ void __launch_bounds__(VC_THREADS) feat_stft_mel_kernel(const float* __restrict__ y, int len, float preemph,
                                                                   float* __restrict__ mag_out, float* __restrict__ mel_out,
                                                                   const float* __restrict__ melw, const int2* __restrict__ melrange,
                                                                   const float2* __restrict__ tw, const float* __restrict__ window,
                                                                   int F, int n_mels, int win, int lpad, int hop, float ref_db,
                                                                   float max_db) {
    __shared__ __align__(16) float2 s0[VC_H];
    __shared__ __align__(16) float2 s1[VC_H];
    const int t = blockIdx.x, tid = threadIdx.x;
    auto sample = [&](int n) -> float {
        const int m = n - lpad;
        if (m < 0 || m >= win) return 0.f;
        int u = t * hop + n - VC_N / 2;                   // np.pad(y, n_fft//2, mode='reflect')
        if (u < 0) u = -u;
        if (u >= len) u = 2 * (len - 1) - u;
        u = min(max(u, 0), len - 1);
        const float v = (u > 0) ? __fsub_rn(y[u], __fmul_rn(preemph, y[u - 1])) : y[0];      // utils.py:39
        return v * window[m];
    };
    float2 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int n = 2 * (tid + 256 * k); v[k] = make_float2(sample(n), sample(n + 1)); }
    fft1024<false>(v, s0, s1, tw);
#pragma unroll
    for (int k = 0; k < 4; ++k) s0[tid + 256 * k] = v[k];
    __syncthreads();
    float* lin = reinterpret_cast<float*>(s1);             // |X[k]|, k <= 1024 (s1 was last read inside the FFT)
    float* mo = mag_out + (size_t)t * F;
    auto emit = [&](int kk, float2 e) {
        const float a = sqrtf(e.x * e.x + e.y * e.y);
        lin[kk] = a;
        const float db = 20.0f * log10f(fmaxf(1e-5f, a));
        mo[kk] = fminf(fmaxf((db - ref_db + max_db) / max_db, 1e-8f), 1.0f);
    };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int kk = tid + 256 * k;
        const float2 zk = v[k];
        float2 zc = s0[(VC_H - kk) & (VC_H - 1)]; zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = cmul(make_float2(d.y, -d.x), tw[kk]);
        emit(kk, make_float2(e.x + o.x, e.y + o.y));
        if (kk == 0) emit(VC_H, make_float2(zk.x - zk.y, 0.f));
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int m = warp; m < n_mels; m += VC_THREADS / 32) {
        const int2 r = melrange[m];
        const float* w = melw + (size_t)m * F;
        float acc = 0.f;
        for (int k = r.x + lane; k < r.y; k += 32) acc = fmaf(w[k], lin[k], acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float db = 20.0f * log10f(fmaxf(1e-5f, acc));
            mel_out[(size_t)t * n_mels + m] = fminf(fmaxf((db - ref_db + max_db) / max_db, 1e-8f), 1.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------- host
void voc_make_tables(float2* tw_dev, float* window_dev, float* wss_dev, int T, int win, int hop, cudaStream_t s) {
    voc_twiddle_kernel<<<(VC_N + 255) / 256, 256, 0, s>>>(tw_dev);
    // periodic Hann of win taps (scipy get_window('hann', win, fftbins=True)) and librosa's window_sumsquare,
    // accumulated in float32 in frame order like the reference
    std::vector<float> w(win);
    for (int n = 0; n < win; ++n) w[n] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)win));
    const int lpad = (VC_N - win) / 2;
    const int n_tot = VC_N + hop * (T - 1);
    std::vector<float> wss(n_tot, 0.f);
    std::vector<float> wsq(VC_N, 0.f);
    for (int n = 0; n < win; ++n) { const double d = 0.5 - 0.5 * std::cos(2.0 * M_PI * (double)n / (double)win); wsq[lpad + n] = (float)(d * d); }
    for (int t = 0; t < T; ++t)
        for (int n = 0; n < VC_N && t * hop + n < n_tot; ++n) wss[t * hop + n] += wsq[n];
    cudaMemcpyAsync(window_dev, w.data(), win * sizeof(float), cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(wss_dev, wss.data(), n_tot * sizeof(float), cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);
}

// librosa.filters.mel(sr, n_fft, n_mels): Slaney scale, fmin 0, fmax sr/2, area-normalised triangles (float64 here,
// float32 on the device); range[m] = [first, last+1) non-zero FFT bin of mel bin m.
void feat_make_mel_basis(int sr, int n_fft, int n_mels, std::vector<float>& w, std::vector<int>& range) {
    const int F = 1 + n_fft / 2;
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    auto hz2mel = [&](double f) { return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp; };
    auto mel2hz = [&](double m) { return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m; };
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz2mel(0.0), m_hi = hz2mel(sr / 2.0);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel2hz(m_lo + (m_hi - m_lo) * (double)i / (double)(n_mels + 1));
    w.assign((size_t)n_mels * F, 0.f);
    range.assign(2 * (size_t)n_mels, 0);
    for (int i = 0; i < n_mels; ++i) {
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int first = -1, last = -1;
        for (int k = 0; k < F; ++k) {
            const double f = (sr / 2.0) * (double)k / (double)(F - 1);
            const double lower = (f - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
            const double upper = (mel_f[i + 2] - f) / (mel_f[i + 2] - mel_f[i + 1]);
            const double v = std::max(0.0, std::min(lower, upper)) * enorm;
            w[(size_t)i * F + k] = (float)v;
            if (v > 0.0) { if (first < 0) first = k; last = k; }
        }
        range[2 * i] = first < 0 ? 0 : first;
        range[2 * i + 1] = first < 0 ? 0 : last + 1;
    }
}

void feat_frame_mse(const float* y, float* mse, int n, int nfr, cudaStream_t s) {
    voc_frame_mse_kernel<<<dim3(nfr, 1), 256, 0, s>>>(y, mse, n, nfr, 2048, 512);
}

void feat_run(const float* y, int len, float preemph, float* mag, float* mel, const float* melw, const int* melrange,
              const float2* tw, const float* window, int T, int F, int n_mels, int win, int hop, float ref_db, float max_db,
              cudaStream_t s) {
    feat_stft_mel_kernel<<<T, VC_THREADS, 0, s>>>(y, len, preemph, mag, mel, melw, reinterpret_cast<const int2*>(melrange), tw,
                                                  window, F, n_mels, win, (VC_N - win) / 2, hop, ref_db, max_db);
}

int voc_launches_per_call(int n_iter) { return 1 + 3 * n_iter + 2 + 3 + 1; }
size_t voc_deemph_scratch_bytes(int B, int T, int hop) { return (size_t)B * ((hop * (T - 1) + DE_LC - 1) / DE_LC) * sizeof(double); }

void voc_run(const VocoderArgs& a, cudaStream_t s) {
    const int T = a.T, F = a.F, B = a.B, win = a.win, hop = a.hop, Ly = hop * (T - 1), lpad = (VC_N - win) / 2;
    const long long n = (long long)B * T * F;
    voc_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a.mag, a.S, a.X, n, a.max_db, a.ref_db, a.power);
    dim3 gframes(T, B), gsamp((Ly + 255) / 256, B);
    for (int it = 0; it <= a.n_iter; ++it) {
        voc_istft_kernel<<<gframes, VC_THREADS, 0, s>>>(a.X, a.frames, a.tw, a.window, T, F, win, lpad);
        voc_ola_kernel<<<gsamp, 256, 0, s>>>(a.frames, a.wss, a.wav, T, win, lpad, hop, Ly, 1.17549435e-38f);
        if (it < a.n_iter)
            voc_stft_phase_kernel<<<gframes, VC_THREADS, 0, s>>>(a.wav, a.S, a.X, a.tw, a.window, T, F, win, lpad, hop, Ly);
    }
    const int nch = (Ly + DE_LC - 1) / DE_LC;
    const dim3 gch((nch + 63) / 64, B);
    voc_deemph_local_kernel<<<gch, 64, 0, s>>>(a.wav, a.deemph, Ly, nch, (double)a.preemphasis);
    voc_deemph_carry_kernel<<<(B + 31) / 32, 32, 0, s>>>(a.deemph, nch, B, (double)a.preemphasis);
    voc_deemph_apply_kernel<<<gch, 64, 0, s>>>(a.wav, a.deemph, Ly, nch, (double)a.preemphasis);
    const int nfr = 1 + Ly / 512;
    voc_frame_mse_kernel<<<dim3(nfr, B), 256, 0, s>>>(a.wav, a.mse, Ly, nfr, 2048, 512);
}

}  // namespace dctts
