// kernels_vocoder.cu -- Griffin-Lim vocoder (`spectrogram2wav`, reference utils.py:67-114) on the GPU.
// This is the first "next" row of SURVEY.md 8(f): the step right after the synthesis path, serial
// per-utterance CPU work in the reference (librosa), 51 inverse + 50 forward STFTs per utterance.
//
// One CTA per STFT frame; a 2048-point radix-2 FFT lives entirely in shared memory (16 KB), the
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

__device__ __forceinline__ int bitrev11(int i) { return (int)(__brev((unsigned)i) >> 21); }

// In-place radix-2 decimation-in-time FFT of s[2048] (input already in bit-reversed order).
// tw[k] = exp(-2 pi i k / 2048), k < 1024; inverse = conjugated twiddles (no 1/N scaling).
__device__ __forceinline__ void fft2048(float2* s, const float2* __restrict__ tw, bool inverse) {
#pragma unroll 1
    for (int st = 0; st < 11; ++st) {
        const int half = 1 << st;
        for (int j = threadIdx.x; j < VC_N / 2; j += VC_THREADS) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> st) << (st + 1)) + pos, i1 = i0 + half;
            float2 w = tw[pos << (10 - st)];
            if (inverse) w.y = -w.y;
            const float2 a = s[i0], b = s[i1];
            const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            s[i0] = make_float2(a.x + t.x, a.y + t.y);
            s[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
}

__global__ void voc_twiddle_kernel(float2* tw) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < VC_N / 2) {
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
__global__ void __launch_bounds__(VC_THREADS) voc_istft_kernel(const float2* __restrict__ X, float* __restrict__ fr,
                                                               const float2* __restrict__ tw, const float* __restrict__ window,
                                                               int T, int F, int win, int lpad) {
    __shared__ float2 s[VC_N];
    const int t = blockIdx.x, b = blockIdx.y;
    const float2* x = X + ((size_t)b * T + t) * F;
    for (int k = threadIdx.x; k < VC_N; k += VC_THREADS) {
        float2 v;
        if (k < F) v = x[k];
        else { v = x[VC_N - k]; v.y = -v.y; }            // spec[-2:0:-1].conj()
        s[bitrev11(k)] = v;
    }
    __syncthreads();
    fft2048(s, tw, true);
    float* o = fr + ((size_t)b * T + t) * win;
    for (int n = threadIdx.x; n < win; n += VC_THREADS)
        o[n] = s[lpad + n].x * (1.0f / VC_N) * window[n];
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
// (utils.py:101-104): X = S * est / max(1e-8, |est|).  grid (T, B).
__global__ void __launch_bounds__(VC_THREADS) voc_stft_phase_kernel(const float* __restrict__ y, const float* __restrict__ S,
                                                                    float2* __restrict__ X, const float2* __restrict__ tw,
                                                                    const float* __restrict__ window, int T, int F, int win,
                                                                    int lpad, int hop, int Ly) {
    __shared__ float2 s[VC_N];
    const int t = blockIdx.x, b = blockIdx.y;
    const float* yb = y + (size_t)b * Ly;
    for (int n = threadIdx.x; n < VC_N; n += VC_THREADS) {
        float v = 0.f;
        if (n >= lpad && n < lpad + win) {
            int u = t * hop + n - VC_N / 2;               // np.pad(y, n_fft//2, mode='reflect')
            if (u < 0) u = -u;
            if (u >= Ly) u = 2 * (Ly - 1) - u;
            v = yb[u] * window[n - lpad];
        }
        s[bitrev11(n)] = make_float2(v, 0.f);
    }
    __syncthreads();
    fft2048(s, tw, false);
    const float* Sb = S + ((size_t)b * T + t) * F;
    float2* x = X + ((size_t)b * T + t) * F;
    for (int k = threadIdx.x; k < F; k += VC_THREADS) {
        const float2 e = s[k];
        const float mag = fmaxf(1e-8f, sqrtf(e.x * e.x + e.y * e.y));
        const float a = Sb[k];
        x[k] = make_float2(a * (e.x / mag), a * (e.y / mag));
    }
}

// scipy.signal.lfilter([1], [1, -c], wav): y[n] = x[n] + c*y[n-1], evaluated in float64 like scipy
__global__ void voc_deemph_kernel(float* __restrict__ y, int Ly, int B, double c) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float* p = y + (size_t)b * Ly;
    double prev = 0.0;
    for (int n = 0; n < Ly; ++n) { prev = (double)p[n] + c * prev; p[n] = (float)prev; }
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

// ---------------------------------------------------------------------------------------- host
void voc_make_tables(float2* tw_dev, float* window_dev, float* wss_dev, int T, int win, int hop, cudaStream_t s) {
    voc_twiddle_kernel<<<(VC_N / 2 + 255) / 256, 256, 0, s>>>(tw_dev);
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

int voc_launches_per_call(int n_iter) { return 1 + 3 * n_iter + 2 + 2; }

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
    voc_deemph_kernel<<<(B + 31) / 32, 32, 0, s>>>(a.wav, Ly, B, (double)a.preemphasis);
    const int nfr = 1 + Ly / 512;
    voc_frame_mse_kernel<<<dim3(nfr, B), 256, 0, s>>>(a.wav, a.mse, Ly, nfr, 2048, 512);
}

}  // namespace dctts
