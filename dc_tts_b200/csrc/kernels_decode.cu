// kernels_decode.cu -- the whole autoregressive decode (reference synthesize.py:45-54: 210 x
// {AudioEnc, Attention, AudioDec}, networks.py:73-212) in ONE persistent launch.
//
// Why: a decode step is 24 dependent conv blocks on one row per utterance; as 46 graph nodes it was
// bound by kernel boundaries and by split-K round trips through L2 (round 1: ~300 us per step, 1.5 %
// of any roofline).  Utterances never interact (networks.py:140-153 is batched per row), so a group
// of G <= 4 utterances can be decoded by one thread-block CLUSTER with no grid-level synchronisation:
//
//   * 16 CTAs per cluster; CTA r owns 1/16 of every block's output channels (for `hc` the same 16
//     channels of the gate and of the info half, so the highway mix is local).
//   * its weight slices for all 24 blocks form one contiguous 1.7 MB stream (packed at commit time,
//     [k/4][column][4] so that one LDS.128 yields four k of one column); the stream is identical for
//     every frame and is pulled from L2 by TMA bulk copies (cp.async.bulk + mbarrier) into a
//     9 x 16 KB shared-memory ring, always ~128 KB ahead of the math.
//   * per block: GEMV of the slice on fp32 FMA (exact fp32, as the reference), the pre-LN slice is
//     written into every peer's shared memory (DSMEM all-gather, 512 B per peer), ONE cluster
//     barrier, then every CTA normalises the whole rows redundantly (LayerNorm, gate, mix): the next
//     block's input sits in local shared memory.  Dilated taps come from the per-layer history in
//     HBM/L2, prefetched one block ahead with cp.async; each CTA appends its slice of the new row.
//   * Quirk Q1: the reference recomputes R under the CURRENT window every step.  While
//     the window of an utterance does not move, the cached rows are exactly what a recompute would
//     give, so only row j is evaluated (1 row per block).  When it moves, the 85-row receptive field
//     of AudioDec is recomputed (85/83/77/59/5/3 rows for C_1, HC_2..HC_6) by a register-tiled fp32
//     GEMM over the same weight stream, pre-LN rows through an L2 scratch, LayerNorm one warp per row.
//
// All waits are bounded (mbarrier waits trap after ~2 s).  Cluster barriers are executed by all
// threads of all CTAs in uniform control flow: every branch that contains one depends only on values
// that are computed identically in every CTA (the attention windows).
#include "kernels_decode.cuh"
#include "tc_ptx.cuh"

#include <math.h>

namespace dctts {
using namespace ptx;

namespace {

constexpr int NC = DEC_NC, GMAX = DEC_GMAX, NT = DEC_THREADS, NWARP = NT / 32;
constexpr int XLD = 512;                       // row pitch of the shared activation buffers
constexpr int ALD = 132;                       // row pitch of the transposed A sub-tile (128 rows + 4)
constexpr int WRK_F = 2 * 16 * ALD;            // max(A staging 2 x 16 x 132, normalised rows GMAX x 2 x 256)
static_assert(WRK_F >= GMAX * 2 * 256, "work buffer too small for the normalised rows");

struct Smem {
    float ring[DEC_NSLOT * DEC_SLOT_F];
    float prm[2][DEC_PRM_F];
    float xtap[2][GMAX][XLD];
    float xcur[2][GMAX][XLD];
    float pre[2][GMAX][XLD];
    float red[GMAX][NT];
    float wrk[WRK_F];
    unsigned long long full[DEC_NSLOT];
    int p_cur[GMAX], p_prev[GMAX], p_next[GMAX], moved[GMAX];
    long long prof[16], prof_last;
};

// lap timer (option decode_prof): thread 0 attributes the cycles since the previous lap to bucket i
#define LAP(i) do { if (P.prof && threadIdx.x == 0) { const long long now_ = clock64(); S.prof[i] += now_ - S.prof_last; S.prof_last = now_; } } while (0)
enum { LP_START = 0, LP_WAIT = 1, LP_GEMV = 2, LP_RELEASE = 3, LP_GATHER = 4, LP_CBAR = 5, LP_LN = 6, LP_MIX = 7, LP_ATT = 8,
       LP_PYR_ATT = 9, LP_PYR_GEMM = 10, LP_PYR_LN = 11, LP_PYR_BAR = 12, LP_FRAME = 13 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;                                   // 0: zero fill (TF zero padding of the causal conv)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- the weight stream ---------------------------------------------------------------------------
struct Stream {
    const float* base;          // this rank's stream
    long long pos;              // next chunk to consume (global index over all frames)
    long long total;            // frames * chunks per frame
};

__device__ __forceinline__ void stream_issue(const DecParams& P, Smem& S, const Stream& st, long long idx) {
    if (idx >= st.total) return;
    const int c = (int)(idx % P.nch), slot = (int)(idx % DEC_NSLOT);
    const uint32_t bytes = (uint32_t)P.C[c].nfl4 * 16u;
    mbar_expect_tx(reinterpret_cast<uint64_t*>(&S.full[slot]), bytes);
    bulk_g2s(S.ring + slot * DEC_SLOT_F, st.base + P.C[c].off, bytes, &S.full[slot]);
}
__device__ __forceinline__ const float* stream_acquire(Smem& S, const Stream& st) {
    const int slot = (int)(st.pos % DEC_NSLOT);
    mbar_wait(reinterpret_cast<uint64_t*>(&S.full[slot]), (uint32_t)((st.pos / DEC_NSLOT) & 1));
    return S.ring + slot * DEC_SLOT_F;
}
// all threads are done with the chunk at st.pos: refill its slot with the chunk DEC_NSLOT ahead
__device__ __forceinline__ void stream_release(const DecParams& P, Smem& S, Stream& st) {
    __syncthreads();
    if (threadIdx.x == 0) stream_issue(P, S, st, st.pos + DEC_NSLOT);
    st.pos++;
}

// ---- prefetch of the next block's parameters (and dilated taps) -----------------------------------
__device__ __forceinline__ void prefetch_params(const DecParams& P, Smem& S, int li, int rank) {
    const DecLayer& l = P.L[li];
    float* dst = S.prm[li & 1];
    const int tid = threadIdx.x;
    cp_async16(dst + tid * 4, P.lnp[li] + tid * 4, true);                        // 1024 floats
    if (tid < l.ns / 4) {                                                         // bias slice, stream column order
        // hc: columns [0,cs) gate of channels rank*cs.., [cs,2cs) info; conv: [0,cs) (+ zero padding handled by the reader)
        const int n = tid * 4;
        const float* src = (l.kind == 1 && n >= l.cs) ? P.bias[li] + l.cout + rank * l.cs + (n - l.cs) : P.bias[li] + rank * l.cs + n;
        // slices are 16-byte aligned only when cs % 4 == 0; otherwise (n_mels / 16 = 5) the reader loads bias from global
        if ((l.cs & 3) == 0) cp_async16(dst + 1024 + n, src, true);
    }
}
// taps (all but the last) of block li at frame j for the G utterances: rows j - (ntaps-1-tap)*rate of the input history
__device__ __forceinline__ void prefetch_taps(const DecParams& P, Smem& S, int li, int j, int b0, int G) {
    const DecLayer& l = P.L[li];
    const int ntap_ld = l.ntaps - 1;
    if (ntap_ld <= 0 || !P.in_hist[li]) return;
    const int per_row = l.cin / 4;                                                // float4 per row
    const int total = G * ntap_ld * per_row;
    for (int i = threadIdx.x; i < total; i += NT) {
        const int c4 = i % per_row, rt = i / per_row, tap = rt % ntap_ld, g = rt / ntap_ld;
        const int t = j - (l.ntaps - 1 - tap) * l.rate;
        const float* src = P.in_hist[li] + ((size_t)(b0 + g) * P.T + (t < 0 ? 0 : t)) * l.ldin + c4 * 4;
        cp_async16(&S.xtap[li & 1][g][tap * 256 + c4 * 4], src, t >= 0);
    }
}

// ---- GEMV of one weight chunk: acc[g] += sum_k x[g][k] * W[k][n] -----------------------------------
template <int NS>
__device__ __forceinline__ void gemv_chunk(const float* __restrict__ w, const float* __restrict__ x, int krows, int G,
                                           float (&acc)[GMAX]) {
    constexpr int NG = NT / NS;
    const int n = threadIdx.x % NS, kq = threadIdx.x / NS;
    const int kper = krows / NG, kb = kq * kper;
#pragma unroll 2
    for (int k = kb; k < kb + kper; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + ((size_t)(k >> 2) * NS + n) * 4);
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g < G) {
                const float4 xv = *reinterpret_cast<const float4*>(x + g * XLD + k);
                acc[g] = fmaf(xv.x, wv.x, acc[g]); acc[g] = fmaf(xv.y, wv.y, acc[g]);
                acc[g] = fmaf(xv.z, wv.z, acc[g]); acc[g] = fmaf(xv.w, wv.w, acc[g]);
            }
        }
    }
}

// global column of stream column n of rank r (hc: gate | info halves of the 2*cout pre-LN row)
__device__ __forceinline__ int pre_col(const DecLayer& l, int rank, int n) {
    if (l.kind == 1) return n < l.cs ? rank * l.cs + n : 256 + rank * l.cs + (n - l.cs);
    return rank * l.cs + n;
}
__device__ __forceinline__ float bias_of(const DecParams& P, const Smem& S, int li, int rank, int n) {
    const DecLayer& l = P.L[li];
    if ((l.cs & 3) == 0) return S.prm[li & 1][1024 + n];
    if (n >= l.cs) return 0.f;
    return __ldg(P.bias[li] + rank * l.cs + n);
}

// ---- one block on ONE row per utterance -------------------------------------------------------------
// in: S.xcur[cb] = the block's input at frame j (all G utterances), S.xtap[li&1] = its dilated taps (prefetched),
// S.prm[li&1] = its parameters.  out: S.xcur[cb^1] = the block's output row; this CTA's channel slice appended
// to the output history.  Returns the new cb.
__device__ int layer_row(const DecParams& P, Smem& S, Stream& st, int li, int j, int b0, int G, int rank, int cb, bool direct_in) {
    const DecLayer& l = P.L[li];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nl_next = (li + 1 == P.nl) ? 0 : li + 1;
    // prefetch for the following block (its taps are rows of earlier frames, already in the history)
    {
        const bool next_frame = (li + 1 == P.nl);
        const int jn = next_frame ? j + 1 : j;
        if (!next_frame || j + 1 < P.steps) {
            prefetch_params(P, S, nl_next, rank);
            prefetch_taps(P, S, nl_next, jn, b0, G);
        }
        cp_async_commit();
    }
    if (direct_in) {
        // first single-row block after a recompute: its input rows (taps AND the current row) were just rewritten in the
        // history by the whole cluster (a cluster barrier precedes this call)
        const int per_row = l.cin / 4, total = G * l.ntaps * per_row;
        for (int i = tid; i < total; i += NT) {
            const int c4 = i % per_row, rt = i / per_row, tap = rt % l.ntaps, g = rt / l.ntaps;
            const int t = j - (l.ntaps - 1 - tap) * l.rate;
            const float* src = P.in_hist[li] + ((size_t)(b0 + g) * P.T + (t < 0 ? 0 : t)) * l.ldin + c4 * 4;
            float* dst = (tap == l.ntaps - 1) ? &S.xcur[cb][g][c4 * 4] : &S.xtap[li & 1][g][tap * 256 + c4 * 4];
            cp_async16(dst, src, t >= 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        cp_async_wait<1>();                 // everything but the group just committed: this block's taps and parameters
    }
    __syncthreads();
    LAP(LP_START);

    float acc[GMAX];
#pragma unroll
    for (int g = 0; g < GMAX; ++g) acc[g] = 0.f;
    for (int c = l.ch0; c < l.ch0 + l.nch; ++c) {
        const DecChunk& ch = P.C[c];
        const float* w = stream_acquire(S, st);
        LAP(LP_WAIT);
        const float* x = (ch.tap == l.ntaps - 1) ? &S.xcur[cb][0][ch.ci0] : &S.xtap[li & 1][0][ch.tap * 256 + ch.ci0];
        if (l.ns == 32) gemv_chunk<32>(w, x, ch.krows, G, acc);
        else if (l.ns == 16) gemv_chunk<16>(w, x, ch.krows, G, acc);
        else gemv_chunk<8>(w, x, ch.krows, G, acc);
        LAP(LP_GEMV);
        stream_release(P, S, st);
        LAP(LP_RELEASE);
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g) S.red[g][tid] = acc[g];
    __syncthreads();
    // final sums of the slice, written into every CTA's `pre` (all-gather through distributed shared memory)
    const int pb = li & 1;
    {
        const int nvals = G * l.ns;                       // <= 128
        const int idx = tid % 128, half = tid / 128;      // two thread halves serve 8 peers each
        if (idx < nvals) {
            const int g = idx / l.ns, n = idx % l.ns;
            const int ng = NT / l.ns;
            float s = bias_of(P, S, li, rank, n);
            for (int q = 0; q < ng; ++q) s += S.red[g][q * l.ns + n];
            const bool real = (l.kind == 1) ? true : (n < l.cs);
            if (real) {
                const uint32_t local = smem_u32(&S.pre[pb][g][pre_col(l, rank, n)]);
#pragma unroll
                for (int p = 0; p < 8; ++p) st_cluster_f32(mapa(local, (uint32_t)(half * 8 + p)), s);
            }
        }
    }
    LAP(LP_GATHER);
    cluster_sync_all();
    LAP(LP_CBAR);
    // LayerNorm of whole rows, redundantly in every CTA: warp -> (utterance, half)
    const int nh = l.kind + 1, C = l.cout;
    float* nrm = S.wrk;                                   // [g][half][256]
    const float* prm = S.prm[li & 1];
    for (int pr = warp; pr < G * nh; pr += NWARP) {
        const int g = pr / nh, hf = pr % nh;
        const float* y = &S.pre[pb][g][hf * 256];
        float v[8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; v[i] = (c < C) ? y[c] : 0.f; s += v[i]; }
        const float fC = (float)C;
        const float mean = warp_sum(s) / fC;
        float qd = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; v[i] = (c < C) ? v[i] - mean : 0.f; qd = fmaf(v[i], v[i], qd); }
        const float inv = 1.0f / sqrtf(warp_sum(qd) / fC + 1e-12f);
        const float* gam = prm + hf * 512; const float* bet = gam + 256;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = lane + 32 * i; if (c < C) nrm[(g * 2 + hf) * 256 + c] = v[i] * inv * gam[c] + bet[c]; }
    }
    __syncthreads();
    LAP(LP_LN);
    const bool last = (li + 1 == P.nl);
    float* oh = P.out_hist[li];
    for (int i = tid; i < G * C; i += NT) {
        const int g = i / C, c = i % C;
        float o;
        if (l.kind == 1) {
            const float h1 = sigmoid_acc(nrm[(g * 2) * 256 + c]);
            o = h1 * nrm[(g * 2 + 1) * 256 + c] + (1.0f - h1) * S.xcur[cb][g][c];
        } else {
            o = nrm[(g * 2) * 256 + c];
            if (l.act == 1) o = fmaxf(o, 0.f);
        }
        const size_t row = (size_t)(b0 + g) * P.T + j;
        if (c / l.cs == rank && oh) oh[row * C + c] = o;           // this CTA's slice of the history row
        if (last) {                                                 // Y = sigmoid(logits), networks.py:210; next frame's AudioEnc input
            o = sigmoid_acc(o);
            if (rank == 0) P.ybuf[row * C + c] = o;
        }
        S.xcur[cb ^ 1][g][c] = o;
    }
    if (last) {                                                     // AudioEnc C_1 reads K = 128 padded channels
        for (int i = tid; i < G * (128 - C); i += NT) S.xcur[cb ^ 1][i / (128 - C)][C + i % (128 - C)] = 0.f;
    }
    LAP(LP_MIX);
    return cb ^ 1;
}

// ---- attention of ONE query row under the 3-key window (networks.py:140-153) -------------------------
// lane holds q[lane*8 .. +8); returns ctx[8] in the same layout and the argmax key (first index among equal maxima)
__device__ __forceinline__ int attend_row(const DecParams& P, const float (&qv)[8], int b, int p, int lane, float (&ctx)[8]) {
    const int d = P.d;
    const int n_lo = min(max(p, 0), P.N - 1), n_hi = min(n_lo + P.win_size, P.N);
    const float scale = rsqrtf((float)d);
    float sc[4];
    for (int n = n_lo; n < n_hi; ++n) {
        const float* k = P.kv + ((size_t)b * P.N + n) * (2 * d) + lane * 8;
        const float4 k0 = ldcg4(k), k1 = ldcg4(k + 4);
        float s = 0.f;
        s = fmaf(qv[0], k0.x, s); s = fmaf(qv[1], k0.y, s); s = fmaf(qv[2], k0.z, s); s = fmaf(qv[3], k0.w, s);
        s = fmaf(qv[4], k1.x, s); s = fmaf(qv[5], k1.y, s); s = fmaf(qv[6], k1.z, s); s = fmaf(qv[7], k1.w, s);
        sc[n - n_lo] = warp_sum(s) * scale;
    }
    const int cnt = n_hi - n_lo;
    float mx = -INFINITY;
    for (int i = 0; i < cnt; ++i) mx = fmaxf(mx, sc[i]);
    float sum = 0.f;
    for (int i = 0; i < cnt; ++i) { sc[i] = expf(sc[i] - mx); sum += sc[i]; }
    float best = -1.f; int besti = 0;
    for (int i = 0; i < cnt; ++i) { sc[i] = sc[i] / sum; if (sc[i] > best) { best = sc[i]; besti = i; } }
#pragma unroll
    for (int i = 0; i < 8; ++i) ctx[i] = 0.f;
    for (int n = n_lo; n < n_hi; ++n) {
        const float* v = P.kv + ((size_t)b * P.N + n) * (2 * d) + d + lane * 8;
        const float4 v0 = ldcg4(v), v1 = ldcg4(v + 4);
        const float p_ = sc[n - n_lo];
        ctx[0] = fmaf(p_, v0.x, ctx[0]); ctx[1] = fmaf(p_, v0.y, ctx[1]); ctx[2] = fmaf(p_, v0.z, ctx[2]); ctx[3] = fmaf(p_, v0.w, ctx[3]);
        ctx[4] = fmaf(p_, v1.x, ctx[4]); ctx[5] = fmaf(p_, v1.y, ctx[5]); ctx[6] = fmaf(p_, v1.z, ctx[6]); ctx[7] = fmaf(p_, v1.w, ctx[7]);
    }
    return n_lo + besti;
}

// ---- recompute path: rows of the AudioDec receptive field ---------------------------------------------
struct RowList { int cnt[GMAX], off[GMAX + 1], M; };
__device__ __forceinline__ RowList make_rows(const Smem& S, int G, int j, int prow) {
    RowList r; r.off[0] = 0;
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
        r.cnt[g] = (g < G) ? (S.moved[g] ? min(prow, j + 1) : 1) : 0;
        r.off[g + 1] = r.off[g] + r.cnt[g];
    }
    r.M = r.off[GMAX];
    return r;
}
__device__ __forceinline__ void row_of(const RowList& r, int m, int j, int& g, int& t) {
    g = 0;
#pragma unroll
    for (int i = 1; i < GMAX; ++i) if (m >= r.off[i]) g = i;
    t = j - r.cnt[g] + 1 + (m - r.off[g]);
}

// register-tiled fp32 GEMM of the row list against this CTA's weight slice; 128 rows x NS columns per pass,
// thread tile 4 rows x TN columns {q, q+8, ..}; A sub-tiles (128 rows x 16 k) transposed through shared memory
template <int TN>
__device__ void pyr_gemm(const DecParams& P, Smem& S, Stream& st, int li, int j, int b0, const RowList& rl, int rank, float* scr) {
    constexpr int NS = 8 * TN;
    const DecLayer& l = P.L[li];
    const int tid = threadIdx.x, q = tid & 7, rg = tid >> 3;
    const int nrb = (rl.M + 127) / 128;                   // <= 3
    float acc[3][4][TN];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < TN; ++c) acc[a][i][c] = 0.f;
    float* As = S.wrk;                                    // [2][16][ALD]
    const int lr = tid & 127, lk = tid >> 7;              // loader role: row, k-quad {lk, lk+2}
    for (int c = l.ch0; c < l.ch0 + l.nch; ++c) {
        const DecChunk& ch = P.C[c];
        const float* w = stream_acquire(S, st);
        const int shift = -(l.ntaps - 1 - ch.tap) * l.rate;
        const int nks = ch.krows / 16;
#pragma unroll
        for (int rb = 0; rb < 3; ++rb) {
            if (rb < nrb) {
                const int m = rb * 128 + lr;
                const float* src = nullptr;
                if (m < rl.M) {
                    int g, t; row_of(rl, m, j, g, t);
                    const int ts = t + shift;
                    if (ts >= 0) src = P.in_hist[li] + ((size_t)(b0 + g) * P.T + ts) * l.ldin + ch.ci0;
                }
                float4 ra[2];
                auto gload = [&](int ks) {
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        ra[i] = src ? ldcg4(src + ks * 16 + (lk + 2 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                auto sstore = [&](int buf) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float* a0 = As + (buf * 16 + (lk + 2 * i) * 4) * ALD + lr;
                        a0[0] = ra[i].x; a0[ALD] = ra[i].y; a0[2 * ALD] = ra[i].z; a0[3 * ALD] = ra[i].w;
                    }
                };
                gload(0);
                __syncthreads();                          // previous users of As are done
                sstore(0);
                __syncthreads();
                for (int ks = 0; ks < nks; ++ks) {
                    const int buf = ks & 1;
                    if (ks + 1 < nks) gload(ks + 1);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        float4 w4[TN];
#pragma unroll
                        for (int cc = 0; cc < TN; ++cc)
                            w4[cc] = *reinterpret_cast<const float4*>(w + ((size_t)(ks * 4 + k4) * NS + q + 8 * cc) * 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float4 a4 = *reinterpret_cast<const float4*>(As + (buf * 16 + k4 * 4 + kk) * ALD + rg * 4);
#pragma unroll
                            for (int cc = 0; cc < TN; ++cc) {
                                const float ww = kk == 0 ? w4[cc].x : (kk == 1 ? w4[cc].y : (kk == 2 ? w4[cc].z : w4[cc].w));
                                acc[rb][0][cc] = fmaf(a4.x, ww, acc[rb][0][cc]);
                                acc[rb][1][cc] = fmaf(a4.y, ww, acc[rb][1][cc]);
                                acc[rb][2][cc] = fmaf(a4.z, ww, acc[rb][2][cc]);
                                acc[rb][3][cc] = fmaf(a4.w, ww, acc[rb][3][cc]);
                            }
                        }
                    }
                    if (ks + 1 < nks) { sstore(buf ^ 1); __syncthreads(); }
                }
            }
        }
        stream_release(P, S, st);
    }
    // pre-LN slice (+ bias) -> scratch rows [m][512]
#pragma unroll
    for (int rb = 0; rb < 3; ++rb) {
        if (rb < nrb) {
#pragma unroll
            for (int cc = 0; cc < TN; ++cc) {
                const int n = q + 8 * cc;
                const float bs = bias_of(P, S, li, rank, n);
                const int col = pre_col(l, rank, n);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = rb * 128 + rg * 4 + i;
                    if (m < rl.M) scr[(size_t)m * 512 + col] = acc[rb][i][cc] + bs;
                }
            }
        }
    }
}

// LayerNorm / gate / highway mix of the recomputed rows: one warp per row over the whole cluster
__device__ void pyr_ln(const DecParams& P, Smem& S, int li, int j, int b0, const RowList& rl, int rank, const float* scr) {
    const DecLayer& l = P.L[li];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* prm = S.prm[li & 1];
    const float fC = 256.f;
    for (int m = rank * NWARP + warp; m < rl.M; m += NC * NWARP) {
        int g, t; row_of(rl, m, j, g, t);
        const float* y = scr + (size_t)m * 512;
        const size_t row = (size_t)(b0 + g) * P.T + t;
        float z[2][8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (hf > l.kind) break;
            const float4 a = ldcg4(y + hf * 256 + lane * 4), b = ldcg4(y + hf * 256 + 128 + lane * 4);
            float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
            const float mean = warp_sum(s) / fC;
            float qd = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] -= mean; qd = fmaf(v[i], v[i], qd); }
            const float inv = 1.0f / sqrtf(warp_sum(qd) / fC + 1e-12f);
            const float* gam = prm + hf * 512; const float* bet = gam + 256;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int c = (i < 4 ? 0 : 128) + lane * 4 + (i & 3); z[hf][i] = v[i] * inv * gam[c] + bet[c]; }
        }
        float o[8];
        if (l.kind == 1) {
            const float* xr = P.in_hist[li] + row * l.ldin;
            const float4 a = ldcg4(xr + lane * 4), b = ldcg4(xr + 128 + lane * 4);
            const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float h1 = sigmoid_acc(z[0][i]); o[i] = h1 * z[1][i] + (1.0f - h1) * x[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = l.act == 1 ? fmaxf(z[0][i], 0.f) : z[0][i];
        }
        float* orow = P.out_hist[li] + row * 256;
        *reinterpret_cast<float4*>(orow + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(orow + 128 + lane * 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

}  // namespace

__global__ void __cluster_dims__(DEC_NC, 1, 1) __launch_bounds__(DEC_THREADS, 1)
decode_cluster_kernel(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x / NC;
    const int b0 = cluster * P.G;
    const int G = min(P.G, P.B - b0);
    float* scr = P.pre_scr + (size_t)cluster * P.G * 85 * 512;

    Stream st;
    st.base = P.wstream + (size_t)rank * P.stream_len;
    st.pos = 0; st.total = (long long)P.steps * P.nch;

    if (tid == 0) {
        for (int s = 0; s < DEC_NSLOT; ++s) mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 2 * GMAX * XLD; i += NT) { (&S.xcur[0][0][0])[i] = 0.f; (&S.xtap[0][0][0])[i] = 0.f; (&S.pre[0][0][0])[i] = 0.f; }
    if (tid < GMAX) { S.p_cur[tid] = 0; S.p_prev[tid] = 0; S.p_next[tid] = 0; S.moved[tid] = 0; }
    if (tid < 16) S.prof[tid] = 0;
    __syncthreads();
    if (tid == 0) for (int s = 0; s < DEC_NSLOT; ++s) stream_issue(P, S, st, s);
    prefetch_params(P, S, 0, rank);
    cp_async_commit();
    cluster_sync_all();                                   // every CTA of the cluster is running: DSMEM is addressable

    int cb = 0;
    int n_moved_frames = 0, n_moved_utt = 0;
    if (tid == 0) S.prof_last = clock64();
    for (int j = 0; j < P.steps; ++j) {
        bool any_moved = false;
        if (tid < GMAX) S.moved[tid] = (tid < G && j > 0 && S.p_cur[tid] != S.p_prev[tid]) ? 1 : 0;
        __syncthreads();
        for (int g = 0; g < G; ++g) any_moved |= (S.moved[g] != 0);
        if (rank == 0 && tid < G) P.p_hist[(size_t)(b0 + tid) * P.T + j] = S.p_cur[tid];

        // AudioEnc (networks.py:81-124): input Y[j-1] (train.py:51), already in xcur[cb]
        for (int li = 0; li < P.n_enc; ++li) cb = layer_row(P, S, st, li, j, b0, G, rank, cb, false);

        // Attention of row j under the current window, redundantly in every CTA: R[j] = [A.V ; Q] (networks.py:140-153)
        __syncthreads();
        if (warp < G) {
            const int g = warp;
            float qv[8], ctx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] = S.xcur[cb][g][lane * 8 + i];
            const int amax = attend_row(P, qv, b0 + g, S.p_cur[g], lane, ctx);
            if (lane == 0) S.p_next[g] = amax;
#pragma unroll
            for (int i = 0; i < 8; ++i) { S.xcur[cb ^ 1][g][lane * 8 + i] = ctx[i]; S.xcur[cb ^ 1][g][P.d + lane * 8 + i] = qv[i]; }
        }
        cb ^= 1;
        __syncthreads();
        LAP(LP_ATT);

        int li = P.n_enc;
        if (any_moved) {
            // ---- recompute the receptive field of the moved utterances under the new window ----
            n_moved_frames++;
            for (int g = 0; g < G; ++g) n_moved_utt += S.moved[g];
            cluster_sync_all();                           // Q[j] slices of all CTAs are in the history
            const RowList ra = make_rows(S, G, j, P.L[P.n_enc].prow);
            const float* Qh = P.out_hist[P.n_enc - 1];
            for (int m = rank * NWARP + warp; m < ra.M; m += NC * NWARP) {
                int g, t; row_of(ra, m, j, g, t);
                const size_t row = (size_t)(b0 + g) * P.T + t;
                float qv[8], ctx[8];
                const float4 q0 = ldcg4(Qh + row * P.d + lane * 8), q1 = ldcg4(Qh + row * P.d + lane * 8 + 4);
                qv[0] = q0.x; qv[1] = q0.y; qv[2] = q0.z; qv[3] = q0.w; qv[4] = q1.x; qv[5] = q1.y; qv[6] = q1.z; qv[7] = q1.w;
                attend_row(P, qv, b0 + g, S.p_cur[g], lane, ctx);
                float* rr = P.rbuf + row * (2 * P.d);
                *reinterpret_cast<float4*>(rr + lane * 8) = make_float4(ctx[0], ctx[1], ctx[2], ctx[3]);
                *reinterpret_cast<float4*>(rr + lane * 8 + 4) = make_float4(ctx[4], ctx[5], ctx[6], ctx[7]);
                *reinterpret_cast<float4*>(rr + P.d + lane * 8) = q0;
                *reinterpret_cast<float4*>(rr + P.d + lane * 8 + 4) = q1;
            }
            cluster_sync_all();
            LAP(LP_PYR_ATT);
            for (; li < P.nl && P.L[li].prow > 1; ++li) {
                const int nl_next = li + 1;               // AudioDec never ends on a recomputed block
                prefetch_params(P, S, nl_next, rank);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                const RowList rl = make_rows(S, G, j, P.L[li].prow);
                if (P.L[li].ns == 32) pyr_gemm<4>(P, S, st, li, j, b0, rl, rank, scr);
                else pyr_gemm<2>(P, S, st, li, j, b0, rl, rank, scr);
                LAP(LP_PYR_GEMM);
                cluster_sync_all();
                LAP(LP_PYR_BAR);
                pyr_ln(P, S, li, j, b0, rl, rank, scr);
                LAP(LP_PYR_LN);
                cluster_sync_all();
                LAP(LP_PYR_BAR);
            }
            cb = layer_row(P, S, st, li, j, b0, G, rank, cb, true);
            ++li;
        }
        for (; li < P.nl; ++li) cb = layer_row(P, S, st, li, j, b0, G, rank, cb, false);

        __syncthreads();
        if (tid < GMAX) { S.p_prev[tid] = S.p_cur[tid]; S.p_cur[tid] = S.p_next[tid]; }
        __syncthreads();
        LAP(LP_FRAME);
    }
    if (rank == 0 && tid < G) P.p_final[b0 + tid] = S.p_cur[tid];
    if (rank == 0 && tid == 0 && P.stats) { P.stats[2 * cluster] = n_moved_frames; P.stats[2 * cluster + 1] = n_moved_utt; }
    if (P.prof && cluster == 0 && rank == 0 && tid < 16) P.prof[tid] = S.prof[tid];
    cp_async_wait<0>();
    cluster_sync_all();                                   // no CTA exits while a peer may still write into its shared memory
}

size_t decode_smem_bytes() { return sizeof(Smem) + 128; }

static cudaError_t decode_prepare() {
    cudaError_t e = cudaFuncSetAttribute(decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem_bytes());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(decode_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

int decode_max_active_clusters() {
    if (decode_prepare() != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(DEC_NC); cfg.blockDim = dim3(DEC_THREADS); cfg.dynamicSmemBytes = decode_smem_bytes();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = DEC_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, decode_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

cudaError_t launch_decode_cluster(const DecParams& p, int n_clusters, cudaStream_t s) {
    cudaError_t e = decode_prepare();                     // per call: the attribute is per device, handles may live on several
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * DEC_NC); cfg.blockDim = dim3(DEC_THREADS); cfg.dynamicSmemBytes = decode_smem_bytes(); cfg.stream = s;
    cfg.attrs = nullptr; cfg.numAttrs = 0;               // cluster dims are compiled in (__cluster_dims__)
    return cudaLaunchKernelEx(&cfg, decode_cluster_kernel, p);
}

}  // namespace dctts
