// kernels_decode.cu -- the whole autoregressive decode (reference synthesize.py:45-54: 210 x
// {AudioEnc, Attention, AudioDec}, networks.py:73-212) in ONE persistent launch.
//
// Why: a decode step is 24 dependent conv blocks on one row per utterance; as 46 graph nodes it was
// bound by kernel boundaries and by split-K round trips through L2 (round 1: ~300 us per step, 1.5 %
// of any roofline).  Utterances never interact (networks.py:140-153 is batched per row), so a group
// of G <= 5 utterances is decoded by one thread-block CLUSTER with no grid-level synchronisation:
//
//   * 16 CTAs per cluster; CTA r owns 1/16 of every block's output channels (for `hc` the same 16
//     channels of the gate and of the info half, so the highway mix is local).
//   * its weight slices for all 24 blocks form one contiguous 1.7 MB stream (packed at commit time).
//     The stream is the same for every frame and is pulled from L2 by TMA bulk copies
//     (cp.async.bulk + mbarrier) into a 3 x 48 KB shared-memory ring.  Every WARP owns the k rows it
//     multiplies: it waits on its own mbarrier, and refills its own 6 KB region the moment it has read
//     it -- the chunk loop has no block-wide synchronisation at all.
//   * per block: GEMV of the slice on fp32 FMA (exact fp32 arithmetic, as the reference), one block
//     barrier for the cross-warp reduction, the pre-LN slice goes to every peer with ONE bulk copy
//     per peer through distributed shared memory, completing on the peer's mbarrier (no hardware
//     cluster barrier on this path); every CTA then normalises the whole rows redundantly (LayerNorm
//     statistics one warp per (utterance, half), gate / highway mix one thread per channel), so the next
//     block's input is in local shared memory.  Dilated taps come from the per-layer history in HBM/L2, prefetched one
//     block ahead with cp.async; each CTA appends its channel slice of the new row.
//   * Quirk Q1: the reference recomputes R under the CURRENT window every step.  While
//     the window of an utterance does not move, the cached rows are exactly what a recompute would
//     give.  When it moves, a PRE-PASS refreshes the rows t < j of its AudioDec receptive field
//     (84/82/76/58/4/2 rows of C_1, HC_2..HC_6) under the new window -- attention one warp per row, a
//     tcgen05 GEMM per utterance (split-fp16 planes staged per 16-channel slab in the no-swizzle K-major
//     layout, the three taps being the same slab read through shifted descriptors), pre-LN rows through
//     an L2 scratch, LayerNorm one warp per row over the whole cluster -- and the ordinary one-row pass
//     then runs for every utterance.  The pre-pass consumes the same weight chunks a second time: the
//     stream is "virtual" (frame, segment, chunk) and both the consumer and the refill cursor walk it.
//
// All waits are bounded (mbarrier waits trap after ~2 s).  Hardware cluster barriers (pre-pass, frame
// end) are executed by all threads of all CTAs in uniform control flow: every branch that contains
// one depends only on the attention windows, which every CTA computes identically.
#include "kernels_decode.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <atomic>
#include <math.h>

namespace dctts {
using namespace ptx;

namespace {

constexpr int NC = DEC_NC, GMAX = DEC_GMAX, NT = DEC_THREADS, NWARP = NT / 32;
constexpr int XLD = 768;                       // row pitch of the input vectors: [tap0 | tap1 | current]
constexpr int PLD = GMAX * 32 + 16;            // pitch between the ranks' slices in `pre` (16 floats of bank skew)
constexpr int TC_RA = 96;                      // pre-pass: rows per k8 group of an A slab plane (<= 96 source rows per utterance)
constexpr int TC_APLANE = 2 * TC_RA * 16;      // bytes of one plane of one 16-channel slab (2 k8 groups)
constexpr int TC_ASTAGE = 2 * TC_APLANE;       // hi + lo planes
constexpr int TC_NSTG = 2;                     // A slab stages
constexpr int WRK_F = TC_NSTG * TC_ASTAGE / 4; // the pre-pass work buffer, shared with the per-frame partial sums
static_assert(WRK_F >= GMAX * NT && WRK_F >= 1024, "work buffer: partial sums of the per-frame path / LayerNorm parameters of the pre-pass");

struct Smem {
    float ring[DEC_NSLOT][NWARP][DEC_REG_F];
    union {                             // never live together: the pre-pass stages A here while no per-frame block is in flight
        float wrk[WRK_F];               // tcgen05 pre-pass A slab stages (its reads run up to 128 + 54 rows past a slab start: xin follows)
        float red[GMAX][NT];            // per-frame path: partial sums per warp; pre-pass: LayerNorm parameters of the block
    };
    float xin[2][GMAX][XLD];
    float pre[2][NC][PLD];
    float outv[2][GMAX * 32];
    float prm[2][DEC_PRM_F];
    unsigned long long fullw[DEC_NSLOT][NWARP];
    unsigned long long gbar[2];
    unsigned long long sbar[TC_NSTG], abar[TC_NSTG], dbar;   // tcgen05 pre-pass: slab stage free / slab stage filled / accumulator complete
    uint32_t tmem_base, pad_;
    float stat[GMAX][2][2];             // per utterance and LN half: mean, 1/sqrt(var + eps)
    uint32_t tc_baddr[96];              // tcgen05 pre-pass: descriptor start field of every weight slab of the current block
    int n_moved_frames, n_moved_utt;
    DecParams P;                        // the kernel's parameter block: indexed per block / chunk on the critical path; in the
                                        // constant bank those indexed loads missed the (instruction-shared) constant cache
    int p_cur[GMAX], p_prev[GMAX], p_next[GMAX], moved[GMAX];
    int fmoved[2];
    long long prof[16], prof_last;
};

static_assert(sizeof(Smem) + 128 <= 232448, "decode kernel: shared memory budget (227 KB per CTA)");

// lap timer (option decode_prof): thread 0 attributes the cycles since the previous lap to bucket i
#define LAP(i) do { if constexpr (PROF) { if (threadIdx.x == 0) { const long long now_ = clock64(); S.prof[i] += now_ - S.prof_last; S.prof_last = now_; } } } while (0)
enum { LP_START = 0, LP_WAIT = 1, LP_GEMV = 2, LP_RELEASE = 3, LP_GATHER = 4, LP_CBAR = 5, LP_LN = 6, LP_MIX = 7, LP_ATT = 8,
       LP_PYR_ATT = 9, LP_PYR_GEMM = 10, LP_PYR_LN = 11, LP_PYR_BAR = 12, LP_FRAME = 13 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// NOTE on `__noinline__` in this file: there is none.  With 227 KB of shared memory per CTA the L1 data cache is ~0 KB, so every
// stack access (ABI spills of a non-inlined call, a dynamically indexed local array) is an L2 round trip of ~500 cycles: the
// first three versions of this kernel spent 60 % of their time there (ptxas must report a 0-byte stack frame).
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
// the per-frame path: ex2.approx + rcp (2 ulp each; far inside the 1e-3 budget, half the instructions)
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

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
// local shared memory -> a peer CTA's shared memory, completing on the PEER's mbarrier
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_cluster), "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint64_t* bar64(unsigned long long* p) { return reinterpret_cast<uint64_t*>(p); }

// ---- the virtual weight stream ---------------------------------------------------------------------
// frame f = AudioEnc chunks, [the AudioDec chunks of the receptive-field blocks, if a window moved in f], AudioDec chunks
struct Cur { int f, seg, c; };
__device__ __forceinline__ void cur_next(const DecParams& P, const Smem& S, Cur& u) {
    u.c++;
    const int end = (u.seg == 0) ? P.nch_enc : (u.seg == 1 ? P.pyr_ch1 : P.nch);
    if (u.c < end) return;
    if (u.seg == 0) {
        if (S.fmoved[u.f & 1]) { u.seg = 1; u.c = P.pyr_ch0; } else { u.seg = 2; u.c = P.nch_enc; }
    } else if (u.seg == 1) { u.seg = 2; u.c = P.nch_enc; }
    else { u.f++; u.seg = 0; u.c = 0; }
}
struct Stream {
    const float* base;          // this rank's packed stream
    Cur prod;                   // chunk that will be loaded into the slot the consumer frees next (the consumer itself walks the blocks' chunk ranges)
    unsigned pos;               // number of chunks consumed so far
    unsigned tcq, tca;          // tcgen05 pre-pass: slabs staged / accumulators completed so far (mbarrier phase parities)
};
// one lane of warp `warp`: load the warp's rows of chunk `u` into its region of `slot`
__device__ __forceinline__ void stream_issue(const DecParams& P, Smem& S, const Stream& st, const Cur& u, int slot, int warp) {
    if (u.f >= P.steps) return;
    const DecChunk& ch = P.C[u.c];
    const uint32_t bytes = (uint32_t)ch.nfl4 * 2u;                   // nfl * 4 bytes / 8 warps
    mbar_expect_tx(bar64(&S.fullw[slot][warp]), bytes);
    const int off = (u.seg == 1) ? ch.off16 : ch.off;                // the pre-pass reads the same rows as split-fp16 MMA slabs
    bulk_g2s(&S.ring[slot][warp][0], st.base + off + warp * (ch.nfl4 >> 1), bytes, &S.fullw[slot][warp]);
}
__device__ __forceinline__ void stream_advance(const DecParams& P, const Smem& S, Stream& st) {
    cur_next(P, S, st.prod); st.pos++;
}
// the calling warp has read its region of the current chunk: refill it with its rows of the chunk 3 ahead
__device__ __forceinline__ void warp_release(const DecParams& P, Smem& S, Stream& st, int warp, int lane) {
    __syncwarp();                                                     // every lane's loads of the region have returned (their FMAs have issued)
    if (lane == 0) stream_issue(P, S, st, st.prod, (int)(st.pos % DEC_NSLOT), warp);
    stream_advance(P, S, st);
}

// ---- prefetch of the next block's parameters and dilated taps -------------------------------------------
__device__ __forceinline__ void prefetch_params(const DecParams& P, Smem& S, int li, int rank) {
    const DecLayer& l = P.L[li];
    float* dst = S.prm[li & 1];
    const int tid = threadIdx.x;
    cp_async16(dst + tid * 4, P.lnp[li] + tid * 4, true);                        // 1024 floats
    // bias slice in stream column order -- hc: columns [0,cs) gate of channels rank*cs.., [cs,2cs) info; conv: [0,cs), zero beyond
    if ((l.cs & 3) == 0) {
        if (tid < l.ns / 4) {
            const int n = tid * 4;
            cp_async16(dst + 1024 + n, P.bias[li] + ((l.kind == 1 && n >= l.cs) ? l.cout + rank * l.cs + (n - l.cs) : rank * l.cs + n), true);
        }
    } else if (tid < l.ns) {                                                     // the n_mels-wide last block (5 channels per CTA)
        dst[1024 + tid] = (tid < l.cs) ? __ldg(P.bias[li] + rank * l.cs + tid) : 0.f;
    }
}
// taps (all but the last) of block li at frame j: rows j - (ntaps-1-tap)*rate of its input history -> xin[buf][g][tap*256..]
// (multi-tap blocks have 256 input channels and 3 taps: pack_decode checks it)
__device__ __forceinline__ void prefetch_taps(const DecParams& P, Smem& S, int li, int j, int b0, int G, int buf) {
    const DecLayer& l = P.L[li];
    if (l.ntaps != 3 || !P.in_hist[li]) return;
    for (int i = threadIdx.x; i < G * 128; i += NT) {
        const int c4 = i & 63, tap = (i >> 6) & 1, g = i >> 7;
        const int t = j - (2 - tap) * l.rate;
        const float* src = P.in_hist[li] + ((size_t)(b0 + g) * P.T + (t < 0 ? 0 : t)) * 256 + c4 * 4;
        cp_async16(&S.xin[buf][g][tap * 256 + c4 * 4], src, t >= 0);
    }
}

// ---- GEMV of the calling warp's k rows of one chunk: acc[g] += sum_k x[g][k] * W[k][n] ----------------------
// wreg: the warp's region ([k/4][column][4]); x: row 0 of the input vectors at the warp's first k; ns = 32 / 16 / 8 columns
// per CTA.  The utterance count GT is a template parameter of the whole kernel (only ONE instantiation runs per launch, so the
// instruction-cache footprint is that of one): no per-utterance branches, all loads of an 8-k step hoisted, two independent
// FMA chains per utterance.  Slots g >= the cluster's real utterance count multiply zeros (their input rows stay zero).
template <int GT>
__device__ __forceinline__ void gemv_warp(const float* __restrict__ wreg, const float* __restrict__ x, int kr8, int ns, float (&acc)[GT]) {
    const int lane = threadIdx.x & 31;
    const int lg = (ns == 32) ? 5 : (ns == 16 ? 4 : 3);
    const int n = lane & (ns - 1), sg = lane >> lg;                  // column, k sub-group inside the warp
    const int kper = kr8 >> (5 - lg);                                // k rows per sub-group, multiple of 8
    const float* w = wreg + ((size_t)((sg * kper) >> 2) * ns + n) * 4;
    const float* xs = x + sg * kper;
    const int wstep = ns * 4;
    // packed fp32 FMA (FFMA2, new on sm_100): one instruction multiplies two consecutive k of the column -- half the issue
    // slots of the fma pipe, which two warps per scheduler otherwise saturate (ncu: the GEMV was fma-pipe bound at G = 5).
    // Four partial sums per utterance (k mod 4 = {0,1} and {2,3} of either float4), added at the end.
    float2 p0[GT], p1[GT];
#pragma unroll
    for (int g = 0; g < GT; ++g) { p0[g] = make_float2(acc[g], 0.f); p1[g] = make_float2(0.f, 0.f); }
#pragma unroll 2
    for (int k = 0; k < kper; k += 8) {
        const float4 w0 = *reinterpret_cast<const float4*>(w);
        const float4 w1 = *reinterpret_cast<const float4*>(w + wstep);
        w += 2 * wstep;
#pragma unroll
        for (int g = 0; g < GT; ++g) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + g * XLD + k);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + g * XLD + k + 4);
            float2 a = p0[g], c = p1[g];
            a = __ffma2_rn(make_float2(x0.x, x0.y), make_float2(w0.x, w0.y), a);
            c = __ffma2_rn(make_float2(x1.x, x1.y), make_float2(w1.x, w1.y), c);
            a = __ffma2_rn(make_float2(x0.z, x0.w), make_float2(w0.z, w0.w), a);
            c = __ffma2_rn(make_float2(x1.z, x1.w), make_float2(w1.z, w1.w), c);
            p0[g] = a; p1[g] = c;
        }
    }
#pragma unroll
    for (int g = 0; g < GT; ++g) acc[g] = (p0[g].x + p0[g].y) + (p1[g].x + p1[g].y);
}


// ---- the same for the 32-column slices (all hc blocks: 16 of the 24) in the PAIR-SPLIT layout ------------------------------
// The loop above is bound by shared-memory wavefronts: per 8 k a warp reads 8 wavefronts of weights and 2 x GT broadcast
// loads of the activations (every lane wants the same 8 k of x).  Here a lane owns TWO columns (2cp, 2cp+1) and HALF of the k
// (kh = lane >> 4 takes k-group 2j + kh), so a step needs ONE activation load per utterance (two addresses per warp, one
// wavefront) for the same number of FMAs: 8 + GT wavefronts instead of 8 + 2 GT.  Weight layout per 8-k block (1 KB, pack_decode):
// [column parity][k-group kh][column pair cp][4 k], so both weight loads of a warp are 512 contiguous bytes.  Partial sums of
// the two k-halves are added with one shuffle per accumulator after the last chunk.
template <int GT>
__device__ __forceinline__ void gemv_warp32(const float* __restrict__ wreg, const float* __restrict__ x, int kr8, float2 (&pe)[GT],
                                            float2 (&po)[GT]) {
    const int lane = threadIdx.x & 31, cp = lane & 15, kh = lane >> 4;
    const float* w = wreg + (kh * 16 + cp) * 4;
    const float* xs = x + kh * 4;
#pragma unroll 2
    for (int k = 0; k < kr8; k += 8) {
        const float4 wa = *reinterpret_cast<const float4*>(w);
        const float4 wb = *reinterpret_cast<const float4*>(w + 128);
        w += 256;
#pragma unroll
        for (int g = 0; g < GT; ++g) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + g * XLD + k);
            float2 a = pe[g], c = po[g];
            a = __ffma2_rn(make_float2(xv.x, xv.y), make_float2(wa.x, wa.y), a);
            c = __ffma2_rn(make_float2(xv.x, xv.y), make_float2(wb.x, wb.y), c);
            a = __ffma2_rn(make_float2(xv.z, xv.w), make_float2(wa.z, wa.w), a);
            c = __ffma2_rn(make_float2(xv.z, xv.w), make_float2(wb.z, wb.w), c);
            pe[g] = a; po[g] = c;
        }
    }
}


// ---- one block on ONE row per utterance -------------------------------------------------------------------
// in: S.xin[cb][g] = [taps | current row] of the block's input, S.prm[li&1] = its parameters (both prefetched).
// out: S.xin[cb^1][g][next_off ..] = the block's output row; this CTA's channel slice appended to the output history.
template <bool PROF, int GT>
__device__ __forceinline__ int layer_row(const DecParams& P, Smem& S, Stream& st, int li, int j, int b0, int G, int rank, int cb, unsigned& lcount) {
    const DecLayer& l = P.L[li];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool last = (li + 1 == P.nl);
    const int nl_next = last ? 0 : li + 1;
    const int pb = (int)(lcount & 1u);
    const uint32_t gpar = (lcount >> 1) & 1u;
    const uint32_t gbytes = (uint32_t)(GT * l.ns * 4);

    cp_async_wait<0>();                     // this block's taps and parameters (issued one block ago)
    __syncthreads();                        // ... and every warp has finished the previous block
    if (tid == 0) mbar_expect_tx(bar64(&S.gbar[pb]), NC * gbytes);
    // prefetch for the following block: its taps are rows of earlier frames, already in the history
    if (!last || j + 1 < P.steps) {
        prefetch_params(P, S, nl_next, rank);
        // the following block's input vector lives in xin[cb^1]; the attention writes the whole AudioDec C_1 input itself
        if (nl_next != P.n_enc) prefetch_taps(P, S, nl_next, last ? j + 1 : j, b0, G, cb ^ 1);
    }
    cp_async_commit();
    LAP(LP_START);

    if (l.ns == 32) {
        float2 pe[GT], po[GT];
#pragma unroll
        for (int g = 0; g < GT; ++g) { pe[g] = make_float2(0.f, 0.f); po[g] = make_float2(0.f, 0.f); }
        for (int c = 0; c < l.nch; ++c) {
            const DecChunk& ch = P.C[l.ch0 + c];
            const int slot = (int)(st.pos % DEC_NSLOT);
            mbar_wait(bar64(&S.fullw[slot][warp]), (st.pos / DEC_NSLOT) & 1u);
            LAP(LP_WAIT);
            const int kr8 = ch.krows >> 3;
            gemv_warp32<GT>(&S.ring[slot][warp][0], &S.xin[cb][0][ch.k0 + warp * kr8], kr8, pe, po);
            LAP(LP_GEMV);
            warp_release(P, S, st, warp, lane);
            LAP(LP_RELEASE);
        }
#pragma unroll
        for (int g = 0; g < GT; ++g) {                              // the two k-halves of the warp (lanes l and l ^ 16)
            float e = pe[g].x + pe[g].y, o = po[g].x + po[g].y;
            e += __shfl_xor_sync(0xffffffffu, e, 16);
            o += __shfl_xor_sync(0xffffffffu, o, 16);
            if (lane < 16) *reinterpret_cast<float2*>(&S.red[g][warp * 32 + 2 * lane]) = make_float2(e, o);
        }
    } else {
        float acc[GT];
#pragma unroll
        for (int g = 0; g < GT; ++g) acc[g] = 0.f;
        for (int c = 0; c < l.nch; ++c) {
            const DecChunk& ch = P.C[l.ch0 + c];
            const int slot = (int)(st.pos % DEC_NSLOT);
            mbar_wait(bar64(&S.fullw[slot][warp]), (st.pos / DEC_NSLOT) & 1u);
            LAP(LP_WAIT);
            const int kr8 = ch.krows >> 3;
            gemv_warp<GT>(&S.ring[slot][warp][0], &S.xin[cb][0][ch.k0 + warp * kr8], kr8, l.ns, acc);
            LAP(LP_GEMV);
            warp_release(P, S, st, warp, lane);
            LAP(LP_RELEASE);
        }
#pragma unroll
        for (int g = 0; g < GT; ++g) S.red[g][tid] = acc[g];
    }
    __syncthreads();
    // final sums of the slice (one thread per value) -> staging -> one bulk copy per peer (all-gather through distributed
    // shared memory).  With one utterance the values fit one warp, which then issues the copies without a second block barrier.
    const int lgns = (l.ns == 32) ? 5 : (l.ns == 16 ? 4 : 3);
    {
        const int nvals = GT << lgns, ng = NT >> lgns;
        float* ov = S.outv[pb];
        if (tid < nvals) {
            const float* bs = S.prm[li & 1] + 1024;
            const int g = tid >> lgns, n = tid & (l.ns - 1);
            const float* rp = &S.red[g][n];
            float s0 = bs[n], s1 = 0.f;
#pragma unroll 4
            for (int q = 0; q < ng; q += 2) { s0 += rp[q << lgns]; s1 += rp[(q + 1) << lgns]; }
            ov[tid] = s0 + s1;
            fence_proxy_async_smem();
        }
        if (GT > 1) __syncthreads(); else __syncwarp();
        if (tid < NC)
            bulk_s2peer(mapa(smem_u32(&S.pre[pb][rank][0]), (uint32_t)tid), ov, gbytes, mapa(smem_u32(&S.gbar[pb]), (uint32_t)tid));
    }
    LAP(LP_GATHER);
    mbar_wait(bar64(&S.gbar[pb]), gpar);
    LAP(LP_CBAR);
    // LayerNorm statistics: one warp per (utterance, half), pivoted single pass (pivot = channel 0: a constant row gives exactly 0,
    // quirk Q4); then gate / highway mix one thread per channel.  Redundantly in every CTA: the next block's input is local.
    {
        const int C = l.cout, cs = l.cs, nh = l.kind + 1;
        const float rC = (C == 256) ? (1.0f / 256.0f) : __frcp_rn((float)C);
        // channel c lives in the slice of rank c / cs at column c % cs (cs = 16, or 5 for the n_mels-wide last block)
        auto pre_off = [&](int c) { const int rk = (cs == 16) ? (c >> 4) : ((c * 205) >> 10); return rk * PLD + (c - rk * cs); };
        for (int pr = warp; pr < GT * nh; pr += NWARP) {
            const int g = (nh == 2) ? (pr >> 1) : pr, hf = (nh == 2) ? (pr & 1) : 0;
            const float* base = &S.pre[pb][0][(g << lgns) + hf * cs];
            const float piv = base[0];
            float sv = 0.f, qv = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = lane + 32 * i;
                if (c < C) { const float d = base[pre_off(c)] - piv; sv += d; qv = fmaf(d, d, qv); }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, o); qv += __shfl_xor_sync(0xffffffffu, qv, o); }
            if (lane == 0) {
                const float m = sv * rC;
                S.stat[g][hf][0] = piv + m;
                S.stat[g][hf][1] = rsqrtf(fmaxf(qv * rC - m * m, 0.f) + 1e-12f);
            }
        }
        __syncthreads();
        LAP(LP_LN);
        const float* prm = S.prm[li & 1];
        const int cur_off = (l.ntaps - 1) * 256;
        const int next_off = (last || li + 1 == P.n_enc) ? 0 : (P.L[li + 1].ntaps - 1) * 256;
        float* oh = P.out_hist[li];
        const int c = tid;
        if (c < C) {
            const int po = pre_off(c);
            const bool mine = (po / PLD == rank) && oh;
            const float g1 = prm[c], b1 = prm[256 + c], g2 = prm[512 + c], b2 = prm[768 + c];
            float o[GT];
#pragma unroll
            for (int g = 0; g < GT; ++g) {
                const float* pr = &S.pre[pb][0][g << lgns];
                o[g] = (pr[po] - S.stat[g][0][0]) * S.stat[g][0][1] * g1 + b1;
            }
            if (l.kind == 1) {
#pragma unroll
                for (int g = 0; g < GT; ++g) {
                    const float* pr = &S.pre[pb][0][g << lgns];
                    const float h2 = (pr[po + cs] - S.stat[g][1][0]) * S.stat[g][1][1] * g2 + b2;
                    const float h1 = sigmoid_fast(o[g]);
                    o[g] = h1 * h2 + (1.0f - h1) * S.xin[cb][g][cur_off + c];
                }
            } else if (l.act == 1) {
#pragma unroll
                for (int g = 0; g < GT; ++g) o[g] = fmaxf(o[g], 0.f);
            }
#pragma unroll
            for (int g = 0; g < GT; ++g) {
                const size_t row = (size_t)(b0 + g) * P.T + j;
                if (mine && g < G) oh[row * C + c] = o[g];           // this CTA's slice of the history row
                if (last) {                                           // Y = sigmoid(logits), networks.py:210; next frame's AudioEnc input
                    o[g] = sigmoid_fast(o[g]);
                    if (rank == 0 && g < G) P.ybuf[row * C + c] = o[g];
                }
                S.xin[cb ^ 1][g][next_off + c] = (g < G) ? o[g] : 0.f;   // unused slots stay zero
            }
        } else if (last && c < 128) {                                 // AudioEnc C_1 reads K = 128 padded channels
#pragma unroll
            for (int g = 0; g < GT; ++g) S.xin[cb ^ 1][g][c] = 0.f;
        }
    }
    LAP(LP_MIX);
    lcount++;
    return cb ^ 1;
}

// ---- attention of ONE query row under the 3-key window (networks.py:140-153) -------------------------------
// lane holds q[lane*8 .. +8); returns ctx[8] in the same layout and the argmax key (first index among equal maxima)
__device__ __forceinline__ int attend_row(const DecParams& P, const float (&qv)[8], int b, int p, int lane, float (&ctx)[8]) {
    const int d = P.d;
    const int n_lo = min(max(p, 0), P.N - 1), n_hi = min(n_lo + P.win_size, P.N);
    const float scale = rsqrtf((float)d);
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n_lo + i;
        if (n < n_hi) {
            const float* k = P.kv + ((size_t)b * P.N + n) * (2 * d) + lane * 8;
            const float4 k0 = ldcg4(k), k1 = ldcg4(k + 4);
            float s = 0.f;
            s = fmaf(qv[0], k0.x, s); s = fmaf(qv[1], k0.y, s); s = fmaf(qv[2], k0.z, s); s = fmaf(qv[3], k0.w, s);
            s = fmaf(qv[4], k1.x, s); s = fmaf(qv[5], k1.y, s); s = fmaf(qv[6], k1.z, s); s = fmaf(qv[7], k1.w, s);
            sc[i] = warp_sum(s) * scale;
        }
    }
    const int cnt = n_hi - n_lo;
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < cnt) mx = fmaxf(mx, sc[i]);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < cnt) { sc[i] = expf(sc[i] - mx); sum += sc[i]; }
    float best = -1.f; int besti = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < cnt) { sc[i] = sc[i] / sum; if (sc[i] > best) { best = sc[i]; besti = i; } }
#pragma unroll
    for (int i = 0; i < 8; ++i) ctx[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n_lo + i;
        if (n < n_hi) {
            const float* v = P.kv + ((size_t)b * P.N + n) * (2 * d) + d + lane * 8;
            const float4 v0 = ldcg4(v), v1 = ldcg4(v + 4);
            const float p_ = sc[i];
            ctx[0] = fmaf(p_, v0.x, ctx[0]); ctx[1] = fmaf(p_, v0.y, ctx[1]); ctx[2] = fmaf(p_, v0.z, ctx[2]); ctx[3] = fmaf(p_, v0.w, ctx[3]);
            ctx[4] = fmaf(p_, v1.x, ctx[4]); ctx[5] = fmaf(p_, v1.y, ctx[5]); ctx[6] = fmaf(p_, v1.z, ctx[6]); ctx[7] = fmaf(p_, v1.w, ctx[7]);
        }
    }
    return n_lo + besti;
}

// ---- pre-pass: refresh the receptive field (rows t < j) of the utterances whose window moved -------------------
// every moved utterance refreshes the SAME rows t_lo .. j-1 (t_lo depends on the block and the frame only), so the row list is
// (bit mask of moved utterances, first row, rows per utterance): no local arrays (see the note on the stack above)
struct PreRows { unsigned mask; int t_lo, n, total; };
__device__ __forceinline__ PreRows pre_rows(const Smem& S, int G, int j, int prow) {
    PreRows r;
    r.mask = 0;
#pragma unroll
    for (int g = 0; g < GMAX; ++g) if (g < G && S.moved[g]) r.mask |= 1u << g;
    r.t_lo = max(0, j - (prow - 1));
    r.n = j - r.t_lo;
    r.total = r.n * __popc(r.mask);
    return r;
}
// scratch row m -> (utterance, time row): the k-th moved utterance owns rows [k*n, (k+1)*n)
__device__ __forceinline__ void pre_row_of(const PreRows& r, int m, int& g, int& t) {
    const int k = m / r.n;
    unsigned mk = r.mask;
    for (int i = 0; i < k; ++i) mk &= mk - 1;                        // drop the k lowest set bits
    g = __ffs(mk) - 1;
    t = r.t_lo + (m - k * r.n);
}
__device__ __forceinline__ int pre_off_of(const PreRows& r, int g) { return r.n * __popc(r.mask & ((1u << g) - 1u)); }
// address of W[k][n] (k = row within the layer's K) inside the ring; the layer's chunks occupy consecutive slots from pos0
// ---- the pre-pass GEMM on the 5th-generation tensor cores ---------------------------------------------------------------
// One utterance, <= 96 source rows.  A = the source rows as split-fp16 planes (hi = fp16(x), lo = fp16(x - hi)), staged per
// 16-channel slab in the NO-SWIZZLE K-major core-matrix layout [k8][row][8 halfs]: rows are consecutive 16-byte chunks, so
// the three taps of the dilated conv are the SAME slab read through descriptors whose start address is shifted by
// tap * rate rows -- staged once, multiplied three times.  B = this CTA's weight columns, pre-packed in the same layout
// ([plane][k8][column][8 halfs], 2 KB per 16-k slab) and streamed through the ring like the fp32 weights.  D = 128 x ns fp32
// in tensor memory; per slab and tap hi*Whi + hi*Wlo + lo*Whi (the dropped lo*lo term is 2^-22 relative).
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;               // between the two 8-wide k groups of one MMA
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;               // between 8-row groups
    d |= 1ull << 46;                                                // descriptor version (sm_100); layout type 0 = no swizzle
    return d;
}
constexpr int TC_LOADW = 3;                                         // warps 0..2 stage A (one thread per source row, <= 96 rows)
constexpr int TC_NISS = 3;                                          // lane 0 of warps 3, 4, 5: one MMA issuer per split-fp16 product
constexpr int TC_COLS = 128;                                        // tensor-memory columns: three 32-column accumulators (power of two)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// (descriptor start-address field, i.e. address >> 4) of every weight slab of block li, tap-major: S.tc_baddr[tap * nslab + ks].
// One division chain per entry, computed by 96 threads in parallel, OFF the single-thread MMA issue paths.
__device__ __forceinline__ void pyr_tc_table(const DecParams& P, Smem& S, int li, unsigned pos0) {
    const DecLayer& l = P.L[li];
    const int nslab = l.cin / 16, spc = l.krows / 16, spr = spc / 8, slab_f = 16 * l.ns;
    const int sid = (int)threadIdx.x;
    if (sid < l.ntaps * nslab) {
        const int c = sid / spc, wi = sid - c * spc, reg = wi / spr, jj = wi - reg * spr;
        S.tc_baddr[sid] = (smem_u32(&S.ring[(pos0 + c) % DEC_NSLOT][reg][jj * slab_f]) & 0x3FFFFu) >> 4;
    }
}

// The slab pipeline has no block barrier.  Loader warps fill stage s and arrive on abar[s].  THREE issuer threads (in three
// warps, so on three schedulers) wait for abar[s]; issuer i issues product i of the split-fp16 scheme for every tap --
// 0: hi x Whi, 1: hi x Wlo, 2: lo x Whi -- into ITS OWN 32-column accumulator and commits to sbar[s] (count 3), which the
// loaders wait for before they overwrite the stage.  One thread issuing all nine MMAs of a slab was the bottleneck of the
// whole pre-pass (ncu: the loaders spent their time waiting for the stage to be released).  Slabs are numbered through the
// whole launch (st.tcq): slab q lives in stage q % TC_NSTG and is that stage's (q / TC_NSTG)-th use, which gives every wait
// its phase parity without any shared counter.
__device__ __forceinline__ void pyr_tc_utt(const DecParams& P, Smem& S, int li, unsigned q0, unsigned acc_use, int b, int t_lo, int n_out,
                                        int rank, float* scr_rows) {
    const DecLayer& l = P.L[li];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int halo = (l.ntaps - 1) * l.rate, n_src = n_out + halo;   // <= 96
    const int nslab = l.cin / 16, ns = l.ns;
    unsigned char* As = reinterpret_cast<unsigned char*>(S.wrk);
    unsigned use = q0 / TC_NSTG;
    int stg = (int)(q0 - use * TC_NSTG);
    const uint32_t tacc = S.tmem_base;
    if (warp < TC_LOADW) {
        const bool loader = tid < n_src;
        const int t_src = t_lo - halo + tid;
        const float* src = P.in_hist[li] + ((size_t)b * P.T + (t_src < 0 ? 0 : t_src)) * l.ldin;
        const bool have = loader && t_src >= 0;                      // rows before the utterance start: TF zero padding
        // The loop is bound by the latency of the activation loads (L2, ~700 cycles), not by the conversion or the MMAs: each
        // thread keeps FOUR slabs (4 x 64 bytes of its row) in flight in registers.  (nslab is 16 or 32.)
        constexpr int PF = 4;
        float4 r[PF][4];
#pragma unroll
        for (int u = 0; u < PF; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) r[u][i] = (have && u < nslab) ? ldcg4(src + u * 16 + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned char* dst0 = As + tid * 16;
#pragma unroll 1
        for (int ks0 = 0; ks0 < nslab; ks0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (ks0 + u < nslab) {
                    if (use > 0) mbar_wait(bar64(&S.sbar[stg]), (use - 1) & 1u);   // the MMAs that read this stage are done
                    if (loader) {
                        unsigned char* dst = dst0 + stg * TC_ASTAGE;
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {              // the two 8-channel k groups of the slab
                            const float4 p0 = r[u][2 * h8], p1 = r[u][2 * h8 + 1];
                            const float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                            __align__(16) __half hi[8];
                            __align__(16) __half lo[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) { hi[i] = __float2half_rn(v[i]); lo[i] = __float2half_rn(v[i] - __half2float(hi[i])); }
                            *reinterpret_cast<uint4*>(dst + h8 * (TC_RA * 16)) = *reinterpret_cast<const uint4*>(hi);
                            *reinterpret_cast<uint4*>(dst + h8 * (TC_RA * 16) + TC_APLANE) = *reinterpret_cast<const uint4*>(lo);
                        }
                    }
                    if (have && ks0 + u + PF < nslab) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) r[u][i] = ldcg4(src + (ks0 + u + PF) * 16 + i * 4);
                    }
                    fence_proxy_async_smem();                         // generic-proxy stores -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar64(&S.abar[stg]));
                    if (++stg == TC_NSTG) { stg = 0; ++use; }
                }
            }
        }
    } else if (warp < TC_LOADW + TC_NISS && lane == 0) {
        const int prod = warp - TC_LOADW;                            // 0: hi x Whi, 1: hi x Wlo, 2: lo x Whi
        const uint32_t idesc = umma_idesc_f16(128, (uint32_t)ns);
        const uint64_t dA0 = umma_desc_noswz(0, TC_RA * 16, 128), dB0 = umma_desc_noswz(0, (uint32_t)ns * 16, 128);
        const uint32_t a_base = ((smem_u32(As) & 0x3FFFFu) >> 4) + (prod == 2 ? (TC_APLANE >> 4) : 0u);
        const uint32_t b_plane = prod == 1 ? (uint32_t)ns * 2 : 0u, tap_step = (uint32_t)l.rate;
        const uint32_t dacc = tacc + (uint32_t)prod * 32u;
        const int ntaps = l.ntaps;
#pragma unroll 1
        for (int ks = 0; ks < nslab; ++ks) {
            mbar_wait(bar64(&S.abar[stg]), use & 1u);                 // the slab is staged (all loader warps arrived)
            tc_fence_after();
            uint32_t aa = a_base + (uint32_t)stg * (TC_ASTAGE >> 4);
            const uint32_t* bt = &S.tc_baddr[ks];
#pragma unroll 1
            for (int tap = 0; tap < ntaps; ++tap, aa += tap_step, bt += nslab)
                tc_mma_f16(dacc, dA0 | aa, dB0 | (*bt + b_plane), idesc, (ks | tap) != 0);
            tc_commit(bar64(&S.sbar[stg]));                           // the stage may be overwritten once all three issuers' MMAs have read it
            if (++stg == TC_NSTG) { stg = 0; ++use; }
        }
        tc_commit(bar64(&S.dbar));                                    // this product's accumulator is complete
    }
    // epilogue: thread == output row (TMEM lane); the three partial accumulators summed, scaled, + bias -> scratch
    if (warp < 4) {
        mbar_wait(bar64(&S.dbar), acc_use & 1u);
        tc_fence_after();
        const int m = tid;
        const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
        const float inv = P.inv_scale[li];
        float v[32];
        if (ns == 32) {
            float w[32];
            tmem_ld32_nowait(taddr, v); tmem_ld32_nowait(taddr + 32, w); tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
            tmem_ld32_nowait(taddr + 64, w); tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += w[i];
        } else {
            float w16[16], x16[16];
            tmem_ld16(taddr, w16); tmem_ld16(taddr + 32, x16);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = w16[i] + x16[i];
            tmem_ld16(taddr + 64, x16);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += x16[i];
        }
        if (m < n_out) {
            float* orow = scr_rows + (size_t)m * 512;
            const float* bs = P.bias[li];
            if (l.kind == 1) {                                        // columns [0,16) gate, [16,32) info of channels rank*16..
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int n = hf * 16 + q4 * 4, col = hf * 256 + rank * 16 + q4 * 4;
                        const float4 bq = __ldg(reinterpret_cast<const float4*>(bs + hf * l.cout + rank * 16 + q4 * 4));
                        *reinterpret_cast<float4*>(orow + col) = make_float4(fmaf(v[n], inv, bq.x), fmaf(v[n + 1], inv, bq.y),
                                                                             fmaf(v[n + 2], inv, bq.z), fmaf(v[n + 3], inv, bq.w));
                    }
            } else {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int col = rank * 16 + q4 * 4;
                    const float4 bq = __ldg(reinterpret_cast<const float4*>(bs + col));
                    *reinterpret_cast<float4*>(orow + col) = make_float4(fmaf(v[q4 * 4], inv, bq.x), fmaf(v[q4 * 4 + 1], inv, bq.y),
                                                                         fmaf(v[q4 * 4 + 2], inv, bq.z), fmaf(v[q4 * 4 + 3], inv, bq.w));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();                                                  // the accumulators have been read: the next utterance may overwrite them
}

// LayerNorm / gate / highway mix of the refreshed rows: one warp per row over the whole cluster (parameters in S.red)
__device__ __forceinline__ void pyr_ln(const DecParams& P, Smem& S, int li, int b0, const PreRows& rl, int rank, const float* scr) {
    const DecLayer& l = P.L[li];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* prm = &S.red[0][0];
    const float fC = 256.f;
    for (int m = rank * NWARP + warp; m < rl.total; m += NC * NWARP) {
        int g, t; pre_row_of(rl, m, g, t);
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

// ---- the pre-pass as ONE out-of-line function: cold code (most frames do not have it) kept out of the per-frame
// instruction stream.  The stream cursors go in and come back by value; everything else lives in shared memory.
template <bool PROF>
__device__ __noinline__ Stream prepass(const DecParams& P, Smem& S, Stream st, int j, int b0, int G, int rank, float* scr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

            // ---- pre-pass: rows t < j of the moved utterances under the new window (uniform branch: see the file header) ----
            if (threadIdx.x == 0) { S.n_moved_frames++; for (int g = 0; g < G; ++g) S.n_moved_utt += S.moved[g]; }
            const PreRows ra = pre_rows(S, G, j, P.L[P.n_enc].prow);
            const float* Qh = P.out_hist[P.n_enc - 1];
            for (int m = rank * NWARP + warp; m < ra.total; m += NC * NWARP) {
                int g, t; pre_row_of(ra, m, g, t);
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
            for (int lp = P.n_enc; lp < P.nl && P.L[lp].prow > 1; ++lp) {
                const DecLayer& l = P.L[lp];
                // every warp waits for ALL regions of the block's chunks (lane w watches region w)
                for (int c = 0; c < l.nch; ++c) {
                    const unsigned pp = st.pos + c;
                    if (lane < NWARP) mbar_wait(bar64(&S.fullw[pp % DEC_NSLOT][lane]), (pp / DEC_NSLOT) & 1u);
                }
                __syncwarp();
                const PreRows rl = pre_rows(S, G, j, l.prow);
                if (rl.n > 0) {
                    pyr_tc_table(P, S, lp, st.pos);
                    __syncthreads();
                    for (int g = 0; g < G; ++g) {
                        if (!((rl.mask >> g) & 1u)) continue;
                        float* rows = scr + (size_t)pre_off_of(rl, g) * 512;
                        pyr_tc_utt(P, S, lp, st.tcq, st.tca, b0 + g, rl.t_lo, rl.n, rank, rows);
                        st.tcq += l.cin / 16; st.tca++;
                    }
                }
                __syncthreads();                          // every warp is done with every region of these chunks
                for (int c = 0; c < l.nch; ++c) {
                    if (lane == 0) { fence_proxy_async_smem(); stream_issue(P, S, st, st.prod, (int)(st.pos % DEC_NSLOT), warp); }
                    stream_advance(P, S, st);
                }
                for (int i = tid; i < 256; i += NT)       // this block's LayerNorm parameters for pyr_ln
                    *reinterpret_cast<float4*>(&S.red[0][0] + i * 4) = __ldg(reinterpret_cast<const float4*>(P.lnp[lp]) + i);
                LAP(LP_PYR_GEMM);
                cluster_sync_all();
                LAP(LP_PYR_BAR);
                pyr_ln(P, S, lp, b0, rl, rank, scr);
                LAP(LP_PYR_LN);
                cluster_sync_all();
                LAP(LP_PYR_BAR);
            }
            return st;
}

template <bool PROF, int GT>
__global__ void __cluster_dims__(DEC_NC, 1, 1) __launch_bounds__(DEC_THREADS, 1)
decode_cluster_kernel(const __grid_constant__ DecParams Pc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const int* src = reinterpret_cast<const int*>(&Pc);
        int* dst = reinterpret_cast<int*>(&S.P);
        for (int i = tid; i < (int)(sizeof(DecParams) / 4); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const DecParams& P = S.P;
    const int rank = (int)cluster_ctarank();
    const int cluster = blockIdx.x / NC;
    const int b0 = cluster * P.G;
    const int G = min(P.G, P.B - b0);
    float* scr = P.pre_scr + (size_t)cluster * P.G * 85 * 512;

    if (tid == 0) {
        for (int s = 0; s < DEC_NSLOT; ++s)
            for (int w = 0; w < NWARP; ++w) mbar_init(bar64(&S.fullw[s][w]), 1);
        mbar_init(bar64(&S.gbar[0]), 1); mbar_init(bar64(&S.gbar[1]), 1);
        for (int i = 0; i < TC_NSTG; ++i) { mbar_init(bar64(&S.sbar[i]), TC_NISS); mbar_init(bar64(&S.abar[i]), TC_LOADW); }
        mbar_init(bar64(&S.dbar), TC_NISS);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<TC_COLS>(&S.tmem_base);          // 128 lanes x 32 fp32 columns: the pre-pass accumulator
    for (int i = tid; i < 2 * GMAX * XLD; i += NT) (&S.xin[0][0][0])[i] = 0.f;
    for (int i = tid; i < 2 * NC * PLD; i += NT) (&S.pre[0][0][0])[i] = 0.f;
    if (tid < GMAX) { S.p_cur[tid] = 0; S.p_prev[tid] = 0; S.p_next[tid] = 0; S.moved[tid] = 0; }
    if (tid < 2) S.fmoved[tid] = 0;
    if (tid < 16) S.prof[tid] = 0;
    if (tid == 0) { S.n_moved_frames = 0; S.n_moved_utt = 0; }
    __syncthreads();

    Stream st;
    st.base = P.wstream + (size_t)rank * P.stream_len;
    st.prod = Cur{0, 0, 0}; st.pos = 0; st.tcq = 0; st.tca = 0;
    for (int s = 0; s < DEC_NSLOT; ++s) {                // the first chunks are AudioEnc chunks of frame 0 (nch_enc > DEC_NSLOT)
        if (lane == 0) stream_issue(P, S, st, st.prod, s, warp);
        cur_next(P, S, st.prod);
    }
    prefetch_params(P, S, 0, rank);
    cp_async_commit();
    cluster_sync_all();                                   // every CTA of the cluster is running: DSMEM and its barriers exist

    int cb = 0;
    unsigned lcount = 0;
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0) S.prof_last = clock64();
    for (int j = 0; j < P.steps; ++j) {
        if (tid < GMAX) S.moved[tid] = (tid < G && j > 0 && S.p_cur[tid] != S.p_prev[tid]) ? 1 : 0;
        if (tid == 0) {
            int any = 0;
            for (int g = 0; g < G; ++g) any |= (j > 0 && S.p_cur[g] != S.p_prev[g]) ? 1 : 0;
            S.fmoved[j & 1] = any;
        }
        __syncthreads();
        const bool any_moved = S.fmoved[j & 1] != 0;
        if (rank == 0 && tid < G) P.p_hist[(size_t)(b0 + tid) * P.T + j] = S.p_cur[tid];

        // AudioEnc (networks.py:81-124; input Y[j-1], train.py:51, already in xin[cb]), Attention, AudioDec (networks.py:166-212):
        // ONE copy of the block body in the instruction stream -- the per-frame loop has to stay inside the instruction cache
        for (int li = 0; li < P.nl; ++li) {
        if (li == P.n_enc) {
        // Attention of row j under the current window, redundantly in every CTA: R[j] = [A.V ; Q] (networks.py:140-153)
        __syncthreads();                                  // Q[j] was written one thread per channel
        if (warp < G) {
            const int g = warp;
            float qv[8], ctx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] = S.xin[cb][g][lane * 8 + i];
            const int amax = attend_row(P, qv, b0 + g, S.p_cur[g], lane, ctx);
            if (lane == 0) S.p_next[g] = amax;
#pragma unroll
            for (int i = 0; i < 8; ++i) { S.xin[cb ^ 1][g][lane * 8 + i] = ctx[i]; S.xin[cb ^ 1][g][P.d + lane * 8 + i] = qv[i]; }
        }
        cb ^= 1;
        LAP(LP_ATT);

        if (any_moved) st = prepass<PROF>(P, S, st, j, b0, G, rank, scr);
        }   // li == n_enc
        cb = layer_row<PROF, GT>(P, S, st, li, j, b0, G, rank, cb, lcount);
        }   // blocks

        __syncthreads();
        if (tid < GMAX) { S.p_prev[tid] = S.p_cur[tid]; S.p_cur[tid] = S.p_next[tid]; }
        cluster_sync_all();                               // this frame's history rows are visible to the whole cluster
        LAP(LP_FRAME);
    }
    if (rank == 0 && tid < G) P.p_final[b0 + tid] = S.p_cur[tid];
    if (rank == 0 && tid == 0 && P.stats) { P.stats[2 * cluster] = S.n_moved_frames; P.stats[2 * cluster + 1] = S.n_moved_utt; }
    if (PROF && P.prof && cluster == 0 && rank == 0 && tid < 16) P.prof[tid] = S.prof[tid];
    cp_async_wait<0>();
    cluster_sync_all();                                   // no CTA exits while a peer may still write into its shared memory
    if (warp == 0) tmem_dealloc<TC_COLS>(S.tmem_base);
}

size_t decode_smem_bytes() { return sizeof(Smem) + 128; }

using DecKernel = void (*)(DecParams);
// one instantiation per (lap timers, utterances per cluster); exactly one of them runs in a launch
static DecKernel decode_kernel_of(bool prof, int G) {
    switch (G) {
        case 1: return prof ? decode_cluster_kernel<true, 1> : decode_cluster_kernel<false, 1>;
        case 2: return prof ? decode_cluster_kernel<true, 2> : decode_cluster_kernel<false, 2>;
        case 3: return prof ? decode_cluster_kernel<true, 3> : decode_cluster_kernel<false, 3>;
        case 4: return prof ? decode_cluster_kernel<true, 4> : decode_cluster_kernel<false, 4>;
        default: return prof ? decode_cluster_kernel<true, 5> : decode_cluster_kernel<false, 5>;
    }
}
static_assert(DEC_GMAX == 5, "decode_kernel_of: one instantiation per utterance count");

static cudaError_t decode_prepare() {
    static std::atomic<bool> done[64];                    // per device: the attributes stick to the (device, function) pair
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev].load(std::memory_order_acquire)) return cudaSuccess;
    for (int G = 1; G <= DEC_GMAX; ++G)
        for (int prof = 0; prof < 2; ++prof) {
            DecKernel k = decode_kernel_of(prof != 0, G);
            e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_smem_bytes());
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
    if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release);
    return e;
}

int decode_max_active_clusters() {
    if (decode_prepare() != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(DEC_NC); cfg.blockDim = dim3(DEC_THREADS); cfg.dynamicSmemBytes = decode_smem_bytes();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = DEC_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, decode_kernel_of(false, DEC_GMAX), &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

cudaError_t launch_decode_cluster(const DecParams& p, int n_clusters, cudaStream_t s) {
    cudaError_t e = decode_prepare();                     // per call: the attribute is per device, handles may live on several
    if (e != cudaSuccess) return e;
    if (p.G < 1 || p.G > DEC_GMAX) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_clusters * DEC_NC); cfg.blockDim = dim3(DEC_THREADS); cfg.dynamicSmemBytes = decode_smem_bytes(); cfg.stream = s;
    cfg.attrs = nullptr; cfg.numAttrs = 0;               // cluster dims are compiled in (__cluster_dims__)
    return cudaLaunchKernelEx(&cfg, decode_kernel_of(p.prof != nullptr, p.G), p);
}

}  // namespace dctts
