// kernels_chain.cu -- a whole CHAIN of reference blocks in one cluster-persistent kernel.
//
// The autoregressive step (reference synthesize.py:48-54) is a strictly sequential chain of
// 13 AudioEnc blocks (networks.py:81-124) on ONE new row per utterance, and -- after the wide
// part of the AudioDec pyramid -- 7 more blocks (networks.py:177-209) on 5,3,1,1,1,1,1 rows.
// Launched block by block that is 40 tiny, latency-bound kernels per step.  Here a
// thread-block cluster of 8 CTAs walks the chain by itself:
//   * every CTA owns 1/8 of the output channels of every block (for `hc` the gate and the info
//     columns of the SAME channels, so the highway mix needs no exchange);
//   * the block's weight slice streams from L2 through a 6-stage cp.async ring (96 KB in
//     flight per SM -- the decode is weight-bandwidth bound);
//   * LayerNorm statistics are merged across the cluster through distributed shared memory,
//     and two barrier.cluster phases per block replace two kernel boundaries.
// One cluster serves up to 16 rows (G utterances x R rows); fp32 throughout (this is the
// latency path -- the wide pyramid rows go to the tcgen05 kernel).
#include "kernels.cuh"
#include "tc_ptx.cuh"

#include <stdexcept>
#include <string>

namespace dctts {

using namespace ptx;

constexpr int CH_NC = 8;          // cluster size
constexpr int CH_THREADS = 256;
constexpr int CH_STAGES = 6;
constexpr int CH_BM = 16;         // rows per cluster
constexpr int CH_BN = 64;         // conv columns per CTA
constexpr int CH_BK = 64;         // reduction slice per stage
constexpr int CH_STAGE_FLOATS = CH_BK * CH_BN;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// conv column of accumulator slot `slot` for CTA `rank` (or -1): hc -> [gate | info] of the CTA's channels
__device__ __forceinline__ int chain_col(const ChainLayer& L, int rank, int slot, int per) {
    if (L.kind == 1) {
        if (slot < per) return rank * per + slot;
        if (slot < 2 * per) return L.C + rank * per + (slot - per);
        return -1;
    }
    const int col = rank * per + slot;
    return (slot < per && col < L.C) ? col : -1;
}

constexpr int CH_MAXCH = 12;      // reduction chunks per block (3 taps x 256 channels)

// The reduction loop of one block for MR (4, 8 or 16) rows: weights from the cp.async ring,
// the whole (pre-transposed) A operand of the block already resident in shared memory.
template <int MR>
__device__ __forceinline__ void chain_gemm(const float* __restrict__ Ws, const float* __restrict__ Afull, int nch,
                                           int c, int kg, float (&acc)[CH_BM], const ChainLayer& L, int rank, int per,
                                           int KC, int tid, bool vec_ok) {
    auto issue_w = [&](int ch) {
        if (ch < nch) {
            const int tap = ch / KC, k0 = (ch - tap * KC) * CH_BK;
            const float* Wt = L.W + (size_t)tap * L.K * L.ldw;
            float* st = const_cast<float*>(Ws) + (size_t)(ch % CH_STAGES) * CH_STAGE_FLOATS;
            if (vec_ok) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = tid + i * CH_THREADS;
                    const int k = idx / 16, slot = (idx % 16) * 4;
                    const int col = chain_col(L, rank, slot, per);
                    const bool ok = (col >= 0) && (k0 + k < L.K);
                    const float* src = ok ? Wt + (size_t)(k0 + k) * L.ldw + col : Wt;
                    cp_async16(st + k * CH_BN + slot, src, ok ? 16 : 0);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int idx = tid + i * CH_THREADS;
                    const int k = idx / CH_BN, slot = idx % CH_BN;
                    const int col = chain_col(L, rank, slot, per);
                    const bool ok = (col >= 0) && (k0 + k < L.K);
                    const float* src = ok ? Wt + (size_t)(k0 + k) * L.ldw + col : Wt;
                    cp_async4(st + k * CH_BN + slot, src, ok ? 4 : 0);
                }
            }
        }
        cp_async_commit();
    };
    // (the first CH_STAGES-1 chunks were issued by the caller's prologue through the same lambda shape)
    for (int ch = 0; ch < nch; ++ch) {
        cp_async_wait<CH_STAGES - 2>();                         // this thread's part of chunk `ch` has landed
        __syncthreads();                                        // ... everybody's; and chunk ch-1 is fully consumed
        issue_w(ch + CH_STAGES - 1);                            // refill the stage chunk ch-1 just released
        const float* st = Ws + (size_t)(ch % CH_STAGES) * CH_STAGE_FLOATS + (kg * 16) * CH_BN + c;
        const float* ab = Afull + ((size_t)ch * CH_BK + kg * 16) * CH_BM;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float ww = st[kk * CH_BN];
            const float4* xr = reinterpret_cast<const float4*>(ab + kk * CH_BM);
#pragma unroll
            for (int g = 0; g < MR / 4; ++g) {
                const float4 x = xr[g];
                acc[4 * g + 0] = fmaf(x.x, ww, acc[4 * g + 0]); acc[4 * g + 1] = fmaf(x.y, ww, acc[4 * g + 1]);
                acc[4 * g + 2] = fmaf(x.z, ww, acc[4 * g + 2]); acc[4 * g + 3] = fmaf(x.w, ww, acc[4 * g + 3]);
            }
        }
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const ChainArgs a) {
    extern __shared__ __align__(16) float dyn[];
    float* Ws = dyn;                                            // [CH_STAGES][CH_BK][CH_BN]
    float* Afull = dyn + CH_STAGES * CH_STAGE_FLOATS;           // [CH_MAXCH][CH_BK][CH_BM], k-major (transposed)
    __shared__ float red[3][CH_BM][CH_BN];
    __shared__ float ytile[CH_BM][CH_BN];
    __shared__ float4 part[CH_NC][CH_BM];
    __shared__ float stat[CH_BM][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    const int b0 = blockIdx.y * a.G;
    const int nb = min(a.G, a.B - b0);
    const int T = a.T;
    const int t_end = a.jptr ? *a.jptr : T - 1;
    const int c = tid % CH_BN, kg = tid / CH_BN;
    const int lrow = tid % CH_BM, lkq = (tid / CH_BM) * 4;

    cluster_arrive();                                           // phase 0: all CTAs of the cluster are running
    cluster_wait();

    for (int li = 0; li < a.nlayers; ++li) {
        const ChainLayer& L = a.L[li];
        const int R = L.R;
        const int M = nb * R;                                   // rows of this block (<= 16)
        const int per = (L.kind == 1) ? L.C / CH_NC : (L.C + CH_NC - 1) / CH_NC;   // channels (hc) / columns per CTA
        const int my_col = chain_col(L, rank, c, per);
        const int KC = (L.K + CH_BK - 1) / CH_BK;
        const int nch = L.ntaps * KC;
        const bool vec_ok = (per % 4) == 0;
        const int n1 = (L.kind == 1) ? per : max(0, min(per, L.C - rank * per));
        const int n2 = (L.kind == 1) ? per : 0;

        // ---- prologue: everything this block needs is put in flight at once ----
        // (1) first weight chunks
        {
            auto issue0 = [&](int ch) {
                if (ch < nch) {
                    const int tap = ch / KC, k0 = (ch - tap * KC) * CH_BK;
                    const float* Wt = L.W + (size_t)tap * L.K * L.ldw;
                    float* st = Ws + (size_t)(ch % CH_STAGES) * CH_STAGE_FLOATS;
                    if (vec_ok) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int idx = tid + i * CH_THREADS;
                            const int k = idx / 16, slot = (idx % 16) * 4;
                            const int col = chain_col(L, rank, slot, per);
                            const bool ok = (col >= 0) && (k0 + k < L.K);
                            const float* src = ok ? Wt + (size_t)(k0 + k) * L.ldw + col : Wt;
                            cp_async16(st + k * CH_BN + slot, src, ok ? 16 : 0);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int idx = tid + i * CH_THREADS;
                            const int k = idx / CH_BN, slot = idx % CH_BN;
                            const int col = chain_col(L, rank, slot, per);
                            const bool ok = (col >= 0) && (k0 + k < L.K);
                            const float* src = ok ? Wt + (size_t)(k0 + k) * L.ldw + col : Wt;
                            cp_async4(st + k * CH_BN + slot, src, ok ? 4 : 0);
                        }
                    }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int s = 0; s < CH_STAGES - 1; ++s) issue0(s);
        }
        // (2) the whole A operand: thread (lrow, lkq) loads its float4 of every chunk
        int lb = 0, lt = -1;
        if (lrow < M) { lb = b0 + lrow / R; lt = t_end - (R - 1) + (lrow % R); }
        float4 areg[CH_MAXCH];
#pragma unroll
        for (int ch = 0; ch < CH_MAXCH; ++ch) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch < nch) {
                const int tap = ch / KC, k = (ch - tap * KC) * CH_BK + lkq;
                const int ts = lt + L.shifts[tap];
                if (lt >= 0 && ts >= 0 && ts < T && k < L.K)    // K is a multiple of 4 on this path
                    v = __ldcg(reinterpret_cast<const float4*>(L.X + ((size_t)lb * T + ts) * L.ldx + k));
            }
            areg[ch] = v;
        }
        // (3) epilogue operands of the (at most two) outputs this thread will write
        const float bsv = my_col >= 0 ? __ldg(L.bias + my_col) : 0.f;
        float eg1[2], eb1[2], eg2[2], eb2[2], ex[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            eg1[u] = eb1[u] = eg2[u] = eb2[u] = ex[u] = 0.f;
            const int idx = tid + u * CH_THREADS;
            if (n1 > 0 && idx < M * n1) {
                const int row = idx / n1, slot = idx - row * n1;
                const int ch = rank * per + slot;
                const int b = b0 + row / R, t = t_end - (R - 1) + (row % R);
                eg1[u] = __ldg(L.g1 + ch); eb1[u] = __ldg(L.b1 + ch);
                if (L.kind == 1) {
                    eg2[u] = __ldg(L.g2 + ch); eb2[u] = __ldg(L.b2 + ch);
                    if (t >= 0) ex[u] = __ldcg(L.X + ((size_t)b * T + t) * L.ldx + ch);
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < CH_MAXCH; ++ch)
            if (ch < nch) {
                float* dst = Afull + ((size_t)ch * CH_BK + lkq) * CH_BM + lrow;
                dst[0] = areg[ch].x; dst[CH_BM] = areg[ch].y; dst[2 * CH_BM] = areg[ch].z; dst[3 * CH_BM] = areg[ch].w;
            }
        // (the first __syncthreads of the reduction loop publishes Afull)

        float acc[CH_BM];
#pragma unroll
        for (int i = 0; i < CH_BM; ++i) acc[i] = 0.f;
        if (M <= 4)      chain_gemm<4>(Ws, Afull, nch, c, kg, acc, L, rank, per, KC, tid, vec_ok);
        else if (M <= 8) chain_gemm<8>(Ws, Afull, nch, c, kg, acc, L, rank, per, KC, tid, vec_ok);
        else             chain_gemm<16>(Ws, Afull, nch, c, kg, acc, L, rank, per, KC, tid, vec_ok);

        // reduce the four k-groups, add the bias: ytile[row][slot] = pre-LN conv output
        if (kg > 0) {
#pragma unroll
            for (int i = 0; i < CH_BM; ++i) red[kg - 1][i][c] = acc[i];
        }
        __syncthreads();
        if (kg == 0) {
#pragma unroll
            for (int i = 0; i < CH_BM; ++i) ytile[i][c] = acc[i] + red[0][i][c] + red[1][i][c] + red[2][i][c] + bsv;
        }
        __syncthreads();

        // LayerNorm partial statistics of this CTA's columns: warp w -> rows 2w, 2w+1
        for (int rr = 0; rr < 2; ++rr) {
            const int row = warp * 2 + rr;
            float s1 = 0.f, s2 = 0.f;
            for (int i = lane; i < n1; i += 32) s1 += ytile[row][i];
            for (int i = lane; i < n2; i += 32) s2 += ytile[row][n1 + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const float m1 = n1 > 0 ? s1 / (float)n1 : 0.f, m2 = n2 > 0 ? s2 / (float)n2 : 0.f;
            float q1 = 0.f, q2 = 0.f;
            for (int i = lane; i < n1; i += 32) { float d = ytile[row][i] - m1; q1 = fmaf(d, d, q1); }
            for (int i = lane; i < n2; i += 32) { float d = ytile[row][n1 + i] - m2; q2 = fmaf(d, d, q2); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { q1 += __shfl_xor_sync(0xffffffffu, q1, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); }
            if (lane < CH_NC) st_cluster_f4(mapa(smem_u32(&part[rank][row]), (uint32_t)lane), s1, q1, s2, q2);
        }
        cluster_arrive();                                       // partial statistics published
        cluster_wait();
        if (tid < 2 * CH_BM) {
            const int row = tid >> 1, hsel = tid & 1;
            float S = 0.f;
            for (int p = 0; p < CH_NC; ++p) { const float4 v = part[p][row]; S += hsel ? v.z : v.x; }
            const float mean = S / (float)L.C;
            float M2 = 0.f;
            for (int p = 0; p < CH_NC; ++p) {
                const float4 v = part[p][row];
                const int np = (L.kind == 1) ? per : max(0, min(per, L.C - p * per));
                if (np > 0 && (hsel == 0 || L.kind == 1)) {
                    const float d = (hsel ? v.z : v.x) / (float)np - mean;
                    M2 += (hsel ? v.w : v.y) + (float)np * d * d;
                }
            }
            stat[row][2 * hsel] = mean;
            stat[row][2 * hsel + 1] = 1.0f / sqrtf(M2 / (float)L.C + 1e-12f);
        }
        __syncthreads();
        // normalise, activate / gate / mix, write the new rows
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = tid + u * CH_THREADS;
            if (n1 > 0 && idx < M * n1) {
                const int row = idx / n1, slot = idx - row * n1;
                const int b = b0 + row / R, t = t_end - (R - 1) + (row % R);
                if (t >= 0) {
                    const int ch = rank * per + slot;
                    const size_t grow = (size_t)b * T + t;
                    float o;
                    if (L.kind == 1) {
                        const float z1 = (ytile[row][slot] - stat[row][0]) * stat[row][1] * eg1[u] + eb1[u];
                        const float z2 = (ytile[row][per + slot] - stat[row][2]) * stat[row][3] * eg2[u] + eb2[u];
                        const float h1 = sigmoid_f(z1);
                        o = h1 * z2 + (1.0f - h1) * ex[u];
                    } else {
                        o = (ytile[row][slot] - stat[row][0]) * stat[row][1] * eg1[u] + eb1[u];
                        if (L.act == 1) o = fmaxf(o, 0.f);
                    }
                    L.out[grow * L.ldo + ch] = o;
                    if (L.out2) L.out2[grow * L.ldo2 + ch] = sigmoid_f(o);
                }
            }
        }
        cluster_arrive();                                       // the block's output rows are visible cluster-wide
        cluster_wait();
    }
}

void launch_chain(const ChainArgs& a, cudaStream_t s) {
    static bool attr_set = false;
    const size_t smem = ((size_t)CH_STAGES * CH_STAGE_FLOATS + (size_t)CH_MAXCH * CH_BK * CH_BM) * sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(chain): ") + cudaGetErrorString(e));
        attr_set = true;
    }
    const int nclusters = (a.B + a.G - 1) / a.G;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CH_NC, (unsigned)nclusters, 1);
    cfg.blockDim = dim3(CH_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CH_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, chain_kernel, a);
    if (e != cudaSuccess) throw std::runtime_error(std::string("chain launch: ") + cudaGetErrorString(e));
}

}  // namespace dctts
