// kernels_rows.cu -- ONE reference block (conv + bias + LayerNorm + activation / highway mix,
// modules.py:91-141 and :143-197) on a FEW rows in ONE launch, for the autoregressive decode.
//
// The decode step is a chain of ~24 such blocks on 1..85 rows per utterance: latency and
// weight-bandwidth bound.  The two-kernel form (split-K GEMM, then a row-wise LN kernel that
// sums the partials) costs two kernel boundaries per block; this kernel does the block in one:
//   * a thread-block cluster of 8 CTAs splits the output channels (for `hc` a CTA owns the gate
//     AND the info columns of its channels, so the highway mix stays local);
//   * each CTA runs the full reduction for its <= 64 conv columns: weights stream from L2
//     through a 6-stage cp.async ring, the (<= 16 x 768) activation operand is fetched once;
//   * LayerNorm statistics are merged across the cluster through distributed shared memory
//     (Chan's mean/M2 merge), so no pre-LN tensor ever goes to global memory.
// fp32 FFMA throughout.  The code is deliberately kept small (rolled loops, one row-count
// variant per instantiation): a first version that walked a whole chain of blocks inside one
// kernel was instruction-fetch bound (190 KB of SASS, IPC 0.25) and slower than 26 launches.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dctts {

using namespace ptx;

constexpr int RB_NC = 8;            // cluster size (CTAs along the channels)
constexpr int RB_THREADS = 256;
constexpr int RB_STAGES = 6;
constexpr int RB_BM = 16;           // rows per cluster
constexpr int RB_BN = 64;           // conv columns per CTA
constexpr int RB_BK = 64;           // reduction slab per stage
constexpr int RB_MAXCH = 12;        // slabs per block (3 taps x 256 channels, or 1 x 512)
constexpr int RB_STAGE_FLOATS = RB_BK * RB_BN;

__device__ __forceinline__ void rb_cp16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void rb_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rb_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float rb_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// conv column behind accumulator slot `slot` of CTA `rank` (-1: unused)
__device__ __forceinline__ int rb_col(int kind, int C, int rank, int slot, int per) {
    if (kind == 1) {
        if (slot < per) return rank * per + slot;
        if (slot < 2 * per) return C + rank * per + (slot - per);
        return -1;
    }
    const int col = rank * per + slot;
    return (slot < per && col < C) ? col : -1;
}

// one 64 x 64 weight slab of this CTA into ring stage ch % RB_STAGES (always commits a group)
__device__ __forceinline__ void rb_issue_w(const RowsBlockArgs& a, float* Ws, int ch, int nch, int KC, int rank, int per, int tid) {
    if (ch < nch) {
        const int tap = ch / KC, k0 = (ch - tap * KC) * RB_BK;
        const float* Wt = a.W + (size_t)tap * a.K * a.ldw;
        float* st = Ws + (size_t)(ch % RB_STAGES) * RB_STAGE_FLOATS;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * RB_THREADS;
            const int k = idx >> 4, slot = (idx & 15) * 4;
            const int col = rb_col(a.kind, a.C, rank, slot, per);
            const bool ok = (col >= 0) && (k0 + k < a.K);
            rb_cp16(st + k * RB_BN + slot, ok ? Wt + (size_t)(k0 + k) * a.ldw + col : Wt, ok ? 16 : 0);
        }
    }
    rb_commit();
}

template <int MR>
__global__ void __launch_bounds__(RB_THREADS, 1) rows_block_kernel(const RowsBlockArgs a) {
    extern __shared__ __align__(16) float rb_dyn[];
    float* Ws = rb_dyn;                                          // [RB_STAGES][64][64]
    float* Afull = rb_dyn + RB_STAGES * RB_STAGE_FLOATS;         // [RB_MAXCH][64][16], k-major
    __shared__ float red[3][RB_BM][RB_BN];
    __shared__ float ytile[RB_BM][RB_BN];
    __shared__ float4 part[RB_NC][RB_BM];
    __shared__ float stat[RB_BM][4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    const int c = tid % RB_BN, kg = tid / RB_BN;
    const int lrow = tid % RB_BM, lkq = (tid / RB_BM) * 4;
    const int L = a.win.L, R = a.win.R;
    const int Mtot = a.win.B * R;
    const int m0 = blockIdx.y * RB_BM;
    const int per = (a.kind == 1) ? a.C / RB_NC : (a.C + RB_NC - 1) / RB_NC;
    const int KC = (a.K + RB_BK - 1) / RB_BK;
    const int nch = a.ntaps * KC;
    const int n1 = (a.kind == 1) ? per : max(0, min(per, a.C - rank * per));
    const int n2 = (a.kind == 1) ? per : 0;

    // weights do not depend on the previous kernel: start streaming before anything else
#pragma unroll
    for (int s = 0; s < RB_STAGES - 1; ++s) rb_issue_w(a, Ws, s, nch, KC, rank, per, tid);

    const int t_end = a.win.jptr ? *a.win.jptr : L - 1;
    auto row_bt = [&](int m, int& b, int& t) { b = m / R; t = t_end - (R - 1) + (m - b * R); };

    // the whole activation operand of these 16 rows: one float4 per thread per slab, all in flight
    int lb = 0, lt = -1;
    if (m0 + lrow < Mtot) row_bt(m0 + lrow, lb, lt);
    float4 areg[RB_MAXCH];
#pragma unroll
    for (int ch = 0; ch < RB_MAXCH; ++ch) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch < nch) {
            const int tap = ch / KC, k = (ch - tap * KC) * RB_BK + lkq;
            const int ts = lt + a.shifts[tap];
            if (lt >= 0 && ts >= 0 && ts < L && k < a.K)
                v = __ldcg(reinterpret_cast<const float4*>(a.X + ((size_t)lb * L + ts) * a.ldx + k));
        }
        areg[ch] = v;
    }
    const int my_col = rb_col(a.kind, a.C, rank, c, per);
    const float bsv = my_col >= 0 ? __ldg(a.bias + my_col) : 0.f;
#pragma unroll
    for (int ch = 0; ch < RB_MAXCH; ++ch)
        if (ch < nch) {
            float* dst = Afull + ((size_t)ch * RB_BK + lkq) * RB_BM + lrow;
            dst[0] = areg[ch].x; dst[RB_BM] = areg[ch].y; dst[2 * RB_BM] = areg[ch].z; dst[3 * RB_BM] = areg[ch].w;
        }

    float acc[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
        rb_wait<RB_STAGES - 2>();                                // my part of slab `ch` has landed
        __syncthreads();                                         // everybody's; slab ch-1 fully consumed; Afull visible
        rb_issue_w(a, Ws, ch + RB_STAGES - 1, nch, KC, rank, per, tid);
        const float* st = Ws + (size_t)(ch % RB_STAGES) * RB_STAGE_FLOATS + (kg * 16) * RB_BN + c;
        const float* ab = Afull + ((size_t)ch * RB_BK + kg * 16) * RB_BM;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float ww = st[kk * RB_BN];
            const float4* xr = reinterpret_cast<const float4*>(ab + kk * RB_BM);
#pragma unroll
            for (int g = 0; g < MR / 4; ++g) {
                const float4 x = xr[g];
                acc[4 * g + 0] = fmaf(x.x, ww, acc[4 * g + 0]); acc[4 * g + 1] = fmaf(x.y, ww, acc[4 * g + 1]);
                acc[4 * g + 2] = fmaf(x.z, ww, acc[4 * g + 2]); acc[4 * g + 3] = fmaf(x.w, ww, acc[4 * g + 3]);
            }
        }
    }
    rb_wait<0>();
    // reduce the four k-groups, add the bias
    if (kg > 0) {
#pragma unroll
        for (int i = 0; i < MR; ++i) red[kg - 1][i][c] = acc[i];
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
        for (int i = 0; i < MR; ++i) ytile[i][c] = acc[i] + red[0][i][c] + red[1][i][c] + red[2][i][c] + bsv;
    }
    __syncthreads();

    // partial LayerNorm statistics of my columns (warp w: rows 2w, 2w+1), published to the whole cluster
    for (int rr = 0; rr < 2; ++rr) {
        const int row = warp * 2 + rr;
        float s1 = 0.f, s2 = 0.f, q1 = 0.f, q2 = 0.f;
        if (row < MR) {
            for (int i = lane; i < n1; i += 32) s1 += ytile[row][i];
            for (int i = lane; i < n2; i += 32) s2 += ytile[row][n1 + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const float m1 = n1 > 0 ? s1 / (float)n1 : 0.f, m2 = n2 > 0 ? s2 / (float)n2 : 0.f;
            for (int i = lane; i < n1; i += 32) { float d = ytile[row][i] - m1; q1 = fmaf(d, d, q1); }
            for (int i = lane; i < n2; i += 32) { float d = ytile[row][n1 + i] - m2; q2 = fmaf(d, d, q2); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { q1 += __shfl_xor_sync(0xffffffffu, q1, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); }
        }
        if (lane < RB_NC) st_cluster_f4(mapa(smem_u32(&part[rank][row]), (uint32_t)lane), s1, q1, s2, q2);
    }
    cluster_arrive();
    cluster_wait();
    if (tid < 2 * RB_BM) {
        const int row = tid >> 1, hsel = tid & 1;
        float S = 0.f;
#pragma unroll
        for (int p = 0; p < RB_NC; ++p) { const float4 v = part[p][row]; S += hsel ? v.z : v.x; }
        const float mean = S / (float)a.C;
        float M2 = 0.f;
#pragma unroll
        for (int p = 0; p < RB_NC; ++p) {
            const float4 v = part[p][row];
            const int np = (a.kind == 1) ? per : max(0, min(per, a.C - p * per));
            if (np > 0 && (hsel == 0 || a.kind == 1)) {
                const float d = (hsel ? v.z : v.x) / (float)np - mean;
                M2 += (hsel ? v.w : v.y) + (float)np * d * d;
            }
        }
        stat[row][2 * hsel] = mean;
        stat[row][2 * hsel + 1] = 1.0f / sqrtf(M2 / (float)a.C + 1e-12f);
    }
    __syncthreads();
    // normalise, activate / gate / mix, store (at most 16 rows x 32 channels per CTA)
    for (int idx = tid; idx < MR * n1; idx += RB_THREADS) {
        const int row = idx / n1, slot = idx - row * n1;
        const int m = m0 + row;
        if (m >= Mtot) continue;
        int b, t; row_bt(m, b, t);
        if (t < 0) continue;
        const int ch = rank * per + slot;
        const size_t grow = (size_t)b * L + t;
        float o;
        if (a.kind == 1) {
            const float z1 = (ytile[row][slot] - stat[row][0]) * stat[row][1] * __ldg(a.g1 + ch) + __ldg(a.b1 + ch);
            const float z2 = (ytile[row][per + slot] - stat[row][2]) * stat[row][3] * __ldg(a.g2 + ch) + __ldg(a.b2 + ch);
            const float h1 = rb_sigmoid(z1);
            o = h1 * z2 + (1.0f - h1) * __ldcg(a.X + grow * a.ldx + ch);
        } else {
            o = (ytile[row][slot] - stat[row][0]) * stat[row][1] * __ldg(a.g1 + ch) + __ldg(a.b1 + ch);
            if (a.act == 1) o = fmaxf(o, 0.f);
        }
        a.out[grow * a.ldo + ch] = o;
        if (a.out2) a.out2[grow * a.ldo2 + ch] = rb_sigmoid(o);
    }
    // nobody may leave while a peer can still write its statistics into my shared memory
    cluster_arrive();
    cluster_wait();
}

bool rows_block_supported(int kind, int K, int C, int ntaps) {
    if (K % 4) return false;
    const int nch = ntaps * ((K + RB_BK - 1) / RB_BK);
    if (nch > RB_MAXCH) return false;
    if (kind == 1) return C % RB_NC == 0 && (C / RB_NC) % 4 == 0 && C / RB_NC <= 32;
    const int per = (C + RB_NC - 1) / RB_NC;
    return per % 4 == 0 && per <= 64;
}

void launch_rows_block(const RowsBlockArgs& a, cudaStream_t s) {
    static bool attr_set = false;
    const size_t smem = ((size_t)RB_STAGES * RB_STAGE_FLOATS + (size_t)RB_MAXCH * RB_BK * RB_BM) * sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rows_block_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_block_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(rows_block_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute(rows_block): ") + cudaGetErrorString(e));
        attr_set = true;
    }
    const int M = a.win.B * a.win.R;
    if (M <= 0) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(RB_NC, (unsigned)((M + RB_BM - 1) / RB_BM), 1);
    cfg.blockDim = dim3(RB_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = RB_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e;
    if (M <= 4)      e = cudaLaunchKernelEx(&cfg, rows_block_kernel<4>, a);
    else if (M <= 8) e = cudaLaunchKernelEx(&cfg, rows_block_kernel<8>, a);
    else             e = cudaLaunchKernelEx(&cfg, rows_block_kernel<16>, a);
    if (e != cudaSuccess) throw std::runtime_error(std::string("rows_block launch: ") + cudaGetErrorString(e));
}

}  // namespace dctts
