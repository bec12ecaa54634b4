// kernels_decode.cuh -- the autoregressive Text2Mel decode loop (reference synthesize.py:45-54) as ONE
// persistent launch: a 16-CTA thread-block cluster per group of <= 5 utterances walks AudioEnc ->
// Attention -> AudioDec for all mel frames, streaming its slice of the 27 MB of decode weights
// from L2 through a TMA-bulk ring.  See kernels_decode.cu for the design.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dctts {

constexpr int DEC_NC = 16;          // CTAs per cluster (non-portable cluster size)
constexpr int DEC_GMAX = 5;         // utterances per cluster (7 clusters of 16 CTAs are co-resident on a B200: 35 >= the benchmark's 32)
constexpr int DEC_THREADS = 256;
constexpr int DEC_NSLOT = 3;        // ring slots
constexpr int DEC_REG_F = 1536;     // floats per warp region of a slot (6 KB): every warp streams and frees its own k-rows
constexpr int DEC_SLOT_F = 8 * DEC_REG_F;   // 48 KB per slot
constexpr int DEC_MAXL = 24;        // 13 AudioEnc + 11 AudioDec blocks
constexpr int DEC_MAXCH = 48;       // weight chunks per frame
constexpr int DEC_PRM_F = 1024 + 64;   // per-layer parameter block: gamma1 | beta1 | gamma2 | beta2 (256 each) | bias slice

struct DecLayer {
    int kind;        // 0 conv1d (LN, optional relu), 1 hc (two LNs, sigmoid gate, highway mix)
    int cin;         // input channels
    int cout;        // output channels (256, or n_mels for the last block)
    int ntaps, rate; // causal taps at t - (ntaps-1-i)*rate
    int act;         // 1 = relu (conv1d)
    int ns;          // weight-slice columns per CTA (multiple of 4)
    int cs;          // channels owned per CTA (per LN half)
    int ch0, nch;    // chunk range of this layer within a frame
    int krows;       // k rows per chunk of this layer
    int prow;        // AudioDec receptive-field rows to recompute when the attention window moved (1 otherwise)
    int ldin;        // leading dimension of the input history
};
struct DecChunk { int off; int off16; short nfl4; short k0; short krows; short layer; };   // off: float offset in a rank's stream (off16: of the
                                                                   // same rows as split-fp16 MMA slabs); nfl4: floats / 4; k0: first k row (tap*cin + ci)

struct DecParams {
    DecLayer L[DEC_MAXL];
    DecChunk C[DEC_MAXCH];
    const float* in_hist[DEC_MAXL];    // (B, T, cin) history of the layer's input (nullptr: lives in shared memory only)
    float* out_hist[DEC_MAXL];         // (B, T, cout) history of the layer's output
    const float* lnp[DEC_MAXL];        // [4][256] gamma1, beta1, gamma2, beta2
    const float* bias[DEC_MAXL];       // [nconv] in the TF column order (gate | info for hc)
    const float* wstream;              // [DEC_NC][stream_len] packed weight slices, chunk by chunk
    const float* kv;                   // (B, N, 2d): K | V
    float* ybuf;                       // (B, T, n_mels)
    float* rbuf;                       // (B, T, 2d)
    float* pre_scr;                    // [clusters][G * max prow][512] pre-LN scratch of the recompute path
    int* p_hist;                       // (B, T) window used at every step
    int* p_final;                      // (B) window after the last step
    float inv_scale[DEC_MAXL];         // 1 / (power-of-two scale of the block's split-fp16 weight planes), tcgen05 pre-pass
    int* stats;                        // [clusters][2]: frames with a window move, utterance-frames recomputed
    long long* prof;                   // optional [16] SM-clock lap timers of cluster 0 / rank 0 (option decode_prof), else nullptr
    int nl, n_enc, nch, nch_enc, pyr_ch0, pyr_ch1, stream_len;   // pyr_ch0..pyr_ch1: chunks of the AudioDec blocks with prow > 1
    int B, G, T, N, d, n_mels, win_size, steps;
};
static_assert(sizeof(DecParams) <= 4000, "DecParams must fit the kernel parameter space");

size_t decode_smem_bytes();
// returns cudaSuccess or the launch / attribute error (the caller decides whether to fall back)
cudaError_t launch_decode_cluster(const DecParams& p, int n_clusters, cudaStream_t s);
// 0 when a 16-CTA cluster with this shared-memory footprint cannot be scheduled on the current device
int decode_max_active_clusters();

}  // namespace dctts
