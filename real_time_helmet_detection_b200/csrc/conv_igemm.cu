// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), used for the forward pass and for
// dgrad (same kernel, tap-flipped / transposed packed weights) of every 3x3 and 1x1 convolution
// of the hourglass (reference: hourglass.py:94-108 `Convolution`, called from `Residual` :111-127,
// `Neck` :176-186, `Head` :189-195 and the merge convs :215-218).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * Wp[tap, cout, cin]
//
// * Activations are NHWC bf16. One CTA tile = 128 output pixels (TN images x TH rows x TW columns,
//   all powers of two) x BLOCK_N output channels.
// * A operand: for every (tap, 64-channel chunk) one 4-D TMA box {64ch, TW, TH, TN} at the tap-shifted
//   pixel coordinate; out-of-image coordinates are zero-filled by TMA, which is exactly the conv's
//   zero padding. The box lands in smem as 128 rows of 128 B = the canonical K-major SWIZZLE_128B
//   UMMA layout, so no im2col buffer ever exists.
// * B operand: packed weights [tap][BLOCK_N][Cin] bf16, one 3-D TMA box {64, BLOCK_N, 1} per k-step.
// * MMA: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16, fp32 accumulators in TMEM,
//   double-buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
//   warps 4-7 = epilogue (TMEM -> registers -> bias / residual add / BN statistics -> global).
// * Persistent: grid = min(tiles, #SM); tiles are taken round-robin.
#include <cuda_bf16.h>

#include "hd_b200.h"
#include "hd_common.h"
#include "hd_ptx.cuh"

namespace hd {

constexpr int kStages = 6;
constexpr int kABytes = 128 * 128;  // 128 pixel rows x 64 bf16
constexpr int kThreads = 256;

struct ConvParams {
    int N, H, W;          // output == input spatial size (stride 1, "same" padding)
    int cin_chunks;       // Cin / 64
    int kh, kw, pad;      // taps
    int pad_y, pad_x;     // generic kernel: tap offsets dy = tap / kw - pad_y, dx = tap % kw - pad_x (== pad for k x k convs;
                          // the space-to-depth stem is a 4 x 1 column of taps with pad_y = 2, pad_x = 0)
    int cout;             // real output channels (<= BLOCK_N)
    int tw_log2, th_log2, tn_log2;
    int tiles_x, tiles_y, tiles_n, num_tiles;
    // outputs
    int out_mode;         // 0: NHWC bf16 (channel stride out_cs) ; 1: NCHW fp32 slice of (B,S,cout,H,W)
    int out_cs;           // channel stride (elements) of the NHWC output
    int stack_idx, num_stack;
    void* out;
    __nv_bfloat16* out2;  // optional second NHWC bf16 copy (channel stride out2_cs, zero padded), mode 1 only
    int out2_cs;
    const float* bias;        // optional [cout]
    const __nv_bfloat16* addend;  // optional NHWC bf16, same shape as out (mode 0)
    float* stat_sum;          // optional [cout] : sum over pixels of the fp32 conv output (bias included)
    float* stat_sqsum;        // optional [cout]
    // optional fused train-mode BN finalize by the last CTA to flush its statistics (see hd_bn_fuse in hd_b200.h)
    const float* bn_gamma; const float* bn_beta; float* bn_rm; float* bn_rv; long long* bn_nbt;
    float bn_momentum, bn_eps, bn_count; float* bn_out; unsigned int* bn_counter;
    // optional per-channel affine + ReLU epilogue (eval-mode BatchNorm folded into the conv, hd_conv2d_igemm_affine):
    // out = relu?( (conv + bias) * ep_scale[c] + ep_shift[c] + addend )
    const float* ep_scale; const float* ep_shift; int ep_relu;
    int dbg;                  // profiling only (hd_set_conv_debug): 1 = epilogue drains TMEM but skips math/stores,
                              // 2 = MMA issue skipped (halo kernel only)
    // optional BatchNorm-BACKWARD statistics of the tensor this launch writes (transposed halo kernel only,
    // hd_conv2d_igemm_bwdstat): the dgrad of a Residual's second conv produces dZ1, whose consumer is the backward of
    // conv1's BN + ReLU; with g = bf16(out) * (bwd_y * bwd_sc + bwd_sh > 0) the epilogue adds sum g to stat_sum and
    // sum g * bwd_y to stat_sqsum (raw moments, as bn_bwd_reduce_kernel) and the last CTA turns them into the apply
    // kernel's coefficients + dgamma / dbeta (bf_*), re-zeroing the sums: the separate reduction pass over dZ1 and Y1
    // (2 x 134 MB at 128x128, B = 32) becomes one extra read of Y1 under the tensor-bound MMA loop
    const __nv_bfloat16* bwd_y; const float* bwd_sc; const float* bwd_sh;
    const float* bf_gamma; const float* bf_mean; const float* bf_rstd; float* bf_coef; float* bf_dgamma; float* bf_dbeta;
    float bf_count;
    int x2_chunks;            // N=64 halo kernel only: 64-channel chunks of a SECOND input that enters as one extra 1x1
                              // tap (out += W2 * x2): the 1x1 skip-branch dgrad fused into the 3x3 dgrad of a Residual
};

// Sum v[0..31] across the 32 lanes of the warp; on return lane l holds the total of element l.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], uint32_t lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float send = upper ? v[i] : v[i + off];
            float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Epilogue of one 128-row accumulator block: TMEM -> registers (32 columns at a time) -> bias / residual addend ->
// bf16 NHWC store, plus the per-channel sum / sum-of-squares (train-mode BN statistics) into s_stat.
// Stores go through a 2 KB per-warp staging buffer: a lane owns one pixel row, so writing its 64 bytes (32 channels)
// directly would make every store instruction touch 32 different 128-byte lines with 16 bytes each (measured: the
// epilogue alone took 78 us of a 123 us kernel). Staged, one instruction writes 8 rows x 64 contiguous bytes.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_rows(const ConvParams& p, uint32_t taddr, bool valid, size_t pix, uint32_t lane,
                                              float* s_stat, bool do_stats, uint8_t* stage, int c_begin = 0,
                                              int c_end = BLOCK_N) {
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + c0 + i);
        }
        if (p.ep_scale) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], __ldg(p.ep_scale + c0 + i), __ldg(p.ep_shift + c0 + i));
        }
        if (p.addend) {
            // coalesced read of the 32 rows x 64 B addend tile through the staging buffer (8 rows x 64 B per
            // instruction), then every lane picks up its own row
            const uint32_t q = lane & 3u;
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const uint32_t r = 8u * sidx + (lane >> 2);
                const unsigned long long rp = __shfl_sync(0xffffffffu, static_cast<unsigned long long>(pix), r);
                const int rv = __shfl_sync(0xffffffffu, static_cast<int>(valid), r);
                uint4 w = make_uint4(0u, 0u, 0u, 0u);
                if (rv) w = __ldg(reinterpret_cast<const uint4*>(p.addend + rp * p.out_cs + c0 + q * 8));
                *reinterpret_cast<uint4*>(stage + r * 64 + ((q ^ ((r >> 1) & 3u)) << 4)) = w;
            }
            __syncwarp();
            const uint32_t swz_a = (lane >> 1) & 3u;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const uint4 u = *reinterpret_cast<const uint4*>(stage + lane * 64 + ((static_cast<uint32_t>(qq) ^ swz_a) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = __bfloat1622float2(h[j]);
                    v[qq * 8 + 2 * j] += f.x;
                    v[qq * 8 + 2 * j + 1] += f.y;
                }
            }
            __syncwarp();
        }
        if (p.ep_relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        {
            const uint32_t swz = (lane >> 1) & 3u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                *reinterpret_cast<uint4*>(stage + lane * 64 + ((static_cast<uint32_t>(q) ^ swz) << 4)) = u;
            }
            __syncwarp();
            const uint32_t q = lane & 3u;
#pragma unroll
            for (int sidx = 0; sidx < 4; ++sidx) {
                const uint32_t r = 8u * sidx + (lane >> 2);
                const uint4 w = *reinterpret_cast<const uint4*>(stage + r * 64 + ((q ^ ((r >> 1) & 3u)) << 4));
                const unsigned long long rp = __shfl_sync(0xffffffffu, static_cast<unsigned long long>(pix), r);
                const int rv = __shfl_sync(0xffffffffu, static_cast<int>(valid), r);
                if (rv)
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + rp * p.out_cs + c0 + q * 8) = w;
            }
            __syncwarp();
        }
        if (do_stats) {
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                v[i] = valid ? v[i] : 0.f;
                sq[i] = v[i] * v[i];
            }
            float s1 = warp_transpose_reduce(v, lane);
            float s2 = warp_transpose_reduce(sq, lane);
            atomicAdd(&s_stat[c0 + lane], s1);
            atomicAdd(&s_stat[BLOCK_N + c0 + lane], s2);
        }
    }
}


// Flush this CTA's per-channel statistics to global memory; the LAST CTA of the grid to do so (atomic ticket) turns the
// complete statistics into the BatchNorm scale / shift / mean / rstd and updates the running statistics - the
// separate finalize launch of hourglass.py:103's train-mode BN disappears. Called by the `nthr` epilogue threads
// (t = 0..nthr-1), which share named barrier 1.
template <int BLOCK_N>
__device__ __forceinline__ void flush_stats(const ConvParams& p, float* s_stat, int t, int nthr, int* s_flag) {
    named_bar_sync(1, nthr);
    if (t < p.cout) {
        atomicAdd(p.stat_sum + t, s_stat[t]);
        atomicAdd(p.stat_sqsum + t, s_stat[BLOCK_N + t]);
    }
    if (p.bn_out == nullptr && p.bf_coef == nullptr) return;
    __threadfence();
    named_bar_sync(1, nthr);
    if (t == 0) *s_flag = (atomicAdd(p.bn_counter, 1u) == gridDim.x - 1) ? 1 : 0;
    named_bar_sync(1, nthr);
    if (*s_flag == 0) return;
    __threadfence();
    if (p.bf_coef != nullptr) {      // BN-backward statistics (see ConvParams::bwd_y): coefficients of dy = a*g + b*y + c
        if (t < p.cout) {
            const float S0 = __ldcg(p.stat_sum + t);
            const float g = p.bf_gamma[t], r = p.bf_rstd[t], m = p.bf_mean[t];
            const float S1 = r * (__ldcg(p.stat_sqsum + t) - m * S0);
            p.bf_coef[t] = g * r;
            p.bf_coef[p.cout + t] = -g * r * r * S1 / p.bf_count;
            p.bf_coef[2 * p.cout + t] = g * r * (m * r * S1 - S0) / p.bf_count;
            if (p.bf_dgamma) p.bf_dgamma[t] = S1;
            if (p.bf_dbeta) p.bf_dbeta[t] = S0;
            p.stat_sum[t] = 0.f;            // accumulators left zeroed for the next reduction on this scratch block
            p.stat_sqsum[t] = 0.f;
        }
        if (t == 0) *p.bn_counter = 0u;
        return;
    }
    if (t < p.cout) {
        const float sum = __ldcg(p.stat_sum + t), sq = __ldcg(p.stat_sqsum + t);
        const float mean = sum / p.bn_count;
        const float var = fmaxf(sq / p.bn_count - mean * mean, 0.f);
        if (p.bn_rm) {
            const float unbiased = p.bn_count > 1.f ? var * (p.bn_count / (p.bn_count - 1.f)) : var;
            p.bn_rm[t] = (1.f - p.bn_momentum) * p.bn_rm[t] + p.bn_momentum * mean;
            p.bn_rv[t] = (1.f - p.bn_momentum) * p.bn_rv[t] + p.bn_momentum * unbiased;
        }
        const float rstd = rsqrtf(var + p.bn_eps);
        const float sc = p.bn_gamma[t] * rstd;
        p.bn_out[t] = sc;
        p.bn_out[p.cout + t] = p.bn_beta[t] - mean * sc;
        p.bn_out[2 * p.cout + t] = mean;
        p.bn_out[3 * p.cout + t] = rstd;
    }
    if (t == 0) {
        if (p.bn_nbt) *p.bn_nbt += 1;
        *p.bn_counter = 0u;
    }
}

// BLOCK_N >= 64 runs with EIGHT epilogue warps (384 threads): warps 4-7 drain the first half of the columns, warps 8-11
// the second half - with few k-steps per tile (1x1 convs, the K=192 stem, N=64) the epilogue is the pacing stage.
template <int BLOCK_N>
__global__ void __launch_bounds__(BLOCK_N >= 64 ? 384 : kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const ConvParams p) {
    pdl_launch_dependents();
    constexpr int kEpiWarps = BLOCK_N >= 64 ? 8 : 4;
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr uint32_t kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;
    constexpr uint32_t kIdesc = umma_idesc_bf16(BLOCK_N, 0, 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);  // [2][BLOCK_N]
    uint8_t* s_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_stat + 2 * BLOCK_N) + 15) & ~uintptr_t(15));   // [epilogue warps][2 KB]

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 32 * kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    if (threadIdx.x < 2 * BLOCK_N) s_stat[threadIdx.x] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    const int ksteps = p.kh * p.kw * p.cin_chunks;
    const int TW = 1 << p.tw_log2, TH = 1 << p.th_log2;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int tn = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx << p.tw_log2, y0 = ty << p.th_log2, n0 = tn << p.tn_log2;
                for (int tap = 0; tap < p.kh * p.kw; ++tap) {
                    const int dy = tap / p.kw - p.pad_y, dx = tap % p.kw - p.pad_x;
                    for (int ck = 0; ck < p.cin_chunks; ++ck) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * kStageBytes;
                        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                        tma_load_4d(sa, &tmap_x, &full_bar[stage], ck * 64, x0 + dx, y0 + dy, n0);
                        tma_load_3d(sa + kABytes, &tmap_w, &full_bar[stage], ck * 64, 0, tap);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        uint32_t stage = 0, phase = 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = umma_smem_desc_sw128(sa, 0, 1024);
                    const uint64_t bdesc = umma_smem_desc_sw128(sa + kABytes, 0, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128B swizzle row: +2 in (addr>>4) units
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (ks > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (ks == ksteps - 1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int ew = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = ew * 32 + (int)lane;     // tile row == TMEM lane
        const bool do_stats = p.stat_sum != nullptr;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int tn = tile / (p.tiles_x * p.tiles_y);
            const int x = (tx << p.tw_log2) + (row & (TW - 1));
            const int y = (ty << p.th_log2) + ((row >> p.tw_log2) & (TH - 1));
            const int n = (tn << p.tn_log2) + (row >> (p.tw_log2 + p.th_log2));
            const bool valid = (x < p.W) && (y < p.H) && (n < p.N);
            const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;

            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;

            if constexpr (BLOCK_N >= 32) {
                constexpr int kHalf = kEpiWarps == 8 ? BLOCK_N / 2 : BLOCK_N;
                const int cb = ((warp - 4) >> 2) * kHalf;
                epilogue_rows<BLOCK_N>(p, taddr, valid, pix, lane, s_stat, do_stats, s_stage + (warp - 4) * 2048, cb,
                                       cb + kHalf);
            } else {
                // BLOCK_N == 16: prediction head (hourglass.py:189-195), fp32 NCHW logits
                uint32_t r[16];
                tmem_ld_x16(taddr, r);
                tmem_ld_wait();
                if (valid) {
                    const size_t hw = static_cast<size_t>(p.H) * p.W;
                    float* o = reinterpret_cast<float*>(p.out) +
                               (static_cast<size_t>(n) * p.num_stack + p.stack_idx) * p.cout * hw +
                               static_cast<size_t>(y) * p.W + x;
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = __uint_as_float(r[i]);
                        if (p.bias && i < p.cout) v[i] += __ldg(p.bias + i);
                        if (i >= p.cout) v[i] = 0.f;
                    }
                    if (p.out_mode == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < p.cout) o[i * hw] = v[i];
                    } else {
                        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_cs;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < p.cout) ob[i] = __float2bfloat16(v[i]);
                    }
                    if (p.out2) {
                        uint4* o2 = reinterpret_cast<uint4*>(p.out2 + pix * p.out2_cs);
                        uint4 u0, u1;
                        u0.x = pack_bf16x2(v[0], v[1]);   u0.y = pack_bf16x2(v[2], v[3]);
                        u0.z = pack_bf16x2(v[4], v[5]);   u0.w = pack_bf16x2(v[6], v[7]);
                        u1.x = pack_bf16x2(v[8], v[9]);   u1.y = pack_bf16x2(v[10], v[11]);
                        u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
                        o2[0] = u0;
                        o2[1] = u1;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
        if (do_stats) flush_stats<BLOCK_N>(p, s_stat, static_cast<int>(threadIdx.x) - 128, 32 * kEpiWarps, reinterpret_cast<int*>(tmem_slot + 1));
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = kStages * (kABytes + BLOCK_N * 128) + 1024 /*align*/ + 256 /*barriers*/ +
                               2 * BLOCK_N * 4 + 8 * 2048 /*store staging*/;
    HD_ENSURE_DYN_SMEM(conv_igemm_kernel<BLOCK_N>, smem_bytes);
    int grid = p.num_tiles < sm_budget() ? p.num_tiles : sm_budget();
    HD_CHECK_CUDA(::hd::launch_k_pdl(p.num_tiles < sm_count(), conv_igemm_kernel<BLOCK_N>, grid,
                                     BLOCK_N >= 64 ? 384 : kThreads,
                                     smem_bytes, stream, tx, tw, p));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// "Halo" kernel for the FLOP-dominant case (3x3, 128 output channels, maps >= 16x16): the generic kernel above
// re-fetches the activation tile for each of the 9 taps and the weights for every 128-pixel tile, which makes it
// L2->SM bandwidth bound (measured 14.3 TB/s, ~54 % of the tensor peak). Here a CTA tile is 16x16 pixels and for each dx
// ONE activation box of 18 rows x 16 columns is loaded; the three dy taps read that box at a row offset of dy*16
// pixels = dy*2048 B (whole 8-row swizzle groups, so only the UMMA descriptor start address moves). L2->SM bytes per
// FLOP drop 2.3x.
constexpr int kHARows = 18 * 16;
constexpr int kHABytes = kHARows * 128;   // 36,864
constexpr int kHAStages = 3;
constexpr int kHBBytes = 128 * 128;       // one (tap, 64-channel chunk) weight tile for 128 output channels
constexpr int kHBStages = 5;

// 12 warps: producer, MMA issuer, TMEM allocator, (idle), and EIGHT epilogue warps - warps 4-7 drain pixel half 0,
// warps 8-11 half 1 (a warp may only touch the TMEM lane quarter warp_idx % 4).
constexpr int kHaloThreads = 384;

// The product is computed TRANSPOSED: D^T[cout, pixel] = W[cout, k] * X^T[k, pixel]. The weights are the A operand
// (M = 128 output channels), the 16x16-pixel activation tile is the B operand (N = 256), so one tcgen05.mma (M128 N256
// K16, 128 cycles) does the work of two M128 N128 ones and the tensor core re-reads 4 KB of weights + 8 KB of
// activations per 128 cycles instead of 2 x (4 + 4) KB: shared-memory operand traffic drops from 128 to 96 B/clk per SM.
// The first halo kernel (pixels as M, two 128-row accumulators) was paced by exactly that port: 116-126 us per launch
// at 128^2 x B32, 108 us with the epilogue switched off; this one: 105-109 us, 92 us without epilogue (round-1 ncu).
// The accumulator is [lane = output channel][column = pixel]:
//   * BN statistics are per-THREAD running sums (a thread owns one channel) - the 62-shuffle transposing reduction of
//     epilogue_rows() disappears;
//   * the NHWC store needs a transpose. The SM's L1/shared-memory data pipe is the contended resource (ncu: tensor-core
//     operand reads 55 % + LSU 37 % of its peak with a naive staged transpose: 2-byte STS, LDS.128, STG.128), so the
//     transpose costs as few wavefronts as possible: one shuffle per pixel pair inside lane pairs, 16 conflict-free
//     32-bit STS per 32x32 block into a [pixel][channel] staging tile, and a TMA bulk-tensor STORE of that tile (which
//     also clips at the image border) instead of LDS + STG.
// What remains between 92 and ~106 us is the epilogue's TMEM read (128 KB of fp32 accumulators per tile at 64 B/clk
// = 2048 of the tile's 9216 MMA cycles), which the tensor core's own accumulator traffic has to share.
// Packed bf16x2 words of a second NHWC tensor of the output's shape (the residual addend, or the consumer's BN input in
// BN-backward-statistics mode) for one 32-channel x 32-pixel block: lane pair (2m, 2m+1) reads channels (c, c+1) of
// pixel 2j (even lane) / 2j+1 (odd lane). Issued right after the block's TMEM load, in front of the math. (Requesting
// them one block AHEAD - before the wait for the accumulator and before the previous block's math, epilogue loop fully
// unrolled - was measured SLOWER: 137 -> 163 us for the statistics dgrad, 131 -> 137 us for the addend dgrad at 128x128.)
__device__ __forceinline__ void load_aux_block(const __nv_bfloat16* __restrict__ aux, const ConvParams& p, int cbase,
                                               uint32_t lane, int n, int x0, int yb, uint32_t (&w)[16]) {
    const int cpair = cbase + static_cast<int>(lane & ~1u);
    const size_t row0 = (static_cast<size_t>(n) * p.H + yb) * p.W;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int x = x0 + ((2 * j) & 15) + static_cast<int>(lane & 1u);
        w[j] = 0u;
        if (yb + (j >> 3) < p.H && x < p.W)
            w[j] = __ldg(reinterpret_cast<const uint32_t*>(aux + (row0 + (j >> 3) * p.W + x) * p.out_cs + cpair));
    }
}

template <bool BWD>
__device__ __forceinline__ void epilogue_block_t(const ConvParams& p, const CUtensorMap* tmap_o, uint32_t taddr, int cbase,
                                                 uint32_t lane, int n, int x0, int yb, uint8_t* stage, float bias,
                                                 float ep_scale, float ep_shift, bool do_stats, float& s1, float& s2,
                                                 const __nv_bfloat16* __restrict__ aux_ptr, float bsc = 0.f, float bsh = 0.f) {
    uint32_t r[32];
    tmem_ld_x32(taddr, r);
    constexpr bool bwd = BWD;       // BN-backward-statistics instance (never with an addend): see ConvParams::bwd_y
    uint32_t aux[16];
    if (aux_ptr) load_aux_block(aux_ptr, p, cbase, lane, n, x0, yb, aux);
    // Lane pair (2m, 2m+1) = channels (c, c+1). For the pixel pair (2j, 2j+1) the even lane ends up with both channels
    // of pixel 2j and the odd lane with both channels of pixel 2j+1 (one shuffle), i.e. one packed bf16x2 word each.
    const uint32_t odd = lane & 1u;
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]) + bias, ep_scale, ep_shift);
    if (!BWD && p.addend) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, aux[j], 1);
            // even lane (channel c): c @ pixel 2j = low half of its own word, c @ 2j+1 = low half of the partner's;
            // odd lane (channel c+1): @ 2j = high half of the partner's word, @ 2j+1 = high half of its own
            const uint32_t w0 = odd ? other : aux[j], w1 = odd ? aux[j] : other;
            v[2 * j] += __uint_as_float(odd ? (w0 & 0xffff0000u) : (w0 << 16));
            v[2 * j + 1] += __uint_as_float(odd ? (w1 & 0xffff0000u) : (w1 << 16));
        }
    }
    if (p.ep_relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    // the previous TMA store of this warp must have finished reading the staging buffer before it is overwritten
    if (lane == 0) bulk_wait_group_read<0>();
    __syncwarp();
    // staging tile [32 pixels][32 channels] bf16 = the store box {32 ch, 16 x, 2 y}: even lanes fill pixel row 2j, odd
    // lanes row 2j+1, 16 consecutive words each -> one conflict-free 128-byte wavefront per instruction
    uint32_t* srow = reinterpret_cast<uint32_t*>(stage) + odd * 16 + (lane >> 1);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[2 * j] : v[2 * j + 1], 1);
        srow[32 * j] = odd ? pack_bf16x2(recv, v[2 * j + 1]) : pack_bf16x2(v[2 * j], recv);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_4d(tmap_o, stage, cbase, x0, yb, n);     // clipped at the image border by the TMA unit
        bulk_commit_group();
    }
    if (bwd) {
        // g = the STORED (bf16-rounded) value where the consumer's ReLU was open; out-of-image pixels contribute nothing
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, aux[j], 1);
            const uint32_t w0 = odd ? other : aux[j], w1 = odd ? aux[j] : other;      // as the addend exchange above
            const float y0 = __uint_as_float(odd ? (w0 & 0xffff0000u) : (w0 << 16));
            const float y1 = __uint_as_float(odd ? (w1 & 0xffff0000u) : (w1 << 16));
            const bool row_ok = yb + (j >> 3) < p.H;
            const int xa = x0 + ((2 * j) & 15);
            const float g0 = __bfloat162float(__float2bfloat16_rn(v[2 * j]));
            const float g1 = __bfloat162float(__float2bfloat16_rn(v[2 * j + 1]));
            const float t0 = (row_ok && xa < p.W && fmaf(y0, bsc, bsh) > 0.f) ? g0 : 0.f;
            const float t1 = (row_ok && xa + 1 < p.W && fmaf(y1, bsc, bsh) > 0.f) ? g1 : 0.f;
            s1 += t0 + t1;
            s2 = fmaf(t0, y0, fmaf(t1, y1, s2));
        }
    } else if (do_stats) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const bool ok = (yb + (i >> 4) < p.H) && (x0 + (i & 15) < p.W);
            const float t = ok ? v[i] : 0.f;
            s1 += t;
            s2 = fmaf(t, t, s2);
        }
    }
}

template <bool BWD>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_igemm_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_o, const ConvParams p) {
    pdl_launch_dependents();
    constexpr int BLOCK_N = 128;                                   // output channels (the MMA's M here)
    constexpr uint32_t kIdesc = umma_idesc_bf16(256, 0, 0);        // M = 128 channels, N = 256 pixels
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                        // activation ring (B operand)
    uint8_t* smem_b = smem + kHAStages * kHABytes;                 // weight ring (A operand)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kHBStages * kHBBytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kHAStages;
    uint64_t* b_full = a_empty + kHAStages;
    uint64_t* b_empty = b_full + kHBStages;
    uint64_t* tmem_full = b_empty + kHBStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);  // [2][128]
    uint8_t* s_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_stat + 2 * BLOCK_N) + 127) & ~uintptr_t(127));   // [8 epilogue warps][2 KB], TMA store source

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_o);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kHAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kHBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < 2 * BLOCK_N) s_stat[threadIdx.x] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on

    if (warp == 0) {
        if (elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int n = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * 16, y0 = ty * 16;
                // 3x3: box of 18 rows per dx, three dy taps inside it; 1x1 (p.kw == 1): one 16-row box, one tap
                const uint32_t a_bytes = static_cast<uint32_t>(16 + 2 * p.pad) * 16u * 128u;
                for (int dx = 0; dx < p.kw; ++dx) {
                    for (int ck = 0; ck < p.cin_chunks; ++ck) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                        tma_load_4d(smem_a + sa * kHABytes, &tmap_x, &a_full[sa], ck * 64, x0 + dx - p.pad, y0 - p.pad, n);
                        for (int dy = 0; dy < p.kh; ++dy) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            mbar_arrive_expect_tx(&b_full[sb], kHBBytes);
                            tma_load_3d(smem_b + sb * kHBBytes, &tmap_w, &b_full[sb], ck * 64, 0, dy * p.kw + dx);
                            if (++sb == kHBStages) { sb = 0; pb ^= 1; }
                        }
                        if (++sa == kHAStages) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            bool first = true;
            for (int dx = 0; dx < p.kw; ++dx) {
                for (int ck = 0; ck < p.cin_chunks; ++ck) {
                    mbar_wait(&a_full[sa], pa);
                    const uint32_t x_addr = smem_u32(smem_a + sa * kHABytes);
                    for (int dy = 0; dy < p.kh; ++dy) {
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t wdesc = umma_smem_desc_sw128(smem_u32(smem_b + sb * kHBBytes), 0, 1024);
                            // 256 pixel rows of the box starting dy rows down: 16 tile rows x 16 columns, 32 KB contiguous
                            const uint64_t xdesc = umma_smem_desc_sw128(x_addr + dy * 2048, 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (p.dbg != 2)
                                    umma_bf16(d_tmem, wdesc + 2 * k, xdesc + 2 * k, kIdesc, (first && k == 0) ? 0u : 1u);
                            umma_commit(&b_empty[sb]);
                        }
                        __syncwarp();
                        first = false;
                        if (++sb == kHBStages) { sb = 0; pb ^= 1; }
                    }
                    if (elect_one()) umma_commit(&a_empty[sa]);
                    __syncwarp();
                    if (++sa == kHAStages) { sa = 0; pa ^= 1; }
                }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;                       // TMEM lane quarter = 32 output channels
        const int h = (warp - 4) >> 2;                 // pixel half: tile rows 8h .. 8h+7 = accumulator columns 128h ..
        const int cbase = ew * 32, ch = cbase + static_cast<int>(lane);
        const bool do_stats = p.stat_sum != nullptr;
        const float bias = (p.bias && ch < p.cout) ? __ldg(p.bias + ch) : 0.f;
        const float ep_scale = p.ep_scale ? __ldg(p.ep_scale + ch) : 1.f, ep_shift = p.ep_scale ? __ldg(p.ep_shift + ch) : 0.f;
        const float bsc = BWD ? __ldg(p.bwd_sc + ch) : 0.f, bsh = BWD ? __ldg(p.bwd_sh + ch) : 0.f;
        const __nv_bfloat16* aux = BWD ? p.bwd_y : p.addend;
        uint8_t* stage = s_stage + (warp - 4) * 2048;
        float s1 = 0.f, s2 = 0.f;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int n = tile / (p.tiles_x * p.tiles_y);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 256 + h * 128;
            if (p.dbg != 1) {
#pragma unroll 1
                for (int b = 0; b < 4; ++b)
                    epilogue_block_t<BWD>(p, &tmap_o, taddr + b * 32, cbase, lane, n, tx * 16, ty * 16 + 8 * h + 2 * b, stage,
                                     bias, ep_scale, ep_shift, do_stats, s1, s2, aux, bsc, bsh);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) bulk_wait_group<0>();           // all output tiles written before the CTA retires
        if (do_stats) {
            atomicAdd(&s_stat[ch], s1);
            atomicAdd(&s_stat[BLOCK_N + ch], s2);
            flush_stats<BLOCK_N>(p, s_stat, static_cast<int>(threadIdx.x) - 128, 256, reinterpret_cast<int*>(tmem_slot + 1));
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// Halo kernel for 64 OUTPUT channels on the large maps (the 256x256 level of PreLayer: the 7x7 stem as four vertical
// taps over the space-to-depth image, hourglass.py:163, and the dgrad of `Residual(64, 128)`, hourglass.py:164).
// The generic kernel above is L2->SM bound there (every tap re-fetches its 128-pixel activation tile: 7 GB of operand
// traffic for the 3x3 128->64 dgrad at 256^2 x B32, 455 us = 0.39 of the tensor peak) and the transposed halo kernel
// would run the tensor core at half rate (M = 64). Here pixels are M again: a CTA tile is 16x16 pixels = two 128-row
// accumulators of 64 fp32 columns; per (dx, 64-channel chunk) ONE activation box of (16 + kh - 1) rows x 16 columns is
// loaded and the kh vertical taps / the two pixel halves are row offsets into it (dy*2048 B / 16 KB: whole swizzle groups).
// Optionally a second input enters as an extra 1x1 tap (x2_chunks): dX = dgrad3x3(dY1) + dgrad1x1(dYs) in ONE pass -
// the intermediate tensor and the second launch of the unfused pair disappear.
// Epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves; a warp owns 32 channels of 2 x 32 pixels per tile, so the
// BN statistics are per-lane running sums over all tiles (64 registers) with ONE transposing reduction per CTA, and the
// bf16 NHWC output leaves through a 64B-swizzled staging tile and a TMA store (conflict-free STS, fully coalesced lines).
constexpr int kN64ARows = 19 * 16;                 // up to 4 vertical taps: 16 + 3 box rows
constexpr int kN64ABytes = kN64ARows * 128;        // 38,912 (a multiple of 1024)
constexpr int kN64AStages = 3;
constexpr int kN64BBytes = 64 * 128;               // one (tap, 64-channel chunk) weight tile for 64 output channels
constexpr int kN64BStages = 6;
constexpr int kN64Threads = 384;

__global__ void __launch_bounds__(kN64Threads, 1)
conv_igemm_n64_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_x2, const __grid_constant__ CUtensorMap tmap_w2,
                      const __grid_constant__ CUtensorMap tmap_o, const ConvParams p) {
    pdl_launch_dependents();
    constexpr int BLOCK_N = 64;
    constexpr uint32_t kIdesc = umma_idesc_bf16(BLOCK_N, 0, 0);    // M = 128 pixels, N = 64 channels
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kN64AStages * kN64ABytes;
    uint8_t* s_stage = smem_b + kN64BStages * kN64BBytes;          // [8 warps][2 halves][2 KB], 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + 8 * 2 * 2048);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kN64AStages;
    uint64_t* b_full = a_empty + kN64AStages;
    uint64_t* b_empty = b_full + kN64BStages;
    uint64_t* tmem_full = b_empty + kN64BStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);       // [2][64]

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_o);
        if (p.x2_chunks) { tma_prefetch_desc(&tmap_x2); tma_prefetch_desc(&tmap_w2); }
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kN64AStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kN64BStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);                     // 2 accumulator sets x 2 pixel halves x 64 columns
    if (threadIdx.x < 2 * BLOCK_N) s_stat[threadIdx.x] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const uint32_t a_bytes = static_cast<uint32_t>(16 + p.kh - 1) * 16u * 128u;
    if (warp == 0) {
        if (elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int n = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * 16, y0 = ty * 16;
                for (int dx = 0; dx < p.kw; ++dx) {
                    for (int ck = 0; ck < p.cin_chunks; ++ck) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        mbar_arrive_expect_tx(&a_full[sa], a_bytes);
                        tma_load_4d(smem_a + sa * kN64ABytes, &tmap_x, &a_full[sa], ck * 64, x0 + dx - p.pad_x, y0 - p.pad_y, n);
                        for (int dy = 0; dy < p.kh; ++dy) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            mbar_arrive_expect_tx(&b_full[sb], kN64BBytes);
                            tma_load_3d(smem_b + sb * kN64BBytes, &tmap_w, &b_full[sb], ck * 64, 0, dy * p.kw + dx);
                            if (++sb == kN64BStages) { sb = 0; pb ^= 1; }
                        }
                        if (++sa == kN64AStages) { sa = 0; pa ^= 1; }
                    }
                }
                for (int ck = 0; ck < p.x2_chunks; ++ck) {         // the fused 1x1 input: a 16-row box, one tap
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    mbar_arrive_expect_tx(&a_full[sa], 16u * 16u * 128u);
                    tma_load_4d(smem_a + sa * kN64ABytes, &tmap_x2, &a_full[sa], ck * 64, x0, y0, n);
                    mbar_wait(&b_empty[sb], pb ^ 1);
                    mbar_arrive_expect_tx(&b_full[sb], kN64BBytes);
                    tma_load_3d(smem_b + sb * kN64BBytes, &tmap_w2, &b_full[sb], ck * 64, 0, 0);
                    if (++sb == kN64BStages) { sb = 0; pb ^= 1; }
                    if (++sa == kN64AStages) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128;
            bool first = true;
            const int groups = p.kw * p.cin_chunks + p.x2_chunks;
            for (int g = 0; g < groups; ++g) {
                const int ntaps = g < p.kw * p.cin_chunks ? p.kh : 1;
                mbar_wait(&a_full[sa], pa);
                const uint32_t x_addr = smem_u32(smem_a + sa * kN64ABytes);
                for (int dy = 0; dy < ntaps; ++dy) {
                    mbar_wait(&b_full[sb], pb);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t wdesc = umma_smem_desc_sw128(smem_u32(smem_b + sb * kN64BBytes), 0, 1024);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            // 128 pixel rows of the box: tile rows 8h .. 8h+7, shifted dy rows down
                            const uint64_t xdesc = umma_smem_desc_sw128(x_addr + (dy + 8 * h) * 2048, 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem + h * 64, xdesc + 2 * k, wdesc + 2 * k, kIdesc, (first && k == 0) ? 0u : 1u);
                        }
                        umma_commit(&b_empty[sb]);
                    }
                    __syncwarp();
                    first = false;
                    if (++sb == kN64BStages) { sb = 0; pb ^= 1; }
                }
                if (elect_one()) umma_commit(&a_empty[sa]);
                __syncwarp();
                if (++sa == kN64AStages) { sa = 0; pa ^= 1; }
            }
            if (elect_one()) umma_commit(&tmem_full[acc]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;                       // TMEM lane quarter: pixels 32*ew .. 32*ew+31 of a 128-pixel half
        const int cw = (warp - 4) >> 2;                // column half: channels 32*cw .. 32*cw+31
        const int cbase = cw * 32;
        const bool do_stats = p.stat_sum != nullptr;
        uint8_t* stage0 = s_stage + (warp - 4) * 4096;
        float bias[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) bias[i] = (p.bias && cbase + i < p.cout) ? __ldg(p.bias + cbase + i) : 0.f;
        float rs1[32], rs2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { rs1[i] = 0.f; rs2[i] = 0.f; }
        const int prow = (ew * 32 + static_cast<int>(lane)) >> 4, pcol = static_cast<int>(lane) & 15;   // within the half
        const uint32_t swz = (lane >> 1) & 3u;         // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int n = tile / (p.tiles_x * p.tiles_y);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 128 + h * 64 + cbase, r);
                const int y = ty * 16 + 8 * h + prow, x = tx * 16 + pcol;
                const bool valid = y < p.H && x < p.W;
                uint4 add[4];
                if (p.addend) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) add[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (valid) {
                        const uint4* ap = reinterpret_cast<const uint4*>(
                            p.addend + ((static_cast<size_t>(n) * p.H + y) * p.W + x) * p.out_cs + cbase);
#pragma unroll
                        for (int q = 0; q < 4; ++q) add[q] = __ldg(ap + q);
                    }
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + bias[i];
                if (p.ep_scale) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], __ldg(p.ep_scale + cbase + i), __ldg(p.ep_shift + cbase + i));
                }
                if (p.addend) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&add[q]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __bfloat1622float2(hh[j]);
                            v[q * 8 + 2 * j] += f.x;
                            v[q * 8 + 2 * j + 1] += f.y;
                        }
                    }
                }
                if (p.ep_relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (do_stats && valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        rs1[i] += v[i];
                        rs2[i] = fmaf(v[i], v[i], rs2[i]);
                    }
                }
                // staging tile [32 pixels][32 channels] bf16, 64-byte rows in the SWIZZLE_64B pattern of the store map
                uint8_t* stage = stage0 + h * 2048;
                if (lane == 0) bulk_wait_group_read<1>();      // the store that last read THIS buffer (two tiles... one tile ago) is done
                __syncwarp();
                if (p.dbg != 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
                        u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
                        u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
                        u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
                        *reinterpret_cast<uint4*>(stage + lane * 64 + ((static_cast<uint32_t>(q) ^ swz) << 4)) = u;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&tmap_o, stage, cbase, tx * 16, ty * 16 + 8 * h + 2 * ew, n);   // clipped at the border
                        bulk_commit_group();
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) bulk_wait_group<0>();           // all output tiles written before the CTA retires
        if (do_stats) {
            const float t1 = warp_transpose_reduce(rs1, lane), t2 = warp_transpose_reduce(rs2, lane);
            atomicAdd(&s_stat[cbase + lane], t1);
            atomicAdd(&s_stat[BLOCK_N + cbase + lane], t2);
            flush_stats<BLOCK_N>(p, s_stat, static_cast<int>(threadIdx.x) - 128, 256, reinterpret_cast<int*>(tmem_slot + 1));
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// test / tuning knobs: per calling thread, like the library's error state (no process-global mutable state)
static thread_local int g_conv_debug = 0;
static thread_local int g_conv_variant = 0;  // 0 auto, 1 generic only, 2 halo whenever eligible

static int launch_conv_halo(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = kHAStages * kHABytes + kHBStages * kHBBytes + 1024 + 256 + 2 * 128 * 4 + 128 + 8 * 2048;
    HD_ENSURE_DYN_SMEM(conv_igemm_halo_kernel<false>, smem_bytes);
    HD_ENSURE_DYN_SMEM(conv_igemm_halo_kernel<true>, smem_bytes);
    // store box of the transposed epilogue: 32 channels x 16 columns x 2 rows, dense (un-swizzled) in shared memory
    alignas(64) CUtensorMap to;
    uint64_t dims[4] = {(uint64_t)p.cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t str[3] = {(uint64_t)p.out_cs * 2, (uint64_t)p.W * p.out_cs * 2, (uint64_t)p.H * p.W * p.out_cs * 2};
    uint32_t box[4] = {32, 16, 2, 1};
    int rc = make_tmap_bf16(&to, p.out, 4, dims, str, box, /*swizzle_bytes=*/0);
    if (rc) return rc;
    int grid = p.num_tiles < sm_budget() ? p.num_tiles : sm_budget();
    if (p.bwd_y)
        HD_CHECK_CUDA(::hd::launch_k_pdl(p.num_tiles < sm_count(), conv_igemm_halo_kernel<true>, grid, kHaloThreads,
                                         smem_bytes, stream, tx, tw, to, p));
    else
        HD_CHECK_CUDA(::hd::launch_k_pdl(p.num_tiles < sm_count(), conv_igemm_halo_kernel<false>, grid, kHaloThreads,
                                         smem_bytes, stream, tx, tw,
                                         to, p));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

static int launch_conv_n64(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& tx2, const CUtensorMap& tw2,
                           const ConvParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = kN64AStages * kN64ABytes + kN64BStages * kN64BBytes + 8 * 2 * 2048 + 256 + 2 * 64 * 4 + 64 + 1024;
    HD_ENSURE_DYN_SMEM(conv_igemm_n64_kernel, smem_bytes);
    alignas(64) CUtensorMap to;
    uint64_t dims[4] = {(uint64_t)p.cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t str[3] = {(uint64_t)p.out_cs * 2, (uint64_t)p.W * p.out_cs * 2, (uint64_t)p.H * p.W * p.out_cs * 2};
    uint32_t box[4] = {32, 16, 2, 1};
    int rc = make_tmap_bf16(&to, p.out, 4, dims, str, box, /*swizzle_bytes=*/64);
    if (rc) return rc;
    int grid = p.num_tiles < sm_budget() ? p.num_tiles : sm_budget();
    HD_CHECK_CUDA(::hd::launch_k_pdl(p.num_tiles < sm_count(), conv_igemm_n64_kernel, grid, kN64Threads, smem_bytes, stream,
                                     tx, tw, tx2, tw2, to, p));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

}  // namespace hd

// Test / tuning knob: 0 auto, 1 generic kernel only, 2 halo kernel whenever the shape is eligible.
extern "C" void hd_set_conv_variant(int v) { hd::g_conv_variant = v; }
// Profiling only: see ConvParams::dbg (results are wrong when non-zero).
extern "C" void hd_set_conv_debug(int v) { hd::g_conv_debug = v; }

static int conv_dispatch(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                         const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin, int cout,
                         int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx, int num_stack,
                         const hd_bn_fuse* bn, const float* ep_scale, const float* ep_shift, int ep_relu,
                         cudaStream_t stream, int vtaps = 0, int pad_top = 0, const void* x2 = nullptr,
                         const void* w2_packed = nullptr, int cin2 = 0, const void* bwd_y = nullptr,
                         const float* bwd_sc = nullptr, const float* bwd_sh = nullptr, const hd_bn_bwd_fuse* bwd_fin = nullptr);

static bool halo_eligible(int N, int H, int W, int cout, int block_n, int ksize) {
    using namespace hd;
    if (!((ksize == 3 || ksize == 1) && block_n == 128 && cout == 128 && H >= 16 && W >= 16 && g_conv_variant != 1)) return false;
    const int ht = ((W + 15) / 16) * ((H + 15) / 16) * N;
    return g_conv_variant == 2 || ht >= sm_count();
}

// 1 when a (ksize x ksize, cout-channel NHWC bf16) convolution of this shape runs on the transposed halo kernel - the one
// that can also produce the consumer's BN-backward statistics (hd_conv2d_igemm_bwdstat).
extern "C" int hd_conv2d_igemm_halo_eligible(int N, int H, int W, int cout, int ksize) {
    return halo_eligible(N, H, W, cout, cout > 64 ? 128 : 64, ksize) ? 1 : 0;
}

// See include/hd_b200.h.
extern "C" int hd_conv2d_igemm_bwdstat(const void* x, const void* w_packed, void* out, int N, int H, int W, int cin, int cout,
                                       int ksize, const void* y, const float* act_scale, const float* act_shift,
                                       float* sums, const hd_bn_bwd_fuse* fin, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(y && act_scale && act_shift && sums, "conv_igemm_bwdstat: null argument");
    HD_REQUIRE(fin && fin->coef && fin->counter && fin->gamma && fin->mean && fin->rstd && fin->count > 0.f,
               "conv_igemm_bwdstat: incomplete finalize block");
    HD_REQUIRE(halo_eligible(N, H, W, cout, 128, ksize),
               "conv_igemm_bwdstat: shape (%d,%d,%d) cout %d k %d does not run on the halo kernel (hd_conv2d_igemm_halo_eligible)",
               N, H, W, cout, ksize);
    return conv_dispatch(x, w_packed, out, nullptr, nullptr, nullptr, sums, sums + cout, N, H, W, cin, cout, 128, ksize, 0, cout,
                         0, 0, 1, nullptr, nullptr, nullptr, 0, stream, 0, 0, nullptr, nullptr, 0, y, act_scale, act_shift, fin);
}

// See include/hd_b200.h for the contract.
extern "C" int hd_conv2d_igemm(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                               const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin,
                               int cout, int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx,
                               int num_stack, cudaStream_t stream) {
    return conv_dispatch(x, w_packed, out, out2, bias, addend, stat_sum, stat_sqsum, N, H, W, cin, cout, block_n, ksize,
                         out_mode, out_cs, out2_cs, stack_idx, num_stack, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int hd_conv2d_igemm_bn(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                                  const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin,
                                  int cout, int block_n, int ksize, int out_mode, int out_cs, int out2_cs,
                                  int stack_idx, int num_stack, const hd_bn_fuse* bn, cudaStream_t stream) {
    return conv_dispatch(x, w_packed, out, out2, bias, addend, stat_sum, stat_sqsum, N, H, W, cin, cout, block_n, ksize,
                         out_mode, out_cs, out2_cs, stack_idx, num_stack, bn, nullptr, nullptr, 0, stream);
}

extern "C" int hd_conv2d_igemm_affine(const void* x, const void* w_packed, void* out, const float* bias,
                                      const void* addend, const float* scale, const float* shift, int relu, int N,
                                      int H, int W, int cin, int cout, int block_n, int ksize, int out_cs,
                                      cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE((scale == nullptr) == (shift == nullptr), "conv_igemm_affine: scale and shift come together");
    HD_REQUIRE(block_n == 128 || block_n == 64, "conv_igemm_affine: block_n=%d unsupported", block_n);
    return conv_dispatch(x, w_packed, out, nullptr, bias, addend, nullptr, nullptr, N, H, W, cin, cout, block_n, ksize, 0,
                         out_cs, 0, 0, 1, nullptr, scale, shift, relu, stream);
}

extern "C" int hd_conv2d_igemm_dual(const void* x, const void* w_packed, const void* x2, const void* w2_packed, void* out,
                                    const void* addend, int N, int H, int W, int cin, int cin2, int cout, int block_n,
                                    int ksize, int out_cs, cudaStream_t stream) {
    return conv_dispatch(x, w_packed, out, nullptr, nullptr, addend, nullptr, nullptr, N, H, W, cin, cout, block_n, ksize, 0,
                         out_cs, 0, 0, 1, nullptr, nullptr, nullptr, 0, stream, 0, 0, x2, w2_packed, cin2);
}

extern "C" int hd_conv2d_igemm_vtaps(const void* x, const void* w_packed, void* out, const float* bias, float* stat_sum,
                                     float* stat_sqsum, int N, int H, int W, int cin, int cout, int block_n, int vtaps,
                                     int pad_top, int out_cs, const hd_bn_fuse* bn, const float* scale,
                                     const float* shift, int relu, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(vtaps >= 1, "conv_igemm_vtaps: vtaps=%d", vtaps);
    return conv_dispatch(x, w_packed, out, nullptr, bias, nullptr, stat_sum, stat_sqsum, N, H, W, cin, cout, block_n, 1, 0,
                         out_cs, 0, 0, 1, bn, scale, shift, relu, stream, vtaps, pad_top);
}

static int conv_dispatch(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                         const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin, int cout,
                         int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx, int num_stack,
                         const hd_bn_fuse* bn, const float* ep_scale, const float* ep_shift, int ep_relu,
                         cudaStream_t stream, int vtaps, int pad_top, const void* x2, const void* w2_packed, int cin2,
                         const void* bwd_y, const float* bwd_sc, const float* bwd_sh, const hd_bn_bwd_fuse* bwd_fin) {
    using namespace hd;
    HD_REQUIRE(bn == nullptr || (stat_sum != nullptr && bn->out && bn->counter && bn->gamma && bn->beta),
               "conv_igemm: fused BN finalize needs statistics, gamma/beta, an output block and a ticket counter");
    HD_REQUIRE(cin % 64 == 0 && cin >= 64 && cin <= 512, "conv_igemm: cin=%d must be a multiple of 64", cin);
    HD_REQUIRE(block_n == 128 || block_n == 64 || block_n == 16, "conv_igemm: block_n=%d unsupported", block_n);
    HD_REQUIRE(cout >= 1 && cout <= block_n, "conv_igemm: cout=%d > block_n=%d", cout, block_n);
    HD_REQUIRE(ksize == 1 || ksize == 3, "conv_igemm: ksize=%d unsupported", ksize);
    HD_REQUIRE(N > 0 && H > 0 && W > 0, "conv_igemm: empty tensor");
    HD_REQUIRE(block_n != 16 || (addend == nullptr && stat_sum == nullptr), "conv_igemm: head variant has no addend/stats");
    HD_REQUIRE(block_n == 16 || (out_mode == 0 && cout % 32 == 0 && out_cs % 8 == 0),
               "conv_igemm: NHWC output needs cout %% 32 == 0");
    HD_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "conv_igemm: stat pointers must come in pairs");

    ConvParams p{};
    p.N = N; p.H = H; p.W = W;
    p.cin_chunks = cin / 64;
    p.kh = p.kw = ksize;
    p.pad = (ksize - 1) / 2;
    p.pad_y = p.pad_x = p.pad;
    if (vtaps > 0) {        // a column of `vtaps` vertical taps (rows y - pad_top .. y - pad_top + vtaps - 1), generic kernel only
        HD_REQUIRE(ksize == 1 && vtaps <= 8 && pad_top >= 0 && pad_top < vtaps, "conv_igemm: vtaps=%d pad_top=%d", vtaps, pad_top);
        p.kh = vtaps; p.kw = 1; p.pad_y = pad_top; p.pad_x = 0;
    }
    p.cout = cout;
    int tw = 1 << ilog2_ceil(W); if (tw > 16) tw = 16;
    int th = 1 << ilog2_ceil(H); if (th > 128 / tw) th = 128 / tw;
    int tn = 128 / (tw * th);
    p.tw_log2 = ilog2_ceil(tw); p.th_log2 = ilog2_ceil(th); p.tn_log2 = ilog2_ceil(tn);
    p.tiles_x = (W + tw - 1) / tw; p.tiles_y = (H + th - 1) / th; p.tiles_n = (N + tn - 1) / tn;
    p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.out_mode = out_mode; p.out_cs = out_cs; p.stack_idx = stack_idx; p.num_stack = num_stack;
    p.out = out; p.out2 = reinterpret_cast<__nv_bfloat16*>(out2); p.out2_cs = out2_cs;
    p.bias = bias; p.addend = reinterpret_cast<const __nv_bfloat16*>(addend);
    p.stat_sum = stat_sum; p.stat_sqsum = stat_sqsum;
    p.dbg = g_conv_debug;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift; p.ep_relu = ep_relu;
    if (bn) {
        p.bn_gamma = bn->gamma; p.bn_beta = bn->beta; p.bn_rm = bn->running_mean; p.bn_rv = bn->running_var;
        p.bn_nbt = bn->num_batches_tracked; p.bn_momentum = bn->momentum; p.bn_eps = bn->eps;
        p.bn_count = static_cast<float>(N) * H * W; p.bn_out = bn->out; p.bn_counter = bn->counter;
    }

    // halo kernel: 3x3 (or 1x1: same kernel, one tap, no halo rows), 128 output channels, map >= 16x16 and enough 16x16
    // tiles to fill the machine
    const bool halo = vtaps == 0 && out_mode == 0 && halo_eligible(N, H, W, cout, block_n, ksize);
    if (halo) {
        tw = 16; th = 16 + 2 * p.pad; tn = 1;       // activation box: 16x16 pixels plus the 3x3 halo rows
        p.tiles_x = (W + 15) / 16; p.tiles_y = (H + 15) / 16; p.tiles_n = N;
        p.num_tiles = ((W + 15) / 16) * ((H + 15) / 16) * N;
    }
    if (bwd_y) {
        HD_REQUIRE(halo && bwd_fin && addend == nullptr && bn == nullptr && ep_scale == nullptr && out_cs == cout,
                   "conv_igemm: BN-backward statistics need the halo kernel, a dense output and no other epilogue option");
        p.bwd_y = reinterpret_cast<const __nv_bfloat16*>(bwd_y); p.bwd_sc = bwd_sc; p.bwd_sh = bwd_sh;
        p.bf_gamma = bwd_fin->gamma; p.bf_mean = bwd_fin->mean; p.bf_rstd = bwd_fin->rstd; p.bf_coef = bwd_fin->coef;
        p.bf_dgamma = bwd_fin->dgamma; p.bf_dbeta = bwd_fin->dbeta; p.bf_count = bwd_fin->count;
        p.bn_counter = bwd_fin->counter;
    }
    // N=64 halo kernel: 64 output channels on a map with enough 16x16 tiles to fill the machine (the 256x256 level)
    bool n64 = false;
    if (block_n == 64 && cout == 64 && out_mode == 0 && out2 == nullptr && H >= 16 && W >= 16 && p.kh <= 4 &&
        g_conv_variant != 1) {
        const int ht = ((W + 15) / 16) * ((H + 15) / 16) * N;
        n64 = g_conv_variant == 2 || ht >= sm_count();
        if (n64) {
            tw = 16; th = 16 + p.kh - 1; tn = 1;
            p.tiles_x = (W + 15) / 16; p.tiles_y = (H + 15) / 16; p.tiles_n = N;
            p.num_tiles = ht;
        }
    }
    HD_REQUIRE(x2 == nullptr || (n64 && w2_packed && cin2 >= 64 && cin2 % 64 == 0 && cin2 <= 256),
               "conv_igemm: a fused second 1x1 input needs the 64-output-channel halo kernel (large map, cout 64)");
    alignas(64) CUtensorMap tmx, tmw;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
        int rc = make_tmap_bf16(&tmx, x, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)block_n, (uint64_t)(p.kh * p.kw)};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)block_n * cin * 2};
        uint32_t box[3] = {64, (uint32_t)block_n, 1};
        int rc = make_tmap_bf16(&tmw, w_packed, 3, dims, str, box);
        if (rc) return rc;
    }
    if (n64) {
        alignas(64) CUtensorMap tmx2 = tmx, tmw2 = tmw;
        if (x2) {
            p.x2_chunks = cin2 / 64;
            uint64_t dims[4] = {(uint64_t)cin2, (uint64_t)W, (uint64_t)H, (uint64_t)N};
            uint64_t str[3] = {(uint64_t)cin2 * 2, (uint64_t)W * cin2 * 2, (uint64_t)H * W * cin2 * 2};
            uint32_t box[4] = {64, 16, 16, 1};
            int rc = make_tmap_bf16(&tmx2, x2, 4, dims, str, box);
            if (rc) return rc;
            uint64_t wdims[3] = {(uint64_t)cin2, (uint64_t)block_n, 1};
            uint64_t wstr[2] = {(uint64_t)cin2 * 2, (uint64_t)block_n * cin2 * 2};
            uint32_t wbox[3] = {64, (uint32_t)block_n, 1};
            rc = make_tmap_bf16(&tmw2, w2_packed, 3, wdims, wstr, wbox);
            if (rc) return rc;
        }
        return launch_conv_n64(tmx, tmw, tmx2, tmw2, p, stream);
    }
    if (halo) return launch_conv_halo(tmx, tmw, p, stream);
    if (block_n == 128) return launch_conv<128>(tmx, tmw, p, stream);
    if (block_n == 64) return launch_conv<64>(tmx, tmw, p, stream);
    return launch_conv<16>(tmx, tmw, p, stream);
}
