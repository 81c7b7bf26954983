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
    int dbg;                  // profiling only (hd_set_conv_debug): 1 = epilogue drains TMEM but skips math/stores,
                              // 2 = MMA issue skipped, 3 = weight tiles loaded only for the first tile
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
    if (p.bn_out == nullptr) return;
    __threadfence();
    named_bar_sync(1, nthr);
    if (t == 0) *s_flag = (atomicAdd(p.bn_counter, 1u) == gridDim.x - 1) ? 1 : 0;
    named_bar_sync(1, nthr);
    if (*s_flag == 0) return;
    __threadfence();
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
                    const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad;
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
    static bool attr_set = false;
    if (!attr_set) {
        HD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem_bytes));
        attr_set = true;
    }
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    conv_igemm_kernel<BLOCK_N><<<grid, BLOCK_N >= 64 ? 384 : kThreads, smem_bytes, stream>>>(tx, tw, p);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// "Halo" variant for the FLOP-dominant case (3x3, 128 output channels, maps >= 16x16): the generic kernel above
// re-fetches the activation tile for each of the 9 taps and the weights for every 128-pixel tile, which makes it
// L2->SM bandwidth bound (measured 14.3 TB/s, ~54 % of the tensor peak). Here a CTA tile is 16x16 pixels (M = 256:
// two 128-row accumulators sharing every weight tile), and for each dx ONE activation box of 18 rows x 16 columns is
// loaded; the three dy taps read that tile at a row offset of dy*16 pixels = dy*2048 B (whole 8-row swizzle groups,
// so only the UMMA descriptor start address moves). L2->SM bytes per FLOP drop 2.3x.
constexpr int kHARows = 18 * 16;
constexpr int kHABytes = kHARows * 128;   // 36,864
constexpr int kHAStages = 3;
constexpr int kHBBytes = 128 * 128;       // one (tap, 64-channel chunk) weight tile for 128 output channels
constexpr int kHBStages = 5;

// 12 warps: producer, MMA issuer, TMEM allocator, (idle), and EIGHT epilogue warps - warps 4-7 drain accumulator half 0,
// warps 8-11 half 1 (a warp may only touch the TMEM lane quarter warp_idx % 4). With four warps the BN-statistics
// epilogue (62 shuffles per 32 columns) was the pacing stage: 148 us with statistics vs 119 us without.
constexpr int kHaloThreads = 384;

__global__ void __launch_bounds__(kHaloThreads, 1)
conv_igemm_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const ConvParams p) {
    constexpr int BLOCK_N = 128;
    constexpr uint32_t kIdesc = umma_idesc_bf16(BLOCK_N, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kHAStages * kHABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kHBStages * kHBBytes);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kHAStages;
    uint64_t* b_full = a_empty + kHAStages;
    uint64_t* b_empty = b_full + kHBStages;
    uint64_t* tmem_full = b_empty + kHBStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);  // [2][128]
    uint8_t* s_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_stat + 2 * BLOCK_N) + 15) & ~uintptr_t(15));   // [8 epilogue warps][2 KB]

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
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

    if (warp == 0) {
        if (elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int n = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * 16, y0 = ty * 16;
                for (int dx = 0; dx < 3; ++dx) {
                    for (int ck = 0; ck < p.cin_chunks; ++ck) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        mbar_arrive_expect_tx(&a_full[sa], kHABytes);
                        tma_load_4d(smem_a + sa * kHABytes, &tmap_x, &a_full[sa], ck * 64, x0 + dx - 1, y0 - 1, n);
                        for (int dy = 0; dy < 3; ++dy) {
                            mbar_wait(&b_empty[sb], pb ^ 1);
                            if (p.dbg == 3 && tile != static_cast<int>(blockIdx.x)) {
                                mbar_arrive(&b_full[sb]);
                                if (++sb == kHBStages) { sb = 0; pb ^= 1; }
                                continue;
                            }
                            mbar_arrive_expect_tx(&b_full[sb], kHBBytes);
                            tma_load_3d(smem_b + sb * kHBBytes, &tmap_w, &b_full[sb], ck * 64, 0, dy * 3 + dx);
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
            for (int dx = 0; dx < 3; ++dx) {
                for (int ck = 0; ck < p.cin_chunks; ++ck) {
                    mbar_wait(&a_full[sa], pa);
                    const uint32_t a_addr = smem_u32(smem_a + sa * kHABytes);
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + sb * kHBBytes), 0, 1024);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                // rows of accumulator h = tile rows 8h..8h+7, shifted down by dy: (dy + 8h) * 16 pixels
                                const uint64_t adesc = umma_smem_desc_sw128(a_addr + (dy + 8 * h) * 2048, 0, 1024);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (p.dbg != 2)
                                        umma_bf16(d_tmem + h * 128, adesc + 2 * k, bdesc + 2 * k, kIdesc,
                                                  (first && k == 0) ? 0u : 1u);
                            }
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
        const int ew = warp & 3;
        const int row = ew * 32 + (int)lane;
        const bool do_stats = p.stat_sum != nullptr;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int n = tile / (p.tiles_x * p.tiles_y);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            {
                const int h = (warp - 4) >> 2;            // accumulator half this warp drains
                const int x = tx * 16 + (row & 15);
                const int y = ty * 16 + 8 * h + (row >> 4);
                const bool valid = (x < p.W) && (y < p.H);
                const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 256 + h * 128;
                if (p.dbg != 1) epilogue_rows<BLOCK_N>(p, taddr, valid, pix, lane, s_stat, do_stats, s_stage + (warp - 4) * 2048);
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
        }
        if (do_stats) flush_stats<BLOCK_N>(p, s_stat, static_cast<int>(threadIdx.x) - 128, 256, reinterpret_cast<int*>(tmem_slot + 1));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) of the halo kernel. The halo kernel is limited by the shared-memory operand port:
// an SS-mode M=128 N=128 K=16 MMA re-reads 4 KB of A and 4 KB of B per 64 cycles (= the 128 B/clk port), and the TMA
// fill traffic shares that port. Here two CTAs of a cluster cooperate on one 16x16-pixel tile (M = 256):
//   * each CTA owns 8 tile rows: its activation box is 10 rows x 16 columns per dx (20 KB instead of 36 KB);
//   * each CTA holds HALF of the output channels of ALL 9 x chunks weight tiles, loaded once and kept resident
//     (147 KB for cin = 128) - weights are never re-streamed and each SM reads only N/2 rows of B per MMA;
//   * the leader CTA's MMA thread issues tcgen05.mma.cta_group::2 (M=256, N=128); accumulators live in both CTAs' TMEM;
//   * TMA completions of both CTAs are signalled on the leader's mbarriers, tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs, epilogue warps of both CTAs release the accumulator
//     on the leader's barrier.
constexpr int kPARows = 10 * 16;
constexpr int kPABytes = kPARows * 128;   // 20,480
constexpr int kPAStages = 3;
constexpr int kPBTile = 64 * 128;         // one (tap, chunk) half-tile: 64 output channels x 64 k

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w64,
                       const ConvParams p) {
    constexpr int BLOCK_N = 128;
    constexpr uint32_t kIdesc = umma_idesc_bf16(BLOCK_N, 0, 0, 256);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_tiles = 9 * p.cin_chunks;
    uint8_t* smem_b = smem;                               // resident weights: b_tiles x 8 KB
    uint8_t* smem_a = smem + 18 * kPBTile;                // activation ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + kPAStages * kPABytes);
    uint64_t* a_full = bars;                              // leader's are used (count 2: one arrive per CTA + tx bytes)
    uint64_t* a_empty = a_full + kPAStages;               // local, arrived by the multicast commit
    uint64_t* b_full = a_empty + kPAStages;               // leader's is used
    uint64_t* tmem_full = b_full + 1;                     // local, multicast commit
    uint64_t* tmem_empty = tmem_full + 2;                 // leader's are used (count 2 x 128)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);
    uint8_t* s_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_stat + 2 * BLOCK_N) + 15) & ~uintptr_t(15));   // [4 epilogue warps][2 KB]

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w64);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kPAStages; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); }
        mbar_init(b_full, 2);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
        fence_mbar_init();
    }
    if (threadIdx.x < 2 * BLOCK_N) s_stat[threadIdx.x] = 0.f;
    cluster_sync();                                       // barriers of both CTAs initialised before any remote use
    if (warp == 2) tmem_alloc_2sm(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // resident weights: this CTA's 64 output channels of every (tap, chunk) tile
            const uint32_t bfull0 = mapa_u32(smem_u32(b_full), 0);
            if (leader) mbar_arrive_expect_tx(b_full, 2u * b_tiles * kPBTile);
            else mbar_arrive_cluster(bfull0);
            for (int tap = 0; tap < 9; ++tap)
                for (int ck = 0; ck < p.cin_chunks; ++ck)
                    tma_load_3d_2sm(smem_b + (tap * p.cin_chunks + ck) * kPBTile, &tmap_w64, bfull0, ck * 64,
                                    64 * static_cast<int>(rank), tap);
            uint32_t sa = 0, pa = 0;
            for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int n = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * 16, y0 = ty * 16 + 8 * static_cast<int>(rank);
                for (int dx = 0; dx < 3; ++dx) {
                    for (int ck = 0; ck < p.cin_chunks; ++ck) {
                        mbar_wait(&a_empty[sa], pa ^ 1);
                        const uint32_t afull0 = mapa_u32(smem_u32(&a_full[sa]), 0);
                        if (leader) mbar_arrive_expect_tx(&a_full[sa], 2u * kPABytes);
                        else mbar_arrive_cluster(afull0);
                        tma_load_4d_2sm(smem_a + sa * kPABytes, &tmap_x, afull0, ck * 64, x0 + dx - 1, y0 - 1, n);
                        if (++sa == kPAStages) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && leader) {
        uint32_t sa = 0, pa = 0, it = 0;
        mbar_wait(b_full, 0);
        tc_fence_after();
        for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128;
            bool first = true;
            for (int dx = 0; dx < 3; ++dx) {
                for (int ck = 0; ck < p.cin_chunks; ++ck) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = smem_u32(smem_a + sa * kPABytes);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint64_t adesc = umma_smem_desc_sw128(a_addr + dy * 2048, 0, 1024);
                            const uint64_t bdesc = umma_smem_desc_sw128(
                                smem_u32(smem_b + ((dy * 3 + dx) * p.cin_chunks + ck) * kPBTile), 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc,
                                              (first && dy == 0 && k == 0) ? 0u : 1u);
                        }
                        umma_commit_2sm(&a_empty[sa], 3);
                    }
                    __syncwarp();
                    first = false;
                    if (++sa == kPAStages) { sa = 0; pa ^= 1; }
                }
            }
            if (elect_one()) umma_commit_2sm(&tmem_full[acc], 3);
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int row = ew * 32 + (int)lane;
        const bool do_stats = p.stat_sum != nullptr;
        uint32_t it = 0;
        for (int tile = cluster_id; tile < p.num_tiles; tile += num_clusters, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int n = tile / (p.tiles_x * p.tiles_y);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int x = tx * 16 + (row & 15);
            const int y = ty * 16 + 8 * static_cast<int>(rank) + (row >> 4);
            const bool valid = (x < p.W) && (y < p.H);
            const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 128;
            epilogue_rows<BLOCK_N>(p, taddr, valid, pix, lane, s_stat, do_stats, s_stage + (warp - 4) * 2048);
            tc_fence_before();
            mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        }
        if (do_stats) flush_stats<BLOCK_N>(p, s_stat, static_cast<int>(threadIdx.x) - 128, 128, reinterpret_cast<int*>(tmem_slot + 1));
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();                                       // nobody exits while the peer may still signal its barriers
    tc_fence_after();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, 256);
}

static int g_conv_debug = 0;
static bool g_pair_default = false;   // the CTA-pair kernel is correct but measured slower than the halo kernel (profiles/README.md)
static int g_conv_variant = 0;  // 0 auto, 1 generic only, 2 halo whenever eligible, 3 CTA-pair whenever eligible

static int launch_conv_pair(const CUtensorMap& tx, const CUtensorMap& tw64, const ConvParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = 18 * kPBTile + kPAStages * kPABytes + 1024 + 256 + 2 * 128 * 4 + 4 * 2048;
    static bool attr_set = false;
    if (!attr_set) {
        HD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem_bytes));
        attr_set = true;
    }
    int clusters = sm_count() / 2;
    if (p.num_tiles < clusters) clusters = p.num_tiles;
    conv_igemm_pair_kernel<<<2 * clusters, kThreads, smem_bytes, stream>>>(tx, tw64, p);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

static int launch_conv_halo(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = kHAStages * kHABytes + kHBStages * kHBBytes + 1024 + 256 + 2 * 128 * 4 + 8 * 2048;
    static bool attr_set = false;
    if (!attr_set) {
        HD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem_bytes));
        attr_set = true;
    }
    int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    conv_igemm_halo_kernel<<<grid, kHaloThreads, smem_bytes, stream>>>(tx, tw, p);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

}  // namespace hd

// Test / tuning knob: 0 auto, 1 generic kernel only, 2 halo kernel whenever the shape is eligible.
extern "C" void hd_set_conv_variant(int v) { hd::g_conv_variant = v; }
// Profiling only: see ConvParams::dbg (results are wrong when non-zero).
extern "C" void hd_set_conv_debug(int v) { hd::g_conv_debug = v; }

extern "C" int hd_conv2d_igemm_bn(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                                  const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin,
                                  int cout, int block_n, int ksize, int out_mode, int out_cs, int out2_cs,
                                  int stack_idx, int num_stack, const hd_bn_fuse* bn, cudaStream_t stream);

// See include/hd_b200.h for the contract.
extern "C" int hd_conv2d_igemm(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                               const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin,
                               int cout, int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx,
                               int num_stack, cudaStream_t stream) {
    return hd_conv2d_igemm_bn(x, w_packed, out, out2, bias, addend, stat_sum, stat_sqsum, N, H, W, cin, cout, block_n,
                              ksize, out_mode, out_cs, out2_cs, stack_idx, num_stack, nullptr, stream);
}

extern "C" int hd_conv2d_igemm_bn(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                                  const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin,
                                  int cout, int block_n, int ksize, int out_mode, int out_cs, int out2_cs,
                                  int stack_idx, int num_stack, const hd_bn_fuse* bn, cudaStream_t stream) {
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
    if (bn) {
        p.bn_gamma = bn->gamma; p.bn_beta = bn->beta; p.bn_rm = bn->running_mean; p.bn_rv = bn->running_var;
        p.bn_nbt = bn->num_batches_tracked; p.bn_momentum = bn->momentum; p.bn_eps = bn->eps;
        p.bn_count = static_cast<float>(N) * H * W; p.bn_out = bn->out; p.bn_counter = bn->counter;
    }

    // halo variant: 3x3, 128 output channels, map >= 16x16 and enough 16x16 tiles to fill the machine
    bool halo = false, pair = false;
    if (ksize == 3 && block_n == 128 && H >= 16 && W >= 16 && out_mode == 0 && g_conv_variant != 1) {
        const int ht = ((W + 15) / 16) * ((H + 15) / 16) * N;
        pair = g_conv_variant == 3 || (g_conv_variant == 0 && g_pair_default && ht >= sm_count());
        halo = !pair && (g_conv_variant == 2 || ht >= sm_count());
        if (halo || pair) {
            tw = 16; th = pair ? 10 : 18; tn = 1;
            p.tiles_x = (W + 15) / 16; p.tiles_y = (H + 15) / 16; p.tiles_n = N;
            p.num_tiles = ht;
        }
    }
    alignas(64) CUtensorMap tmx, tmw;
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
        int rc = make_tmap_bf16(&tmx, x, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)block_n, (uint64_t)(ksize * ksize)};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)block_n * cin * 2};
        uint32_t box[3] = {64, (uint32_t)block_n, 1};
        int rc = make_tmap_bf16(&tmw, w_packed, 3, dims, str, box);
        if (rc) return rc;
    }
    if (pair) {
        alignas(64) CUtensorMap tmw64;
        uint64_t dims[3] = {(uint64_t)cin, (uint64_t)block_n, (uint64_t)(ksize * ksize)};
        uint64_t str[2] = {(uint64_t)cin * 2, (uint64_t)block_n * cin * 2};
        uint32_t box[3] = {64, 64, 1};
        int rc = make_tmap_bf16(&tmw64, w_packed, 3, dims, str, box);
        if (rc) return rc;
        return launch_conv_pair(tmx, tmw64, p, stream);
    }
    if (halo) return launch_conv_halo(tmx, tmw, p, stream);
    if (block_n == 128) return launch_conv<128>(tmx, tmw, p, stream);
    if (block_n == 64) return launch_conv<64>(tmx, tmw, p, stream);
    return launch_conv<16>(tmx, tmw, p, stream);
}
