// Weight-gradient of the 3x3 / 1x1 convolutions on tcgen05 tensor cores (sm_100a).
// Autograd counterpart of nn.Conv2d inside hourglass.py:100 (`Convolution.convolution`), i.e.
//
//   dW[cout, cin, tap] = sum_{pixel} dY[pixel, cout] * X[pixel + tap, cin]
//
// GEMM view: M = cout (128), N = cin (64 | 128), K = pixels (B*H*W, split across CTAs).
// Both operands are "MN-major" for UMMA: NHWC keeps the channel (M resp. N) contiguous and the
// reduction dimension (pixels) strided, so the same 4-D TMA boxes as the forward kernel are used and
// only the smem descriptors / instruction descriptor change (a_major = b_major = MN).
//
// Work split: grid = tap_groups x ksplit. A CTA owns one tap group (the three dx taps of one dy row for a
// 3x3 kernel, or the single tap of a 1x1) and every ksplit-th 128-pixel tile; it keeps the group's
// accumulators (3 x N fp32 columns) resident in TMEM over its whole pixel range, then writes one fp32
// partial [tap][cout][cin] to the workspace; after a grid-wide barrier every CTA sums its slice of the partials
// into the OIHW gradient (no second launch).
#include <cuda_bf16.h>

#include <cstdlib>

#include "hd_common.h"
#include "hd_ptx.cuh"

namespace hd {

constexpr int kWgThreads = 256;
constexpr int kWgABytes = 2 * 128 * 128;  // dY tile: 128 pixels x 128 cout (two 64-channel atoms)
// dY buffers / X ring stages per variant (shared memory: <= 227 KB with the 1 KB alignment slack, barriers and 16 KB of
// epilogue staging; HD_WGRAD_SHALLOW at build time restores the first version's 2 / 3-4)
#ifdef HD_WGRAD_SHALLOW
__host__ __device__ constexpr int wg_a_bufs(int) { return 2; }
__host__ __device__ constexpr int wg_b_stages(int block_n, int hrows) { return (block_n > 128 || hrows != 0) ? 3 : 4; }
#else
__host__ __device__ constexpr int wg_a_bufs(int block_n) { return block_n == 64 ? 3 : 2; }
__host__ __device__ constexpr int wg_b_stages(int block_n, int hrows) {
    return block_n == 64 ? (hrows != 0 ? 5 : 6) : ((block_n > 128 || hrows != 0) ? 3 : 4);
}
#endif
__host__ __device__ constexpr int wg_smem_bytes(int block_n, int hrows) {
    return wg_a_bufs(block_n) * kWgABytes + wg_b_stages(block_n, hrows) * (hrows != 0 ? hrows : 128) * block_n * 2 + 1024 + 256 +
           4 * 4096;
}
static_assert(wg_smem_bytes(64, 0) <= 232448 && wg_smem_bytes(64, 160) <= 232448 && wg_smem_bytes(64, 176) <= 232448 &&
                  wg_smem_bytes(128, 0) <= 232448 && wg_smem_bytes(128, 160) <= 232448 && wg_smem_bytes(192, 0) <= 232448,
              "conv_wgrad: shared memory");

struct WgradParams {
    int N, H, W;
    int kw, pad, pad_x;        // taps per row; tap offsets dy = tap / kw - pad, dx = tap % kw - pad_x (pad_x == pad unless the
                               // taps are a vertical column: space-to-depth stem, kw = 1, pad = 2, pad_x = 0)
    int taps_per_group;        // 3 (3x3) or 1 (1x1)
    int groups;                // 3 or 1
    int ksplit;
    int tw_log2, th_log2, tn_log2;
    int tiles_x, tiles_y, num_tiles;
    float* ws;                 // [ksplit][taps][128][BLOCK_N]
    int taps;
    int hx0, hy0;              // halo variants: box origin relative to the pixel tile (3x3: -1, -1 + the CTA's dx group;
                               // space-to-depth stem, four vertical taps in ONE 11-row box: 0, -2)
    // in-kernel split-K reduction (replaces the separate wgrad_reduce launch): after a grid-wide barrier on sync[0]
    // every CTA sums its slice of the partials straight into the OIHW gradient
    float* grad;
    int cout, cin_real, accumulate, stem_perm;
    unsigned int* sync;        // [2] zero on entry, zero on exit: arrivals | CTAs done reducing
};

// HALO (3x3 kernels on maps with a 16x8 pixel tile): a CTA owns the three dy taps of one dx COLUMN. The X operand of a
// pixel tile is then ONE box of (8+2) rows x 16 columns at the dx-shifted x coordinate, and the three dy taps are
// the same shared-memory tile read at a row offset of dy*16 pixels = dy*2048 bytes (a whole number of 8-row swizzle
// groups, so the UMMA descriptor just starts later). X traffic per tile drops from 3 x 32 KB to 40 KB.
// HROWS = pixel rows x 16 of that box: 0 (no halo: one box per tap), 160 (3x3: 8 + 2 rows), 176 (the stem's column of four
// vertical taps, 8 + 3 rows: its X operand was fetched four times per tile - 1.3 GB of L2 -> SM traffic per launch at
// 256x256, B = 32, which is what paced the kernel - and is now fetched once).
template <int BLOCK_N, int HROWS>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradParams p) {
    pdl_launch_dependents();
    constexpr bool HALO = HROWS != 0;
    constexpr int kBRows = HALO ? HROWS : 128;
    constexpr int kBBytes = kBRows * BLOCK_N * 2;
    constexpr int kNChunks = BLOCK_N / 64;
    // operand pipeline depth: the 64-input-channel variants (the 256x256 level: K = 2.1 M pixels per launch) were paced
    // by the ~1.9 us latency of a loaded HBM against two dY tiles in flight; their small X tiles leave room for a third
    // dY buffer and a deeper X ring
    constexpr int kABufs = wg_a_bufs(BLOCK_N);
    constexpr int kWgBStages = wg_b_stages(BLOCK_N, HROWS);
    constexpr uint32_t kIdesc = umma_idesc_bf16(BLOCK_N, 1, 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                       // kABufs buffers
    uint8_t* smem_b = smem + kABufs * kWgABytes;  // kWgBStages buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kWgBStages * kBBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgBStages;
    uint64_t* a_full = bars + 2 * kWgBStages;
    uint64_t* a_empty = a_full + kABufs;
    uint64_t* tmem_full = a_empty + kABufs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const int group = blockIdx.x % p.groups;
    const int split = blockIdx.x / p.groups;
    const uint32_t tmem_cols_needed = p.taps_per_group * BLOCK_N;
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmap_dy);
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < kWgBStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kABufs; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // prologue above overlaps the previous kernel's tail (programmatic dependent launch)

    if (warp == 0) {
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, it = 0;
            for (int tile = split; tile < p.num_tiles; tile += p.ksplit, ++it) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int tn = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx << p.tw_log2, y0 = ty << p.th_log2, n0 = tn << p.tn_log2;
                const uint32_t ab = it % kABufs, aphase = (it / kABufs) & 1;
                mbar_wait(&a_empty[ab], aphase ^ 1);
                mbar_arrive_expect_tx(&a_full[ab], kWgABytes);
                tma_load_4d(smem_a + ab * kWgABytes, &tmap_dy, &a_full[ab], 0, x0, y0, n0);
                tma_load_4d(smem_a + ab * kWgABytes + 128 * 128, &tmap_dy, &a_full[ab], 64, x0, y0, n0);
                if constexpr (HALO) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kBBytes);
#pragma unroll
                    for (int c = 0; c < kNChunks; ++c)
                        tma_load_4d(smem_b + stage * kBBytes + c * kBRows * 128, &tmap_x, &full_bar[stage], c * 64,
                                    x0 + group + p.hx0, y0 + p.hy0, n0);
                    if (++stage == kWgBStages) { stage = 0; phase ^= 1; }
                } else {
                    for (int t = 0; t < p.taps_per_group; ++t) {
                        const int tap = group * p.taps_per_group + t;
                        const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad_x;
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], kBBytes);
#pragma unroll
                        for (int c = 0; c < kNChunks; ++c)
                            tma_load_4d(smem_b + stage * kBBytes + c * 128 * 128, &tmap_x, &full_bar[stage], c * 64,
                                        x0 + dx, y0 + dy, n0);
                        if (++stage == kWgBStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        uint32_t stage = 0, phase = 0, it = 0;
        for (int tile = split; tile < p.num_tiles; tile += p.ksplit, ++it) {
            const uint32_t ab = it % kABufs, aphase = (it / kABufs) & 1;
            mbar_wait(&a_full[ab], aphase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem_a + ab * kWgABytes);
            if constexpr (HALO) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sb = smem_u32(smem_b + stage * kBBytes);
                    const uint64_t adesc = umma_smem_desc_sw128(sa, 128 * 128, 1024);
                    const uint64_t bdesc = umma_smem_desc_sw128(sb, kBRows * 128, 1024);
                    constexpr int kTapsHalo = HROWS == 176 ? 4 : 3;
#pragma unroll
                    for (int t = 0; t < kTapsHalo; ++t) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16(tmem_base + t * BLOCK_N, adesc + 128 * k, bdesc + 128 * (k + t), kIdesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == kWgBStages) { stage = 0; phase ^= 1; }
            } else
            for (int t = 0; t < p.taps_per_group; ++t) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sb = smem_u32(smem_b + stage * kBBytes);
                    // MN-major, 128B swizzle: LBO = stride between 64-channel atoms, SBO = stride between
                    // 8-pixel groups along K.
                    const uint64_t adesc = umma_smem_desc_sw128(sa, 128 * 128, 1024);
                    const uint64_t bdesc = umma_smem_desc_sw128(sb, 128 * 128, 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // 16 pixels along K = 16 rows x 128 B = 2048 B -> +128 in (addr >> 4) units
                        umma_bf16(tmem_base + t * BLOCK_N, adesc + 128 * k, bdesc + 128 * k, kIdesc,
                                  (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == kWgBStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&a_empty[ab]);
            __syncwarp();
        }
        if (elect_one()) umma_commit(tmem_full);
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp & 3;
        const int row = ew * 32 + (int)lane;  // cout
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        // Each lane owns one cout row of the accumulator; rows are 4*BLOCK_N bytes apart in the workspace, so the
        // 32x32 fp32 block is transposed through a 4 KB per-warp staging buffer (the dead dY buffer 0 region is not
        // reusable: other CTAs' pipelines are independent, so a dedicated region after the barriers is used) and
        // written as 4 rows x 128 contiguous bytes per instruction.
        float* stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot) + 16 + 15) & ~uintptr_t(15)) + ew * 1024;
        for (int t = 0; t < p.taps_per_group; ++t) {
            const int tap = HALO ? t * p.kw + group : group * p.taps_per_group + t;
            float* dst0 = p.ws + ((static_cast<size_t>(split) * p.taps + tap) * 128 + ew * 32) * BLOCK_N;
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t r[32];
                tmem_ld_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + t * BLOCK_N + c0, r);
                tmem_ld_wait();
                // stage[row = lane][32 floats], 16-byte pieces XOR-swizzled by the row
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(stage + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                        make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                    __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                __syncwarp();
                const int q = lane & 7;
#pragma unroll
                for (int sidx = 0; sidx < 8; ++sidx) {
                    const int rr = 4 * sidx + (lane >> 3);
                    const float4 w = *reinterpret_cast<const float4*>(stage + rr * 32 + ((q ^ (rr & 7)) << 2));
                    *reinterpret_cast<float4*>(dst0 + static_cast<size_t>(rr) * BLOCK_N + c0 + q * 4) = w;
                }
                __syncwarp();
            }
        }
        __threadfence();       // this CTA's partials are visible device-wide before it announces itself below
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);

    // ---- split-K reduction inside the kernel. The separate reduce launch (576 small CTAs) was the kernel the
    // weight-gradient stream stalled on: whenever a persistent convolution of the main stream owns every SM (all
    // registers) its CTAs cannot be placed at all - the trace showed the 12 us reduction taking 280-470 us at 256x256
    // and delaying every weight gradient queued behind it. Here the grid (<= one CTA per SM, all co-resident once
    // scheduled; nothing this kernel waits for depends on it) meets at a counter, then CTA b sums slice b of the
    // partials (they are L2-resident: 29 MB per layer) in the fixed order k = 0..ksplit-1 - deterministic - into the
    // OIHW gradient.
    if (p.sync == nullptr) return;
    const unsigned G = gridDim.x;
    if (threadIdx.x == 0) {
        atomicAdd(&p.sync[0], 1u);
        unsigned spins = 0;
        while (atomicAdd(&p.sync[0], 0u) < G) {
            __nanosleep(64);
            if (++spins > (1u << 25)) __trap();       // ~2 s: a lost CTA must surface as an error, not as a hang
        }
        __threadfence();
    }
    __syncthreads();
    {
        const int total = p.taps * p.cout * p.cin_real;
        const int chunk = (total + static_cast<int>(G) - 1) / static_cast<int>(G);
        const int begin = static_cast<int>(blockIdx.x) * chunk;
        const int end = begin + chunk < total ? begin + chunk : total;
        const size_t stride = static_cast<size_t>(p.taps) * 128 * BLOCK_N;
        for (int idx = begin + static_cast<int>(threadIdx.x); idx < end; idx += kWgThreads) {
            const int ci = idx % p.cin_real;
            const int co = (idx / p.cin_real) % p.cout;
            const int tap = idx / (p.cin_real * p.cout);
            const float* src = p.ws + (static_cast<size_t>(tap) * 128 + co) * BLOCK_N + ci;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int k = 0;
            for (; k + 8 <= p.ksplit; k += 8) {
                const float a0 = __ldcg(src + (k + 0) * stride), a1 = __ldcg(src + (k + 1) * stride);
                const float a2 = __ldcg(src + (k + 2) * stride), a3 = __ldcg(src + (k + 3) * stride);
                const float a4 = __ldcg(src + (k + 4) * stride), a5 = __ldcg(src + (k + 5) * stride);
                const float a6 = __ldcg(src + (k + 6) * stride), a7 = __ldcg(src + (k + 7) * stride);
                s0 += a0 + a4; s1 += a1 + a5; s2 += a2 + a6; s3 += a3 + a7;
            }
            for (; k < p.ksplit; ++k) s0 += __ldcg(src + k * stride);
            const float sum = (s0 + s1) + (s2 + s3);
            float* g = p.grad + (static_cast<size_t>(co) * p.cin_real + ci) * p.taps + tap;
            if (p.stem_perm == 1) {
                const int c = ci % 3, kk = ci / 3;   // kk = ky*7 + kx
                g = p.grad + (static_cast<size_t>(co) * 3 + c) * 49 + kk;
            } else if (p.stem_perm == 2) {           // space-to-depth stem (stem.cu): tap = dy, ci = dx*12 + (c*2+sy)*2 + sx
                const int dx = ci / 12, q = ci % 12, c = q >> 2, sy = (q >> 1) & 1, sx = q & 1;
                const int ky = 2 * tap + sy - 1, kx = 2 * dx + sx - 1;
                if (ky < 0 || kx < 0) continue;      // the zero-weight taps of the 8x8 -> 7x7 embedding
                g = p.grad + (static_cast<size_t>(co) * 3 + c) * 49 + ky * 7 + kx;
            }
            *g = p.accumulate ? (*g + sum) : sum;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&p.sync[1], 1u) == G - 1) {   // last CTA out: leave the words zeroed
        p.sync[0] = 0u;
        p.sync[1] = 0u;
        __threadfence();
    }
}

// grad[co][ci][tap] (+)= sum_s ws[s][tap][co][ci_pad]   (OIHW fp32, the layout of nn.Conv2d.weight.grad)
// stem_perm: the GEMM K index of the 7x7 stem is k = (ky*7 + kx)*3 + c (im2col order, stem.cu); scatter it
// back to the OIHW position [co][c][ky][kx].
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ grad, int ksplit, int taps,
                                    int cout, int cin, int rows_pad, int cin_pad, int accumulate, int stem_perm) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over taps * cout * cin, ci fastest
    const int total = taps * cout * cin;
    if (idx >= total) return;
    const int ci = idx % cin;
    const int co = (idx / cin) % cout;
    const int tap = idx / (cin * cout);
    const size_t stride = static_cast<size_t>(taps) * rows_pad * cin_pad;
    const float* src = ws + (static_cast<size_t>(tap) * rows_pad + co) * cin_pad + ci;
    // four independent partial sums: eight loads in flight per thread (the loop was a chain of dependent adds behind
    // one load each: 12 us for 29 MB, and this kernel is what the weight-gradient stream waits on)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 8 <= ksplit; k += 8) {
        const float a0 = __ldcs(src + (k + 0) * stride), a1 = __ldcs(src + (k + 1) * stride);
        const float a2 = __ldcs(src + (k + 2) * stride), a3 = __ldcs(src + (k + 3) * stride);
        const float a4 = __ldcs(src + (k + 4) * stride), a5 = __ldcs(src + (k + 5) * stride);
        const float a6 = __ldcs(src + (k + 6) * stride), a7 = __ldcs(src + (k + 7) * stride);
        s0 += a0 + a4; s1 += a1 + a5; s2 += a2 + a6; s3 += a3 + a7;
    }
    for (; k < ksplit; ++k) s0 += __ldcs(src + k * stride);
    const float s = (s0 + s1) + (s2 + s3);
    float* g = grad + (static_cast<size_t>(co) * cin + ci) * taps + tap;
    if (stem_perm == 1) {
        const int c = ci % 3, kk = ci / 3;   // kk = ky*7 + kx
        g = grad + (static_cast<size_t>(co) * 3 + c) * 49 + kk;
    } else if (stem_perm == 2) {             // space-to-depth stem (stem.cu): tap = dy, ci = dx*12 + (c*2+sy)*2 + sx
        const int dx = ci / 12, q = ci % 12, c = q >> 2, sy = (q >> 1) & 1, sx = q & 1;
        const int ky = 2 * tap + sy - 1, kx = 2 * dx + sx - 1;
        if (ky < 0 || kx < 0) return;        // the zero-weight taps of the 8x8 -> 7x7 embedding
        g = grad + (static_cast<size_t>(co) * 3 + c) * 49 + ky * 7 + kx;
    }
    *g = accumulate ? (*g + s) : s;
}

template <int BLOCK_N, int HROWS>
static int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, cudaStream_t stream) {
    constexpr int smem_bytes = wg_smem_bytes(BLOCK_N, HROWS);
    HD_ENSURE_DYN_SMEM((conv_wgrad_kernel<BLOCK_N, HROWS>), smem_bytes);
    HD_CHECK_CUDA(::hd::launch_k_pdl(p.groups * p.ksplit < sm_count() / 2, conv_wgrad_kernel<BLOCK_N, HROWS>,
                                     p.groups * p.ksplit,
                                     kWgThreads, smem_bytes, stream, tdy, tx, p));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

}  // namespace hd

static int wgrad_ksplit(int N, int H, int W, int ksize, int sms);
extern "C" int hd_conv2d_wgrad_ksplit(int N, int H, int W, int ksize) {
    return wgrad_ksplit(N, H, W, ksize, hd::sm_count());      // the upper bound the workspace is sized for
}
static int wgrad_ksplit(int N, int H, int W, int ksize, int sms) {
    using namespace hd;
    int tw = 1 << ilog2_ceil(W); if (tw > 16) tw = 16;
    int th = 1 << ilog2_ceil(H); if (th > 128 / tw) th = 128 / tw;
    int tn = 128 / (tw * th);
    int tiles = ((W + tw - 1) / tw) * ((H + th - 1) / th) * ((N + tn - 1) / tn);
    int groups = ksize == 3 ? 3 : 1;
    int ks = sms / groups;
    if (ks < 1) ks = 1;
    if (ks > tiles) ks = tiles;
    return ks;
}

extern "C" size_t hd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int cin, int ksize) {
    int ks = hd_conv2d_wgrad_ksplit(N, H, W, ksize);
    // at least 4 taps: the space-to-depth stem (stem_perm 2) runs as ksize 1 with four vertical taps
    const int taps = (ksize == 1 && cin == 64) ? 4 : ksize * ksize;
    return static_cast<size_t>(ks) * taps * 128 * cin * sizeof(float) + 256;      // + the two barrier words (see below)
}

// See include/hd_b200.h.
extern "C" int hd_conv2d_wgrad_sync(const void* x, const void* dy, float* grad_w, void* workspace, int N, int H, int W,
                                    int cin, int cin_real, int cout, int ksize, int accumulate, int stem_perm,
                                    unsigned int* sync_words, cudaStream_t stream);

// `workspace`: hd_conv2d_wgrad_workspace_bytes(); its last 256 bytes hold the kernel's two barrier words, zeroed here.
extern "C" int hd_conv2d_wgrad(const void* x, const void* dy, float* grad_w, void* workspace, int N, int H, int W,
                               int cin, int cin_real, int cout, int ksize, int accumulate, int stem_perm,
                               cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(workspace != nullptr, "conv_wgrad: null workspace");
    const size_t bytes = hd_conv2d_wgrad_workspace_bytes(N, H, W, cin, ksize);
    unsigned int* sync = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(workspace) + bytes - 256);
    HD_CHECK_CUDA(cudaMemsetAsync(sync, 0, 2 * sizeof(unsigned int), stream));
    return hd_conv2d_wgrad_sync(x, dy, grad_w, workspace, N, H, W, cin, cin_real, cout, ksize, accumulate, stem_perm, sync,
                                stream);
}

// Same with caller-owned barrier words: two unsigned ints that are ZERO on entry; the kernel leaves them zero again, so
// a sequence of weight-gradient launches on one stream (the network's backward pass) needs no memset in between.
extern "C" int hd_conv2d_wgrad_sync(const void* x, const void* dy, float* grad_w, void* workspace, int N, int H, int W,
                                    int cin, int cin_real, int cout, int ksize, int accumulate, int stem_perm,
                                    unsigned int* sync_words, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(sync_words != nullptr, "conv_wgrad: null barrier words");
    // cout == 64: the second 64-channel atom of the dY tile is fetched out of bounds and zero-filled by TMA.
    HD_REQUIRE(cout == 128 || cout == 64, "conv_wgrad: cout=%d (tensor-core path needs 64 or 128)", cout);
    HD_REQUIRE(cin == 64 || cin == 128 || (cin == 192 && ksize == 1), "conv_wgrad: cin=%d unsupported", cin);
    HD_REQUIRE(stem_perm != 1 || (cin == 192 && cin_real == 147 && ksize == 1), "conv_wgrad: stem_perm 1 needs K=147/192");
    HD_REQUIRE(stem_perm != 2 || (cin == 64 && cin_real == 48 && ksize == 1 && cout == 64),
               "conv_wgrad: stem_perm 2 (space-to-depth stem) needs cin 64 (48 real), cout 64");
    HD_REQUIRE(cin_real >= 1 && cin_real <= cin, "conv_wgrad: cin_real=%d", cin_real);
    HD_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: ksize=%d unsupported", ksize);
    HD_REQUIRE(N > 0 && H > 0 && W > 0, "conv_wgrad: empty tensor");
    WgradParams p{};
    p.N = N; p.H = H; p.W = W;
    p.kw = ksize; p.pad = (ksize - 1) / 2; p.pad_x = p.pad;
    p.taps = ksize * ksize;
    p.groups = ksize == 3 ? 3 : 1;
    p.taps_per_group = ksize == 3 ? 3 : 1;
    if (stem_perm == 2) {       // a column of four vertical taps (rows y-2 .. y+1) kept resident by ONE tap group
        p.kw = 1; p.pad = 2; p.pad_x = 0; p.taps = 4; p.groups = 1; p.taps_per_group = 4;
    }
    int tw = 1 << ilog2_ceil(W); if (tw > 16) tw = 16;
    int th = 1 << ilog2_ceil(H); if (th > 128 / tw) th = 128 / tw;
    int tn = 128 / (tw * th);
    p.tw_log2 = ilog2_ceil(tw); p.th_log2 = ilog2_ceil(th); p.tn_log2 = ilog2_ceil(tn);
    p.tiles_x = (W + tw - 1) / tw; p.tiles_y = (H + th - 1) / th;
    p.num_tiles = p.tiles_x * p.tiles_y * ((N + tn - 1) / tn);
    p.ksplit = wgrad_ksplit(N, H, W, ksize, sm_budget());
    p.ws = reinterpret_cast<float*>(workspace);
    p.grad = grad_w; p.cout = cout; p.cin_real = cin_real; p.accumulate = accumulate; p.stem_perm = stem_perm;
    // Default: partials + a second, separate reduction launch. HD_WGRAD_FUSED_REDUCE=1 reduces inside the kernel behind
    // the grid-wide barrier instead - built to stop the separate launch from starving behind persistent convolutions,
    // but MEASURED SLOWER (round 2, same box: 12.46 / 12.54 vs 12.43 / 12.37 ms per step; 2 stacks 10.30 vs 10.17): CTAs
    // spinning at the barrier hold their SMs, and 147 CTAs reduce more slowly than 576 small ones.
    static const bool fused = getenv("HD_WGRAD_FUSED_REDUCE") != nullptr;
    p.sync = fused ? sync_words : nullptr;
    const bool halo = ksize == 3 && tw == 16 && th == 8;   // the dy taps become row offsets of one 10-row X tile
    static const bool no_stem_halo = getenv("HD_NO_STEM_WGRAD_HALO") != nullptr;
    const bool halo4 = stem_perm == 2 && tw == 16 && th == 8 && !no_stem_halo;   // four vertical taps: one 11-row X tile
    p.hx0 = halo4 ? 0 : -1; p.hy0 = halo4 ? -2 : -1;

    alignas(64) CUtensorMap tdy, tx;
    {
        uint64_t dims[4] = {(uint64_t)cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cout * 2, (uint64_t)W * cout * 2, (uint64_t)H * W * cout * 2};
        uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
        int rc = make_tmap_bf16(&tdy, dy, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
        uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)(halo ? th + 2 : (halo4 ? th + 3 : th)), (uint32_t)tn};
        int rc = make_tmap_bf16(&tx, x, 4, dims, str, box);
        if (rc) return rc;
    }
    int rc = (cin == 192)   ? launch_wgrad<192, 0>(tdy, tx, p, stream)
             : (cin == 128) ? (halo ? launch_wgrad<128, 160>(tdy, tx, p, stream) : launch_wgrad<128, 0>(tdy, tx, p, stream))
             : halo4        ? launch_wgrad<64, 176>(tdy, tx, p, stream)
                            : (halo ? launch_wgrad<64, 160>(tdy, tx, p, stream) : launch_wgrad<64, 0>(tdy, tx, p, stream));
    if (rc) return rc;
    if (p.sync == nullptr) {
        const int total = p.taps * cout * cin_real;
        HD_CHECK_CUDA(::hd::launch_k(wgrad_reduce_kernel, (total + 255) / 256, 256, 0, stream, p.ws, grad_w, p.ksplit,
                                     p.taps, cout, cin_real, 128, cin, accumulate, stem_perm));
        HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    }
    return HD_OK;
}
