// HBM-bound elementwise / reduction kernels around the convolutions, all on NHWC bf16 activations with
// 16-byte (8-channel) vector accesses and fp32 math:
//   * train/eval BatchNorm finalize (hourglass.py:103 nn.BatchNorm2d: eps 1e-5, momentum 0.1, biased var to
//     normalise, unbiased var into running_var), BN+activation apply, residual tail relu(bn(y2) + skip)
//     (hourglass.py:125-127), 2x2 max-pool (:72), nearest-upsample + add (:147,:155-156);
//   * their backward counterparts (the autograd graph PyTorch would build for those modules).
// Per-channel BN statistics themselves are produced by the conv epilogue (conv_igemm.cu).
#include <cuda_bf16.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "hd_b200.h"
#include "hd_common.h"

namespace hd {

struct F8 {
    float v[8];
};

__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    F8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ F8 cvt8(const uint4& u) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    F8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}

__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// 8 consecutive per-channel constants from shared memory with two 128-bit loads (the per-element scalar loads of an
// earlier version made these kernels LDS-issue bound instead of HBM bound)
__device__ __forceinline__ void lds8(const float* s, int c0, float (&o)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(s + c0);
    const float4 b = *reinterpret_cast<const float4*>(s + c0 + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// pixels per loop iteration of the single-BN backward kernels (HD_BN_UNROLL = 1 | 2 | 4)
static int bn_bwd_unroll() {
    static const int u = getenv("HD_BN_NO_UNROLL") ? 1 : (getenv("HD_BN_UNROLL") ? atoi(getenv("HD_BN_UNROLL")) : 2);
    return u;
}

static inline int grid_for(size_t work, int block, int max_blocks) {
    size_t g = (work + block - 1) / block;
    if (g > (size_t)max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return (int)g;
}

// Traversal order and grid of the streaming kernels that sit between two convolutions (round 2, last experiment).
// A convolution writes its output tiles front to back, so when it ends the TAIL of that tensor is what the 126 MB L2 still
// holds; a consumer that also walks front to back starts at the evicted head and has displaced the tail by the time it gets
// there (ncu: these kernels read every byte from DRAM). Walking BACK to front the consumer finds the tail in L2, and what it
// writes last - the head of its own output - is what the next convolution reads first. That only works if the whole grid
// sweeps the tensor ONCE: grid = the number of CTAs that are resident at the same time (occupancy x SMs) instead of a
// fixed 16 per SM (2.7 passes at 128x128, each touching every 2.7th line).
// MEASURED (same box, graph replay, ms per step): front-to-back 11.91 / 11.86; reversed 11.75; reversed + single sweep
// 11.84 / 11.85; single sweep alone 11.94. The reuse is real but small (the L2 keeps less of a 134 MB write stream than
// its size suggests), and the single-sweep grid costs what it gains (fewer CTAs in flight beside a weight-gradient CTA).
//   HD_EW_REVERSE=0 : front to back (first version)      HD_EW_SWEEP1=1 : one-sweep grids (default: fixed CTAs per SM)
static bool ew_reverse() {
    const char* e = getenv("HD_EW_REVERSE");
    return !(e != nullptr && e[0] == '0');
}
static bool ew_sweep1() {
    const char* e = getenv("HD_EW_SWEEP1");
    return e != nullptr && e[0] == '1';
}
// CTAs of `kernel` (256 threads, `smem` dynamic bytes) resident on the device at once; `fallback` when not sweeping once
template <typename K>
static int sweep_grid(K kernel, size_t work_items, int items_per_block, size_t smem, int fallback_per_sm) {
    int per_sm = fallback_per_sm;
    if (ew_sweep1()) {
        static std::mutex mu;
        static std::map<std::pair<const void*, size_t>, int> cache;     // occupancy per (kernel, dynamic smem)
        const std::pair<const void*, size_t> key(reinterpret_cast<const void*>(kernel), smem);
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it == cache.end()) {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ <= 0)
                occ = fallback_per_sm;
            it = cache.emplace(key, occ).first;
        }
        per_sm = it->second;
    }
    return grid_for(work_items, items_per_block, sm_count() * per_sm);
}

// ---------------------------------------------------------------------------------------------- BN finalize
// bnp layout per BN layer (fp32, 6*C): scale | shift | mean | rstd | (unused) | (unused)
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sqsum, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked, float momentum, float eps,
                                   int training, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd, int C) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    float mean, var;
    if (training) {
        mean = sum[c] / count;
        var = fmaxf(sqsum[c] / count - mean * mean, 0.f);
        if (running_mean) {
            const float unbiased = count > 1.f ? var * (count / (count - 1.f)) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - mean * sc;
    save_mean[c] = mean;
    save_rstd[c] = rstd;
}

// ---------------------------------------------------------------------------------------------- forward
// z = act(y * scale + shift)
template <bool RELU>
__global__ void bn_act_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                              const float* __restrict__ shift, __nv_bfloat16* __restrict__ z, size_t nvec, int C, int rev) {
    pdl_prologue();
    __shared__ __align__(16) float s_sc[256], s_sh[256];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        s_sc[i] = scale[i];
        s_sh[i] = shift[i];
    }
    __syncthreads();
    // blockDim (256) is a multiple of C/8, so a thread always works on the same 8 channels (rev: vector nvec-1-i, i.e.
    // the mirrored channel group - nvec is a multiple of C/8)
    const int cv = static_cast<int>(threadIdx.x % (C >> 3));
    const int c0 = (rev ? (C >> 3) - 1 - cv : cv) << 3;
    float sc[8], sh[8];
    lds8(s_sc, c0, sc);
    lds8(s_sh, c0, sh);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t e = rev ? nvec - 1 - i : i;
        F8 a = load8(y + e * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = fmaf(a.v[j], sc[j], sh[j]);
            a.v[j] = RELU ? fmaxf(t, 0.f) : t;
        }
        store8(z + e * 8, a);
    }
}

// out = relu(y2 * s2 + b2 + skip), skip = x (identity) or ys * ss + bs (1x1 conv + BN)
template <bool SKIP_BN>
__global__ void bn_add_relu_kernel(const __nv_bfloat16* __restrict__ y2, const float* __restrict__ s2,
                                   const float* __restrict__ b2, const __nv_bfloat16* __restrict__ skip,
                                   const float* __restrict__ ss, const float* __restrict__ bs,
                                   __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ mask, size_t nvec, int C, int rev) {
    pdl_prologue();
    __shared__ __align__(16) float p[4][256];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        p[0][i] = s2[i];
        p[1][i] = b2[i];
        p[2][i] = SKIP_BN ? ss[i] : 1.f;
        p[3][i] = SKIP_BN ? bs[i] : 0.f;
    }
    __syncthreads();
    const int cv = static_cast<int>(threadIdx.x % (C >> 3));
    const int c0 = (rev ? (C >> 3) - 1 - cv : cv) << 3;          // see bn_act_kernel
    float k0[8], k1[8], k2[8], k3[8];
    lds8(p[0], c0, k0);
    lds8(p[1], c0, k1);
    lds8(p[2], c0, k2);
    lds8(p[3], c0, k3);
    for (size_t i0 = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < nvec;
         i0 += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t i = rev ? nvec - 1 - i0 : i0;
        F8 a = load8(y2 + i * 8);
        F8 k = load8(skip + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = fmaf(a.v[j], k0[j], k1[j]);
            float sv = SKIP_BN ? fmaf(k.v[j], k2[j], k3[j]) : k.v[j];
            a.v[j] = fmaxf(t + sv, 0.f);
        }
        store8(out + i * 8, a);
        if (mask) {     // ReLU mask of the STORED (bf16-rounded) output, one bit per channel: backward reads 1 byte, not 16
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) m |= (__bfloat162float(__float2bfloat16_rn(a.v[j])) > 0.f ? 1u : 0u) << j;
            mask[i] = static_cast<uint8_t>(m);
        }
    }
}

// 2x2 max pool, stride 2 (H, W even)
__global__ void maxpool2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                int W, int C) {
    pdl_prologue();
    const int cvec = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        size_t pix = i / cvec;
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        const __nv_bfloat16* base = x + ((static_cast<size_t>(n) * H + 2 * oy) * W + 2 * ox) * C + cv * 8;
        F8 a = load8(base), b = load8(base + C), c = load8(base + static_cast<size_t>(W) * C),
           d = load8(base + static_cast<size_t>(W) * C + C);
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] = fmaxf(fmaxf(a.v[j], b.v[j]), fmaxf(c.v[j], d.v[j]));
        store8(y + i * 8, a);
    }
}

// Residual tail with a BN'd skip branch FUSED with the 2x2 max pool that follows it (PreLayer, hourglass.py:165-166):
//   out = relu(y2*s2+b2 + ys*ss+bs) is never stored; pooled = max over the 2x2 window of the bf16-rounded out values,
//   idx = position (0..3, row-major) of the FIRST maximum - what max_pool2d's backward routes the gradient to.
// At 256x256, B=32 this saves writing (537 MB) and re-reading (537 MB) the block output; backward needs neither: the
// ReLU mask is rebuilt from y2 / ys and the pool backward reads `idx` (67 MB).
__global__ void __launch_bounds__(256, 2)
bn_add_relu_pool_kernel(const __nv_bfloat16* __restrict__ y2, const float* __restrict__ s2, const float* __restrict__ b2,
                        const __nv_bfloat16* __restrict__ ys, const float* __restrict__ ss, const float* __restrict__ bs,
                        __nv_bfloat16* __restrict__ pooled, uint8_t* __restrict__ idx, int N, int H, int W, int C) {
    pdl_prologue();
    __shared__ __align__(16) float p[4][256];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        p[0][i] = s2[i]; p[1][i] = b2[i]; p[2][i] = ss[i]; p[3][i] = bs[i];
    }
    __syncthreads();
    const int cvec = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const int c0 = static_cast<int>(threadIdx.x % cvec) << 3;
    float k0[8], k1[8], k2[8], k3[8];
    lds8(p[0], c0, k0); lds8(p[1], c0, k1); lds8(p[2], c0, k2); lds8(p[3], c0, k3);
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        size_t pix = i / cvec;
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        const size_t o00 = ((static_cast<size_t>(n) * H + 2 * oy) * W + 2 * ox) * C + c0;
        const size_t offs[4] = {o00, o00 + C, o00 + static_cast<size_t>(W) * C, o00 + static_cast<size_t>(W) * C + C};
        F8 best;
        uint32_t bi[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const F8 a = load8(y2 + offs[k]), b = load8(ys + offs[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float o = __bfloat162float(__float2bfloat16_rn(
                    fmaxf(fmaf(a.v[j], k0[j], k1[j]) + fmaf(b.v[j], k2[j], k3[j]), 0.f)));
                if (k == 0 || o > best.v[j]) {
                    best.v[j] = o;
                    bi[j >> 2] = (bi[j >> 2] & ~(0xffu << (8 * (j & 3)))) | (static_cast<uint32_t>(k) << (8 * (j & 3)));
                }
            }
        }
        store8(pooled + i * 8, best);
        *reinterpret_cast<uint2*>(idx + i * 8) = make_uint2(bi[0], bi[1]);
    }
}

// Max-pool backward from stored argmax indices (see above): dx[window k] = (k == idx ? dpool : 0) [+ add1] [+ add2]
__global__ void maxpool2_bwd_idx_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dpool,
                                        const __nv_bfloat16* __restrict__ add1, const __nv_bfloat16* __restrict__ add2,
                                        __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
    pdl_prologue();
    const int cvec = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        size_t pix = i / cvec;
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        const size_t o00 = ((static_cast<size_t>(n) * H + 2 * oy) * W + 2 * ox) * C + cv * 8;
        const size_t offs[4] = {o00, o00 + C, o00 + static_cast<size_t>(W) * C, o00 + static_cast<size_t>(W) * C + C};
        const F8 g = load8(dpool + i * 8);
        const uint2 w = *reinterpret_cast<const uint2*>(idx + i * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            F8 r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t b = ((j < 4 ? w.x : w.y) >> (8 * (j & 3))) & 0xffu;
                r.v[j] = b == static_cast<uint32_t>(k) ? g.v[j] : 0.f;
            }
            if (add1) {
                const F8 a = load8(add1 + offs[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) r.v[j] += a.v[j];
            }
            if (add2) {
                const F8 a = load8(add2 + offs[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) r.v[j] += a.v[j];
            }
            store8(dx + offs[k], r);
        }
    }
}

// out[n,y,x,:] = up1[n,y,x,:] + low[n,y/2,x/2,:]   (nearest x2 upsample + add); H, W are the OUTPUT sizes
__global__ void upsample_add_kernel(const __nv_bfloat16* __restrict__ up1, const __nv_bfloat16* __restrict__ low,
                                    __nv_bfloat16* __restrict__ out, int N, int H, int W, int C) {
    pdl_prologue();
    const int cvec = C >> 3;
    const size_t nvec = static_cast<size_t>(N) * H * W * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        size_t pix = i / cvec;
        const int x = pix % W;
        const int y = (pix / W) % H;
        const int n = pix / (static_cast<size_t>(W) * H);
        F8 a = load8(up1 + i * 8);
        F8 b = load8(low + ((static_cast<size_t>(n) * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * C + cv * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] += b.v[j];
        store8(out + i * 8, a);
    }
}

// ---------------------------------------------------------------------------------------------- backward
// Per-channel reductions for BN backward behind a ReLU:  g = dout * mask
//   sums[0][c] += sum g ; sums[1][c] += sum g * y ; sums[2][c] += sum g * ys   (ys optional)   -- RAW moments; the
// finalize kernel turns them into sum g * yhat = rstd * (sum g*y - mean * sum g). Keeping mean / rstd out of the
// streaming loop leaves it with no per-channel constants at all (mask from `out`) or just scale/shift (REMASK:
// the ReLU mask of a plain conv+BN+ReLU is recomputed as y*scale+shift > 0, which saves reading the activated tensor),
// so the kernel stays at ~50 registers and runs at HBM speed.
template <bool SECOND, bool REMASK, int U = 1>
__global__ void __launch_bounds__(256, SECOND ? 2 : (U >= 4 ? 2 : (U == 2 ? 3 : 4)))   // the two-BN variants need > 64 registers
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                     const float* __restrict__ act_scale, const float* __restrict__ act_shift,
                     const float* __restrict__ act_scale_s, const float* __restrict__ act_shift_s,
                     const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ ys,
                     float* __restrict__ sums, size_t npix, int C, const hd_bn_bwd_fuse fin,
                     const uint8_t* __restrict__ mbits, int rev) {
    pdl_prologue();
    extern __shared__ __align__(16) float red[];  // [3][C]
    const int cvec = C >> 3;
    const int lane_c = threadIdx.x % cvec;          // which 8-channel vector
    const int row = threadIdx.x / cvec;             // pixel lane inside the block
    const int rows = blockDim.x / cvec;
    const int c0 = lane_c << 3;
    // REMASK && SECOND: two-branch residual tail relu(bn(y) + bn_s(ys)) - the mask is rebuilt from both conv outputs,
    // which this kernel reads anyway, exactly as bn_add_relu_kernel<true> computed it
    float asc[8], ash[8], asc2[8], ash2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        asc[j] = REMASK ? act_scale[c0 + j] : 0.f;
        ash[j] = REMASK ? act_shift[c0 + j] : 0.f;
        asc2[j] = (REMASK && SECOND) ? act_scale_s[c0 + j] : 0.f;
        ash2[j] = (REMASK && SECOND) ? act_shift_s[c0 + j] : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
    float a0[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = a2[j] = 0.f;
    const size_t pstride = static_cast<size_t>(gridDim.x) * rows;
    size_t pix = static_cast<size_t>(blockIdx.x) * rows + row;
    if (!SECOND && U > 1) {
        // U pixels per iteration with all their 16-byte loads issued before the first use: beside a weight-gradient
        // CTA only two of these CTAs fit on an SM (instead of eight), so the bytes in flight per THREAD decide the
        // bandwidth this kernel gets in the concurrent windows of the backward pass (measured: 2 pixels -0.19 ms/step).
        for (; pix + (U - 1) * pstride < npix; pix += U * pstride) {
            uint4 rg[U], ry[U], ro[U];
            uint32_t mm[U];
#pragma unroll
            for (int h = 0; h < U; ++h) {
                const size_t pp = pix + h * pstride;
                const size_t o = (rev ? npix - 1 - pp : pp) * C + c0;       // back to front: see ew_reverse()
                rg[h] = ldg16(dout + o);
                ry[h] = ldg16(y + o);
                ro[h] = make_uint4(0, 0, 0, 0);
                mm[h] = 0;
                if (!REMASK) {
                    if (mbits) mm[h] = mbits[o >> 3];
                    else ro[h] = ldg16(out + o);
                }
            }
#pragma unroll
            for (int h = 0; h < U; ++h) {
                const F8 g = cvt8(rg[h]), yy = cvt8(ry[h]);
                F8 o;
                if (!REMASK) {
                    if (mbits) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) o.v[j] = static_cast<float>((mm[h] >> j) & 1u);
                    } else {
                        o = cvt8(ro[h]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float pre = REMASK ? fmaf(yy.v[j], asc[j], ash[j]) : o.v[j];
                    const float gj = pre > 0.f ? g.v[j] : 0.f;
                    a0[j] += gj;
                    a1[j] = fmaf(gj, yy.v[j], a1[j]);
                }
            }
        }
    }
    for (; pix < npix; pix += pstride) {
        const size_t off = (rev ? npix - 1 - pix : pix) * C + c0;
        F8 g = load8(dout + off);
        F8 yy = load8(y + off);
        F8 o, y2;
        if (!REMASK) {
            if (mbits) {        // stored ReLU mask bits (bn_add_relu_kernel) instead of the activated tensor
                const uint32_t m = mbits[off >> 3];
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] = static_cast<float>((m >> j) & 1u);
            } else {
                o = load8(out + off);
            }
        }
        if (SECOND) y2 = load8(ys + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float pre = REMASK ? fmaf(yy.v[j], asc[j], ash[j]) : o.v[j];
            if (REMASK && SECOND) pre += fmaf(y2.v[j], asc2[j], ash2[j]);
            const float gj = pre > 0.f ? g.v[j] : 0.f;
            a0[j] += gj;
            a1[j] = fmaf(gj, yy.v[j], a1[j]);
            if (SECOND) a2[j] = fmaf(gj, y2.v[j], a2[j]);
        }
    }
    __syncthreads();
    // block reduction with a tiny shared footprint (3*C floats) so that several of these CTAs can share an SM with a
    // persistent tcgen05 conv CTA when the executor overlaps them on two streams.
    // lanes l and l^16 (cvec == 16) or l^8, l^16 (cvec == 8) hold the same channels of different pixel rows
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        for (int o2 = cvec; o2 < 32; o2 <<= 1) {
            a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o2);
            a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o2);
            if (SECOND) a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], o2);
        }
    }
    if ((threadIdx.x & 31) < cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&red[c0 + j], a0[j]);
            atomicAdd(&red[C + c0 + j], a1[j]);
            if (SECOND) atomicAdd(&red[2 * C + c0 + j], a2[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(sums + c, red[c]);
        atomicAdd(sums + C + c, red[C + c]);
        if (SECOND) atomicAdd(sums + 2 * C + c, red[2 * C + c]);
    }
    if (fin.coef == nullptr) return;
    // fused finalize: the last block to add its partial sums builds dy = a*g + b*y + c and dgamma / dbeta
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float S0 = __ldcg(sums + c);
        {
            const float g = fin.gamma[c], r = fin.rstd[c], m = fin.mean[c];
            const float S1 = r * (__ldcg(sums + C + c) - m * S0);
            fin.coef[c] = g * r;
            fin.coef[C + c] = -g * r * r * S1 / fin.count;
            fin.coef[2 * C + c] = g * r * (m * r * S1 - S0) / fin.count;
            if (fin.dgamma) fin.dgamma[c] = S1;
            if (fin.dbeta) fin.dbeta[c] = S0;
        }
        if (SECOND) {
            const float g = fin.gamma_s[c], r = fin.rstd_s[c], m = fin.mean_s[c];
            const float S1 = r * (__ldcg(sums + 2 * C + c) - m * S0);
            fin.coef_s[c] = g * r;
            fin.coef_s[C + c] = -g * r * r * S1 / fin.count;
            fin.coef_s[2 * C + c] = g * r * (m * r * S1 - S0) / fin.count;
            if (fin.dgamma_s) fin.dgamma_s[c] = S1;
            if (fin.dbeta_s) fin.dbeta_s[c] = S0;
        }
        // leave the accumulators zeroed for the next reduction that uses this scratch block (stream-ordered)
        sums[c] = 0.f;
        sums[C + c] = 0.f;
        if (SECOND) sums[2 * C + c] = 0.f;
    }
    if (threadIdx.x == 0) *fin.counter = 0u;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reduce + finalize + apply of one single-BN backward in ONE launch, for the small maps of the deep hourglass levels
// (<= 32x32 at batch 32: tensors of <= 8 MB that stay in L2). Those levels are latency-bound - a reduce launch (~15 us)
// and an apply launch (~12 us) per BatchNorm regardless of size - and the two kernels differ only by a grid-wide
// dependency: the coefficients. Here a grid of <= 64 CTAs (always co-resident: nothing this kernel waits for can be
// waiting for it) accumulates the sums, the last CTA to arrive builds the coefficients and raises an epoch flag, every
// CTA waits for it and applies them, re-reading its elements from L2.
//   mask source: REMASK ? (y * act_scale + act_shift > 0) : stored bits (mbits).   WRITE_G: also emit g = dout * mask.
template <bool REMASK, bool WRITE_G>
__global__ void __launch_bounds__(256, 4)
bn_bwd_fused_small_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ mbits,
                          const float* __restrict__ act_scale, const float* __restrict__ act_shift,
                          const __nv_bfloat16* __restrict__ y, float* __restrict__ sums, __nv_bfloat16* __restrict__ dy,
                          __nv_bfloat16* __restrict__ gout, size_t npix, int C, const hd_bn_bwd_fuse fin,
                          unsigned int* __restrict__ epoch) {
    pdl_prologue();
    __shared__ __align__(16) float red[2 * 256];
    __shared__ __align__(16) float kco[3][256];
    __shared__ unsigned int s_e0;
    __shared__ int s_last;
    if (threadIdx.x == 0) s_e0 = atomicAdd(epoch, 0u);          // read BEFORE this CTA's arrival below
    const int cvec = C >> 3;
    const int lane_c = threadIdx.x % cvec, row = threadIdx.x / cvec, rows = blockDim.x / cvec;
    const int c0 = lane_c << 3;
    float asc[8], ash[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        asc[j] = REMASK ? act_scale[c0 + j] : 0.f;
        ash[j] = REMASK ? act_shift[c0 + j] : 0.f;
    }
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) red[i] = 0.f;
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
    const size_t pstride = static_cast<size_t>(gridDim.x) * rows;
    for (size_t pix = static_cast<size_t>(blockIdx.x) * rows + row; pix < npix; pix += pstride) {
        const size_t off = pix * C + c0;
        const F8 g = load8(dout + off), yy = load8(y + off);
        const uint32_t m = REMASK ? 0u : mbits[off >> 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool on = REMASK ? fmaf(yy.v[j], asc[j], ash[j]) > 0.f : ((m >> j) & 1u) != 0u;
            const float gj = on ? g.v[j] : 0.f;
            a0[j] += gj;
            a1[j] = fmaf(gj, yy.v[j], a1[j]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        for (int o2 = cvec; o2 < 32; o2 <<= 1) {
            a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o2);
            a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o2);
        }
    }
    if ((threadIdx.x & 31) < cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&red[c0 + j], a0[j]);
            atomicAdd(&red[C + c0 + j], a1[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(sums + c, red[c]);
        atomicAdd(sums + C + c, red[C + c]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float S0 = __ldcg(sums + c);
            const float g = fin.gamma[c], r = fin.rstd[c], m = fin.mean[c];
            const float S1 = r * (__ldcg(sums + C + c) - m * S0);
            fin.coef[c] = g * r;
            fin.coef[C + c] = -g * r * r * S1 / fin.count;
            fin.coef[2 * C + c] = g * r * (m * r * S1 - S0) / fin.count;
            if (fin.dgamma) fin.dgamma[c] = S1;
            if (fin.dbeta) fin.dbeta[c] = S0;
            sums[c] = 0.f;
            sums[C + c] = 0.f;
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            *fin.counter = 0u;
            __threadfence();
            atomicAdd(epoch, 1u);                   // release: the coefficients are complete
        }
    }
    // every CTA (the last one included) waits for the epoch to move, then applies
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (atomicAdd(epoch, 0u) == s_e0) {
            __nanosleep(32);
            if (++spins > (1u << 25)) __trap();    // ~1 s: surface a lost CTA as an error instead of hanging
        }
        __threadfence();
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        kco[0][c] = __ldcg(fin.coef + c);
        kco[1][c] = __ldcg(fin.coef + C + c);
        kco[2][c] = __ldcg(fin.coef + 2 * C + c);
    }
    __syncthreads();
    float ka[8], kb[8], kc[8];
    lds8(kco[0], c0, ka);
    lds8(kco[1], c0, kb);
    lds8(kco[2], c0, kc);
    for (size_t pix = static_cast<size_t>(blockIdx.x) * rows + row; pix < npix; pix += pstride) {
        const size_t off = pix * C + c0;
        F8 g = load8(dout + off);
        const F8 yy = load8(y + off);
        const uint32_t m = REMASK ? 0u : mbits[off >> 3];
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool on = REMASK ? fmaf(yy.v[j], asc[j], ash[j]) > 0.f : ((m >> j) & 1u) != 0u;
            const float gj = on ? g.v[j] : 0.f;
            g.v[j] = gj;
            r.v[j] = fmaf(ka[j], gj, fmaf(kb[j], yy.v[j], kc[j]));
        }
        store8(dy + off, r);
        if (WRITE_G) store8(gout + off, g);
    }
}

// From the reduction sums build the per-channel affine form of the BN input gradient
//   dy = a * g + b * y + c      with  a = gamma*rstd, b = -gamma*rstd^2*S1/M, c = gamma*rstd*(mean*rstd*S1 - S0)/M
// (S0 = sum g, S1 = sum g*yhat, recovered from the raw moment the reduce kernel accumulates)
// and the parameter gradients dgamma = S1, dbeta = S0 (accumulated when `accumulate`).
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ s0, const float* __restrict__ s1, float count,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ rstd, float* __restrict__ coef,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate, int C) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float g = gamma[c], r = rstd[c], m = mean[c];
    const float S0 = s0[c];
    const float S1 = r * (s1[c] - m * S0);      // s1 holds the raw moment sum g*y: sum g*yhat = rstd*(sum g*y - mean*sum g)
    coef[c] = g * r;
    coef[C + c] = -g * r * r * S1 / count;
    coef[2 * C + c] = g * r * (m * r * S1 - S0) / count;
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + S1 : S1;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + S0 : S0;
}

// g = dout * (out > 0);  dy = a*g + b*y + c ; optionally dys = as*g + bs*ys + cs ; optionally gout = g
template <bool SECOND, bool WRITE_G, int U = 1>
__global__ void __launch_bounds__(256, SECOND ? 2 : (U >= 4 ? 2 : (U == 2 ? 3 : 4))) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                    const float* __restrict__ act_scale, const float* __restrict__ act_shift,
                                    const float* __restrict__ act_scale_s, const float* __restrict__ act_shift_s,
                                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef,
                                    __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ ys,
                                    const float* __restrict__ coef_s, __nv_bfloat16* __restrict__ dys,
                                    __nv_bfloat16* __restrict__ gout, size_t nvec, int C,
                                    const uint8_t* __restrict__ mbits) {
    pdl_prologue();
    __shared__ __align__(16) float p[10][256];
    const bool remask = out == nullptr && mbits == nullptr;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        p[6][i] = remask ? act_scale[i] : 0.f;
        p[7][i] = remask ? act_shift[i] : 0.f;
        p[8][i] = (remask && SECOND) ? act_scale_s[i] : 0.f;
        p[9][i] = (remask && SECOND) ? act_shift_s[i] : 0.f;
        p[0][i] = coef[i];
        p[1][i] = coef[C + i];
        p[2][i] = coef[2 * C + i];
        if (SECOND) {
            p[3][i] = coef_s[i];
            p[4][i] = coef_s[C + i];
            p[5][i] = coef_s[2 * C + i];
        }
    }
    __syncthreads();
    const int c0 = static_cast<int>(threadIdx.x % (C >> 3)) << 3;   // loop-invariant: blockDim is a multiple of C/8
    float ka[8], kb[8], kc[8];
    lds8(p[0], c0, ka);
    lds8(p[1], c0, kb);
    lds8(p[2], c0, kc);
    const size_t istride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (!SECOND && U > 1) {       // U vectors per iteration, loads first (see bn_bwd_reduce_kernel)
        float asc[8], ash[8];
        if (remask) { lds8(p[6], c0, asc); lds8(p[7], c0, ash); }
        for (; i + (U - 1) * istride < nvec; i += U * istride) {
            uint4 rg[U], ry[U], ro[U];
            uint32_t mm[U];
#pragma unroll
            for (int h = 0; h < U; ++h) {
                const size_t ii = i + h * istride;
                rg[h] = ldg16(dout + ii * 8);
                ry[h] = ldg16(y + ii * 8);
                ro[h] = make_uint4(0, 0, 0, 0);
                mm[h] = 0;
                if (!remask) {
                    if (mbits) mm[h] = mbits[ii];
                    else ro[h] = ldg16(out + ii * 8);
                }
            }
#pragma unroll
            for (int h = 0; h < U; ++h) {
                F8 g = cvt8(rg[h]);
                const F8 yy = cvt8(ry[h]);
                F8 o, r;
                if (remask) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o.v[j] = fmaf(yy.v[j], asc[j], ash[j]);
                } else if (mbits) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o.v[j] = static_cast<float>((mm[h] >> j) & 1u);
                } else {
                    o = cvt8(ro[h]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gj = o.v[j] > 0.f ? g.v[j] : 0.f;
                    g.v[j] = gj;
                    r.v[j] = fmaf(ka[j], gj, fmaf(kb[j], yy.v[j], kc[j]));
                }
                const size_t ii = i + h * istride;
                store8(dy + ii * 8, r);
                if (WRITE_G) store8(gout + ii * 8, g);
            }
        }
    }
    for (; i < nvec; i += istride) {
        F8 g = load8(dout + i * 8);
        F8 yy = load8(y + i * 8);
        F8 o, r, r2, y2;
        if (SECOND) y2 = load8(ys + i * 8);
        if (remask) {
            float asc[8], ash[8];
            lds8(p[6], c0, asc);
            lds8(p[7], c0, ash);
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = fmaf(yy.v[j], asc[j], ash[j]);
            if (SECOND) {       // two-branch residual tail: relu(bn(y) + bn_s(ys)), as bn_add_relu_kernel<true>
                lds8(p[8], c0, asc);
                lds8(p[9], c0, ash);
#pragma unroll
                for (int j = 0; j < 8; ++j) o.v[j] += fmaf(y2.v[j], asc[j], ash[j]);
            }
        } else if (mbits) {
            const uint32_t m = mbits[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = static_cast<float>((m >> j) & 1u);
        } else {
            o = load8(out + i * 8);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gj = o.v[j] > 0.f ? g.v[j] : 0.f;
            g.v[j] = gj;
            r.v[j] = fmaf(ka[j], gj, fmaf(kb[j], yy.v[j], kc[j]));
        }
        if (SECOND) {
            float sa[8], sb[8], sc2[8];
            lds8(p[3], c0, sa);
            lds8(p[4], c0, sb);
            lds8(p[5], c0, sc2);
#pragma unroll
            for (int j = 0; j < 8; ++j) r2.v[j] = fmaf(sa[j], g.v[j], fmaf(sb[j], y2.v[j], sc2[j]));
        }
        store8(dy + i * 8, r);
        if (SECOND) store8(dys + i * 8, r2);
        if (WRITE_G) store8(gout + i * 8, g);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-branch BN backward BEHIND the fused relu + 2x2 max pool of PreLayer's Residual(64,128) (bn_add_relu_pool_kernel):
//   g[window k] = (k == idx ? dpool : 0) * (bn(y) + bn_s(ys) > 0)
// is never materialised. Both kernels walk the POOLED tensor - one thread = 8 channels of one 2x2 window, all eight
// 16-byte loads of y / ys in flight before the first use - and read dpool (134 MB at 256x256, B = 32) + idx (67 MB)
// where the unfused sequence wrote the routed gradient (537 MB, hd_maxpool2_bwd_idx) and read it back twice.
// Same arithmetic, per element, as bn_bwd_reduce_kernel<true,true> / bn_bwd_apply_kernel<true,false> on that tensor.
__device__ __forceinline__ void pool_window(uint32_t i, int cvec_log2, int Ho, int Wo, int W, int C, int c0, size_t offs[4]) {
    const uint32_t pix = i >> cvec_log2;
    const uint32_t ox = pix % Wo, t = pix / Wo;
    const uint32_t oy = t % Ho, n = t / Ho;
    const size_t o00 = ((static_cast<size_t>(n) * (2 * Ho) + 2 * oy) * W + 2 * ox) * C + c0;
    offs[0] = o00; offs[1] = o00 + C; offs[2] = o00 + static_cast<size_t>(W) * C; offs[3] = offs[2] + C;
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_pool_kernel(const __nv_bfloat16* __restrict__ dpool, const uint8_t* __restrict__ idx,
                          const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ sc_s,
                          const float* __restrict__ sh_s, const __nv_bfloat16* __restrict__ y,
                          const __nv_bfloat16* __restrict__ ys, float* __restrict__ sums, uint32_t nvec, int Ho, int Wo,
                          int C, int cvec_log2, const hd_bn_bwd_fuse fin) {
    pdl_prologue();
    __shared__ __align__(16) float red[3 * 256];
    const int cvec = 1 << cvec_log2, W = 2 * Wo;
    const int c0 = static_cast<int>(threadIdx.x & (cvec - 1)) << 3;
    float k0[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { k0[j] = sc[c0 + j]; k1[j] = sh[c0 + j]; k2[j] = sc_s[c0 + j]; k3[j] = sh_s[c0 + j]; }
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
    float a0[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = a2[j] = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        size_t offs[4];
        pool_window(i, cvec_log2, Ho, Wo, W, C, c0, offs);
        uint4 ra[4], rb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { ra[k] = ldg16(y + offs[k]); rb[k] = ldg16(ys + offs[k]); }
        const F8 g = load8(dpool + static_cast<size_t>(i) * 8);
        const uint2 w = *reinterpret_cast<const uint2*>(idx + static_cast<size_t>(i) * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const F8 a = cvt8(ra[k]), b = cvt8(rb[k]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t sel = ((j < 4 ? w.x : w.y) >> (8 * (j & 3))) & 0xffu;
                const float pre = fmaf(a.v[j], k0[j], k1[j]) + fmaf(b.v[j], k2[j], k3[j]);
                const float gj = (sel == static_cast<uint32_t>(k) && pre > 0.f) ? g.v[j] : 0.f;
                a0[j] += gj;
                a1[j] = fmaf(gj, a.v[j], a1[j]);
                a2[j] = fmaf(gj, b.v[j], a2[j]);
            }
        }
    }
    __syncthreads();
    // block reduction + "last CTA builds the coefficients", as in bn_bwd_reduce_kernel<true, .>
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        for (int o2 = cvec; o2 < 32; o2 <<= 1) {
            a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o2);
            a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o2);
            a2[j] += __shfl_xor_sync(0xffffffffu, a2[j], o2);
        }
    }
    if (static_cast<int>(threadIdx.x & 31) < cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&red[c0 + j], a0[j]);
            atomicAdd(&red[C + c0 + j], a1[j]);
            atomicAdd(&red[2 * C + c0 + j], a2[j]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) atomicAdd(sums + c, red[c]);
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float S0 = __ldcg(sums + c);
        {
            const float g = fin.gamma[c], r = fin.rstd[c], m = fin.mean[c];
            const float S1 = r * (__ldcg(sums + C + c) - m * S0);
            fin.coef[c] = g * r;
            fin.coef[C + c] = -g * r * r * S1 / fin.count;
            fin.coef[2 * C + c] = g * r * (m * r * S1 - S0) / fin.count;
            if (fin.dgamma) fin.dgamma[c] = S1;
            if (fin.dbeta) fin.dbeta[c] = S0;
        }
        {
            const float g = fin.gamma_s[c], r = fin.rstd_s[c], m = fin.mean_s[c];
            const float S1 = r * (__ldcg(sums + 2 * C + c) - m * S0);
            fin.coef_s[c] = g * r;
            fin.coef_s[C + c] = -g * r * r * S1 / fin.count;
            fin.coef_s[2 * C + c] = g * r * (m * r * S1 - S0) / fin.count;
            if (fin.dgamma_s) fin.dgamma_s[c] = S1;
            if (fin.dbeta_s) fin.dbeta_s[c] = S0;
        }
        sums[c] = 0.f;
        sums[C + c] = 0.f;
        sums[2 * C + c] = 0.f;
    }
    if (threadIdx.x == 0) *fin.counter = 0u;
}

__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_pool_kernel(const __nv_bfloat16* __restrict__ dpool, const uint8_t* __restrict__ idx,
                         const float* __restrict__ sc, const float* __restrict__ sh, const float* __restrict__ sc_s,
                         const float* __restrict__ sh_s, const __nv_bfloat16* __restrict__ y,
                         const __nv_bfloat16* __restrict__ ys, const float* __restrict__ coef,
                         const float* __restrict__ coef_s, __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dys,
                         uint32_t nvec, int Ho, int Wo, int C, int cvec_log2) {
    pdl_prologue();
    __shared__ __align__(16) float p[10][256];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        p[0][i] = coef[i]; p[1][i] = coef[C + i]; p[2][i] = coef[2 * C + i];
        p[3][i] = coef_s[i]; p[4][i] = coef_s[C + i]; p[5][i] = coef_s[2 * C + i];
        p[6][i] = sc[i]; p[7][i] = sh[i]; p[8][i] = sc_s[i]; p[9][i] = sh_s[i];
    }
    __syncthreads();
    const int cvec = 1 << cvec_log2, W = 2 * Wo;
    const int c0 = static_cast<int>(threadIdx.x & (cvec - 1)) << 3;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        size_t offs[4];
        pool_window(i, cvec_log2, Ho, Wo, W, C, c0, offs);
        uint4 ra[4], rb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { ra[k] = ldg16(y + offs[k]); rb[k] = ldg16(ys + offs[k]); }
        const F8 g = load8(dpool + static_cast<size_t>(i) * 8);
        const uint2 w = *reinterpret_cast<const uint2*>(idx + static_cast<size_t>(i) * 8);
        // pass 1: which of the 4 x 8 window elements carry the pooled gradient (bit k*8+j)
        uint32_t on = 0;
        {
            float k0[8], k1[8], k2[8], k3[8];
            lds8(p[6], c0, k0); lds8(p[7], c0, k1); lds8(p[8], c0, k2); lds8(p[9], c0, k3);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const F8 a = cvt8(ra[k]), b = cvt8(rb[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t sel = ((j < 4 ? w.x : w.y) >> (8 * (j & 3))) & 0xffu;
                    const float pre = fmaf(a.v[j], k0[j], k1[j]) + fmaf(b.v[j], k2[j], k3[j]);
                    on |= (sel == static_cast<uint32_t>(k) && pre > 0.f ? 1u : 0u) << (k * 8 + j);
                }
            }
        }
        // pass 2: dy = a*g + b*y + c per branch. (The compiler barriers keep the three constant sets from being live at the
        // same time: 80 floats of per-channel constants + eight 16-byte loads do not fit 128 registers.)
        asm volatile("" ::: "memory");
        {
            float ka[8], kb[8], kc[8];
            lds8(p[0], c0, ka); lds8(p[1], c0, kb); lds8(p[2], c0, kc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const F8 a = cvt8(ra[k]);
                F8 r;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gj = ((on >> (k * 8 + j)) & 1u) ? g.v[j] : 0.f;
                    r.v[j] = fmaf(ka[j], gj, fmaf(kb[j], a.v[j], kc[j]));
                }
                store8(dy + offs[k], r);
            }
        }
        asm volatile("" ::: "memory");
        {
            float ka[8], kb[8], kc[8];
            lds8(p[3], c0, ka); lds8(p[4], c0, kb); lds8(p[5], c0, kc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const F8 b = cvt8(rb[k]);
                F8 r;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gj = ((on >> (k * 8 + j)) & 1u) ? g.v[j] : 0.f;
                    r.v[j] = fmaf(ka[j], gj, fmaf(kb[j], b.v[j], kc[j]));
                }
                store8(dys + offs[k], r);
            }
        }
    }
}

// dx = route(dpool) [+ add1] [+ add2]; the pooled gradient goes to the first maximum of each 2x2 window in
// row-major scan order (PyTorch max_pool2d backward semantics).
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dpool,
                                    const __nv_bfloat16* __restrict__ add1, const __nv_bfloat16* __restrict__ add2,
                                    __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
    pdl_prologue();
    const int cvec = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        size_t pix = i / cvec;
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        const size_t o00 = ((static_cast<size_t>(n) * H + 2 * oy) * W + 2 * ox) * C + cv * 8;
        const size_t offs[4] = {o00, o00 + C, o00 + static_cast<size_t>(W) * C, o00 + static_cast<size_t>(W) * C + C};
        F8 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = load8(x + offs[k]);
        F8 g = load8(dpool + i * 8);
        F8 r[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int best = 0;
            float bv = v[0].v[j];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k].v[j] > bv) { bv = v[k].v[j]; best = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k].v[j] = (k == best) ? g.v[j] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (add1) {
                F8 a = load8(add1 + offs[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[k].v[j] += a.v[j];
            }
            if (add2) {
                F8 a = load8(add2 + offs[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[k].v[j] += a.v[j];
            }
            store8(dx + offs[k], r[k]);
        }
    }
}

// dlow[n,y,x,:] = sum of the 2x2 block of dout (backward of nearest x2 upsample); H, W are dout's sizes
__global__ void sum2x2_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dlow, int N, int H,
                              int W, int C) {
    pdl_prologue();
    const int cvec = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * cvec;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = i % cvec;
        size_t pix = i / cvec;
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        const __nv_bfloat16* base = dout + ((static_cast<size_t>(n) * H + 2 * oy) * W + 2 * ox) * C + cv * 8;
        F8 a = load8(base), b = load8(base + C), c = load8(base + static_cast<size_t>(W) * C),
           d = load8(base + static_cast<size_t>(W) * C + C);
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] = (a.v[j] + b.v[j]) + (c.v[j] + d.v[j]);
        store8(dlow + i * 8, a);
    }
}

// out = a + b (+ c)
__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           const __nv_bfloat16* __restrict__ c, __nv_bfloat16* __restrict__ out, size_t nvec) {
    pdl_prologue();
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        F8 x = load8(a + i * 8), y = load8(b + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) x.v[j] += y.v[j];
        if (c) {
            F8 z = load8(c + i * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) x.v[j] += z.v[j];
        }
        store8(out + i * 8, x);
    }
}

// out[c] (+)= sum over pixels of x[pix, c]   (bias gradients)
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, size_t npix, int C,
                              int cs) {
    pdl_prologue();
    extern __shared__ float red[];
    const int cvec = C >> 3;
    const int lane_c = threadIdx.x % cvec, row = threadIdx.x / cvec, rows = blockDim.x / cvec;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (size_t pix = static_cast<size_t>(blockIdx.x) * rows + row; pix < npix;
         pix += static_cast<size_t>(gridDim.x) * rows) {
        F8 v = load8(x + pix * cs + lane_c * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += v.v[j];
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        for (int o = cvec; o < 32; o <<= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
    if ((threadIdx.x & 31) < cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&red[lane_c * 8 + j], a[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + c, red[c]);
}

}  // namespace hd

using namespace hd;
typedef const void* cvp;
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFW(p) reinterpret_cast<__nv_bfloat16*>(p)

static inline int ew_blocks(size_t nvec) { return grid_for(nvec, 256, sm_count() * 16); }

extern "C" int hd_bn_finalize(const float* sum, const float* sqsum, float count, const float* gamma,
                              const float* beta, float* running_mean, float* running_var,
                              long long* num_batches_tracked, float momentum, float eps, int training, float* scale,
                              float* shift, float* save_mean, float* save_rstd, int C, cudaStream_t stream) {
    HD_REQUIRE(C > 0 && C <= 256, "bn_finalize: C=%d", C);
    HD_REQUIRE(training || (running_mean && running_var), "bn_finalize: eval mode needs running statistics");
    HD_CHECK_CUDA(::hd::launch_k(bn_finalize_kernel, (C + 127) / 128, 128, 0, stream, sum, sqsum, count, gamma, beta,
                                 running_mean, running_var, num_batches_tracked, momentum, eps, training, scale, shift,
                                 save_mean, save_rstd, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_act(cvp y, const float* scale, const float* shift, void* z, long long npix, int C, int relu,
                         cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && C <= 256, "bn_act: C=%d", C);
    const size_t nvec = static_cast<size_t>(npix) * (C / 8);
    if (nvec == 0) return HD_OK;
    if (relu)
        HD_CHECK_CUDA(::hd::launch_k(bn_act_kernel<true>, sweep_grid(bn_act_kernel<true>, nvec, 256, 0, 16), 256, 0, stream,
                                     BF(y), scale, shift, BFW(z), nvec, C, ew_reverse() ? 1 : 0));
    else
        HD_CHECK_CUDA(::hd::launch_k(bn_act_kernel<false>, sweep_grid(bn_act_kernel<false>, nvec, 256, 0, 16), 256, 0, stream,
                                     BF(y), scale, shift, BFW(z), nvec, C, ew_reverse() ? 1 : 0));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_add_relu_mask(cvp y2, const float* s2, const float* b2, cvp skip, const float* ss, const float* bs,
                                   void* out, void* mask_out, long long npix, int C, cudaStream_t stream) {
    uint8_t* mask = reinterpret_cast<uint8_t*>(mask_out);
    HD_REQUIRE(C % 8 == 0 && C <= 256, "bn_add_relu: C=%d", C);
    HD_REQUIRE((ss == nullptr) == (bs == nullptr), "bn_add_relu: skip scale/shift must come in pairs");
    const size_t nvec = static_cast<size_t>(npix) * (C / 8);
    if (nvec == 0) return HD_OK;
    if (ss)
        HD_CHECK_CUDA(::hd::launch_k(bn_add_relu_kernel<true>, sweep_grid(bn_add_relu_kernel<true>, nvec, 256, 0, 16), 256, 0,
                                     stream, BF(y2), s2, b2, BF(skip), ss, bs, BFW(out), mask, nvec, C, ew_reverse() ? 1 : 0));
    else
        HD_CHECK_CUDA(::hd::launch_k(bn_add_relu_kernel<false>, sweep_grid(bn_add_relu_kernel<false>, nvec, 256, 0, 16), 256, 0,
                                     stream, BF(y2), s2, b2, BF(skip), ss, bs, BFW(out), mask, nvec, C, ew_reverse() ? 1 : 0));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_add_relu(cvp y2, const float* s2, const float* b2, cvp skip, const float* ss, const float* bs,
                              void* out, long long npix, int C, cudaStream_t stream) {
    return hd_bn_add_relu_mask(y2, s2, b2, skip, ss, bs, out, nullptr, npix, C, stream);
}

extern "C" int hd_maxpool2(cvp x, void* y, int N, int H, int W, int C, cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2: shape (%d,%d,%d,%d)", N, H, W, C);
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(maxpool2_kernel, ew_blocks(nvec), 256, 0, stream, BF(x), BFW(y), N, H, W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_upsample2_add(cvp up1, cvp low, void* out, int N, int H, int W, int C, cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "upsample2_add: shape (%d,%d,%d,%d)", N, H, W, C);
    const size_t nvec = static_cast<size_t>(N) * H * W * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(upsample_add_kernel, ew_blocks(nvec), 256, 0, stream, BF(up1), BF(low), BFW(out), N, H,
                                 W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_bwd_reduce(cvp dout, cvp out, const float* act_scale, const float* act_shift, cvp y,
                                const float* mean, const float* rstd, cvp ys, const float* mean_s,
                                const float* rstd_s, float* sums, long long npix, int C, cudaStream_t stream) {
    (void)mean; (void)rstd; (void)mean_s; (void)rstd_s;   // the sums are raw moments; hd_bn_bwd_finalize applies mean / rstd
    HD_REQUIRE(out != nullptr || ys == nullptr, "bn_bwd_reduce: a two-branch tail without `out` needs hd_bn_bwd_reduce_fin");
    return hd_bn_bwd_reduce_fin(dout, out, act_scale, act_shift, nullptr, nullptr, y, ys, sums, npix, C, nullptr, stream);
}

static int bn_bwd_reduce_impl(cvp dout, cvp out, const uint8_t* mbits, const float* act_scale, const float* act_shift,
                              const float* act_scale_s, const float* act_shift_s, cvp y, cvp ys, float* sums,
                              long long npix, int C, const hd_bn_bwd_fuse* fin_in, cudaStream_t stream);

extern "C" int hd_bn_bwd_reduce_fin(cvp dout, cvp out, const float* act_scale, const float* act_shift,
                                    const float* act_scale_s, const float* act_shift_s, cvp y, cvp ys, float* sums,
                                    long long npix, int C, const hd_bn_bwd_fuse* fin_in, cudaStream_t stream) {
    return bn_bwd_reduce_impl(dout, out, nullptr, act_scale, act_shift, act_scale_s, act_shift_s, y, ys, sums, npix, C,
                              fin_in, stream);
}

extern "C" int hd_bn_bwd_reduce_fin_mask(cvp dout, cvp mask, cvp y, float* sums, long long npix, int C,
                                         const hd_bn_bwd_fuse* fin_in, cudaStream_t stream) {
    HD_REQUIRE(mask != nullptr, "bn_bwd_reduce_fin_mask: null mask");
    return bn_bwd_reduce_impl(dout, nullptr, reinterpret_cast<const uint8_t*>(mask), nullptr, nullptr, nullptr, nullptr, y,
                              nullptr, sums, npix, C, fin_in, stream);
}

static int bn_bwd_reduce_impl(cvp dout, cvp out, const uint8_t* mbits, const float* act_scale, const float* act_shift,
                              const float* act_scale_s, const float* act_shift_s, cvp y, cvp ys, float* sums,
                              long long npix, int C, const hd_bn_bwd_fuse* fin_in, cudaStream_t stream) {
    hd_bn_bwd_fuse fin{};
    if (fin_in) fin = *fin_in;
    HD_REQUIRE(fin.coef == nullptr || (fin.counter && fin.gamma && fin.mean && fin.rstd && fin.count > 0.f),
               "bn_bwd_reduce: fused finalize needs gamma / mean / rstd / count and a ticket counter");
    HD_REQUIRE(fin.coef == nullptr || ys == nullptr || (fin.coef_s && fin.gamma_s && fin.mean_s && fin.rstd_s),
               "bn_bwd_reduce: fused finalize of the skip branch needs its gamma / mean / rstd / coef");
    HD_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0, "bn_bwd_reduce: C=%d", C);
    HD_REQUIRE(out != nullptr || mbits != nullptr || (act_scale && act_shift),
               "bn_bwd_reduce: need `out`, mask bits or the activation scale/shift");
    HD_REQUIRE(out != nullptr || mbits != nullptr || ys == nullptr || (act_scale_s && act_shift_s),
               "bn_bwd_reduce: rebuilding the mask of a two-branch tail needs the skip branch's scale/shift too");
    HD_REQUIRE(mbits == nullptr || ys == nullptr, "bn_bwd_reduce: mask bits are for single-BN tails");
    if (npix == 0) return HD_OK;
    const int rows = 256 / (C / 8);
    const size_t smem = 3 * static_cast<size_t>(C) * sizeof(float);
    const size_t np = static_cast<size_t>(npix);
    const int unroll = bn_bwd_unroll();      // pixels per iteration of the single-BN variants
    const uint8_t* nomask = nullptr;
    const int rev = ew_reverse() ? 1 : 0;
    int blocks = grid_for(np, rows * 4, sm_count() * 8);
#define HD_RED(U_, RM_, OUT_, MB_)                                                                                   \
    blocks = sweep_grid(bn_bwd_reduce_kernel<false, RM_, U_>, np, rows * 4, smem, 8);                                \
    HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<false, RM_, U_>, blocks, 256, smem, stream, BF(dout), OUT_, act_scale, \
                                 act_shift, act_scale_s, act_shift_s, BF(y), nullptr, sums, np, C, fin, MB_, rev))
    if (unroll > 1 && !ys && mbits) {
        if (unroll >= 4) { HD_RED(4, false, nullptr, mbits); } else { HD_RED(2, false, nullptr, mbits); }
    } else if (unroll > 1 && !ys && out) {
        if (unroll >= 4) { HD_RED(4, false, BF(out), nomask); } else { HD_RED(2, false, BF(out), nomask); }
    } else if (unroll > 1 && !ys) {
        if (unroll >= 4) { HD_RED(4, true, nullptr, nomask); } else { HD_RED(2, true, nullptr, nomask); }
    }
#undef HD_RED
    else if (mbits)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<false, false>, blocks, 256, smem, stream, BF(dout), nullptr,
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), nullptr, sums, np, C, fin,
                                     mbits, rev));
    else if (ys && out)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<true, false>, blocks, 256, smem, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), BF(ys), sums, np, C, fin,
                                     static_cast<const uint8_t*>(nullptr), rev));
    else if (ys)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<true, true>, blocks, 256, smem, stream, BF(dout), nullptr,
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), BF(ys), sums, np, C, fin,
                                     static_cast<const uint8_t*>(nullptr), rev));
    else if (out)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<false, false>, blocks, 256, smem, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), nullptr, sums, np, C, fin,
                                     static_cast<const uint8_t*>(nullptr), rev));
    else
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_kernel<false, true>, blocks, 256, smem, stream, BF(dout), nullptr,
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), nullptr, sums, np, C, fin,
                                     static_cast<const uint8_t*>(nullptr), rev));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

// See include/hd_b200.h. One launch = reduce + coefficients + apply (single BN; mask bits or recomputed mask).
extern "C" int hd_bn_bwd_fused_small(cvp dout, cvp mask, const float* act_scale, const float* act_shift, cvp y, float* sums,
                                     void* dy, void* gout, long long npix, int C, const hd_bn_bwd_fuse* fin,
                                     unsigned int* epoch, cudaStream_t stream) {
    HD_REQUIRE(fin && fin->coef && fin->counter && fin->gamma && fin->mean && fin->rstd && fin->count > 0.f && epoch,
               "bn_bwd_fused_small: incomplete finalize block");
    HD_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0, "bn_bwd_fused_small: C=%d", C);
    HD_REQUIRE(mask != nullptr || (act_scale && act_shift), "bn_bwd_fused_small: need mask bits or the activation scale/shift");
    HD_REQUIRE(mask != nullptr || gout == nullptr, "bn_bwd_fused_small: the masked-gradient output goes with stored mask bits");
    if (npix == 0) return HD_OK;
    const int rows = 256 / (C / 8);
    int blocks = static_cast<int>((npix + rows * 8 - 1) / (rows * 8));      // >= 8 pixels per thread-row and phase
    if (blocks > 64) blocks = 64;                                            // co-residency of the whole grid
    if (blocks < 1) blocks = 1;
    const size_t np = static_cast<size_t>(npix);
    const uint8_t* mb = reinterpret_cast<const uint8_t*>(mask);
    if (mask && gout)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_fused_small_kernel<false, true>, blocks, 256, 0, stream, BF(dout), mb, act_scale,
                                     act_shift, BF(y), sums, BFW(dy), BFW(gout), np, C, *fin, epoch));
    else if (mask)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_fused_small_kernel<false, false>, blocks, 256, 0, stream, BF(dout), mb, act_scale,
                                     act_shift, BF(y), sums, BFW(dy), nullptr, np, C, *fin, epoch));
    else
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_fused_small_kernel<true, false>, blocks, 256, 0, stream, BF(dout), mb, act_scale,
                                     act_shift, BF(y), sums, BFW(dy), nullptr, np, C, *fin, epoch));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_bwd_finalize(const float* s0, const float* s1, float count, const float* gamma,
                                  const float* mean, const float* rstd, float* coef, float* dgamma, float* dbeta,
                                  int accumulate, int C, cudaStream_t stream) {
    HD_REQUIRE(C > 0 && C <= 256, "bn_bwd_finalize: C=%d", C);
    HD_CHECK_CUDA(::hd::launch_k(bn_bwd_finalize_kernel, (C + 127) / 128, 128, 0, stream, s0, s1, count, gamma, mean,
                                 rstd, coef, dgamma, dbeta, accumulate, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

static int bn_bwd_apply_impl(cvp dout, cvp out, const uint8_t* mbits, const float* act_scale, const float* act_shift,
                             const float* act_scale_s, const float* act_shift_s, cvp y, const float* coef, void* dy,
                             cvp ys, const float* coef_s, void* dys, void* gout, long long npix, int C,
                             cudaStream_t stream);

extern "C" int hd_bn_bwd_apply(cvp dout, cvp out, const float* act_scale, const float* act_shift,
                               const float* act_scale_s, const float* act_shift_s, cvp y, const float* coef, void* dy,
                               cvp ys, const float* coef_s, void* dys, void* gout, long long npix, int C,
                               cudaStream_t stream) {
    return bn_bwd_apply_impl(dout, out, nullptr, act_scale, act_shift, act_scale_s, act_shift_s, y, coef, dy, ys, coef_s,
                             dys, gout, npix, C, stream);
}

extern "C" int hd_bn_bwd_apply_mask(cvp dout, cvp mask, cvp y, const float* coef, void* dy, void* gout, long long npix,
                                    int C, cudaStream_t stream) {
    HD_REQUIRE(mask != nullptr, "bn_bwd_apply_mask: null mask");
    return bn_bwd_apply_impl(dout, nullptr, reinterpret_cast<const uint8_t*>(mask), nullptr, nullptr, nullptr, nullptr, y,
                             coef, dy, nullptr, nullptr, nullptr, gout, npix, C, stream);
}

static int bn_bwd_apply_impl(cvp dout, cvp out, const uint8_t* mbits, const float* act_scale, const float* act_shift,
                             const float* act_scale_s, const float* act_shift_s, cvp y, const float* coef, void* dy,
                             cvp ys, const float* coef_s, void* dys, void* gout, long long npix, int C,
                             cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && C <= 256, "bn_bwd_apply: C=%d", C);
    HD_REQUIRE(out != nullptr || mbits != nullptr || (act_scale && act_shift),
               "bn_bwd_apply: need `out`, mask bits or the activation scale/shift");
    HD_REQUIRE(out != nullptr || mbits != nullptr || ys == nullptr || (act_scale_s && act_shift_s),
               "bn_bwd_apply: rebuilding the mask of a two-branch tail needs the skip branch's scale/shift too");
    const size_t nvec = static_cast<size_t>(npix) * (C / 8);
    if (nvec == 0) return HD_OK;
    int blocks = ew_blocks(nvec);
    const int unroll = bn_bwd_unroll();
#define HD_APP(U_, G_)                                                                                               \
    blocks = sweep_grid(bn_bwd_apply_kernel<false, G_, U_>, nvec, 256, 0, 16);                                       \
    HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_kernel<false, G_, U_>, blocks, 256, 0, stream, BF(dout), BF(out), act_scale, \
                                 act_shift, act_scale_s, act_shift_s, BF(y), coef, BFW(dy), nullptr, nullptr, nullptr,  \
                                 G_ ? BFW(gout) : nullptr, nvec, C, mbits))
    if (unroll > 1 && !ys && gout) {
        if (unroll >= 4) { HD_APP(4, true); } else { HD_APP(2, true); }
    } else if (unroll > 1 && !ys) {
        if (unroll >= 4) { HD_APP(4, false); } else { HD_APP(2, false); }
    }
#undef HD_APP
    else if (ys && gout)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_kernel<true, true>, blocks, 256, 0, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), coef, BFW(dy), BF(ys),
                                     coef_s, BFW(dys), BFW(gout), nvec, C, mbits));
    else if (ys)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_kernel<true, false>, blocks, 256, 0, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), coef, BFW(dy), BF(ys),
                                     coef_s, BFW(dys), nullptr, nvec, C, mbits));
    else if (gout)
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_kernel<false, true>, blocks, 256, 0, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), coef, BFW(dy), nullptr,
                                     nullptr, nullptr, BFW(gout), nvec, C, mbits));
    else
        HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_kernel<false, false>, blocks, 256, 0, stream, BF(dout), BF(out),
                                     act_scale, act_shift, act_scale_s, act_shift_s, BF(y), coef, BFW(dy), nullptr,
                                     nullptr, nullptr, nullptr, nvec, C, mbits));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_maxpool2_bwd(cvp x, cvp dpool, cvp add1, cvp add2, void* dx, int N, int H, int W, int C,
                               cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: shape (%d,%d,%d,%d)", N, H, W, C);
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(maxpool2_bwd_kernel, ew_blocks(nvec), 256, 0, stream, BF(x), BF(dpool), BF(add1),
                                 BF(add2), BFW(dx), N, H, W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_add_relu_pool2(cvp y2, const float* s2, const float* b2, cvp ys, const float* ss, const float* bs,
                                    void* pooled, void* idx, int N, int H, int W, int C, cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && H % 2 == 0 && W % 2 == 0,
               "bn_add_relu_pool2: shape (%d,%d,%d,%d)", N, H, W, C);
    HD_REQUIRE(y2 && ys && s2 && b2 && ss && bs && pooled && idx, "bn_add_relu_pool2: null argument");
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(bn_add_relu_pool_kernel, ew_blocks(nvec), 256, 0, stream, BF(y2), s2, b2, BF(ys), ss,
                                 bs,
                                 BFW(pooled), reinterpret_cast<uint8_t*>(idx), N, H, W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

static int pool_dual_shape(const char* who, int N, int H, int W, int C, uint32_t* nvec, int* cvec_log2) {
    HD_REQUIRE((C == 64 || C == 128 || C == 256) && H % 2 == 0 && W % 2 == 0 && N >= 0, "%s: shape (%d,%d,%d,%d)", who, N, H, W, C);
    const size_t nv = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    HD_REQUIRE(nv < (1ull << 31), "%s: %zu pooled vectors exceed the 32-bit index of this kernel", who, nv);
    *nvec = static_cast<uint32_t>(nv);
    *cvec_log2 = C == 64 ? 3 : (C == 128 ? 4 : 5);
    return HD_OK;
}

extern "C" int hd_bn_bwd_reduce_pool_fin(cvp dpool, cvp idx, const float* sc, const float* sh, const float* sc_s,
                                         const float* sh_s, cvp y, cvp ys, float* sums, int N, int H, int W, int C,
                                         const hd_bn_bwd_fuse* fin, cudaStream_t stream) {
    HD_REQUIRE(dpool && idx && sc && sh && sc_s && sh_s && y && ys && sums, "bn_bwd_reduce_pool: null argument");
    HD_REQUIRE(fin && fin->coef && fin->coef_s && fin->counter && fin->gamma && fin->mean && fin->rstd && fin->gamma_s &&
                   fin->mean_s && fin->rstd_s && fin->count > 0.f,
               "bn_bwd_reduce_pool: incomplete finalize block");
    uint32_t nvec = 0;
    int lg = 0;
    if (int rc = pool_dual_shape("bn_bwd_reduce_pool", N, H, W, C, &nvec, &lg)) return rc;
    if (nvec == 0) return HD_OK;
    const int blocks = grid_for(nvec, 256 * 2, sm_count() * 2);
    HD_CHECK_CUDA(::hd::launch_k(bn_bwd_reduce_pool_kernel, blocks, 256, 0, stream, BF(dpool),
                                 reinterpret_cast<const uint8_t*>(idx), sc, sh, sc_s, sh_s, BF(y), BF(ys), sums, nvec, H / 2,
                                 W / 2, C, lg, *fin));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_bn_bwd_apply_pool(cvp dpool, cvp idx, const float* sc, const float* sh, const float* sc_s,
                                    const float* sh_s, cvp y, cvp ys, const float* coef, const float* coef_s, void* dy,
                                    void* dys, int N, int H, int W, int C, cudaStream_t stream) {
    HD_REQUIRE(dpool && idx && sc && sh && sc_s && sh_s && y && ys && coef && coef_s && dy && dys,
               "bn_bwd_apply_pool: null argument");
    uint32_t nvec = 0;
    int lg = 0;
    if (int rc = pool_dual_shape("bn_bwd_apply_pool", N, H, W, C, &nvec, &lg)) return rc;
    if (nvec == 0) return HD_OK;
    const int blocks = grid_for(nvec, 256 * 2, sm_count() * 2);
    HD_CHECK_CUDA(::hd::launch_k(bn_bwd_apply_pool_kernel, blocks, 256, 0, stream, BF(dpool),
                                 reinterpret_cast<const uint8_t*>(idx), sc, sh, sc_s, sh_s, BF(y), BF(ys), coef, coef_s,
                                 BFW(dy), BFW(dys), nvec, H / 2, W / 2, C, lg));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_maxpool2_bwd_idx(cvp idx, cvp dpool, cvp add1, cvp add2, void* dx, int N, int H, int W, int C,
                                   cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd_idx: shape (%d,%d,%d,%d)", N, H, W, C);
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(maxpool2_bwd_idx_kernel, ew_blocks(nvec), 256, 0, stream,
                                 reinterpret_cast<const uint8_t*>(idx), BF(dpool), BF(add1), BF(add2), BFW(dx), N, H, W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_sum2x2(cvp dout, void* dlow, int N, int H, int W, int C, cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "sum2x2: shape (%d,%d,%d,%d)", N, H, W, C);
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(sum2x2_kernel, ew_blocks(nvec), 256, 0, stream, BF(dout), BFW(dlow), N, H, W, C));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_add(cvp a, cvp b, cvp c, void* out, long long nelem, cudaStream_t stream) {
    HD_REQUIRE(nelem % 8 == 0, "add: nelem %% 8 != 0");
    const size_t nvec = static_cast<size_t>(nelem) / 8;
    if (nvec == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(add_kernel, ew_blocks(nvec), 256, 0, stream, BF(a), BF(b), BF(c), BFW(out), nvec));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_colsum(cvp x, float* out, long long npix, int C, int cs, cudaStream_t stream) {
    HD_REQUIRE(C % 8 == 0 && C <= 256 && 256 % (C / 8) == 0 && cs >= C, "colsum: C=%d cs=%d", C, cs);
    if (npix == 0) return HD_OK;
    const int rows = 256 / (C / 8);
    const int blocks = grid_for(static_cast<size_t>(npix), rows * 4, sm_count() * 4);
    HD_CHECK_CUDA(::hd::launch_k(colsum_kernel, blocks, 256, static_cast<size_t>(C) * sizeof(float), stream, BF(x), out,
                                 static_cast<size_t>(npix), C, cs));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------ eval-mode BN fold
namespace hd {
struct BnFoldJob {
    const float* gamma; const float* beta; const float* mean; const float* var;
    float* out; int channels; float eps;
};
__global__ void bn_fold_all_kernel(const BnFoldJob* __restrict__ jobs) {
    pdl_prologue();
    const BnFoldJob j = jobs[blockIdx.x];
    for (int c = threadIdx.x; c < j.channels; c += blockDim.x) {
        const float sc = j.gamma[c] * rsqrtf(j.var[c] + j.eps);
        j.out[c] = sc;
        j.out[j.channels + c] = j.beta[c] - j.mean[c] * sc;
    }
}
}  // namespace hd

extern "C" int hd_bn_fold_all(const void* jobs_host, int njobs, void* jobs_dev, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(njobs > 0 && jobs_dev, "bn_fold_all: empty job table");
    // small pageable -> device copy, stream-ordered (the runtime stages the source before returning); jobs_host == NULL:
    // the caller has already put the table into jobs_dev
    if (jobs_host)
        HD_CHECK_CUDA(cudaMemcpyAsync(jobs_dev, jobs_host, static_cast<size_t>(njobs) * sizeof(BnFoldJob),
                                      cudaMemcpyHostToDevice, stream));
    HD_CHECK_CUDA(::hd::launch_k(bn_fold_all_kernel, njobs, 128, 0, stream,
                                 reinterpret_cast<const BnFoldJob*>(jobs_dev)));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
