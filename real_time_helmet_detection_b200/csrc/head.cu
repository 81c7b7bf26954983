// Backward of the 1x1 prediction head (hourglass.py:189-195: conv 128 -> num_cls+4, bias, linear).
// cout = 6 is far below a UMMA tile, and the op is HBM-bound (reads the 128-channel feature map once), so this is
// a fused SIMT kernel:   g = dlogits (+ extra) ;  dfeat[p, k] = sum_c g[p, c] * W[c, k] ;
//                        dW[c, k] += sum_p g[p, c] * feat[p, k] ;  dbias[c] += sum_p g[p, c].
#include <cuda_bf16.h>

#include <cstdlib>

#include "hd_common.h"

namespace hd {

constexpr int kHeadMaxC = 16;
constexpr int kHeadTile = 64;   // pixels per CTA iteration

// 256 threads: 16 lanes (8 feature channels each) x 16 pixel rows; a tile of 64 pixels is handled as 4 pixels per
// thread (4 independent 16-byte loads in flight). The per-pixel gradient vectors g[c] are staged coalesced in
// shared memory first. dW partials stay in registers for the CTA's whole pixel range.
template <int MAXC>
__global__ void __launch_bounds__(256, MAXC <= 8 ? 2 : 1)
head_bwd_kernel(const float* __restrict__ dlogits, long long bs, int HW, int cout,
                const __nv_bfloat16* __restrict__ extra, int extra_cs, const __nv_bfloat16* __restrict__ feat,
                const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dw,
                float* __restrict__ dbias, long long npix) {
    pdl_prologue();
    constexpr int GP = MAXC <= 8 ? 8 : 16;                 // per-pixel gradient vector padded to whole float4s
    __shared__ __align__(16) float s_w[MAXC][128];
    __shared__ __align__(16) float s_g[kHeadTile][GP];
    __shared__ float s_red[16][MAXC * 8 + 1];
    for (int i = threadIdx.x; i < MAXC * 128; i += blockDim.x)
        s_w[i / 128][i % 128] = (i / 128) < cout ? __bfloat162float(wp[i]) : 0.f;
    const int lane_c = threadIdx.x & 15, row = threadIdx.x >> 4;
    const int k0 = lane_c * 8;
    float acc[MAXC][8];
    float accb[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        accb[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
    }
    const long long ntiles = (npix + kHeadTile - 1) / kHeadTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * kHeadTile;
        __syncthreads();
        for (int i = threadIdx.x; i < MAXC * kHeadTile; i += blockDim.x) {
            const int c = i / kHeadTile, lp = i - c * kHeadTile;
            const long long pix = p0 + lp;
            float g = 0.f;
            if (c < cout && pix < npix) {
                const long long n = pix / HW, p = pix - n * HW;
                g = dlogits[n * bs + static_cast<long long>(c) * HW + p];
                if (extra) g += __bfloat162float(extra[pix * extra_cs + c]);
            }
            s_g[lp][c] = g;
        }
        __syncthreads();
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long pix = p0 + row + 16 * q;
            u[q] = pix < npix ? *reinterpret_cast<const uint4*>(feat + pix * 128 + k0) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int lp = row + 16 * q;
            const long long pix = p0 + lp;
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
            float f[8], d[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 t = __bfloat1622float2(h[j]);
                f[2 * j] = t.x;
                f[2 * j + 1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = 0.f;
            // the pixel's gradient vector and the weight rows come in as 16-byte shared-memory reads (2 + 2 per channel
            // instead of 1 + 8 scalar ones): the kernel was instruction-issue bound (ncu r02: 228 instructions per
            // pixel and thread, 36 % issue utilisation, 0.30 of the HBM rate)
            float gv[GP];
#pragma unroll
            for (int v4 = 0; v4 < GP / 4; ++v4) {
                const float4 t = *reinterpret_cast<const float4*>(&s_g[lp][4 * v4]);
                gv[4 * v4] = t.x; gv[4 * v4 + 1] = t.y; gv[4 * v4 + 2] = t.z; gv[4 * v4 + 3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                const float g = gv[c];
                accb[c] += g;
                const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][k0]);
                const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][k0 + 4]);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    d[j] = fmaf(g, wv[j], d[j]);
                    acc[c][j] = fmaf(g, f[j], acc[c][j]);
                }
            }
            if (pix < npix) {
                uint4 o;
                __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(d[2 * j], d[2 * j + 1]);
                *reinterpret_cast<uint4*>(dfeat + pix * 128 + k0) = o;
            }
        }
    }
    // reduce the 16 pixel rows of the block, one channel-vector (lane_c) at a time
    for (int lc = 0; lc < 16; ++lc) {
        __syncthreads();
        if (lane_c == lc) {
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) s_red[row][c * 8 + j] = acc[c][j];
        }
        __syncthreads();
        if (threadIdx.x < cout * 8) {
            float s = 0.f;
            for (int r = 0; r < 16; ++r) s += s_red[r][threadIdx.x];
            const int c = threadIdx.x >> 3, j = threadIdx.x & 7;
            atomicAdd(dw + c * 128 + lc * 8 + j, s);
        }
    }
    __syncthreads();
    if (lane_c == 0) {
#pragma unroll
        for (int c = 0; c < MAXC; ++c) s_red[row][c] = accb[c];
    }
    __syncthreads();
    if (threadIdx.x < cout) {
        float s = 0.f;
        for (int r = 0; r < 16; ++r) s += s_red[r][threadIdx.x];
        atomicAdd(dbias + threadIdx.x, s);
    }
}


// Round-2 variant for cout <= 8 (the 6-channel CenterNet head): 32 lanes x 4 feature channels per pixel, 8 pixels per
// thread and tile. The first version (8 channels per thread, below) kept 48 accumulators + 48 weights per thread, spilled
// at its 128-register budget and was instruction-issue bound (ncu: 222 instructions per pixel and thread, 132 us =
// 0.29 of the HBM rate). Here a thread holds 24 accumulators and 24 weights in registers, the bias gradient is summed where
// the gradient vectors are staged, and the block reduction is shared-memory atomics instead of 16 barrier rounds.
template <int MAXC>
__global__ void __launch_bounds__(256, 2)
head_bwd4_kernel(const float* __restrict__ dlogits, long long bs, int HW, int cout,
                 const __nv_bfloat16* __restrict__ extra, int extra_cs, const __nv_bfloat16* __restrict__ feat,
                 const __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dw,
                 float* __restrict__ dbias, long long npix) {
    pdl_prologue();
    static_assert(MAXC <= 8, "gradient vectors are staged as two float4");
    __shared__ __align__(16) float s_g[kHeadTile][8];
    __shared__ float s_dw[MAXC * 128];
    __shared__ float s_db[MAXC];
    const int lane_c = threadIdx.x & 31, row = threadIdx.x >> 5;      // 32 channel groups x 8 pixel rows
    const int k0 = lane_c * 4;
    float w[MAXC][4], acc[MAXC][4];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w[c][j] = c < cout ? __bfloat162float(wp[c * 128 + k0 + j]) : 0.f;
            acc[c][j] = 0.f;
        }
    for (int i = threadIdx.x; i < MAXC * 128; i += blockDim.x) s_dw[i] = 0.f;
    if (threadIdx.x < MAXC) s_db[threadIdx.x] = 0.f;
    // staging: element i of a tile is (c = i / 64, pixel = i % 64); a thread visits i = tid and tid + 256, i.e. it always
    // meets the same one or two channels -> two private bias-gradient partials
    float db0 = 0.f, db1 = 0.f;
    const long long ntiles = (npix + kHeadTile - 1) / kHeadTile;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p0 = tile * kHeadTile;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < (8 * kHeadTile) / 256; ++it) {
            const int i = threadIdx.x + it * 256;
            const int c = i / kHeadTile, lp = i - c * kHeadTile;
            const long long pix = p0 + lp;
            float g = 0.f;
            if (c < cout && pix < npix) {
                const int n = static_cast<int>(pix / HW);
                const int p = static_cast<int>(pix - static_cast<long long>(n) * HW);
                g = dlogits[n * bs + static_cast<long long>(c) * HW + p];
                if (extra) g += __bfloat162float(extra[pix * extra_cs + c]);
            }
            s_g[lp][c] = g;
            if (it == 0) db0 += g; else db1 += g;
        }
        __syncthreads();
        uint2 u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const long long pix = p0 + row + 8 * q;
            u[q] = pix < npix ? __ldg(reinterpret_cast<const uint2*>(feat + pix * 128 + k0)) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int lp = row + 8 * q;
            const long long pix = p0 + lp;
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
            const float2 f01 = __bfloat1622float2(h[0]), f23 = __bfloat1622float2(h[1]);
            const float f[4] = {f01.x, f01.y, f23.x, f23.y};
            const float4 ga = *reinterpret_cast<const float4*>(&s_g[lp][0]);
            const float4 gb = *reinterpret_cast<const float4*>(&s_g[lp][4]);
            const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    d[j] = fmaf(gv[c], w[c][j], d[j]);
                    acc[c][j] = fmaf(gv[c], f[j], acc[c][j]);
                }
            if (pix < npix) {
                uint2 o;
                __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
                oh[0] = __floats2bfloat162_rn(d[0], d[1]);
                oh[1] = __floats2bfloat162_rn(d[2], d[3]);
                *reinterpret_cast<uint2*>(dfeat + pix * 128 + k0) = o;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&s_dw[c * 128 + k0 + j], acc[c][j]);
    {
        const int c0 = threadIdx.x / kHeadTile, c1 = (threadIdx.x + 256) / kHeadTile;
        if (c0 < MAXC) atomicAdd(&s_db[c0], db0);
        if (c1 < MAXC) atomicAdd(&s_db[c1], db1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cout * 128; i += blockDim.x) atomicAdd(dw + i, s_dw[i]);
    if (threadIdx.x < cout) atomicAdd(dbias + threadIdx.x, s_db[threadIdx.x]);
}

}  // namespace hd

// dlogits: fp32 NCHW slice with batch stride `bs` (elements) and channel stride H*W; feat/dfeat: NHWC bf16 (128 ch);
// wp: packed forward weights [>=cout][128] bf16; dw [cout][128] and dbias [cout] are ACCUMULATED into.
extern "C" int hd_head_backward(const float* dlogits, long long bs, const void* extra, int extra_cs, const void* feat,
                                const void* wp, void* dfeat, float* dw, float* dbias, int N, int H, int W, int cout,
                                cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(cout >= 1 && cout <= kHeadMaxC, "head_backward: cout=%d", cout);
    HD_REQUIRE(N > 0 && H > 0 && W > 0, "head_backward: empty tensor");
    const long long npix = static_cast<long long>(N) * H * W;
    long long g = (npix + kHeadTile - 1) / kHeadTile;
    const long long cap = static_cast<long long>(sm_count()) * 2;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    static const bool old_kernel = getenv("HD_HEAD_BWD_V1") != nullptr;
    if (cout <= 6 && !old_kernel) {
        long long g4 = (npix + kHeadTile - 1) / kHeadTile;
        const long long cap4 = static_cast<long long>(sm_count()) * 2;
        if (g4 > cap4) g4 = cap4;
        HD_CHECK_CUDA(::hd::launch_k(head_bwd4_kernel<6>, static_cast<unsigned>(g4), 256, 0, stream, dlogits, bs, H * W, cout,
                                     reinterpret_cast<const __nv_bfloat16*>(extra), extra_cs,
                                     reinterpret_cast<const __nv_bfloat16*>(feat),
                                     reinterpret_cast<const __nv_bfloat16*>(wp),
                                     reinterpret_cast<__nv_bfloat16*>(dfeat), dw, dbias, npix));
    } else if (cout <= 6)
        HD_CHECK_CUDA(::hd::launch_k(head_bwd_kernel<6>, static_cast<unsigned>(g), 256, 0, stream,  dlogits, bs, H * W,
                                     cout, reinterpret_cast<const __nv_bfloat16*>(extra), extra_cs,
                                     reinterpret_cast<const __nv_bfloat16*>(feat),
                                     reinterpret_cast<const __nv_bfloat16*>(wp),
                                     reinterpret_cast<__nv_bfloat16*>(dfeat), dw, dbias, npix));
    else if (cout <= 8)
        HD_CHECK_CUDA(::hd::launch_k(head_bwd_kernel<8>, static_cast<unsigned>(g), 256, 0, stream,  dlogits, bs, H * W,
                                     cout, reinterpret_cast<const __nv_bfloat16*>(extra), extra_cs,
                                     reinterpret_cast<const __nv_bfloat16*>(feat),
                                     reinterpret_cast<const __nv_bfloat16*>(wp),
                                     reinterpret_cast<__nv_bfloat16*>(dfeat), dw, dbias, npix));
    else
        HD_CHECK_CUDA(::hd::launch_k(head_bwd_kernel<16>, static_cast<unsigned>(g), 256, 0, stream,  dlogits, bs, H * W,
                                     cout, reinterpret_cast<const __nv_bfloat16*>(extra), extra_cs,
                                     reinterpret_cast<const __nv_bfloat16*>(feat),
                                     reinterpret_cast<const __nv_bfloat16*>(wp),
                                     reinterpret_cast<__nv_bfloat16*>(dfeat), dw, dbias, npix));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
