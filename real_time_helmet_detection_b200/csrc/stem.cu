// Stem support kernels. The 7x7 stride-2 stem convolution (hourglass.py:163, 3 -> 64 channels, bias + BN + ReLU)
// has K = 147, too ragged for a direct TMA/UMMA mapping, so the image is expanded once into a
// K-padded patch matrix  P[n, oy, ox, k],  k = (ky*7 + kx)*3 + c  (k < 147, zero up to 192)
// in bf16 NHWC; the stem is then the 1x1 case of the tcgen05 implicit-GEMM kernel (cin = 192) and its weight
// gradient the 1x1 case of the wgrad kernel. The patch matrix is kept for the backward pass.
#include <cuda_bf16.h>

#include "hd_common.h"

namespace hd {

constexpr int kStemK = 147, kStemKPad = 192;

// One CTA = a TOH x TOW tile of output pixels. The (2*TOH+5) x (2*TOW+5) x 3 input window is staged once in shared
// memory (zero-filled outside the image = the conv's padding), then every thread emits 16-byte vectors of the patch
// matrix: consecutive threads write consecutive 8-k vectors of one pixel, so the 384-byte rows are written coalesced.
constexpr int kTOH = 8, kTOW = 32;
constexpr int kPH = 2 * kTOH + 5, kPW = 2 * kTOW + 5;

__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int N, int H, int W) {
    pdl_prologue();
    __shared__ float patch[3][kPH][kPW + 1];
    __shared__ short koff[kStemKPad];      // k -> offset inside the patch of the pixel at (0,0); -1 for the zero padding
    const int Ho = H >> 1, Wo = W >> 1;
    const int tiles_x = (Wo + kTOW - 1) / kTOW, tiles_y = (Ho + kTOH - 1) / kTOH;
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int oy0 = ty * kTOH, ox0 = tx * kTOW;
    const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
    for (int k = threadIdx.x; k < kStemKPad; k += blockDim.x) {
        short o = -1;
        if (k < kStemK) {
            const int c = k % 3, kk = k / 3, ky = kk / 7, kx = kk - ky * 7;
            o = static_cast<short>((c * kPH + ky) * (kPW + 1) + kx);
        }
        koff[k] = o;
    }
    for (int i = threadIdx.x; i < 3 * kPH * kPW; i += blockDim.x) {
        const int px = i % kPW, py = (i / kPW) % kPH, c = i / (kPW * kPH);
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + iy) * W + ix);
        patch[c][py][px] = v;
    }
    __syncthreads();
    const float* pbase = &patch[0][0][0];
    constexpr int kVec = kStemKPad / 8;   // 24 vectors per pixel
    // thread -> one FIXED 8-k vector (its 8 patch offsets live in registers) of 10 pixels per pass; 240 of the 256
    // threads work here. Consecutive threads still write consecutive 16-byte vectors of a pixel's 384-byte row.
    constexpr int kPixPerPass = 256 / kVec;   // 10
    const int kv = threadIdx.x % kVec, slot = threadIdx.x / kVec;
    if (slot < kPixPerPass) {
        int off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) off[j] = koff[kv * 8 + j];
        for (int lp = slot; lp < kTOH * kTOW; lp += kPixPerPass) {
            const int lx = lp % kTOW, ly = lp / kTOW;
            const int oy = oy0 + ly, ox = ox0 + lx;
            if (oy >= Ho || ox >= Wo) continue;
            const int pofs = (2 * ly) * (kPW + 1) + 2 * lx;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? pbase[off[j] + pofs] : 0.f;
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * Ho + oy) * Wo + ox) * kStemKPad + kv * 8) = u;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Space-to-depth formulation of the same stem (the default): a 7x7 stride-2 pad-3 convolution over (3, H, W) is a
// 4x4 stride-1 convolution over the space-to-depth image S[c, sy, sx](Y, X) = img[c, 2Y+sy, 2X+sx] (12 channels):
//     out[oy, ox] = sum_{dy,dx in 0..3} sum_{c,sy,sx} w[c, 2dy+sy-1, 2dx+sx-1] * S[c,sy,sx](oy+dy-2, ox+dx-2)
// (taps with ky or kx = -1 have weight 0). Only the HORIZONTAL taps are unfolded into channels,
//     U[n, Y, X, dx*12 + (c*2+sy)*2 + sx] = S[c,sy,sx](Y, X+dx-2),   48 of 64 channels used,
// so U is 128 B per pixel = 268 MB at B=32 instead of the 805 MB im2col patch matrix, and the four VERTICAL taps are
// row-shifted TMA boxes of the implicit-GEMM kernel (hd_conv2d_igemm_vtaps: K = 4 x 64) and of the wgrad kernel.
constexpr int kUTOH = 4, kUTOW = 64;
constexpr int kUPH = 2 * kUTOH, kUPW = 2 * kUTOW + 6;

__global__ void __launch_bounds__(256) stem_unfold_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int N, int H, int W) {
    pdl_prologue();
    __shared__ float patch[3][kUPH][kUPW + 1];
    const int Ho = H >> 1, Wo = W >> 1;
    const int tiles_x = (Wo + kUTOW - 1) / kUTOW, tiles_y = (Ho + kUTOH - 1) / kUTOH;
    const int tile = blockIdx.x;
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int oy0 = ty * kUTOH, ox0 = tx * kUTOW;
    const int iy0 = 2 * oy0, ix0 = 2 * (ox0 - 2);
    for (int i = threadIdx.x; i < 3 * kUPH * kUPW; i += blockDim.x) {
        const int px = i % kUPW, py = (i / kUPW) % kUPH, c = i / (kUPW * kUPH);
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if (iy < H && ix >= 0 && ix < W) v = __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + iy) * W + ix);
        patch[c][py][px] = v;
    }
    __syncthreads();
    const float* pbase = &patch[0][0][0];
    // thread -> one fixed 8-k vector of 32 pixels per pass; vectors 6 and 7 (k >= 48) are the zero padding
    const int kv = threadIdx.x & 7, slot = threadIdx.x >> 3;
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kv * 8 + j;
        const int dx = k / 12, r = k % 12, c = r >> 2, sy = (r >> 1) & 1, sx = r & 1;
        off[j] = k < 48 ? (c * kUPH + sy) * (kUPW + 1) + 2 * dx + sx : -1;
    }
    for (int lp = slot; lp < kUTOH * kUTOW; lp += 32) {
        const int lx = lp % kUTOW, ly = lp / kUTOW;
        const int oy = oy0 + ly, ox = ox0 + lx;
        if (oy >= Ho || ox >= Wo) continue;
        const int pofs = (2 * ly) * (kUPW + 1) + 2 * lx;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? pbase[off[j] + pofs] : 0.f;
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * Ho + oy) * Wo + ox) * 64 + kv * 8) = u;
    }
}

// w: [64][3][7][7] fp32 -> out: [1][64][192] bf16 in the im2col K order
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cout * kStemKPad) return;
    const int k = idx % kStemKPad, co = idx / kStemKPad;
    float v = 0.f;
    if (k < kStemK) {
        const int c = k % 3, kk = k / 3;
        v = w[(static_cast<size_t>(co) * 3 + c) * 49 + kk];
    }
    out[idx] = __float2bfloat16(v);
}

}  // namespace hd

extern "C" int hd_stem_im2col(const float* x, void* patches, int N, int H, int W, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "stem_im2col: shape (%d,3,%d,%d)", N, H, W);
    const int Ho = H / 2, Wo = W / 2;
    const long long tiles = static_cast<long long>(N) * ((Ho + kTOH - 1) / kTOH) * ((Wo + kTOW - 1) / kTOW);
    HD_REQUIRE(tiles < (1ll << 31), "stem_im2col: too many tiles");
    HD_CHECK_CUDA(::hd::launch_k(stem_im2col_kernel, static_cast<unsigned>(tiles), 256, 0, stream, x,
                                 reinterpret_cast<__nv_bfloat16*>(patches), N, H, W));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_stem_pack_weight(const float* w, void* out, int cout, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(cout > 0 && cout <= 64, "stem_pack_weight: cout=%d", cout);
    const int total = cout * kStemKPad;
    HD_CHECK_CUDA(::hd::launch_k(stem_pack_weight_kernel, (total + 255) / 256, 256, 0, stream, w,
                                 reinterpret_cast<__nv_bfloat16*>(out), cout));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_stem_unfold(const float* x, void* unfolded, int N, int H, int W, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "stem_unfold: shape (%d,3,%d,%d)", N, H, W);
    const int Ho = H / 2, Wo = W / 2;
    const long long tiles = static_cast<long long>(N) * ((Ho + kUTOH - 1) / kUTOH) * ((Wo + kUTOW - 1) / kUTOW);
    HD_REQUIRE(tiles < (1ll << 31), "stem_unfold: too many tiles");
    HD_CHECK_CUDA(::hd::launch_k(stem_unfold_kernel, static_cast<unsigned>(tiles), 256, 0, stream, x,
                                 reinterpret_cast<__nv_bfloat16*>(unfolded), N, H, W));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
