// Stem support kernels. The 7x7 stride-2 stem convolution (hourglass.py:163, 3 -> 64 channels, bias + BN + ReLU)
// has K = 147, too ragged for a direct TMA/UMMA mapping, so the image is expanded once into a
// K-padded patch matrix  P[n, oy, ox, k],  k = (ky*7 + kx)*3 + c  (k < 147, zero up to 192)
// in bf16 NHWC; the stem is then the 1x1 case of the tcgen05 implicit-GEMM kernel (cin = 192) and its weight
// gradient the 1x1 case of the wgrad kernel. The patch matrix is kept for the backward pass.
#include <cuda_bf16.h>

#include "hd_common.h"

namespace hd {

constexpr int kStemK = 147, kStemKPad = 192;

__global__ void stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H,
                                   int W) {
    const int Ho = H >> 1, Wo = W >> 1;
    const size_t nvec = static_cast<size_t>(N) * Ho * Wo * (kStemKPad / 8);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int kv = i % (kStemKPad / 8);
        size_t pix = i / (kStemKPad / 8);
        const int ox = pix % Wo;
        const int oy = (pix / Wo) % Ho;
        const int n = pix / (static_cast<size_t>(Wo) * Ho);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kv * 8 + j;
            float val = 0.f;
            if (k < kStemK) {
                const int c = k % 3, kk = k / 3;
                const int ky = kk / 7, kx = kk - ky * 7;
                const int iy = 2 * oy + ky - 3, ix = 2 * ox + kx - 3;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                    val = __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + iy) * W + ix);
            }
            v[j] = val;
        }
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + i * 8) = u;
    }
}

// w: [64][3][7][7] fp32 -> out: [1][64][192] bf16 in the im2col K order
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout) {
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
    const size_t nvec = static_cast<size_t>(N) * (H / 2) * (W / 2) * (kStemKPad / 8);
    size_t g = (nvec + 255) / 256;
    const size_t cap = static_cast<size_t>(sm_count()) * 16;
    if (g > cap) g = cap;
    stem_im2col_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(patches), N,
                                                                   H, W);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_stem_pack_weight(const float* w, void* out, int cout, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(cout > 0 && cout <= 64, "stem_pack_weight: cout=%d", cout);
    const int total = cout * kStemKPad;
    stem_pack_weight_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), cout);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
