// Layout conversion kernels: OIHW fp32 master weights -> packed bf16 UMMA operand layouts, and
// NCHW fp32 <-> NHWC bf16 activations (boundary of the nn.Module call signature, hourglass.py:223).
#include <cuda_bf16.h>

#include "hd_common.h"

namespace hd {

// mode 0 (forward): out[tap][r = co][k = ci] = w[co][ci][tap]
// mode 1 (dgrad)  : out[tap][r = ci][k = co] = w[co][ci][taps-1-tap]   (180-degree rotated taps, transposed)
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                   int taps, int rows_pad, int k_pad, int mode) {
    pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = taps * rows_pad * k_pad;
    if (idx >= total) return;
    const int k = idx % k_pad;
    const int r = (idx / k_pad) % rows_pad;
    const int tap = idx / (k_pad * rows_pad);
    float v = 0.f;
    if (mode == 0) {
        if (r < cout && k < cin) v = w[(static_cast<size_t>(r) * cin + k) * taps + tap];
    } else {
        if (r < cin && k < cout) v = w[(static_cast<size_t>(k) * cin + r) * taps + (taps - 1 - tap)];
    }
    out[idx] = __float2bfloat16(v);
}

// x: [N, C, H, W] fp32 -> y: [N, H, W, c_pad] bf16 (channels >= C zero-filled)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H,
                                    int W, int c_pad) {
    pdl_prologue();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * H * W * c_pad;
    if (idx >= total) return;
    const int c = idx % c_pad;
    const size_t pix = idx / c_pad;
    const int wv = pix % W;
    const int h = (pix / W) % H;
    const int n = pix / (static_cast<size_t>(W) * H);
    float v = 0.f;
    if (c < C) v = x[((static_cast<size_t>(n) * C + c) * H + h) * W + wv];
    y[idx] = __float2bfloat16(v);
}

// x: [N, H, W, c_stride] bf16 -> y: [N, C, H, W] fp32
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int N, int C, int H,
                                    int W, int c_stride) {
    pdl_prologue();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * C * H * W;
    if (idx >= total) return;
    const int wv = idx % W;
    const int h = (idx / W) % H;
    const int c = (idx / (static_cast<size_t>(W) * H)) % C;
    const int n = idx / (static_cast<size_t>(W) * H * C);
    y[idx] = __bfloat162float(x[((static_cast<size_t>(n) * H + h) * W + wv) * c_stride + c]);
}

// All conv weights of a network in ONE launch: `jobs` is a device table, `starts[j]` the first flat output element of
// job j (ascending), `total` the number of output elements over all jobs.
struct PackJob {
    const float* w;
    __nv_bfloat16* out;
    int cout, cin, taps, rows_pad, k_pad, mode;   // mode 0 forward, 1 dgrad, 2 stem im2col (K = (ky*7+kx)*3+c, padded), 3 stem space-to-depth
    long long start;
};

__global__ void pack_all_kernel(const PackJob* __restrict__ jobs, int njobs, long long total) {
    pdl_prologue();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int lo = 0, hi = njobs - 1;          // last job with start <= idx
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].start <= idx) lo = mid; else hi = mid - 1;
    }
    const PackJob j = jobs[lo];
    const int e = static_cast<int>(idx - j.start);
    const int k = e % j.k_pad;
    const int r = (e / j.k_pad) % j.rows_pad;
    const int tap = e / (j.k_pad * j.rows_pad);
    float v = 0.f;
    if (j.mode == 0) {
        if (r < j.cout && k < j.cin) v = j.w[(static_cast<size_t>(r) * j.cin + k) * j.taps + tap];
    } else if (j.mode == 1) {
        if (r < j.cin && k < j.cout) v = j.w[(static_cast<size_t>(k) * j.cin + r) * j.taps + (j.taps - 1 - tap)];
    } else if (j.mode == 2) {
        if (r < j.cout && k < 147) v = j.w[(static_cast<size_t>(r) * 3 + (k % 3)) * 49 + (k / 3)];
    } else {        // mode 3: space-to-depth stem (stem.cu): tap = dy, k = dx*12 + (c*2+sy)*2 + sx -> w[r][c][2dy+sy-1][2dx+sx-1]
        if (r < j.cout && k < 48) {
            const int dx = k / 12, q = k % 12, c = q >> 2, sy = (q >> 1) & 1, sx = q & 1;
            const int ky = 2 * tap + sy - 1, kx = 2 * dx + sx - 1;
            if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = j.w[(static_cast<size_t>(r) * 3 + c) * 49 + ky * 7 + kx];
        }
    }
    j.out[e] = __float2bfloat16(v);
}

}  // namespace hd

// jobs: device array of hd::PackJob (layout documented in csrc/pack.cu); used by the network executor.
extern "C" int hd_pack_all_weights(const void* jobs, int njobs, long long total, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(njobs > 0 && total > 0, "pack_all_weights: empty job table");
    const long long blocks = (total + 255) / 256;
    HD_REQUIRE(blocks < (1ll << 31), "pack_all_weights: too many elements");
    HD_CHECK_CUDA(::hd::launch_k(pack_all_kernel, static_cast<unsigned>(blocks), 256, 0, stream,
                                 reinterpret_cast<const PackJob*>(jobs), njobs, total));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_pack_conv_weight(const float* w_oihw, void* out, int cout, int cin, int ksize, int rows_pad,
                                   int k_pad, int mode, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(mode == 0 || mode == 1, "pack_conv_weight: mode=%d", mode);
    const int rows = mode == 0 ? cout : cin, kdim = mode == 0 ? cin : cout;
    HD_REQUIRE(rows_pad >= rows && k_pad >= kdim, "pack_conv_weight: padding smaller than the matrix");
    const int total = ksize * ksize * rows_pad * k_pad;
    HD_CHECK_CUDA(::hd::launch_k(pack_weight_kernel, (total + 255) / 256, 256, 0, stream, w_oihw,
                                 reinterpret_cast<__nv_bfloat16*>(out), cout, cin, ksize * ksize, rows_pad, k_pad,
                                 mode));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int c_pad,
                                        cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(c_pad >= C, "nchw_to_nhwc: c_pad < C");
    const size_t total = static_cast<size_t>(N) * H * W * c_pad;
    if (total == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(nchw_to_nhwc_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, x,
                                 reinterpret_cast<__nv_bfloat16*>(y), N, C, H, W, c_pad));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_nhwc_bf16_to_nchw_f32(const void* x, float* y, int N, int C, int H, int W, int c_stride,
                                        cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(c_stride >= C, "nhwc_to_nchw: c_stride < C");
    const size_t total = static_cast<size_t>(N) * C * H * W;
    if (total == 0) return HD_OK;
    HD_CHECK_CUDA(::hd::launch_k(nhwc_to_nchw_kernel, (unsigned)((total + 255) / 256), 256, 0, stream,
                                 reinterpret_cast<const __nv_bfloat16*>(x), y, N, C, H, W, c_stride));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
