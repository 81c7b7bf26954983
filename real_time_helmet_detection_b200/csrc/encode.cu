// Ground-truth encoder and image normalisation on the device - the step BEFORE the hot path.
//
//   hd_encode_targets : padded box list -> CenterNet targets. Replaces the per-image NumPy loop of
//                       transform.py:4-45 (`box2hm`) + :48-70 (`gaussian2D`, `draw_gaussian`) that data.py:108-115 runs
//                       inside collate_fn, and the four per-stack H2D copies of train.py:115-118.
//   hd_normalize_u8   : uint8 HWC images -> normalised fp32 NCHW, replacing TF.to_tensor + Normalize (data.py:118,
//                       utils.py:55-68); the host->device copy shrinks 4x (bytes instead of floats).
//
// The reference draws box after box (scatter, read-modify-write, order dependent). Here every output cell GATHERS over
// its image's boxes (kept in shared memory), which needs no atomics and is deterministic:
//   heat[c][y][x]      = max over boxes j with label c and |x-ix_j|<=r_j, |y-iy_j|<=r_j of exp(-(dx^2+dy^2)/(2 sigma_j^2))
//   mask/offset/size   = values of the LAST box whose centre cell is (y, x)      (later boxes overwrite, as in the loop)
// All box arithmetic is done in fp64 exactly as Python evaluates it (the reference computes in Python floats and
// stores into float32 arrays), without FMA contraction, so offsets/sizes/mask are bit-identical and the Gaussian
// differs from numpy's by at most the last fp64 bit of exp() before the rounding to fp32.
#include <cstdint>
#include <cmath>

#include "hd_b200.h"
#include "hd_common.h"

namespace hd {

constexpr int kEncMaxBoxes = 128;
constexpr int kEncMaxCls = 8;

struct EncBox {
    double two_sigma2;   // (2*sigma)*sigma, sigma = radius/3 (0 for a degenerate box: 0/0 = NaN, as numpy)
    float ox, oy, sx, sy;
    int ix, iy, r, label;  // label < 0: skipped
};

__global__ void __launch_bounds__(256)
encode_targets_kernel(const float* __restrict__ boxes, const int* __restrict__ labels, int nmax, int h, int w, int ncls,
                      int scale_factor, int normalized, float* __restrict__ heat, float* __restrict__ off,
                      float* __restrict__ size, float* __restrict__ mask, int* __restrict__ err) {
    pdl_prologue();
    __shared__ EncBox s_box[kEncMaxBoxes];
    const int b = blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = cell < h * w;
    const int y = live ? cell / w : 0, x = live ? cell - y * w : 0;
    float m[kEncMaxCls];
    bool is_nan[kEncMaxCls];
#pragma unroll
    for (int c = 0; c < kEncMaxCls; ++c) { m[c] = 0.f; is_nan[c] = false; }
    float mk = 0.f, ox = 0.f, oy = 0.f, sx = 0.f, sy = 0.f;
    // any number of boxes per image (the reference's box2hm has no cap): chunks of kEncMaxBoxes through shared memory,
    // in list order, so "the last box at a cell wins" (transform.py:25-37) is preserved across chunks
    for (int j0 = 0; j0 < nmax; j0 += kEncMaxBoxes) {
    const int nchunk = min(kEncMaxBoxes, nmax - j0);
    for (int jj = threadIdx.x; jj < nchunk; jj += blockDim.x) {
        const int j = j0 + jj;
        EncBox e;
        e.label = labels[static_cast<size_t>(b) * nmax + j];
        const float* bp = boxes + (static_cast<size_t>(b) * nmax + j) * 4;
        const double sf = static_cast<double>(scale_factor);
        const double x0 = static_cast<double>(bp[0]) / sf, y0 = static_cast<double>(bp[1]) / sf;
        const double x1 = static_cast<double>(bp[2]) / sf, y1 = static_cast<double>(bp[3]) / sf;
        const double cx = __dadd_rn(x1, x0) / 2.0, cy = __dadd_rn(y1, y0) / 2.0;     // transform.py:21
        e.ix = static_cast<int>(cx); e.iy = static_cast<int>(cy);                    // int(): truncation, :24
        double ox = __dsub_rn(cx, static_cast<double>(e.ix)), oy = __dsub_rn(cy, static_cast<double>(e.iy));
        double sx = __dsub_rn(x1, x0), sy = __dsub_rn(y1, y0);
        if (normalized) {                                                            // :33-35
            ox /= sf; oy /= sf;
            sx /= static_cast<double>(w); sy /= static_cast<double>(h);
        }
        e.ox = static_cast<float>(ox); e.oy = static_cast<float>(oy);
        e.sx = static_cast<float>(sx); e.sy = static_cast<float>(sy);
        const double ax = __dsub_rn(cx, x0), ay = __dsub_rn(cy, y0);
        const double radius = sqrt(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));  // :42
        e.r = static_cast<int>(radius);
        const double sigma = radius / 3.0;
        e.two_sigma2 = __dmul_rn(__dmul_rn(2.0, sigma), sigma);                      // :52 `2 * sigma * sigma`
        // a centre outside the map is an IndexError (or a silent wrap-around for negative indices) in the reference:
        // such a box is skipped here and counted in *err
        const bool inside = e.ix < w && e.iy < h && e.ix >= 0 && e.iy >= 0 && !(isnan(cx) || isnan(cy));
        if (e.label >= 0 && (!inside || e.label >= ncls)) {
            if (err && blockIdx.x == 0) atomicAdd(err, 1);    // every block of the image re-derives the boxes
            e.label = -1;
        }
        s_box[jj] = e;
    }
    __syncthreads();
    for (int j = 0; live && j < nchunk; ++j) {
        const EncBox& e = s_box[j];
        if (e.label < 0) continue;
        const int dx = x - e.ix, dy = y - e.iy;
        if (dx == 0 && dy == 0) { mk = 1.f; ox = e.ox; oy = e.oy; sx = e.sx; sy = e.sy; }
        if (dx < -e.r || dx > e.r || dy < -e.r || dy > e.r) continue;
        const double g64 = exp(static_cast<double>(-(dx * dx + dy * dy)) / e.two_sigma2);   // :52
        const float g = static_cast<float>(g64);
#pragma unroll
        for (int c = 0; c < kEncMaxCls; ++c) {
            if (c == e.label) {
                if (isnan(g)) is_nan[c] = true;     // np.maximum propagates NaN (degenerate zero-size box)
                m[c] = fmaxf(m[c], g);
            }
        }
    }
    __syncthreads();
    }   // box chunks
    if (!live) return;
    const size_t hw = static_cast<size_t>(h) * w;
#pragma unroll
    for (int c = 0; c < kEncMaxCls; ++c)
        if (c < ncls) heat[(static_cast<size_t>(b) * ncls + c) * hw + cell] = is_nan[c] ? nanf("") : m[c];
    off[(static_cast<size_t>(b) * 2 + 0) * hw + cell] = ox;
    off[(static_cast<size_t>(b) * 2 + 1) * hw + cell] = oy;
    size[(static_cast<size_t>(b) * 2 + 0) * hw + cell] = sx;
    size[(static_cast<size_t>(b) * 2 + 1) * hw + cell] = sy;
    mask[static_cast<size_t>(b) * hw + cell] = mk;
}

// 4 pixels (12 bytes) per thread: three 32-bit loads, three float4 stores (one per colour plane).
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint32_t* __restrict__ img, float* __restrict__ out, long long quads_per_img, float m0, float m1,
                    float m2, float s0, float s1, float s2) {
    pdl_prologue();
    const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= quads_per_img) return;
    const int b = blockIdx.y;
    const uint32_t* p = img + (static_cast<size_t>(b) * quads_per_img + q) * 3;
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    uint8_t by[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        by[i] = (w0 >> (8 * i)) & 0xffu; by[4 + i] = (w1 >> (8 * i)) & 0xffu; by[8 + i] = (w2 >> (8 * i)) & 0xffu;
    }
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    const size_t plane = static_cast<size_t>(quads_per_img) * 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float4 v;
        float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i)   // to_tensor: float(u8) / 255 ; Normalize: (x - mean) / std, all fp32, IEEE division
            vv[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(by[3 * i + c]), 255.f), mean[c]), stdv[c]);
        *reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * 3 + c) * plane + q * 4) = v;
    }
}

}  // namespace hd

extern "C" int hd_encode_targets(const float* boxes, const int* labels, int B, int nmax, int h, int w, int num_cls,
                                 int scale_factor, int normalized, float* heat, float* offset, float* size, float* mask,
                                 int* err_count, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(B > 0 && h > 0 && w > 0, "encode_targets: empty output");
    HD_REQUIRE(nmax >= 0, "encode_targets: nmax=%d", nmax);
    HD_REQUIRE(num_cls >= 1 && num_cls <= kEncMaxCls, "encode_targets: num_cls=%d unsupported (max %d)", num_cls, kEncMaxCls);
    HD_REQUIRE(scale_factor >= 1, "encode_targets: scale_factor=%d", scale_factor);
    HD_REQUIRE(heat && offset && size && mask && (nmax == 0 || (boxes && labels)), "encode_targets: null pointer");
    dim3 grid((h * w + 255) / 256, B);
    HD_CHECK_CUDA(::hd::launch_k(encode_targets_kernel, grid, 256, 0, stream, boxes, labels, nmax, h, w, num_cls,
                                 scale_factor, normalized, heat, offset, size, mask, err_count));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_normalize_u8(const void* img_nhwc_u8, float* out_nchw, int B, int H, int W, const float* mean3,
                               const float* std3, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(B > 0 && H > 0 && W > 0, "normalize_u8: empty image");
    HD_REQUIRE((static_cast<long long>(H) * W) % 4 == 0, "normalize_u8: H*W=%lld must be a multiple of 4",
               static_cast<long long>(H) * W);
    HD_REQUIRE(img_nhwc_u8 && out_nchw && mean3 && std3, "normalize_u8: null pointer");
    HD_REQUIRE((reinterpret_cast<uintptr_t>(img_nhwc_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0,
               "normalize_u8: pointers must be 4 / 16 byte aligned");
    const long long quads = static_cast<long long>(H) * W / 4;
    dim3 grid(static_cast<unsigned>((quads + 255) / 256), B);
    HD_CHECK_CUDA(::hd::launch_k(normalize_u8_kernel, grid, 256, 0, stream,
                                 reinterpret_cast<const uint32_t*>(img_nhwc_u8), out_nchw, quads, mean3[0], mean3[1],
                                 mean3[2], std3[0], std3[1], std3[2]));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
