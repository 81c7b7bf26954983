// Fused CenterNet training losses (reference: loss.py:42-69 NormedL1Loss / FocalLoss, loss.py:18-32
// LossCalculator.forward; caller-side sigmoid train.py:107-111).
//
// One forward kernel keeps every per-pixel intermediate in registers and reduces five sums
//   S_mask, S_pos, S_neg, S_off, S_size
// with warp shuffles + one atomic per block; a 1-thread finalize turns them into
//   hm = -(S_pos + S_neg) / (B * clamp(S_mask, 1)),  off = S_off / (B * np),  size = S_size / (B * np),
//   total = w_hm*hm + w_off*off + w_size*size          (per-sample sum -> batch mean -> / num_pos, loss.py:48-50,66-69)
// One backward kernel re-reads the same inputs and writes d(total)/d(input) * grad_out.
// `from_logits` = 1 fuses the head activation: heat-map channels (and offset/size with normalized_coord) are
// raw logits, sigmoid is applied in registers and the chain rule p(1-p) folded into the gradient.
// HBM-bound: algorithmic bytes per pixel = (C+4 preds + C+5 targets) * 4 B read (+ (C+4)*4 B written in backward).
#include "hd_common.h"

namespace hd {

struct LossArgs {
    const float* hm;   long long hm_bs;    // (B, C, H, W) view: batch stride, channel stride = HW
    const float* off;  long long off_bs;   // (B, 2, H, W)
    const float* size; long long size_bs;  // (B, 2, H, W)
    const float* ghm;  const float* goff; const float* gsize; const float* gmask;  // contiguous fp32
    int B, C, HW;
    float alpha, beta, eps;
    int from_logits;        // sigmoid on the heat-map inside the kernel
    int sigmoid_reg;        // sigmoid on offset / size inside the kernel (normalized_coord with from_logits)
};

__device__ __forceinline__ float pow_ab(float x, float e) {
    if (e == 2.f) return x * x;
    if (e == 4.f) { float t = x * x; return t * t; }
    if (e == 1.f) return x;
    return powf(x, e);
}
// d/dx x^e
__device__ __forceinline__ float dpow_ab(float x, float e) {
    if (e == 2.f) return 2.f * x;
    if (e == 4.f) return 4.f * x * x * x;
    if (e == 1.f) return 1.f;
    return e * powf(x, e - 1.f);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float block_sum(float v, float* scratch) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = l < (blockDim.x >> 5) ? scratch[l] : 0.f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

__global__ void loss_fwd_kernel(const LossArgs a, float* __restrict__ sums) {
    pdl_prologue();
    __shared__ float scratch[32];
    float s_mask = 0.f, s_pos = 0.f, s_neg = 0.f, s_off = 0.f, s_size = 0.f;
    const long long total = static_cast<long long>(a.B) * a.HW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / a.HW);
        const int p = static_cast<int>(i - static_cast<long long>(b) * a.HW);
        const float m = a.gmask[i];
        s_mask += m;
        for (int c = 0; c < a.C; ++c) {
            float pr = a.hm[b * a.hm_bs + static_cast<long long>(c) * a.HW + p];
            if (a.from_logits) pr = sigmoidf_(pr);
            const float gt = a.ghm[(static_cast<long long>(b) * a.C + c) * a.HW + p];
            s_pos += logf(pr + a.eps) * pow_ab(1.f - pr, a.alpha) * m;
            s_neg += logf(1.f - pr + a.eps) * pow_ab(pr, a.alpha) * pow_ab(1.f - gt, a.beta) * (1.f - m);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float po = a.off[b * a.off_bs + static_cast<long long>(c) * a.HW + p];
            float ps = a.size[b * a.size_bs + static_cast<long long>(c) * a.HW + p];
            if (a.sigmoid_reg) { po = sigmoidf_(po); ps = sigmoidf_(ps); }
            const float go = a.goff[(static_cast<long long>(b) * 2 + c) * a.HW + p];
            const float gs = a.gsize[(static_cast<long long>(b) * 2 + c) * a.HW + p];
            s_off += fabsf(po * m - go * m);
            s_size += fabsf(ps * m - gs * m);
        }
    }
    float r;
    r = block_sum(s_mask, scratch); if (threadIdx.x == 0) atomicAdd(sums + 0, r);
    r = block_sum(s_pos, scratch);  if (threadIdx.x == 0) atomicAdd(sums + 1, r);
    r = block_sum(s_neg, scratch);  if (threadIdx.x == 0) atomicAdd(sums + 2, r);
    r = block_sum(s_off, scratch);  if (threadIdx.x == 0) atomicAdd(sums + 3, r);
    r = block_sum(s_size, scratch); if (threadIdx.x == 0) atomicAdd(sums + 4, r);
}

// out: [hm, offset, size, total, inv_norm]
__global__ void loss_finalize_kernel(const float* __restrict__ sums, float* __restrict__ out, float B, float w_hm,
                                     float w_off, float w_size) {
    pdl_prologue();
    const float np = fminf(fmaxf(sums[0], 1.f), 1e30f);
    const float inv = 1.f / np;
    const float hm = -((sums[1] / B) + (sums[2] / B)) * inv;
    const float off = (sums[3] / B) * inv;
    const float sz = (sums[4] / B) * inv;
    out[0] = hm;
    out[1] = off;
    out[2] = sz;
    out[3] = hm * w_hm + off * w_off + sz * w_size;
    out[4] = inv / B;
}

struct LossBwdArgs {
    float* d_hm;   long long d_hm_bs;
    float* d_off;  long long d_off_bs;
    float* d_size; long long d_size_bs;
    const float* fwd_out;     // [.., inv_norm] from the forward finalize
    const float* grad_out;    // scalar upstream gradient (device), may be null (== 1)
    float w_hm, w_off, w_size;
};

__global__ void loss_bwd_kernel(const LossArgs a, const LossBwdArgs g) {
    pdl_prologue();
    const float up = g.grad_out ? *g.grad_out : 1.f;
    const float inv = g.fwd_out[4] * up;
    const long long total = static_cast<long long>(a.B) * a.HW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / a.HW);
        const int p = static_cast<int>(i - static_cast<long long>(b) * a.HW);
        const float m = a.gmask[i];
        for (int c = 0; c < a.C; ++c) {
            float pr = a.hm[b * a.hm_bs + static_cast<long long>(c) * a.HW + p];
            if (a.from_logits) pr = sigmoidf_(pr);
            const float gt = a.ghm[(static_cast<long long>(b) * a.C + c) * a.HW + p];
            const float q = 1.f - pr;
            // d/dp [ log(p+eps) (1-p)^alpha ] and d/dp [ log(1-p+eps) p^alpha ]
            const float dpos = pow_ab(q, a.alpha) / (pr + a.eps) - dpow_ab(q, a.alpha) * logf(pr + a.eps);
            const float dneg = -pow_ab(pr, a.alpha) / (q + a.eps) + dpow_ab(pr, a.alpha) * logf(q + a.eps);
            float d = -(dpos * m + dneg * pow_ab(1.f - gt, a.beta) * (1.f - m)) * inv * g.w_hm;
            if (a.from_logits) d *= pr * q;
            g.d_hm[b * g.d_hm_bs + static_cast<long long>(c) * a.HW + p] = d;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float po = a.off[b * a.off_bs + static_cast<long long>(c) * a.HW + p];
            float ps = a.size[b * a.size_bs + static_cast<long long>(c) * a.HW + p];
            float jo = 1.f, js = 1.f;
            if (a.sigmoid_reg) {
                po = sigmoidf_(po); ps = sigmoidf_(ps);
                jo = po * (1.f - po); js = ps * (1.f - ps);
            }
            const float go = a.goff[(static_cast<long long>(b) * 2 + c) * a.HW + p];
            const float gs = a.gsize[(static_cast<long long>(b) * 2 + c) * a.HW + p];
            const float eo = po * m - go * m, es = ps * m - gs * m;
            const float so = eo > 0.f ? 1.f : (eo < 0.f ? -1.f : 0.f);
            const float ss = es > 0.f ? 1.f : (es < 0.f ? -1.f : 0.f);
            g.d_off[b * g.d_off_bs + static_cast<long long>(c) * a.HW + p] = so * m * inv * g.w_off * jo;
            g.d_size[b * g.d_size_bs + static_cast<long long>(c) * a.HW + p] = ss * m * inv * g.w_size * js;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Vectorised variants (H*W % 4 == 0, 16-byte aligned planes - every call of the training path): a thread owns 4
// consecutive pixels of one image, every tensor plane is read (and every gradient plane written) with one 16-byte access,
// index arithmetic is 32-bit, and the five block reductions share one shared-memory round. The scalar kernels above
// spent ~460 instructions per pixel (64-bit divisions, 13 scalar loads, five separate reductions) and ran at 0.2 of the
// HBM rate; these are bound by the transcendental math of the focal term (2 logf + 1 expf per heat-map element).
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int C>
__global__ void __launch_bounds__(256)
loss_fwd4_kernel(const LossArgs a, float* __restrict__ sums) {
    pdl_prologue();
    __shared__ float scratch[5][8];
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};                 // mask, pos, neg, off, size
    const int q_per_img = a.HW >> 2;
    const int total = a.B * q_per_img;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / q_per_img;
        const int p = (i - b * q_per_img) << 2;
        const float4 m4 = ld4(a.gmask + static_cast<size_t>(b) * a.HW + p);
        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
        s[0] += (m[0] + m[1]) + (m[2] + m[3]);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float4 pr4 = ld4(a.hm + b * a.hm_bs + static_cast<long long>(c) * a.HW + p);
            const float4 gt4 = ld4(a.ghm + (static_cast<size_t>(b) * C + c) * a.HW + p);
            const float prv[4] = {pr4.x, pr4.y, pr4.z, pr4.w}, gtv[4] = {gt4.x, gt4.y, gt4.z, gt4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float pr = prv[j];
                if (a.from_logits) pr = sigmoidf_(pr);
                s[1] += logf(pr + a.eps) * pow_ab(1.f - pr, a.alpha) * m[j];
                s[2] += logf(1.f - pr + a.eps) * pow_ab(pr, a.alpha) * pow_ab(1.f - gtv[j], a.beta) * (1.f - m[j]);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float4 po4 = ld4(a.off + b * a.off_bs + static_cast<long long>(c) * a.HW + p);
            const float4 ps4 = ld4(a.size + b * a.size_bs + static_cast<long long>(c) * a.HW + p);
            const float4 go4 = ld4(a.goff + (static_cast<size_t>(b) * 2 + c) * a.HW + p);
            const float4 gs4 = ld4(a.gsize + (static_cast<size_t>(b) * 2 + c) * a.HW + p);
            const float pov[4] = {po4.x, po4.y, po4.z, po4.w}, psv[4] = {ps4.x, ps4.y, ps4.z, ps4.w};
            const float gov[4] = {go4.x, go4.y, go4.z, go4.w}, gsv[4] = {gs4.x, gs4.y, gs4.z, gs4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float po = pov[j], ps = psv[j];
                if (a.sigmoid_reg) { po = sigmoidf_(po); ps = sigmoidf_(ps); }
                s[3] += fabsf(po * m[j] - gov[j] * m[j]);
                s[4] += fabsf(ps * m[j] - gsv[j] * m[j]);
            }
        }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (l == 0) scratch[k][w] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += scratch[threadIdx.x][i];
        atomicAdd(sums + threadIdx.x, r);
    }
}

template <int C>
__global__ void __launch_bounds__(256)
loss_bwd4_kernel(const LossArgs a, const LossBwdArgs g) {
    pdl_prologue();
    const float up = g.grad_out ? *g.grad_out : 1.f;
    const float inv = g.fwd_out[4] * up;
    const int q_per_img = a.HW >> 2;
    const int total = a.B * q_per_img;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / q_per_img;
        const int p = (i - b * q_per_img) << 2;
        const float4 m4 = ld4(a.gmask + static_cast<size_t>(b) * a.HW + p);
        const float m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float4 pr4 = ld4(a.hm + b * a.hm_bs + static_cast<long long>(c) * a.HW + p);
            const float4 gt4 = ld4(a.ghm + (static_cast<size_t>(b) * C + c) * a.HW + p);
            const float prv[4] = {pr4.x, pr4.y, pr4.z, pr4.w}, gtv[4] = {gt4.x, gt4.y, gt4.z, gt4.w};
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float pr = prv[j];
                if (a.from_logits) pr = sigmoidf_(pr);
                const float q = 1.f - pr;
                const float dpos = pow_ab(q, a.alpha) / (pr + a.eps) - dpow_ab(q, a.alpha) * logf(pr + a.eps);
                const float dneg = -pow_ab(pr, a.alpha) / (q + a.eps) + dpow_ab(pr, a.alpha) * logf(q + a.eps);
                d[j] = -(dpos * m[j] + dneg * pow_ab(1.f - gtv[j], a.beta) * (1.f - m[j])) * inv * g.w_hm;
                if (a.from_logits) d[j] *= pr * q;
            }
            *reinterpret_cast<float4*>(g.d_hm + b * g.d_hm_bs + static_cast<long long>(c) * a.HW + p) =
                make_float4(d[0], d[1], d[2], d[3]);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float4 po4 = ld4(a.off + b * a.off_bs + static_cast<long long>(c) * a.HW + p);
            const float4 ps4 = ld4(a.size + b * a.size_bs + static_cast<long long>(c) * a.HW + p);
            const float4 go4 = ld4(a.goff + (static_cast<size_t>(b) * 2 + c) * a.HW + p);
            const float4 gs4 = ld4(a.gsize + (static_cast<size_t>(b) * 2 + c) * a.HW + p);
            const float pov[4] = {po4.x, po4.y, po4.z, po4.w}, psv[4] = {ps4.x, ps4.y, ps4.z, ps4.w};
            const float gov[4] = {go4.x, go4.y, go4.z, go4.w}, gsv[4] = {gs4.x, gs4.y, gs4.z, gs4.w};
            float dof[4], dsz[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float po = pov[j], ps = psv[j];
                float jo = 1.f, js = 1.f;
                if (a.sigmoid_reg) {
                    po = sigmoidf_(po); ps = sigmoidf_(ps);
                    jo = po * (1.f - po); js = ps * (1.f - ps);
                }
                const float eo = po * m[j] - gov[j] * m[j], es = ps * m[j] - gsv[j] * m[j];
                const float so = eo > 0.f ? 1.f : (eo < 0.f ? -1.f : 0.f);
                const float ss = es > 0.f ? 1.f : (es < 0.f ? -1.f : 0.f);
                dof[j] = so * m[j] * inv * g.w_off * jo;
                dsz[j] = ss * m[j] * inv * g.w_size * js;
            }
            *reinterpret_cast<float4*>(g.d_off + b * g.d_off_bs + static_cast<long long>(c) * a.HW + p) =
                make_float4(dof[0], dof[1], dof[2], dof[3]);
            *reinterpret_cast<float4*>(g.d_size + b * g.d_size_bs + static_cast<long long>(c) * a.HW + p) =
                make_float4(dsz[0], dsz[1], dsz[2], dsz[3]);
        }
    }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool vec_ok(const LossArgs& a) {
    return a.C == 2 && a.HW % 4 == 0 && a.hm_bs % 4 == 0 && a.off_bs % 4 == 0 && a.size_bs % 4 == 0 && al16(a.hm) &&
           al16(a.off) && al16(a.size) && al16(a.ghm) && al16(a.goff) && al16(a.gsize) && al16(a.gmask) &&
           static_cast<long long>(a.B) * a.HW < (1ll << 31);
}

}  // namespace hd

using namespace hd;

static int fill_args(LossArgs& a, const float* hm, long long hm_bs, const float* off, long long off_bs,
                     const float* size, long long size_bs, const float* ghm, const float* goff, const float* gsize,
                     const float* gmask, int B, int C, int H, int W, float alpha, float beta, int from_logits,
                     int sigmoid_reg) {
    HD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "loss: empty input (B=%d C=%d H=%d W=%d)", B, C, H, W);
    a.hm = hm; a.hm_bs = hm_bs; a.off = off; a.off_bs = off_bs; a.size = size; a.size_bs = size_bs;
    a.ghm = ghm; a.goff = goff; a.gsize = gsize; a.gmask = gmask;
    a.B = B; a.C = C; a.HW = H * W; a.alpha = alpha; a.beta = beta; a.eps = 1e-7f;
    a.from_logits = from_logits; a.sigmoid_reg = sigmoid_reg;
    return HD_OK;
}

static int loss_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = static_cast<long long>(sm_count()) * 8;
    return static_cast<int>(g < cap ? g : cap);
}

// See include/hd_b200.h. sums: 5 floats of scratch (zeroed here); out: 5 floats [hm, offset, size, total, inv_norm].
extern "C" int hd_loss_forward(const float* hm, long long hm_bs, const float* off, long long off_bs, const float* size,
                               long long size_bs, const float* ghm, const float* goff, const float* gsize,
                               const float* gmask, int B, int C, int H, int W, float alpha, float beta, float w_hm,
                               float w_off, float w_size, int from_logits, int sigmoid_reg, float* sums, float* out,
                               cudaStream_t stream) {
    LossArgs a;
    int rc = fill_args(a, hm, hm_bs, off, off_bs, size, size_bs, ghm, goff, gsize, gmask, B, C, H, W, alpha, beta,
                       from_logits, sigmoid_reg);
    if (rc) return rc;
    HD_CHECK_CUDA(cudaMemsetAsync(sums, 0, 5 * sizeof(float), stream));
    if (vec_ok(a))
        HD_CHECK_CUDA(::hd::launch_k(loss_fwd4_kernel<2>, loss_grid(static_cast<long long>(B) * H * W / 4), 256, 0, stream,
                                     a, sums));
    else
        HD_CHECK_CUDA(::hd::launch_k(loss_fwd_kernel, loss_grid(static_cast<long long>(B) * H * W), 256, 0, stream, a,
                                     sums));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    HD_CHECK_CUDA(::hd::launch_k(loss_finalize_kernel, 1, 1, 0, stream, sums, out, static_cast<float>(B), w_hm, w_off,
                                 w_size));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}

extern "C" int hd_loss_backward(const float* hm, long long hm_bs, const float* off, long long off_bs,
                                const float* size, long long size_bs, const float* ghm, const float* goff,
                                const float* gsize, const float* gmask, int B, int C, int H, int W, float alpha,
                                float beta, float w_hm, float w_off, float w_size, int from_logits, int sigmoid_reg,
                                const float* fwd_out, const float* grad_out, float* d_hm, long long d_hm_bs,
                                float* d_off, long long d_off_bs, float* d_size, long long d_size_bs,
                                cudaStream_t stream) {
    LossArgs a;
    int rc = fill_args(a, hm, hm_bs, off, off_bs, size, size_bs, ghm, goff, gsize, gmask, B, C, H, W, alpha, beta,
                       from_logits, sigmoid_reg);
    if (rc) return rc;
    LossBwdArgs g;
    g.d_hm = d_hm; g.d_hm_bs = d_hm_bs; g.d_off = d_off; g.d_off_bs = d_off_bs; g.d_size = d_size;
    g.d_size_bs = d_size_bs; g.fwd_out = fwd_out; g.grad_out = grad_out;
    g.w_hm = w_hm; g.w_off = w_off; g.w_size = w_size;
    const bool vec = vec_ok(a) && g.d_hm_bs % 4 == 0 && g.d_off_bs % 4 == 0 && g.d_size_bs % 4 == 0 && al16(g.d_hm) &&
                     al16(g.d_off) && al16(g.d_size);
    if (vec)
        HD_CHECK_CUDA(::hd::launch_k(loss_bwd4_kernel<2>, loss_grid(static_cast<long long>(B) * H * W / 4), 256, 0, stream,
                                     a, g));
    else
        HD_CHECK_CUDA(::hd::launch_k(loss_bwd_kernel, loss_grid(static_cast<long long>(B) * H * W), 256, 0, stream, a, g));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
