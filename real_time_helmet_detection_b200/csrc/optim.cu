// Fused multi-tensor Adam (the step that immediately follows the gradient all-reduce).
// Reference: optim.py:3-12 builds torch.optim.Adam(network.parameters(), lr) and train.py:128-139 drives it through
// GradScaler (unscale + inf check + step) and zero_grad. Here ONE launch updates all 226 parameter tensors:
// optional 1/grad_scale unscaling and found_inf skip (GradScaler's `_step_supports_amp_scaling` protocol), exp_avg /
// exp_avg_sq update, bias correction and the parameter update, in fp32 with torch.optim.Adam's formulas.
#include "hd_common.h"

namespace hd {

struct AdamJob {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    long long chunk_start;   // index of this tensor's first 1024-element chunk in the flat chunk space
};

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamJob* __restrict__ jobs, int njobs, long long nchunks,
                                                         float lr, float beta1, float beta2, float eps,
                                                         const float* __restrict__ step_dev,
                                                         const float* __restrict__ grad_scale,
                                                         const float* __restrict__ found_inf) {
    pdl_prologue();
    if (found_inf && *found_inf != 0.f) return;          // GradScaler: skip the whole step on inf / nan gradients
    const long long chunk = blockIdx.x;
    if (chunk >= nchunks) return;
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].chunk_start <= chunk) lo = mid; else hi = mid - 1;
    }
    const AdamJob j = jobs[lo];
    const float t = *step_dev + 1.f;
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float step_size = lr / bc1;
    const float inv_scale = grad_scale ? 1.f / *grad_scale : 1.f;
    const long long base = (chunk - j.chunk_start) * 1024;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i < j.n) {
            const float g = j.g[i] * inv_scale;
            float m = j.m[i], v = j.v[i];
            m = m + (1.f - beta1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
            v = beta2 * v + (1.f - beta2) * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
            const float denom = sqrtf(v) / bc2_sqrt + eps;
            j.p[i] = j.p[i] - step_size * (m / denom);             // param.addcdiv_(exp_avg, denom, -step_size)
            j.m[i] = m;
            j.v[i] = v;
        }
    }
}

__global__ void adam_advance_step_kernel(float* step_dev, const float* found_inf) {
    pdl_prologue();
    if (!(found_inf && *found_inf != 0.f)) *step_dev += 1.f;
}

}  // namespace hd

// jobs_host: njobs records {float* p; const float* g; float* m; float* v; long long n; long long chunk_start}
// (48 bytes each, chunk_start ascending, chunks of 1024 elements); jobs_dev: device scratch of the same size.
// step_dev: device float holding the number of optimizer steps taken so far (incremented here unless found_inf).
extern "C" int hd_adam_step(const void* jobs_host, int njobs, void* jobs_dev, long long nchunks, float lr, float beta1,
                            float beta2, float eps, float* step_dev, const float* grad_scale, const float* found_inf,
                            cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(jobs_host && jobs_dev && njobs > 0 && nchunks > 0 && step_dev, "adam_step: bad arguments");
    HD_REQUIRE(nchunks < (1ll << 31), "adam_step: too many chunks");
    HD_CHECK_CUDA(cudaMemcpyAsync(jobs_dev, jobs_host, static_cast<size_t>(njobs) * sizeof(AdamJob),
                                  cudaMemcpyHostToDevice, stream));
    HD_CHECK_CUDA(::hd::launch_k(adam_multi_kernel, static_cast<unsigned>(nchunks), 256, 0, stream,
                                  reinterpret_cast<const AdamJob*>(jobs_dev), njobs, nchunks, lr, beta1, beta2, eps,
                                 step_dev, grad_scale, found_inf));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    HD_CHECK_CUDA(::hd::launch_k(adam_advance_step_kernel, 1, 1, 0, stream, step_dev, found_inf));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
