// Host-side helpers shared by all translation units of libhd_b200.so:
// status codes, thread-local error string, TMA tensor-map encoding (driver entry point
// fetched through the runtime, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>

#define HD_OK 0
#define HD_ERR_INVALID (-22)   // -EINVAL: bad shape / argument
#define HD_ERR_CUDA (-5)       // -EIO: CUDA runtime / driver error
#define HD_ERR_UNSUPPORTED (-38)  // -ENOSYS

namespace hd {

char* err_buf();                      // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

#define HD_CHECK_CUDA(...)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (__VA_ARGS__);                                                       \
        if (_e != cudaSuccess)                                                                \
            return ::hd::fail(HD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #__VA_ARGS__, \
                              cudaGetErrorString(_e));                                        \
    } while (0)

#define HD_REQUIRE(cond, ...)                                      \
    do {                                                           \
        if (!(cond)) return ::hd::fail(HD_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: remember it per call site and
// device (bit i of the mask = device i), so that one process driving several GPUs opts in on each of them.
#define HD_ENSURE_DYN_SMEM(kernel, bytes)                                                                    \
    do {                                                                                                     \
        static std::atomic<unsigned long long> _hd_done{0};                                                  \
        int _hd_dev = 0;                                                                                     \
        HD_CHECK_CUDA(cudaGetDevice(&_hd_dev));                                                              \
        const unsigned long long _hd_bit = 1ull << (_hd_dev & 63);                                           \
        if (!(_hd_done.load(std::memory_order_acquire) & _hd_bit)) {                                         \
            HD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
            _hd_done.fetch_or(_hd_bit, std::memory_order_release);                                           \
        }                                                                                                    \
    } while (0)

// Encode a tiled bf16 tensor map (shared-memory swizzle of 128 / 64 / 0 bytes) with zero OOB fill.
// dims/strides are innermost-first; strides[0] is implied (2 bytes) and strides_bytes has rank-1 entries.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes = 128);

int sm_count();
// SMs the persistent tensor-core kernels may fill: sm_count() minus the calling thread's reservation (SmReserve).
// Inside the hourglass the executor runs two lanes (the `up1` branch beside the spine); a persistent conv / wgrad grid
// of one CTA per SM owns every SM for its whole duration (all registers, ~200 KB of shared memory), so the other lane's
// kernels - the latency-bound chain of the deep levels, 16-64 CTAs each - cannot start until it ends and the lanes
// serialise at kernel granularity. Leaving a few SMs free lets that chain run underneath the big kernels.
int sm_budget();
int sm_reserve_set(int n);
struct SmReserve {
    int old;
    explicit SmReserve(int n) : old(sm_reserve_set(n)) {}
    ~SmReserve() { sm_reserve_set(old); }
};
void count_launch();   // bumps the kernel-launch counter read by hd_launch_count()
bool pdl_enabled();    // programmatic dependent launch on (default) / off (hd_set_pdl(0) or HD_NO_PDL=1)
bool pdl_allowed(unsigned grid_blocks);   // ... for a launch of this many CTAs under the calling thread's PdlScope
int pdl_scope_set(int v);
// Scoped per-thread override: the training step turns PDL off for its launches. With two execution lanes and the
// weight-gradient side stream the SMs are already kept busy across kernel boundaries, and early-scheduled CTAs that
// only wait take slots from the concurrently running kernels (measured: 13.4 ms/step without, 13.7 ms with PDL),
// while the latency-bound paths - batch-1 inference, decode - gain 7-8 % from it.
// Mode 2 (the training step): only launches of fewer CTAs than the machine has SMs - the latency-bound kernels of the
// deep hourglass levels, where the next kernel's early CTAs find idle SMs instead of competing for busy ones.
struct PdlScope {
    int old;
    explicit PdlScope(int mode) : old(pdl_scope_set(mode)) {}
    ~PdlScope() { pdl_scope_set(old); }
};

// Every kernel of the library is launched through this helper with the "programmatic stream serialization" attribute
// (programmatic dependent launch, PDL): a kernel's CTAs may be scheduled - and run their prologue: barrier init, TMEM
// allocation, tensor-map prefetch - while the previous kernel of the stream is still draining its last wave. Every
// kernel therefore begins with pdl_prologue() (or pdl_launch_dependents() ... pdl_wait() around its prologue):
// `griddepcontrol.wait` returns only when the preceding grid has completed and flushed its memory, so no global memory
// is touched before that. A step is ~280 dependent launches; this removes most of the launch gaps between them.
//
// Exception (launch_k_pdl with early = false): a tensor-core kernel whose grid fills the machine. Its CTAs own a whole
// SM (200+ KB of shared memory); scheduled early they would sit idle on SMs that the concurrently running weight-
// gradient kernel of the side stream could use (measured: the train step got 1.3 % slower with PDL on every kernel).
// HD_TRACE=1: every launch is bracketed by timing events on its stream; hd_trace_dump() prints (stream, start, end,
// kernel) rows relative to the first launch - a poor man's timeline across the executor's streams (profiling aid only:
// the events themselves perturb scheduling a little).
bool trace_enabled();
void trace_begin(const void* func, cudaStream_t stream);
void trace_end(cudaStream_t stream);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_pdl(bool early, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t stream, Args&&... args) {
    if (trace_enabled()) {
        trace_begin(reinterpret_cast<const void*>(kernel), stream);
        cudaLaunchConfig_t c2 = {};
        c2.gridDim = grid; c2.blockDim = block; c2.dynamicSmemBytes = smem; c2.stream = stream;
        cudaError_t e = cudaLaunchKernelEx(&c2, kernel, static_cast<Args&&>(args)...);
        trace_end(stream);
        return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (early && pdl_allowed(grid.x * grid.y * grid.z)) ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
    return launch_k_pdl(true, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }
#endif

static inline int ilog2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace hd
