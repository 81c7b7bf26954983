// Host-side helpers shared by all translation units of libhd_b200.so:
// status codes, thread-local error string, TMA tensor-map encoding (driver entry point
// fetched through the runtime, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#define HD_OK 0
#define HD_ERR_INVALID (-22)   // -EINVAL: bad shape / argument
#define HD_ERR_CUDA (-5)       // -EIO: CUDA runtime / driver error
#define HD_ERR_UNSUPPORTED (-38)  // -ENOSYS

namespace hd {

char* err_buf();                      // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);

#define HD_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::hd::fail(HD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
                              cudaGetErrorString(_e));                                        \
    } while (0)

#define HD_REQUIRE(cond, ...)                                      \
    do {                                                           \
        if (!(cond)) return ::hd::fail(HD_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// Encode a tiled bf16 tensor map with 128B swizzle and zero OOB fill.
// dims/strides are innermost-first; strides[0] is implied (2 bytes) and strides_bytes has rank-1 entries.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128 = true);

int sm_count();
void count_launch();   // bumps the kernel-launch counter read by hd_launch_count()

static inline int ilog2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace hd
