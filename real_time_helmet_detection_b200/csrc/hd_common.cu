#include "hd_common.h"

#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>
#include <cstdlib>

namespace hd {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                         : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(HD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
                    (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                    rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return HD_OK;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

static int g_pdl = -1;
static thread_local int t_pdl_scope = -1;     // PdlScope override of the calling thread (-1: none)
int pdl_scope_set(int v) { int old = t_pdl_scope; t_pdl_scope = v; return old; }
bool pdl_enabled() {
    if (t_pdl_scope >= 0 && g_pdl != 0) return t_pdl_scope != 0;      // scope 2: decided per launch in pdl_allowed()
    if (g_pdl < 0) {
        const char* e = getenv("HD_NO_PDL");
        g_pdl = (e && e[0] == '1') ? 0 : 1;
    }
    return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }
int sm_count();
bool pdl_allowed(unsigned grid_blocks) {
    if (!pdl_enabled()) return false;
    return t_pdl_scope != 2 || grid_blocks < static_cast<unsigned>(sm_count());
}

// ---------------------------------------------------------------------------------------------- launch trace (HD_TRACE)
struct TraceRow { const void* func; cudaStream_t stream; cudaEvent_t a, b; };
static std::vector<TraceRow> g_trace;
static std::mutex g_trace_mu;
bool trace_enabled() {
    static const bool on = getenv("HD_TRACE") != nullptr;
    return on;
}
void trace_begin(const void* func, cudaStream_t stream) {
    TraceRow r{func, stream, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, stream);
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace.push_back(r);
}
void trace_end(cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    for (auto it = g_trace.rbegin(); it != g_trace.rend(); ++it)
        if (it->stream == stream) { cudaEventRecord(it->b, stream); break; }
}
void trace_dump() {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    if (g_trace.empty()) return;
    cudaDeviceSynchronize();
    std::vector<cudaStream_t> streams;
    for (auto& r : g_trace) {
        size_t si = 0;
        while (si < streams.size() && streams[si] != r.stream) ++si;
        if (si == streams.size()) streams.push_back(r.stream);
        float t0 = 0.f, t1 = 0.f;
        cudaEventElapsedTime(&t0, g_trace.front().a, r.a);
        cudaEventElapsedTime(&t1, g_trace.front().a, r.b);
        const char* name = "?";
        cudaFuncGetName(&name, r.func);
        fprintf(stderr, "[hd_trace] s%zu %9.3f %9.3f %7.1f us  %s\n", si, t0, t1, (t1 - t0) * 1e3f, name);
    }
    for (auto& r : g_trace) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_trace.clear();
}

int sm_count() {
    static std::atomic<int> cache[64];      // per device (zero-initialised)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

static thread_local int t_sm_reserve = 0;
int sm_reserve_set(int n) { int old = t_sm_reserve; t_sm_reserve = n < 0 ? 0 : n; return old; }
int sm_budget() {
    const int n = sm_count() - t_sm_reserve;
    return n < 8 ? 8 : n;
}

}  // namespace hd

extern "C" const char* hd_last_error(void) { return hd::err_buf(); }

extern "C" int hd_version(void) { return 1; }

namespace hd { long long launches(); }
// Number of kernels this library has launched in this process (bench.py reports it as gpu_launches).
extern "C" long long hd_launch_count(void) { return hd::launches(); }
namespace hd { void set_pdl(int on); }
extern "C" void hd_set_pdl(int on) { hd::set_pdl(on); }
namespace hd { void trace_dump(); }
extern "C" void hd_trace_dump(void) { hd::trace_dump(); }
