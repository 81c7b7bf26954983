// Fused CenterNet decode + NMS, one CTA per image, no host synchronisation.
// Reference: transform.py:73-110 `hm2box` and evaluate.py:126-182 `Prediction.forward` /
// `nonmaximum_supression` (torchvision.ops.nms semantics, class-agnostic).
//
// Per stack: head activation (sigmoid) -> 3x3 equality peak test -> joint top-k over (C,H,W) by radix select
// -> bitonic sort of the k survivors -> offset/size gather + box assembly -> `score >= conf_th` prefix.
// Then over the concatenated stacks: stable sort by score, all-pairs IoU bit-matrix in shared memory,
// serial sweep by one warp, compacted output in score order.
//
// Determinism: equal scores are ordered by ascending flat index (top-k) / ascending candidate position (NMS),
// the tie-break the oracle (oracle/decode_ref.py) fixes too. Box arithmetic follows the reference's operation
// order in fp32 without FMA contraction so coordinates are bit-identical.
// Heat-map values must be >= 0 (probabilities or GT heat-maps), as in every reference call site.
#include "hd_common.h"

namespace hd {

constexpr int kDecThreads = 1024;
constexpr int kMaxCand = 1024;          // S * topk upper bound
constexpr int kBigFloats = 32768;       // 128 KB region: score map (when it fits) / NMS bit matrix

struct DecodeArgs {
    const float* heat; const float* off; const float* wh;
    long long bs_heat, ss_heat, bs_off, ss_off, bs_wh, ss_wh;   // batch / stack strides (elements)
    int B, S, C, H, W, K;
    float scale_factor, conf_th, nms_th;
    int normalized, apply_sigmoid, do_nms;
    float* scratch;        // [B][C*H*W] floats, used when the map does not fit the shared region
    float* out_boxes;      // [B][S*K][4]
    long long* out_cls;    // [B][S*K]
    float* out_scores;     // [B][S*K]
    int* out_count;        // [B]
};

__device__ __forceinline__ float dsigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned fkey(float v) {  // order-preserving map for v >= 0 (and general floats)
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = warp_sums[lane];
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        warp_sums[lane] = si - s;          // exclusive warp offsets
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[w] + inc - v;
}

// In-place bitonic sort, descending, of n (power of two <= 1024) 64-bit keys in shared memory.
__device__ __forceinline__ void bitonic_desc(unsigned long long* keys, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            const int i = threadIdx.x;
            if (i < n) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = keys[i], b = keys[l];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[l] = a; }
                }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDecThreads, 1) decode_nms_kernel(const DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char dsm[];
    float* big = reinterpret_cast<float*>(dsm);                                   // kBigFloats floats
    unsigned* pk_s = reinterpret_cast<unsigned*>(big + kBigFloats);               // kBigFloats/32 words
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(pk_s + kBigFloats / 32);  // 1024
    float* cbox = reinterpret_cast<float*>(keys + 1024);                          // [kMaxCand][4]
    float* cscore = cbox + 4 * kMaxCand;                                          // [kMaxCand]
    int* ccls = reinterpret_cast<int*>(cscore + kMaxCand);                        // [kMaxCand]
    int* order = ccls + kMaxCand;                                                 // [kMaxCand]
    int* hist = order + kMaxCand;                                                 // 256
    int* wsum = hist + 256;                                                       // 33
    int* misc = wsum + 40;                                                        // small scalars

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int HW = a.H * a.W;
    const int CHW = a.C * HW;
    const bool in_smem = CHW <= kBigFloats;
    float* v = in_smem ? big : a.scratch + static_cast<size_t>(b) * CHW;
    // peak bits: shared when they fit, else reuse the tail of the global scratch row (as floats are 32-bit too)
    unsigned* pk = in_smem ? pk_s : reinterpret_cast<unsigned*>(a.scratch + static_cast<size_t>(a.B) * CHW) +
                                        static_cast<size_t>(b) * ((CHW + 31) / 32);
    const int CHW_pad = (CHW + 31) & ~31;
    int ncand = 0;  // candidates gathered so far (uniform across the block)

    for (int s = 0; s < a.S; ++s) {
        const float* heat = a.heat + b * a.bs_heat + s * a.ss_heat;
        const float* offp = a.off + b * a.bs_off + s * a.ss_off;
        const float* whp = a.wh + b * a.bs_wh + s * a.ss_wh;
        // ---- 1. activation
        for (int i = tid; i < CHW; i += kDecThreads) {
            float x = heat[i];
            v[i] = a.apply_sigmoid ? dsigmoid(x) : x;
        }
        __syncthreads();
        // ---- 2. peak test (equality with the 3x3 max, -inf padding) + count of positive peaks
        int my_pos = 0;
        for (int i = tid; i < CHW_pad; i += kDecThreads) {
            bool peak = false;
            if (i < CHW) {
                const int c = i / HW, r = i - c * HW;
                const int y = r / a.W, x = r - y * a.W;
                const float me = v[i];
                float m = me;
                const float* pl = v + c * HW;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= a.H) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx;
                        if (xx < 0 || xx >= a.W) continue;
                        m = fmaxf(m, pl[yy * a.W + xx]);
                    }
                }
                peak = (m == me);
                if (peak && me > 0.f) ++my_pos;
            }
            const unsigned w = __ballot_sync(0xffffffffu, peak);
            if ((tid & 31) == 0) pk[i >> 5] = w;
        }
        int npos;
        block_excl_scan(my_pos, wsum, &npos);
        // peak-map value of element i: pv = peak ? v[i] : 0 ; only pv > 0 take part in the radix select
        const bool want_fill = !(a.conf_th > 0.f);     // zero-score fillers survive only when conf_th <= 0
        const int K = a.K;
        unsigned T = 0x80000000u;                      // key(0.0f)
        int need_eq = 0;                               // how many elements with key == T to take
        int n_gt;                                      // how many elements with key > T
        if (npos >= K) {
            unsigned prefix = 0;
            int remaining = K;
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                for (int i = tid; i < CHW; i += kDecThreads) {
                    const float val = v[i];
                    if (val > 0.f && ((pk[i >> 5] >> (i & 31)) & 1u)) {
                        const unsigned key = fkey(val);
                        if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
                            atomicAdd(&hist[(key >> shift) & 255u], 1);
                    }
                }
                __syncthreads();
                if (tid < 32) {
                    // warp-parallel scan of the 256 bins from the top: lane l owns bins 255-8l .. 248-8l
                    int cnt[8], sum = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        cnt[j] = hist[255 - (tid * 8 + j)];
                        sum += cnt[j];
                    }
                    int inc = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (tid >= o) inc += t;
                    }
                    const unsigned hit = __ballot_sync(0xffffffffu, inc >= remaining);
                    const int owner = hit ? __ffs(hit) - 1 : 31;
                    if (tid == owner) {
                        int acc = inc - sum, j = 0;
                        for (; j < 7; ++j) {
                            if (acc + cnt[j] >= remaining) break;
                            acc += cnt[j];
                        }
                        misc[0] = 255 - (tid * 8 + j);
                        misc[1] = remaining - acc;
                    }
                }
                __syncthreads();
                prefix |= static_cast<unsigned>(misc[0]) << shift;
                remaining = misc[1];
                __syncthreads();
            }
            T = prefix;
            need_eq = remaining;
            n_gt = K - need_eq;
        } else {
            n_gt = npos;
            need_eq = want_fill ? (K - npos) : 0;
        }
        // ---- 3. collect: everything with key > T (unordered), then the first need_eq elements with key == T
        if (tid == 0) { misc[2] = 0; misc[3] = 0; }
        __syncthreads();
        int my_eq = 0;
        for (int i = tid; i < CHW; i += kDecThreads) {
            const float val = v[i];
            const bool peak = (pk[i >> 5] >> (i & 31)) & 1u;
            const float pv = (peak && val > 0.f) ? val : 0.f;
            const unsigned key = fkey(pv);
            if (key > T) {
                const int slot = atomicAdd(&misc[2], 1);
                keys[slot] = (static_cast<unsigned long long>(key) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(i));
            } else if (key == T) {
                ++my_eq;
            }
        }
        int total_eq;
        block_excl_scan(my_eq, wsum, &total_eq);
        if (need_eq > 0) {
            if (total_eq == need_eq) {
                for (int i = tid; i < CHW; i += kDecThreads) {
                    const float val = v[i];
                    const bool peak = (pk[i >> 5] >> (i & 31)) & 1u;
                    const float pv = (peak && val > 0.f) ? val : 0.f;
                    if (fkey(pv) == T) {
                        const int slot = n_gt + atomicAdd(&misc[3], 1);
                        keys[slot] = (static_cast<unsigned long long>(T) << 32) |
                                     (0xFFFFFFFFu - static_cast<unsigned>(i));
                    }
                }
            } else {
                // ordered selection: lowest indices first (block scan per 1024-element stripe)
                int taken = 0;
                for (int base = 0; base < CHW && taken < need_eq; base += kDecThreads) {
                    const int i = base + tid;
                    int flag = 0;
                    if (i < CHW) {
                        const float val = v[i];
                        const bool peak = (pk[i >> 5] >> (i & 31)) & 1u;
                        const float pv = (peak && val > 0.f) ? val : 0.f;
                        flag = fkey(pv) == T;
                    }
                    int tot;
                    const int rank = taken + block_excl_scan(flag, wsum, &tot);
                    if (flag && rank < need_eq)
                        keys[n_gt + rank] = (static_cast<unsigned long long>(T) << 32) |
                                            (0xFFFFFFFFu - static_cast<unsigned>(i));
                    taken += tot;
                }
            }
        }
        __syncthreads();
        const int nsel = n_gt + need_eq;               // <= K
        int npad = 1;
        while (npad < nsel) npad <<= 1;
        if (tid >= nsel && tid < npad) keys[tid] = 0ull;
        __syncthreads();
        bitonic_desc(keys, npad);
        // ---- 4. gather + boxes + threshold (kept entries are a prefix because scores are sorted)
        int keep = 0;
        if (tid < nsel) {
            const unsigned long long kk = keys[tid];
            const int idx = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(kk & 0xFFFFFFFFull));
            const float val = v[idx];
            const bool peak = (pk[idx >> 5] >> (idx & 31)) & 1u;
            const float score = (peak && val > 0.f) ? val : 0.f;
            const int cls = idx / HW, r = idx - cls * HW;
            const int y = r / a.W, x = r - y * a.W;
            float xo = offp[r], yo = offp[HW + r], ws = whp[r], hs = whp[HW + r];
            if (a.normalized) {
                if (a.apply_sigmoid) { xo = dsigmoid(xo); yo = dsigmoid(yo); ws = dsigmoid(ws); hs = dsigmoid(hs); }
                xo = __fmul_rn(xo, a.scale_factor);
                yo = __fmul_rn(yo, a.scale_factor);
                ws = __fmul_rn(ws, static_cast<float>(a.W));
                hs = __fmul_rn(hs, static_cast<float>(a.H));
            }
            const float cx = __fadd_rn(static_cast<float>(x), xo), cy = __fadd_rn(static_cast<float>(y), yo);
            const float hw_ = __fmul_rn(ws, 0.5f), hh_ = __fmul_rn(hs, 0.5f);
            keep = score >= a.conf_th;
            if (keep && ncand + tid < kMaxCand) {
                float* bx = cbox + 4 * (ncand + tid);
                bx[0] = __fmul_rn(__fsub_rn(cx, hw_), a.scale_factor);
                bx[1] = __fmul_rn(__fsub_rn(cy, hh_), a.scale_factor);
                bx[2] = __fmul_rn(__fadd_rn(cx, hw_), a.scale_factor);
                bx[3] = __fmul_rn(__fadd_rn(cy, hh_), a.scale_factor);
                cscore[ncand + tid] = score;
                ccls[ncand + tid] = cls;
            }
        }
        int nkeep;
        block_excl_scan(keep, wsum, &nkeep);
        ncand += nkeep;
        __syncthreads();
    }

    // ------------------------------------------------------------------------------------------------ NMS
    const int N = ncand;
    float* ob = a.out_boxes + static_cast<size_t>(b) * a.S * a.K * 4;
    long long* oc = a.out_cls + static_cast<size_t>(b) * a.S * a.K;
    float* os = a.out_scores + static_cast<size_t>(b) * a.S * a.K;
    if (!a.do_nms) {
        for (int i = tid; i < N; i += kDecThreads) {
            ob[4 * i] = cbox[4 * i]; ob[4 * i + 1] = cbox[4 * i + 1];
            ob[4 * i + 2] = cbox[4 * i + 2]; ob[4 * i + 3] = cbox[4 * i + 3];
            oc[i] = ccls[i];
            os[i] = cscore[i];
        }
        if (tid == 0) a.out_count[b] = N;
        return;
    }
    // stable sort by descending score (ties: ascending candidate position)
    int npad = 1;
    while (npad < N) npad <<= 1;
    if (tid < npad)
        keys[tid] = tid < N ? ((static_cast<unsigned long long>(fkey(cscore[tid])) << 32) |
                               (0xFFFFFFFFu - static_cast<unsigned>(tid)))
                            : 0ull;
    __syncthreads();
    if (a.S > 1) bitonic_desc(keys, npad);
    if (tid < N) order[tid] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[tid] & 0xFFFFFFFFull));
    __syncthreads();
    // suppression bit matrix: bit j of row i set <=> j > i (in sorted order) and IoU(i, j) > nms_th
    const int words = (N + 63) >> 6;
    unsigned long long* mat = reinterpret_cast<unsigned long long*>(big);   // [N][words] <= 1024*16*8 = 128 KB
    for (int t = tid; t < N * words; t += kDecThreads) {
        const int i = t / words, wj = t - i * words;
        const float* bi = cbox + 4 * order[i];
        const float ax1 = bi[0], ay1 = bi[1], ax2 = bi[2], ay2 = bi[3];
        const float aarea = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
        unsigned long long bits = 0ull;
        const int j0 = wj << 6;
        for (int jj = 0; jj < 64; ++jj) {
            const int j = j0 + jj;
            if (j <= i || j >= N) continue;
            const float* bj = cbox + 4 * order[j];
            const float w = fmaxf(0.f, __fsub_rn(fminf(ax2, bj[2]), fmaxf(ax1, bj[0])));
            const float h = fmaxf(0.f, __fsub_rn(fminf(ay2, bj[3]), fmaxf(ay1, bj[1])));
            const float inter = __fmul_rn(w, h);
            const float barea = __fmul_rn(__fsub_rn(bj[2], bj[0]), __fsub_rn(bj[3], bj[1]));
            const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
            if (iou > a.nms_th) bits |= 1ull << jj;
        }
        mat[t] = bits;
    }
    __syncthreads();
    // serial sweep by warp 0: lane l owns removed-word l (N <= 1024 -> <= 16 words)
    if (tid < 32) {
        unsigned long long removed = 0ull;
        int nk = 0;
        for (int i = 0; i < N; ++i) {
            const unsigned long long rw = __shfl_sync(0xffffffffu, removed, i >> 6);
            if (!((rw >> (i & 63)) & 1ull)) {
                if (tid == 0) {
                    const int src = order[i];
                    ob[4 * nk] = cbox[4 * src]; ob[4 * nk + 1] = cbox[4 * src + 1];
                    ob[4 * nk + 2] = cbox[4 * src + 2]; ob[4 * nk + 3] = cbox[4 * src + 3];
                    oc[nk] = ccls[src];
                    os[nk] = cscore[src];
                }
                ++nk;
                if (tid < words) removed |= mat[i * words + tid];
            }
        }
        if (tid == 0) a.out_count[b] = nk;
    }
}

constexpr size_t kDecSmem = kBigFloats * 4 + kBigFloats / 8 + 1024 * 8 + kMaxCand * (16 + 4 + 4 + 4) + 256 * 4 +
                            40 * 4 + 64;

}  // namespace hd

extern "C" size_t hd_decode_scratch_bytes(int B, int C, int H, int W) {
    const size_t chw = static_cast<size_t>(C) * H * W;
    if (chw <= static_cast<size_t>(hd::kBigFloats)) return 16;
    return static_cast<size_t>(B) * (chw + (chw + 31) / 32) * sizeof(float);
}

// See include/hd_b200.h.
extern "C" int hd_decode_nms(const float* heat, long long bs_heat, long long ss_heat, const float* off,
                             long long bs_off, long long ss_off, const float* wh, long long bs_wh, long long ss_wh,
                             int B, int S, int C, int H, int W, int topk, float scale_factor, float conf_th,
                             float nms_th, int normalized, int apply_sigmoid, int do_nms, void* scratch,
                             float* out_boxes, long long* out_cls, float* out_scores, int* out_count,
                             cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(B > 0 && S > 0 && C > 0 && H > 0 && W > 0, "decode: empty input");
    HD_REQUIRE(topk >= 1, "decode: topk=%d", topk);
    HD_REQUIRE(static_cast<long long>(topk) <= static_cast<long long>(C) * H * W,
               "decode: selected index k out of range (topk=%d > C*H*W=%lld)", topk,
               static_cast<long long>(C) * H * W);
    HD_REQUIRE(static_cast<long long>(S) * topk <= kMaxCand, "decode: S*topk=%d exceeds the fused limit %d",
               S * topk, kMaxCand);
    DecodeArgs a{};
    a.heat = heat; a.off = off; a.wh = wh;
    a.bs_heat = bs_heat; a.ss_heat = ss_heat; a.bs_off = bs_off; a.ss_off = ss_off; a.bs_wh = bs_wh; a.ss_wh = ss_wh;
    a.B = B; a.S = S; a.C = C; a.H = H; a.W = W; a.K = topk;
    a.scale_factor = scale_factor; a.conf_th = conf_th; a.nms_th = nms_th;
    a.normalized = normalized; a.apply_sigmoid = apply_sigmoid; a.do_nms = do_nms;
    a.scratch = reinterpret_cast<float*>(scratch);
    a.out_boxes = out_boxes; a.out_cls = out_cls; a.out_scores = out_scores; a.out_count = out_count;
    static bool attr_set = false;
    if (!attr_set) {
        HD_CHECK_CUDA(cudaFuncSetAttribute(decode_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kDecSmem)));
        attr_set = true;
    }
    decode_nms_kernel<<<B, kDecThreads, kDecSmem, stream>>>(a);
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
