// Fused CenterNet decode + NMS without any host synchronisation.
// Reference: transform.py:73-110 `hm2box` and evaluate.py:126-182 `Prediction.forward` /
// `nonmaximum_supression` (torchvision.ops.nms semantics, class-agnostic).
//
// Two launches for the whole batch:
//   1. decode_peaks_kernel  (grid over all pixels of all images / stacks / classes): head activation (sigmoid),
//      3x3 equality peak test, and compaction of the positive peaks into a per-(image, stack) candidate list of
//      64-bit keys  (order-preserving score bits << 32 | ~flat_index)  with one warp-aggregated atomic per warp.
//      Because the reference applies `score >= conf_th` AFTER top-k, and a threshold commutes with taking a sorted
//      prefix, peaks below a positive conf_th are dropped right here (typically 2,900 -> 50 candidates).
//   2. decode_select_nms_kernel (one CTA per image): per stack, exact top-k of the candidate keys (8-bit radix
//      select with early exit, only when there are more than k), rank sort of the <= k survivors, offset/size gather,
//      box assembly in the reference's fp32 operation order, threshold prefix; then over the concatenated stacks a
//      stable sort by score, the all-pairs IoU bit matrix in shared memory, a serial sweep by one warp and the
//      compacted output in score order.
//
// Determinism: keys are unique, so the result does not depend on the order in which candidates were appended; equal
// scores are ordered by ascending flat index (top-k) / ascending candidate position (NMS) - the tie-break the oracle
// (oracle/decode_ref.py) fixes too. Zero-score fillers (only kept when conf_th <= 0, as in the reference) are the
// lowest flat indices that are not positive peaks. Box arithmetic uses __f*_rn intrinsics (no FMA contraction) so
// coordinates are bit-identical to the reference. Heat-map values must be >= 0 (probabilities / GT heat-maps).
#include <cmath>
#include <cstdlib>

#include "hd_common.h"

namespace hd {

constexpr int kSelThreads = 1024;
constexpr int kMaxCand = 1024;            // S * topk upper bound (fused NMS limit)
constexpr int kNmsWords = kMaxCand / 64;  // 16

struct DecodeArgs {
    const float* heat; const float* off; const float* wh;
    long long bs_heat, ss_heat, bs_off, ss_off, bs_wh, ss_wh;   // batch / stack strides (elements)
    int B, S, C, H, W, K;
    float scale_factor, conf_th, nms_th;
    float logit_th;                  // logit(conf_th) - margin (see decode_peaks_kernel); -inf when conf_th <= 0
    int normalized, apply_sigmoid, do_nms;
    // scratch (per image-stack pair p = b*S + s)
    int* cand_count;                 // [B*S]
    unsigned long long* cand_keys;   // [B*S][C*H*W]
    unsigned char* is_pos;           // [B*S][C*H*W]  (written only when fillers may be needed)
    float* out_boxes;      // [B][S*K][4]
    long long* out_cls;    // [B][S*K]
    float* out_scores;     // [B][S*K]
    int* out_count;        // [B]
};

__device__ __forceinline__ float dsigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned fkey(float v) {  // order-preserving map float -> uint
    unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// ------------------------------------------------------------------------------------------------ kernel 1
// Peak test of ONE heat-map element per lane + warp-aggregated append to the candidate list of (image, stack) p.
// Must be called by all 32 lanes of a warp with the same p (`inside` false for padding lanes).
__device__ __forceinline__ void peak_test_append(const DecodeArgs& a, int p, int c, int y, int x, bool inside) {
    const int b = p / a.S, s = p - b * a.S;
    const int HW = a.H * a.W;
    const float* pl = a.heat + b * a.bs_heat + s * a.ss_heat + static_cast<long long>(c) * HW;
    const bool fill = !(a.conf_th > 0.f);
    bool take = false;
    float me = 0.f;
    // With a positive confidence threshold only elements with sigmoid(logit) >= conf_th can ever be emitted (the
    // reference thresholds after top-k, and a threshold commutes with a sorted prefix), so everything else skips the
    // nine sigmoids of the peak test: one compare against logit(conf_th) minus a margin far above the rounding of expf.
    // (A typical map keeps ~50 of 32,768 elements.) The exact `sigmoid(me) >= conf_th` test still follows below.
    bool maybe = inside;
    if (inside && !fill) {
        const float raw = pl[y * a.W + x];
        maybe = a.apply_sigmoid ? raw >= a.logit_th : raw >= a.conf_th;
    }
    if (maybe) {
        me = pl[y * a.W + x];
        if (a.apply_sigmoid) me = dsigmoid(me);
        float m = me;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= a.H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if ((dy == 0 && dx == 0) || xx < 0 || xx >= a.W) continue;
                float v = pl[yy * a.W + xx];
                if (a.apply_sigmoid) v = dsigmoid(v);
                m = fmaxf(m, v);
            }
        }
        const bool pos_peak = (m == me) && me > 0.f;
        if (fill) a.is_pos[static_cast<size_t>(p) * a.C * HW + static_cast<size_t>(c) * HW + y * a.W + x] = pos_peak;
        take = pos_peak && (fill || me >= a.conf_th);
    }
    // warp-aggregated append
    const unsigned lane = threadIdx.x & 31u;
    const unsigned vote = __ballot_sync(0xffffffffu, take);
    if (vote) {
        int base = 0;
        const int leader = __ffs(vote) - 1;
        if (static_cast<int>(lane) == leader) base = atomicAdd(a.cand_count + p, __popc(vote));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (take) {
            const int slot = base + __popc(vote & ((1u << lane) - 1u));
            const unsigned idx = static_cast<unsigned>(c) * HW + y * a.W + x;
            a.cand_keys[static_cast<size_t>(p) * a.C * HW + slot] =
                (static_cast<unsigned long long>(fkey(me)) << 32) | (0xFFFFFFFFu - idx);
        }
    }
}

// grid = (ceil(W/32), ceil(H/8), B*S*C), block = (32, 8): one thread per heat-map element.
__global__ void __launch_bounds__(256) decode_peaks_kernel(const DecodeArgs a) {
    pdl_prologue();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = blockIdx.z;
    const int c = z % a.C, p = z / a.C;          // p = b*S + s
    peak_test_append(a, p, c, y, x, x < a.W && y < a.H);
}

// ------------------------------------------------------------------------------------------------ kernel 2
// Block-wide exclusive scan of one int per thread (blockDim.x = 32 * nwarps <= 1024 threads).
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total, int nwarps) {
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
        int s = lane < nwarps ? warp_sums[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[w] + inc - v;
}

// Descending rank sort of n <= 1024 unique 64-bit keys: thread i places keys[i] at its rank. src -> dst.
__device__ __forceinline__ void rank_sort_desc(const unsigned long long* src, unsigned long long* dst, int n) {
    __syncthreads();
    if (threadIdx.x < n) {
        const unsigned long long me = src[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += src[j] > me;
        dst[rank] = me;
    }
    __syncthreads();
}

// Per-image part: per stack top-k / gather / boxes / threshold, then NMS over the concatenated stacks. One CTA.
__device__ __forceinline__ void select_nms_image(const DecodeArgs& a, const int b, unsigned char* dsm) {
    unsigned long long* mat = reinterpret_cast<unsigned long long*>(dsm);             // [kMaxCand][kNmsWords] 128 KB
    unsigned long long* keys = mat + kMaxCand * kNmsWords;                              // [1024] selected
    unsigned long long* sorted = keys + 1024;                                           // [1024]
    float* cbox = reinterpret_cast<float*>(sorted + 1024);                              // [kMaxCand][4]
    float* cscore = cbox + 4 * kMaxCand;
    int* ccls = reinterpret_cast<int*>(cscore + kMaxCand);
    int* order = ccls + kMaxCand;
    int* hist = order + kMaxCand;    // 256
    int* wsum = hist + 256;          // 33 (+pad)
    int* misc = wsum + 40;

    const int tid = threadIdx.x;
    const int NT = blockDim.x;        // 1024
    const int HW = a.H * a.W;
    const int CHW = a.C * HW;
    const int K = a.K;
    const bool fill = !(a.conf_th > 0.f);
    int ncand = 0;
    // (Tried in round 2: warps 4.. leaving early when a positive threshold leaves <= 128 candidates, to make the ~20
    // block-wide barriers cheaper - 20.5 us instead of 12.9-15.2 us: the pair-parallel IoU stage wants all 32 warps.)
    for (int s = 0; s < a.S; ++s) {
        const int p = b * a.S + s;
        const float* offp = a.off + b * a.bs_off + s * a.ss_off;
        const float* whp = a.wh + b * a.bs_wh + s * a.ss_wh;
        const unsigned long long* ck = a.cand_keys + static_cast<size_t>(p) * CHW;
        const int n = min(a.cand_count[p], CHW);
        int nsel;
        if (n <= K) {
            for (int i = tid; i < n; i += NT) keys[i] = ck[i];
            nsel = n;
            if (fill && n < K) {
                // zero-score fillers: the lowest flat indices that are not positive peaks, in ascending order
                const unsigned char* isp = a.is_pos + static_cast<size_t>(p) * CHW;
                const int need = K - n;
                int taken = 0;
                for (int base = 0; base < CHW && taken < need; base += NT) {
                    const int i = base + tid;
                    const int flag = (i < CHW) && !isp[i];
                    int tot;
                    const int rank = taken + block_excl_scan(flag, wsum, &tot, NT >> 5);
                    if (flag && rank < need)
                        keys[n + rank] = (static_cast<unsigned long long>(fkey(0.f)) << 32) |
                                         (0xFFFFFFFFu - static_cast<unsigned>(i));
                    taken += tot;
                }
                nsel = K;
            }
        } else {
            // exact k-th largest 64-bit key by 8-bit radix select (keys are unique); stop as soon as a bin is taken whole
            unsigned long long prefix = 0ull;
            int remaining = K;
            int shift = 56;
            for (; shift >= 0; shift -= 8) {
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                for (int i = tid; i < n; i += NT) {
                    const unsigned long long key = ck[i];
                    if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
                        atomicAdd(&hist[static_cast<unsigned>(key >> shift) & 255u], 1);
                }
                __syncthreads();
                if (tid < 32) {
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
                        misc[4] = cnt[j];
                    }
                }
                __syncthreads();
                prefix |= static_cast<unsigned long long>(misc[0]) << shift;
                remaining = misc[1];
                const bool whole_bin = misc[4] == remaining;
                __syncthreads();
                if (whole_bin) break;     // every key with this prefix is selected: threshold = prefix (low bits 0)
            }
            if (tid == 0) misc[2] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += NT) {
                const unsigned long long key = ck[i];
                if (key >= prefix) keys[atomicAdd(&misc[2], 1)] = key;
            }
            nsel = K;
        }
        rank_sort_desc(keys, sorted, nsel);
        // gather + boxes + threshold (kept entries are a prefix because scores are sorted)
        int keep = 0;
        if (tid < nsel) {
            const unsigned long long kk = sorted[tid];
            const int idx = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(kk & 0xFFFFFFFFull));
            const float score = fkey_inv(static_cast<unsigned>(kk >> 32));
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
        // the kept entries are a prefix (scores are sorted), so only their count is needed: one barrier instead of a scan
        ncand += __syncthreads_count(keep);
    }

    // self-cleaning scratch: the candidate counters of this image are dead from here on; leaving them at zero spares the
    // next call its memset launch (contract of hd_decode_nms: counters zero on entry, zero on exit)
    if (tid < a.S) a.cand_count[b * a.S + tid] = 0;
    // ------------------------------------------------------------------------------------------------ NMS
    const int N = ncand;
    float* ob = a.out_boxes + static_cast<size_t>(b) * a.S * K * 4;
    long long* oc = a.out_cls + static_cast<size_t>(b) * a.S * K;
    float* os = a.out_scores + static_cast<size_t>(b) * a.S * K;
    if (!a.do_nms) {
        for (int i = tid; i < N; i += NT) {
            ob[4 * i] = cbox[4 * i]; ob[4 * i + 1] = cbox[4 * i + 1];
            ob[4 * i + 2] = cbox[4 * i + 2]; ob[4 * i + 3] = cbox[4 * i + 3];
            oc[i] = ccls[i];
            os[i] = cscore[i];
        }
        if (tid == 0) a.out_count[b] = N;
        return;
    }
    // stable sort by descending score (ties: ascending candidate position); already sorted for one stack
    if (a.S > 1) {
        if (tid < N)
            keys[tid] = (static_cast<unsigned long long>(fkey(cscore[tid])) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(tid));
        rank_sort_desc(keys, sorted, N);
        if (tid < N) order[tid] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(sorted[tid] & 0xFFFFFFFFull));
    } else if (tid < N) {
        order[tid] = tid;
    }
    __syncthreads();
    // suppression bit matrix: bit j of row i set <=> j > i (in sorted order) and IoU(i, j) > nms_th.
    // One THREAD per (i, j) pair: a warp takes 32 consecutive j of one row and a ballot assembles the 32-bit word
    // (N = 53 candidates: 106 warp tasks over 32 warps instead of 53 threads looping over 64 candidates each).
    unsigned* mat32 = reinterpret_cast<unsigned*>(mat);                 // [N][jw]
    const int jw = (N + 31) >> 5;
    {
        const int lane = tid & 31, wid = tid >> 5;
        for (int t = wid; t < N * jw; t += NT >> 5) {
            const int i = t / jw, w = t - i * jw;
            const int j = (w << 5) + lane;
            bool sup = false;
            if (j > i && j < N) {
                const float* bi = cbox + 4 * order[i];
                const float* bj = cbox + 4 * order[j];
                const float ax1 = bi[0], ay1 = bi[1], ax2 = bi[2], ay2 = bi[3];
                const float aarea = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
                const float ww = fmaxf(0.f, __fsub_rn(fminf(ax2, bj[2]), fmaxf(ax1, bj[0])));
                const float hh = fmaxf(0.f, __fsub_rn(fminf(ay2, bj[3]), fmaxf(ay1, bj[1])));
                const float inter = __fmul_rn(ww, hh);
                const float barea = __fmul_rn(__fsub_rn(bj[2], bj[0]), __fsub_rn(bj[3], bj[1]));
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
                sup = iou > a.nms_th;
            }
            const unsigned bits = __ballot_sync(0xffffffffu, sup);
            if (lane == 0) mat32[t] = bits;
        }
    }
    __syncthreads();
    // serial sweep by warp 0 (lane l owns removed-word l; N <= 1024 = 32 words): it only records WHICH candidates
    // survive - the next row's word is fetched before the current one is needed, and nothing is written to global
    // memory inside the dependent chain. All threads then write the kept boxes in parallel.
    int* kept = reinterpret_cast<int*>(keys);                           // [N] candidate indices in output order
    if (tid < 32) {
        unsigned removed = 0u;
        int nk = 0;
        unsigned next = (N > 0 && tid < jw) ? mat32[tid] : 0u;
        for (int i = 0; i < N; ++i) {
            const unsigned row = next;
            if (i + 1 < N && tid < jw) next = mat32[(i + 1) * jw + tid];
            const unsigned rw = __shfl_sync(0xffffffffu, removed, i >> 5);
            if (!((rw >> (i & 31)) & 1u)) {
                if (tid == 0) kept[nk] = order[i];
                ++nk;
                removed |= row;
            }
        }
        if (tid == 0) misc[3] = nk;
    }
    __syncthreads();
    const int nk = misc[3];
    for (int k = tid; k < nk; k += NT) {
        const int src = kept[k];
        *reinterpret_cast<float4*>(ob + 4 * k) = *reinterpret_cast<const float4*>(cbox + 4 * src);
        oc[k] = ccls[src];
        os[k] = cscore[src];
    }
    if (tid == 0) a.out_count[b] = nk;
}

__global__ void __launch_bounds__(kSelThreads, 1) decode_select_nms_kernel(const DecodeArgs a) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char dsm[];
    select_nms_image(a, blockIdx.x, dsm);
}

// ------------------------------------------------------------------------------------------------ fused single launch
// One thread-block CLUSTER per image does both steps: its kClusterSize CTAs scan disjoint slices of the image's
// S * C * H * W logits (peak test + candidate append, as kernel 1), a cluster barrier (release / acquire at cluster scope:
// the appended keys and counters are visible to the whole cluster) ends the scan, and CTA 0 of the cluster carries on
// with the select / NMS part. Compared with the two-launch path this drops one kernel boundary, but see hd_decode_nms:
// measured 25 us vs 13 us at batch 1, so it is opt-in (HD_DECODE_FUSED=1) and the two launches stay the default.
constexpr int kClusterSize = 4;

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSelThreads, 1)
decode_fused_kernel(const DecodeArgs a) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char dsm[];
    unsigned rank, cid;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    const int b = static_cast<int>(cid);
    const int HW = a.H * a.W;
    const int per_stack = a.C * HW;
    // every thread of the cluster takes elements e = rank * NT + tid, + cluster-wide stride; whole warps stay together
    const int stride = kClusterSize * static_cast<int>(blockDim.x);
    for (int s = 0; s < a.S; ++s) {
        const int p = b * a.S + s;
        for (int e0 = 0; e0 < per_stack; e0 += stride) {
            const int e = e0 + static_cast<int>(rank * blockDim.x + threadIdx.x);
            const bool inside = e < per_stack;
            const int ee = inside ? e : 0;
            const int c = ee / HW, r = ee - c * HW;
            const int y = r / a.W, x = r - y * a.W;
            peak_test_append(a, p, c, y, x, inside);
        }
    }
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (rank != 0) return;
    select_nms_image(a, b, dsm);
}

constexpr size_t kSelSmem = static_cast<size_t>(kMaxCand) * kNmsWords * 8 + 2 * 1024 * 8 +
                            kMaxCand * (16 + 4 + 4 + 4) + 256 * 4 + 40 * 4 + 64;

static inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace hd

// scratch layout: [counts: B*S ints][keys: B*S*CHW u64][is_pos: B*S*CHW u8]
extern "C" size_t hd_decode_scratch_bytes(int B, int S, int C, int H, int W) {
    using namespace hd;
    const size_t chw = static_cast<size_t>(C) * H * W, ps = static_cast<size_t>(B) * S;
    return align256(ps * sizeof(int)) + align256(ps * chw * 8) + align256(ps * chw) + 256;
}

// Zero the candidate counters at the head of a decode scratch buffer (needed once per buffer; see hd_decode_nms).
extern "C" int hd_decode_scratch_init(void* scratch, int B, int S, cudaStream_t stream) {
    using namespace hd;
    HD_REQUIRE(scratch != nullptr && B > 0 && S > 0, "decode_scratch_init: bad argument");
    HD_CHECK_CUDA(cudaMemsetAsync(scratch, 0, align256(static_cast<size_t>(B) * S * sizeof(int)), stream));
    return HD_OK;
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
    HD_REQUIRE(static_cast<long long>(B) * S * C <= 65535, "decode: B*S*C=%lld exceeds the grid limit",
               static_cast<long long>(B) * S * C);
    HD_REQUIRE(scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "decode: scratch must be 8-byte aligned");
    DecodeArgs a{};
    a.heat = heat; a.off = off; a.wh = wh;
    a.bs_heat = bs_heat; a.ss_heat = ss_heat; a.bs_off = bs_off; a.ss_off = ss_off; a.bs_wh = bs_wh; a.ss_wh = ss_wh;
    a.B = B; a.S = S; a.C = C; a.H = H; a.W = W; a.K = topk;
    a.scale_factor = scale_factor; a.conf_th = conf_th; a.nms_th = nms_th;
    a.normalized = normalized; a.apply_sigmoid = apply_sigmoid; a.do_nms = do_nms;
    const size_t chw = static_cast<size_t>(C) * H * W, ps = static_cast<size_t>(B) * S;
    unsigned char* sp = reinterpret_cast<unsigned char*>(scratch);
    a.cand_count = reinterpret_cast<int*>(sp);
    a.cand_keys = reinterpret_cast<unsigned long long*>(sp + align256(ps * sizeof(int)));
    a.is_pos = sp + align256(ps * sizeof(int)) + align256(ps * chw * 8);
    a.out_boxes = out_boxes; a.out_cls = out_cls; a.out_scores = out_scores; a.out_count = out_count;
    a.logit_th = -INFINITY;
    if (conf_th > 0.f && conf_th < 1.f) {
        // logit(conf_th) in double, minus an absolute margin of 1e-3 (+ 1e-3 relative): sigmoid'(x) <= 1/4, so the
        // float sigmoid of anything below that is below conf_th by >= ~1e-4 * conf_th (1-conf_th), i.e. by thousands of ulps
        const double lt = log(static_cast<double>(conf_th) / (1.0 - static_cast<double>(conf_th)));
        a.logit_th = static_cast<float>(lt - 1e-3 - 1e-3 * fabs(lt));
    } else if (conf_th >= 1.f) {
        a.logit_th = 8.f;           // sigmoid(x) rounds to 1.0f only for x > ~16.6; keep everything above 8
    }
    // Single launch (one 4-CTA cluster per image, decode_fused_kernel): built and tested, but MEASURED SLOWER on the B200
    // at batch 1 - 25.2 us vs 12.9 us for the two launches below on the same box (round 2): the grid-wide first kernel
    // scans with 128 CTAs, the cluster has 4, and the kernel boundary it saves costs less than that. Opt-in only.
    static const bool fused = getenv("HD_DECODE_FUSED") != nullptr;
    if (conf_th > 0.f && fused) {
        HD_ENSURE_DYN_SMEM(decode_fused_kernel, static_cast<int>(kSelSmem));
        HD_CHECK_CUDA(::hd::launch_k(decode_fused_kernel, B * kClusterSize, kSelThreads, kSelSmem, stream, a));
        HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
        return HD_OK;
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, static_cast<unsigned>(ps * C));
    HD_CHECK_CUDA(::hd::launch_k(decode_peaks_kernel, grid, dim3(32, 8), 0, stream, a));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    HD_ENSURE_DYN_SMEM(decode_select_nms_kernel, static_cast<int>(kSelSmem));
    HD_CHECK_CUDA(::hd::launch_k(decode_select_nms_kernel, B, kSelThreads, kSelSmem, stream, a));
    HD_CHECK_CUDA(cudaGetLastError()); ::hd::count_launch();
    return HD_OK;
}
