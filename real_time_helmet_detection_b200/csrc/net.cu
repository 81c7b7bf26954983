// Whole-network executor: StackedHourglass.forward (hourglass.py:223-237) and its autograd as one native,
// statically planned sequence of kernel launches over a caller-provided HBM arena.
//
// The module tree of the reference is walked in C++ (PreLayer :159-173, Hourglass :130-156 recursively,
// Neck :176-186, Head :189-195, merge convs :215-218/:235); every `Convolution` is one "unit" whose parameter,
// buffer and gradient pointers arrive in an `hd_unit_ptrs` table filled by the Python nn.Module
// (real_time_helmet_detection_b200/hourglass.py keeps the reference's state_dict layout). Unit order:
//   pre.0 | pre.1.{conv1,conv2,skip} | pre.3.{conv1,conv2} | pre.4.{conv1,conv2} |
//   per stack i: hourglass (26 units, construction order) | neck.1 | neck.2.{conv1,conv2} | head |
//                [merge_feature.i, merge_prediction.i]  (i < S-1)
//
// HBM layout of the arena: [persistent: packed bf16 weights, BN statistics / scale-shift, wgrad split-K scratch]
// [forward region: every activation the backward pass needs, bump-allocated in execution order]
// [backward region: gradient temporaries, stack-allocated per block and released on return].
// All launches go to one stream; nothing synchronises with the host.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <map>
#include <vector>

#include "hd_b200.h"
#include "hd_common.h"

namespace hd {

typedef __nv_bfloat16 bf16;

struct Unit {
    int cin, cout, k;
    bool bias, bn;
    int kind;  // 0 regular, 1 stem (7x7 s2), 2 head, 3 merge_prediction (cin = out_ch)
    // persistent device buffers
    bf16* wp = nullptr;    // forward operand
    bf16* wpd = nullptr;   // dgrad operand
    float* stats = nullptr;  // [2][cout]
    float* bnp = nullptr;    // [4][cout] scale, shift, mean, rstd
    // geometry of the last forward
    long long npix = 0;
};

struct ResSaved {
    int u1, u2, us;  // unit ids (us = -1: identity skip)
    bf16 *X, *Y1, *Z1, *Y2, *Ys, *Out;
    uint8_t* Mask;   // training, identity skip: ReLU mask bits of Out (1 bit per element) for the BatchNorm backward
    int H, W, cin, cout;
};

struct HgSaved {
    int up1, low1, low3;   // residual ids
    int low2_res, low2_hg; // one of them >= 0
    bf16 *x, *out;
    int H, W;
};

struct Arena {
    uint8_t* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    void* alloc(size_t bytes) {
        off = (off + 255) & ~size_t(255);
        void* p = base ? base + off : nullptr;
        off += bytes;
        if (off > peak) peak = off;
        return p;
    }
    bool ok() const { return base == nullptr || peak <= cap; }
};

}  // namespace hd

using namespace hd;

struct hd_net {
    int S, in_ch, out_ch;
    std::vector<Unit> units;
    std::vector<ResSaved> res;
    std::vector<HgSaved> hgs;
    // plan state (valid after a forward)
    int B = 0, H = 0, W = 0;
    bool trained_fwd = false;
    Arena persist, fw, bw;
    size_t persist_bytes = 0;
    bool dry = false;
    int rc = 0;
    cudaStream_t stream = nullptr;
    const hd_unit_ptrs* up = nullptr;
    // saved top-level tensors
    bf16 *patches = nullptr, *Y0 = nullptr, *Z0 = nullptr, *R1pool = nullptr;
    uint8_t* R1idx = nullptr;   // argmax of the PreLayer max pool (training: fused into the residual tail before it)
    int r_pre1 = -1, r_pre3 = -1, r_pre4 = -1;
    struct StackSaved {
        int hg_root, neck_res;
        int u_neck, u_head, u_mf, u_mp;
        bf16 *x_in, *hg_out, *Yn, *F1, *F2, *pred64, *T;
    };
    std::vector<StackSaved> stacks;
    // staged backward (hd_net_backward_stage): stage 1 = the stacks (head, neck, hourglass), stage 2 = PreLayer + join.
    // `dX_pre` carries the gradient w.r.t. the PreLayer output across the stage boundary.
    bf16* dX_pre = nullptr;
    int bwd_stage = 0;                      // 0: no backward in flight, 1: stage 1 enqueued
    bool comm_overlap = false;              // stage 1 handed a communication stream: a collective runs under stage 2
    cudaStream_t side_keep = nullptr;
    void* wgrad_ws = nullptr;
    size_t wgrad_ws_bytes = 0;
    float* small = nullptr;   // scratch for BN-backward sums / coefficients
    void* pack_jobs_dev = nullptr;          // device copy of the weight-pack job table
    void* fold_jobs_dev = nullptr;          // device copy of the eval-mode BN fold table
    unsigned int* tickets = nullptr;        // one zeroed word per unit: "last CTA finalizes the BN" ticket counters
    unsigned int* wgrad_sync = nullptr;     // two zeroed words: grid barrier of the weight-gradient kernels
    // weight-gradient kernels run on a side stream so that they overlap the HBM-bound BN-backward kernels of the
    // main stream; their dY operands live in a bump-only region (`wg`) that is never reused within one backward pass
    Arena wg;
    // second execution lane: every hourglass level runs its `up1` residual (and that residual's backward) on `alt`
    // while the spine (pool -> low1 -> ... -> low3) continues on the caller's stream, so that one lane's HBM-bound
    // BN / ReLU kernels overlap the other lane's tensor-core convolutions. The lane owns a stream, a backward
    // stack arena and a BN-backward scratch block; `swap_lane` makes it the current one.
    struct LaneState { cudaStream_t stream = nullptr; Arena bw; float* small = nullptr; } alt;
    cudaStream_t alt_stream = nullptr;
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> events;
    size_t ev_next = 0;
    // eval-mode forwards may reuse the packed bf16 weights and folded BN constants of the previous eval forward
    // (hd_net_set_static_weights): inference with frozen parameters then skips two launches (~40 us of a 0.5 ms pass)
    bool static_weights = false, eval_packed = false;
    const void* eval_packed_ws = nullptr;
    // HD_PHASE_TIMING=1: timing events on the caller's stream at the phase boundaries of forward / backward; the table
    // is printed to stderr at the end of every backward pass (profiling aid, costs a device sync)
    struct Phase { const char* name; cudaEvent_t ev; };
    std::vector<Phase> phases;
    // pinned staging slots for the job tables of forwards recorded into a CUDA graph (see upload_table)
    uint8_t* pinned = nullptr;
    int pinned_next = 0;
    std::map<const void*, std::vector<uint8_t>> uploaded;   // device table -> host copy of what was last uploaded there
    const void* last_ws = nullptr;
    size_t last_ws_bytes = 0;
    ~hd_net() {
        if (pinned) cudaFreeHost(pinned);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        if (side) cudaStreamDestroy(side);
        if (alt_stream) cudaStreamDestroy(alt_stream);
    }
};

static void phase_mark(hd_net* n, const char* name) {
    static const bool on = getenv("HD_PHASE_TIMING") != nullptr;
    if (!on || n->dry || n->rc != 0) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, n->stream);
    n->phases.push_back({name, e});
}
static void phase_report(hd_net* n) {
    if (n->phases.size() < 2) return;
    cudaEventSynchronize(n->phases.back().ev);
    float total = 0.f;
    cudaEventElapsedTime(&total, n->phases.front().ev, n->phases.back().ev);
    fprintf(stderr, "[hd_net phases] total %.3f ms:", total);
    for (size_t i = 1; i < n->phases.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, n->phases[i - 1].ev, n->phases[i].ev);
        fprintf(stderr, "  %s %.3f", n->phases[i].name, ms);
    }
    fprintf(stderr, "\n");
    for (auto& ph : n->phases) cudaEventDestroy(ph.ev);
    n->phases.clear();
}

// The stem runs in its space-to-depth formulation (csrc/stem.cu: x-unfolded 64-channel tensor + four vertical taps);
// HD_STEM_IM2COL=1 selects the first version (147 -> 192 im2col patch matrix, 3x the traffic).
static bool stem_s2d() {
    static const bool im2col = getenv("HD_STEM_IM2COL") != nullptr;
    return !im2col;
}

// SMs left free by the persistent conv / wgrad kernels while the hourglass runs its two lanes (HD_SM_RESERVE). Measured
// on the B200 (round 2, same box, config 2): 0 / 8 / 16 / 24 / 32 reserved SMs -> 12.39 / 12.50 / 12.49 / 12.49 / 12.42 ms
// per step, i.e. no gain - the free SMs are taken by the big lane's own elementwise kernels - so the default is 0.
static int hg_sm_reserve() {
    static const int r = getenv("HD_SM_RESERVE") ? atoi(getenv("HD_SM_RESERVE")) : 0;
    return r;
}

static const hd_unit_ptrs kNullUnit{};
static inline const hd_unit_ptrs& UP(const hd_net* n, int ui) { return n->up ? n->up[ui] : kNullUnit; }

#define RUN(expr)                                   \
    do {                                            \
        if (!n->dry && n->rc == 0) n->rc = (expr);  \
    } while (0)

static inline size_t act_bytes(int B, int H, int W, int C) { return static_cast<size_t>(B) * H * W * C * sizeof(bf16); }

// ------------------------------------------------------------------------------------------------ construction
static int add_unit(hd_net* n, int cin, int cout, int k, bool bias, bool bn, int kind = 0) {
    Unit u;
    u.cin = cin; u.cout = cout; u.k = k; u.bias = bias; u.bn = bn; u.kind = kind;
    n->units.push_back(u);
    return static_cast<int>(n->units.size()) - 1;
}

static int add_residual(hd_net* n, int cin, int cout) {
    ResSaved r{};
    r.u1 = add_unit(n, cin, cout, 3, false, true);
    r.u2 = add_unit(n, cout, cout, 3, false, true);
    r.us = cin != cout ? add_unit(n, cin, cout, 1, false, true) : -1;
    r.cin = cin; r.cout = cout;
    n->res.push_back(r);
    return static_cast<int>(n->res.size()) - 1;
}

static int add_hourglass(hd_net* n, int depth, int ch) {
    HgSaved h{};
    h.up1 = add_residual(n, ch, ch);
    h.low1 = add_residual(n, ch, ch);
    h.low2_res = h.low2_hg = -1;
    if (depth > 1) h.low2_hg = add_hourglass(n, depth - 1, ch);
    else h.low2_res = add_residual(n, ch, ch);
    h.low3 = add_residual(n, ch, ch);
    n->hgs.push_back(h);
    return static_cast<int>(n->hgs.size()) - 1;
}

extern "C" int hd_net_create(int num_stack, int in_ch, int out_ch, hd_net** out) {
    HD_REQUIRE(num_stack >= 1 && num_stack <= 8, "net_create: num_stack=%d", num_stack);
    HD_REQUIRE(in_ch == 128, "net_create: hourglass_inch=%d (the sm_100a kernels are specialised for 128 channels)", in_ch);
    HD_REQUIRE(out_ch >= 5 && out_ch <= 16, "net_create: out_ch=%d (num_cls+4 must be in [5,16])", out_ch);
    hd_net* n = new (std::nothrow) hd_net();
    HD_REQUIRE(n != nullptr, "net_create: out of host memory");
    n->S = num_stack; n->in_ch = in_ch; n->out_ch = out_ch;
    add_unit(n, 3, 64, 7, true, true, 1);
    n->r_pre1 = add_residual(n, 64, 128);
    n->r_pre3 = add_residual(n, 128, 128);
    n->r_pre4 = add_residual(n, 128, in_ch);
    for (int i = 0; i < num_stack; ++i) {
        hd_net::StackSaved s{};
        s.hg_root = add_hourglass(n, 4, in_ch);
        s.u_neck = add_unit(n, in_ch, in_ch, 1, true, true);
        s.neck_res = add_residual(n, in_ch, in_ch);
        s.u_head = add_unit(n, in_ch, out_ch, 1, true, false, 2);
        s.u_mf = s.u_mp = -1;
        if (i < num_stack - 1) {
            s.u_mf = add_unit(n, in_ch, in_ch, 1, true, false);
            s.u_mp = add_unit(n, out_ch, in_ch, 1, true, false, 3);
        }
        n->stacks.push_back(s);
    }
    *out = n;
    return HD_OK;
}

extern "C" void hd_net_destroy(hd_net* n) { delete n; }
extern "C" void hd_net_set_static_weights(hd_net* n, int on) {
    n->static_weights = on != 0;
    n->eval_packed = false;      // the next eval forward packs once more, later ones reuse that
}
extern "C" int hd_net_num_units(const hd_net* n) { return static_cast<int>(n->units.size()); }

// ------------------------------------------------------------------------------------------------ helpers
static inline int hd_sm_count() { return hd::sm_count(); }
static inline int block_n_for(int cout) { return cout > 64 ? 128 : (cout > 16 ? 64 : 16); }
static inline int pad64(int c) { return (c + 63) / 64 * 64; }

static void plan_persistent(hd_net* n) {
    Arena& a = n->persist;
    size_t stats_total = 0;
    for (Unit& u : n->units) stats_total += 2 * static_cast<size_t>(u.cout);
    // statistics and ticket counters are contiguous: one memset per forward clears both
    float* stats = reinterpret_cast<float*>(a.alloc((stats_total + n->units.size() + 64) * sizeof(float)));
    n->tickets = stats ? reinterpret_cast<unsigned int*>(stats + stats_total) : nullptr;
    size_t so = 0;
    for (Unit& u : n->units) {
        u.stats = stats ? stats + so : nullptr;
        so += 2 * u.cout;
        u.bnp = reinterpret_cast<float*>(a.alloc(4 * u.cout * sizeof(float)));
        const int taps = u.kind == 1 ? 1 : u.k * u.k;
        if (u.kind == 1) {
            u.wp = reinterpret_cast<bf16*>(a.alloc(static_cast<size_t>(64) * 256 * 2));    // [64][192] or [4 taps][64][64]
            u.wpd = nullptr;
        } else {
            u.wp = reinterpret_cast<bf16*>(a.alloc(static_cast<size_t>(taps) * block_n_for(u.cout) * pad64(u.cin) * 2));
            u.wpd = u.kind == 2 ? nullptr
                                : reinterpret_cast<bf16*>(a.alloc(static_cast<size_t>(taps) * block_n_for(u.cin) *
                                                                  pad64(u.cout) * 2));
        }
    }
    n->small = reinterpret_cast<float*>(a.alloc(16 * 256 * sizeof(float)));
    n->alt.small = reinterpret_cast<float*>(a.alloc(16 * 256 * sizeof(float)));
    n->pack_jobs_dev = a.alloc(2 * n->units.size() * 64);
    n->fold_jobs_dev = a.alloc(n->units.size() * 64);
    // the per-forward memset also clears the first spare words behind the tickets: the two barrier words of the
    // weight-gradient kernels' in-kernel split-K reduction (zero on entry, left zero by every launch)
    n->wgrad_sync = n->tickets ? n->tickets + n->units.size() : nullptr;
    n->persist_bytes = (stats_total + n->units.size() + 16) * sizeof(float);
}

// Host mirror of hd::PackJob (csrc/pack.cu).
struct PackJobHost {
    const float* w;
    bf16* out;
    int cout, cin, taps, rows_pad, k_pad, mode;
    long long start;
};
extern "C" int hd_pack_all_weights(const void* jobs, int njobs, long long total, cudaStream_t stream);

// One launch repacks every conv weight (fp32 OIHW master -> bf16 UMMA operands). The job table lives in the
// persistent arena and is uploaded with every forward (parameters may have moved).
// Host job table -> device, stream-ordered. Eager execution: a pageable source (the runtime stages it before
// returning, so the host vector may die). While the stream is being CAPTURED into a CUDA graph a pageable copy is
// illegal and the source must outlive the graph: the table goes into a pinned slot that stays untouched for the
// lifetime of the network (the memcpy node re-reads it on every replay; the pointers in it are the captured ones).
constexpr int kPinnedSlots = 32;
constexpr size_t kPinnedSlotBytes = 8192;
static void upload_table(hd_net* n, void* dst, const void* src, size_t bytes) {
    if (n->dry || n->rc != 0) return;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(n->stream, &st) != cudaSuccess) { n->rc = fail(HD_ERR_CUDA, "net: cudaStreamIsCapturing failed"); return; }
    if (st == cudaStreamCaptureStatusActive) {
        // A table identical to the one already uploaded to this slot (the eager warm-up pass that precedes every capture
        // uses the same parameters and arena) needs no copy node at all. That matters beyond the saved node: a host-to-
        // device copy node shares the copy engine with the application's own H2D traffic - the next batch being staged
        // while this graph replays - and queued behind a 115 MB batch copy it stalled the first kernels of the step by
        // ~1.4 ms (measured: 13.7 vs 12.3 ms per step with pinned-host staging).
        auto it = n->uploaded.find(dst);
        if (it != n->uploaded.end() && it->second.size() == bytes && memcmp(it->second.data(), src, bytes) == 0) return;
        if (!n->pinned || n->pinned_next >= kPinnedSlots || bytes > kPinnedSlotBytes) {
            n->rc = fail(HD_ERR_UNSUPPORTED, "net: cannot record this forward into a CUDA graph (%s)",
                         !n->pinned ? "run one eager forward first" : "more than 16 graphs captured on one network");
            return;
        }
        uint8_t* slot = n->pinned + static_cast<size_t>(n->pinned_next++) * kPinnedSlotBytes;
        memcpy(slot, src, bytes);
        src = slot;
    }
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, n->stream) != cudaSuccess)
        n->rc = fail(HD_ERR_CUDA, "net_forward: upload of a job table failed");
    else
        n->uploaded[dst].assign(static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(src) + bytes);
}

static void pack_weights(hd_net* n, bool need_dgrad) {
    std::vector<PackJobHost> jobs;
    long long total = 0;
    for (size_t i = 0; i < n->units.size(); ++i) {
        Unit& u = n->units[i];
        const hd_unit_ptrs& p = UP(n, static_cast<int>(i));
        PackJobHost j{};
        j.w = p.w; j.cout = u.cout; j.cin = u.cin;
        if (u.kind == 1) {
            if (stem_s2d()) {
                j.out = u.wp; j.taps = 4; j.rows_pad = 64; j.k_pad = 64; j.mode = 3;
                j.start = total; total += 4ll * 64 * 64; jobs.push_back(j);
            } else {
                j.out = u.wp; j.taps = 1; j.rows_pad = 64; j.k_pad = 192; j.mode = 2;
                j.start = total; total += 64ll * 192; jobs.push_back(j);
            }
            continue;
        }
        j.taps = u.k * u.k;
        j.out = u.wp; j.rows_pad = block_n_for(u.cout); j.k_pad = pad64(u.cin); j.mode = 0;
        j.start = total; total += static_cast<long long>(j.taps) * j.rows_pad * j.k_pad; jobs.push_back(j);
        if (need_dgrad && u.wpd) {
            j.out = u.wpd; j.rows_pad = block_n_for(u.cin); j.k_pad = pad64(u.cout); j.mode = 1;
            j.start = total; total += static_cast<long long>(j.taps) * j.rows_pad * j.k_pad; jobs.push_back(j);
        }
    }
    upload_table(n, n->pack_jobs_dev, jobs.data(), jobs.size() * sizeof(PackJobHost));
    RUN(hd_pack_all_weights(n->pack_jobs_dev, static_cast<int>(jobs.size()), total, n->stream));
}

// Eval mode: every BatchNorm (running statistics) is folded to scale / shift once per forward, in ONE launch; the
// convolutions then apply it (+ residual addend, + ReLU) in their epilogue (hd_conv2d_igemm_affine), so a `Convolution`
// is one launch instead of conv + finalize + bn_act and the eval forward has ~50 launches instead of ~125.
struct BnFoldJobHost {
    const float* gamma; const float* beta; const float* mean; const float* var;
    float* out; int channels; float eps;
};
static void fold_bn(hd_net* n) {
    std::vector<BnFoldJobHost> jobs;
    for (size_t i = 0; i < n->units.size(); ++i) {
        Unit& u = n->units[i];
        if (!u.bn) continue;
        const hd_unit_ptrs& p = UP(n, static_cast<int>(i));
        jobs.push_back(BnFoldJobHost{p.gamma, p.beta, p.running_mean, p.running_var, u.bnp, u.cout, 1e-5f});
    }
    if (jobs.empty()) return;
    upload_table(n, n->fold_jobs_dev, jobs.data(), jobs.size() * sizeof(BnFoldJobHost));
    RUN(hd_bn_fold_all(nullptr, static_cast<int>(jobs.size()), n->fold_jobs_dev, n->stream));
}

static cudaEvent_t next_event(hd_net* n) {
    if (n->events.empty()) {
        n->events.resize(128);
        for (cudaEvent_t& e : n->events)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                n->rc = fail(HD_ERR_CUDA, "net: cudaEventCreate failed");
                return nullptr;
            }
    }
    return n->events[n->ev_next++ % n->events.size()];
}

// Event recorded on the CURRENT lane's stream (nullptr in dry mode).
static cudaEvent_t mark_ready(hd_net* n) {
    if (n->dry || n->rc != 0) return nullptr;
    cudaEvent_t e = next_event(n);
    if (e && cudaEventRecord(e, n->stream) != cudaSuccess) n->rc = fail(HD_ERR_CUDA, "net: cudaEventRecord failed");
    return e;
}

static void wait_on(hd_net* n, cudaStream_t waiter, cudaEvent_t e) {
    if (n->dry || n->rc != 0 || !e) return;
    if (cudaStreamWaitEvent(waiter, e, 0) != cudaSuccess) n->rc = fail(HD_ERR_CUDA, "net: cudaStreamWaitEvent failed");
}

// Make the other lane current (stream, backward stack arena, BN-backward scratch).
static void swap_lane(hd_net* n) {
    std::swap(n->stream, n->alt.stream);
    std::swap(n->bw, n->alt.bw);
    std::swap(n->small, n->alt.small);
}

// conv (+ bias) -> raw output + BN statistics; then finalize the BN of this unit
// training: y = raw conv output (+ bias, + addend), BN statistics / finalize fused; the caller applies the BN.
// eval, BN layer: y = relu?(bn(conv) + addend) in one launch (the caller must NOT apply the BN again).
static void conv_unit(hd_net* n, int ui, const bf16* x, bf16* y, int B, int H, int W, const bf16* addend, int training,
                      int eval_relu = 0) {
    Unit& u = n->units[ui];
    const hd_unit_ptrs& p = UP(n, ui);
    const int cin_gemm = u.kind == 1 ? 192 : pad64(u.cin);
    const int k = u.kind == 1 ? 1 : u.k;
    const bool stats = u.bn && training;
    u.npix = static_cast<long long>(B) * H * W;
    if (u.kind == 1 && stem_s2d()) {
        // x = the x-unfolded space-to-depth tensor (64 channels): four vertical taps, rows y-2 .. y+1
        hd_bn_fuse bn{};
        bn.gamma = p.gamma; bn.beta = p.beta; bn.running_mean = p.running_mean; bn.running_var = p.running_var;
        bn.num_batches_tracked = p.num_batches_tracked; bn.momentum = 0.1f; bn.eps = 1e-5f;
        bn.out = u.bnp; bn.counter = n->tickets + ui;
        if (stats)
            RUN(hd_conv2d_igemm_vtaps(x, u.wp, y, p.b, u.stats, u.stats + u.cout, B, H, W, 64, u.cout, 64, 4, 2, u.cout, &bn,
                                      nullptr, nullptr, 0, n->stream));
        else
            RUN(hd_conv2d_igemm_vtaps(x, u.wp, y, p.b, nullptr, nullptr, B, H, W, 64, u.cout, 64, 4, 2, u.cout, nullptr,
                                      u.bnp, u.bnp + u.cout, eval_relu, n->stream));
        return;
    }
    if (stats) {
        // train mode: the conv's last CTA finalizes the BN (scale/shift/mean/rstd + running statistics) itself
        hd_bn_fuse bn{};
        bn.gamma = p.gamma; bn.beta = p.beta; bn.running_mean = p.running_mean; bn.running_var = p.running_var;
        bn.num_batches_tracked = p.num_batches_tracked; bn.momentum = 0.1f; bn.eps = 1e-5f;
        bn.out = u.bnp; bn.counter = n->tickets + ui;
        RUN(hd_conv2d_igemm_bn(x, u.wp, y, nullptr, u.bias ? p.b : nullptr, addend, u.stats, u.stats + u.cout, B, H, W,
                               cin_gemm, u.cout, block_n_for(u.cout), k, 0, u.cout, 0, 0, 1, &bn, n->stream));
        return;
    }
    if (u.bn) {
        RUN(hd_conv2d_igemm_affine(x, u.wp, y, u.bias ? p.b : nullptr, addend, u.bnp, u.bnp + u.cout, eval_relu, B, H, W,
                                   cin_gemm, u.cout, block_n_for(u.cout), k, u.cout, n->stream));
        return;
    }
    RUN(hd_conv2d_igemm(x, u.wp, y, nullptr, u.bias ? p.b : nullptr, addend, nullptr, nullptr, B, H, W, cin_gemm, u.cout,
                        block_n_for(u.cout), k, 0, u.cout, 0, 0, 1, n->stream));
}

// pooled / pool_idx (training, skip-BN residual only): the block output is consumed by a 2x2 max pool and nothing else, so
// the tail kernel emits the pooled tensor + argmax directly and the un-pooled output is never materialised (r.Out = null).
static bf16* residual_fwd(hd_net* n, int ri, bf16* X, int B, int H, int W, int training, bf16* pooled = nullptr,
                          uint8_t* pool_idx = nullptr) {
    ResSaved& r = n->res[ri];
    r.X = X; r.H = H; r.W = W;
    const size_t bytes = act_bytes(B, H, W, r.cout);
    const long long npix = static_cast<long long>(B) * H * W;
    r.Y1 = reinterpret_cast<bf16*>(n->fw.alloc(bytes));
    r.Z1 = reinterpret_cast<bf16*>(n->fw.alloc(bytes));
    r.Y2 = reinterpret_cast<bf16*>(n->fw.alloc(bytes));
    r.Ys = r.us >= 0 ? reinterpret_cast<bf16*>(n->fw.alloc(bytes)) : nullptr;
    r.Out = pooled ? nullptr : reinterpret_cast<bf16*>(n->fw.alloc(bytes));
    static const bool no_mask = getenv("HD_NO_MASK_BITS") != nullptr;
    r.Mask = (training && r.us < 0 && !no_mask) ? reinterpret_cast<uint8_t*>(n->fw.alloc(static_cast<size_t>(npix) * r.cout / 8))
                                                 : nullptr;
    Unit &u1 = n->units[r.u1], &u2 = n->units[r.u2];
    if (!training) {
        // relu(bn1(conv1)) -> Z1 ; [bn_s(conv_s) -> Ys] ; relu(bn2(conv2) + skip) -> Out : 2-3 launches
        conv_unit(n, r.u1, X, r.Z1, B, H, W, nullptr, 0, 1);
        const bf16* skip = X;
        if (r.us >= 0) {
            conv_unit(n, r.us, X, r.Ys, B, H, W, nullptr, 0, 0);
            skip = r.Ys;
        }
        conv_unit(n, r.u2, r.Z1, r.Out, B, H, W, skip, 0, 1);
        return r.Out;
    }
    conv_unit(n, r.u1, X, r.Y1, B, H, W, nullptr, training);
    RUN(hd_bn_act(r.Y1, u1.bnp, u1.bnp + u1.cout, r.Z1, npix, r.cout, 1, n->stream));
    conv_unit(n, r.u2, r.Z1, r.Y2, B, H, W, nullptr, training);
    if (r.us >= 0) {
        Unit& us = n->units[r.us];
        conv_unit(n, r.us, X, r.Ys, B, H, W, nullptr, training);
        if (pooled) {
            RUN(hd_bn_add_relu_pool2(r.Y2, u2.bnp, u2.bnp + u2.cout, r.Ys, us.bnp, us.bnp + us.cout, pooled, pool_idx, B,
                                     H, W, r.cout, n->stream));
            return pooled;
        }
        RUN(hd_bn_add_relu(r.Y2, u2.bnp, u2.bnp + u2.cout, r.Ys, us.bnp, us.bnp + us.cout, r.Out, npix, r.cout,
                           n->stream));
    } else {
        RUN(hd_bn_add_relu_mask(r.Y2, u2.bnp, u2.bnp + u2.cout, X, nullptr, nullptr, r.Out, r.Mask, npix, r.cout,
                                n->stream));
    }
    return r.Out;
}

static bf16* hourglass_fwd(hd_net* n, int hi, bf16* x, int B, int H, int W, int training) {
    HgSaved& h = n->hgs[hi];
    h.x = x; h.H = H; h.W = W;
    const int C = n->in_ch;
    // up1 on the other lane (it only needs x), the spine on this one; join before the upsample-add
    cudaEvent_t x_ready = mark_ready(n);
    swap_lane(n);
    wait_on(n, n->stream, x_ready);
    bf16* up1 = residual_fwd(n, h.up1, x, B, H, W, training);
    cudaEvent_t up1_done = mark_ready(n);
    swap_lane(n);
    bf16* pool = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H / 2, W / 2, C)));
    RUN(hd_maxpool2(x, pool, B, H, W, C, n->stream));
    bf16* low1 = residual_fwd(n, h.low1, pool, B, H / 2, W / 2, training);
    bf16* low2 = h.low2_hg >= 0 ? hourglass_fwd(n, h.low2_hg, low1, B, H / 2, W / 2, training)
                                : residual_fwd(n, h.low2_res, low1, B, H / 2, W / 2, training);
    bf16* low3 = residual_fwd(n, h.low3, low2, B, H / 2, W / 2, training);
    h.out = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H, W, C)));
    wait_on(n, n->stream, up1_done);
    RUN(hd_upsample2_add(up1, low3, h.out, B, H, W, C, n->stream));
    return h.out;
}

static void forward_impl(hd_net* n, const float* x, float* logits, int B, int H, int W, int training) {
    const int C = n->in_ch;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
    cudaEvent_t packed = nullptr;
    if (!n->dry) {
        if (n->rc == 0 && cudaMemsetAsync(n->units[0].stats, 0, n->persist_bytes, n->stream) != cudaSuccess)
            n->rc = fail(HD_ERR_CUDA, "net_forward: memset of the BN statistics failed");
        if (n->rc == 0 && (cudaMemsetAsync(n->small, 0, 16 * 256 * sizeof(float), n->stream) != cudaSuccess ||
                           cudaMemsetAsync(n->alt.small, 0, 16 * 256 * sizeof(float), n->stream) != cudaSuccess))
            n->rc = fail(HD_ERR_CUDA, "net_forward: memset of the BN-backward scratch failed");
        const bool reuse = !training && n->static_weights && n->eval_packed && n->eval_packed_ws == n->persist.base;
        if (!reuse) {
            // weight packing (+ BN folding) on the second lane, beside the stem's unfold kernel, which needs no weights
            cudaEvent_t start = mark_ready(n);
            swap_lane(n);
            wait_on(n, n->stream, start);
            pack_weights(n, training != 0);
            if (!training) fold_bn(n);
            packed = mark_ready(n);
            swap_lane(n);
        }
        n->eval_packed = !training;
        n->eval_packed_ws = n->persist.base;
    }
    phase_mark(n, "fwd:start");
    // ---- PreLayer (hourglass.py:159-173)
    Unit& u0 = n->units[0];
    n->patches = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H2, W2, stem_s2d() ? 64 : 192)));
    n->Y0 = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H2, W2, 64)));
    n->Z0 = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H2, W2, 64)));
    if (stem_s2d()) RUN(hd_stem_unfold(x, n->patches, B, H, W, n->stream));
    else RUN(hd_stem_im2col(x, n->patches, B, H, W, n->stream));
    wait_on(n, n->stream, packed);      // the first convolution needs the packed weights
    if (training) {
        conv_unit(n, 0, n->patches, n->Y0, B, H2, W2, nullptr, training);
        RUN(hd_bn_act(n->Y0, u0.bnp, u0.bnp + 64, n->Z0, static_cast<long long>(B) * H2 * W2, 64, 1, n->stream));
    } else {
        conv_unit(n, 0, n->patches, n->Z0, B, H2, W2, nullptr, 0, 1);
    }
    phase_mark(n, "stem");
    static const bool no_pool_fuse = getenv("HD_NO_POOL_FUSE") != nullptr;
    const bool fuse_pool = training && !no_pool_fuse && n->res[n->r_pre1].us >= 0;
    bf16* r1 = nullptr;
    if (fuse_pool) {
        n->R1pool = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, 128)));
        n->R1idx = reinterpret_cast<uint8_t*>(n->fw.alloc(static_cast<size_t>(B) * H4 * W4 * 128));
        residual_fwd(n, n->r_pre1, n->Z0, B, H2, W2, training, n->R1pool, n->R1idx);
    } else {
        n->R1idx = nullptr;
        r1 = residual_fwd(n, n->r_pre1, n->Z0, B, H2, W2, training);
    }
    phase_mark(n, "pre1@256");
    if (!fuse_pool) {
        n->R1pool = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, 128)));
        RUN(hd_maxpool2(r1, n->R1pool, B, H2, W2, 128, n->stream));
    }
    bf16* r3 = residual_fwd(n, n->r_pre3, n->R1pool, B, H4, W4, training);
    bf16* xcur = residual_fwd(n, n->r_pre4, r3, B, H4, W4, training);
    phase_mark(n, "pool+pre3,4");
    // ---- stacks (hourglass.py:226-235)
    const long long npix4 = static_cast<long long>(B) * H4 * W4;
    for (int i = 0; i < n->S; ++i) {
        hd_net::StackSaved& s = n->stacks[i];
        s.x_in = xcur;
        {
            SmReserve lanes(training ? hg_sm_reserve() : 0);      // two lanes inside: see hd_common.h
            s.hg_out = hourglass_fwd(n, s.hg_root, xcur, B, H4, W4, training);
        }
        phase_mark(n, "hourglass");
        Unit& un = n->units[s.u_neck];
        s.Yn = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, C)));
        s.F1 = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, C)));
        if (training) {
            conv_unit(n, s.u_neck, s.hg_out, s.Yn, B, H4, W4, nullptr, training);
            RUN(hd_bn_act(s.Yn, un.bnp, un.bnp + C, s.F1, npix4, C, 1, n->stream));
        } else {
            conv_unit(n, s.u_neck, s.hg_out, s.F1, B, H4, W4, nullptr, 0, 1);
        }
        s.F2 = residual_fwd(n, s.neck_res, s.F1, B, H4, W4, training);
        Unit& uh = n->units[s.u_head];
        const bool merge = i < n->S - 1;
        s.pred64 = merge ? reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, 64))) : nullptr;
        if (merge && !n->dry && n->rc == 0 &&
            cudaMemsetAsync(s.pred64, 0, act_bytes(B, H4, W4, 64), n->stream) != cudaSuccess)
            n->rc = fail(HD_ERR_CUDA, "net_forward: memset failed");
        RUN(hd_conv2d_igemm(s.F2, uh.wp, logits, s.pred64, UP(n, s.u_head).b, nullptr, nullptr, nullptr, B, H4, W4, C,
                            n->out_ch, 16, 1, 1, 0, 64, i, n->S, n->stream));
        if (merge) {
            // x = x + merge_feature(feature) + merge_prediction(prediction)   (hourglass.py:235)
            s.T = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, C)));
            bf16* xn = reinterpret_cast<bf16*>(n->fw.alloc(act_bytes(B, H4, W4, C)));
            conv_unit(n, s.u_mf, s.F2, s.T, B, H4, W4, xcur, training);
            conv_unit(n, s.u_mp, s.pred64, xn, B, H4, W4, s.T, training);
            xcur = xn;
        }
        phase_mark(n, "neck+head");
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Weight gradient on the side stream, ordered after `ready` (the dY producer) of the main stream.
static void wgrad_unit(hd_net* n, int ui, const bf16* x, const bf16* dy, int B, int H, int W, cudaEvent_t ready) {
    Unit& u = n->units[ui];
    const hd_unit_ptrs& p = UP(n, ui);
    wait_on(n, n->side, ready);
    if (u.kind == 1 && stem_s2d())
        RUN(hd_conv2d_wgrad_sync(x, dy, p.dw, n->wgrad_ws, B, H, W, 64, 48, 64, 1, 0, 2, n->wgrad_sync, n->side));
    else if (u.kind == 1)
        RUN(hd_conv2d_wgrad_sync(x, dy, p.dw, n->wgrad_ws, B, H, W, 192, 147, 64, 1, 0, 1, n->wgrad_sync, n->side));
    else
        RUN(hd_conv2d_wgrad_sync(x, dy, p.dw, n->wgrad_ws, B, H, W, pad64(u.cin), u.cin, u.cout, u.k, 0, 0, n->wgrad_sync,
                                 n->side));
}

static void dgrad_unit(hd_net* n, int ui, const bf16* dy, bf16* dx, int B, int H, int W, const bf16* addend) {
    Unit& u = n->units[ui];
    RUN(hd_conv2d_igemm(dy, u.wpd, dx, nullptr, nullptr, addend, nullptr, nullptr, B, H, W, pad64(u.cout), u.cin,
                        block_n_for(u.cin), u.k, 0, u.cin, 0, 0, 1, n->stream));
}

// BN (+ReLU) backward of one unit: g = dout * (out > 0) -> dy (and the skip branch / g when requested)
// pool_idx != nullptr: `dout` is the gradient of the 2x2-POOLED block output and pool_idx the argmax the fused forward
// tail stored (two-branch tails only): the BN-backward kernels route it themselves (hd_bn_bwd_*_pool)
// The "last CTA finalizes" block of unit ui's BN backward on the CURRENT lane's scratch (n->small: sums [3][256] |
// coef [3][256] | coef_s [3][256] | ticket, epoch). `sums` and the ticket are zeroed once per forward pass and re-zeroed
// by every fused finalize after use.
static hd_bn_bwd_fuse bn_bwd_fin(hd_net* n, int ui, int us) {
    Unit& u = n->units[ui];
    const hd_unit_ptrs& p = UP(n, ui);
    const int C = u.cout;
    hd_bn_bwd_fuse fin{};
    fin.gamma = p.gamma; fin.mean = u.bnp + 2 * C; fin.rstd = u.bnp + 3 * C; fin.coef = n->small + 3 * 256;
    fin.dgamma = p.dgamma; fin.dbeta = p.dbeta;
    fin.count = static_cast<float>(u.npix);
    fin.counter = reinterpret_cast<unsigned int*>(n->small + 9 * 256);
    if (us >= 0) {
        const Unit& s = n->units[us];
        const hd_unit_ptrs& ps = UP(n, us);
        fin.gamma_s = ps.gamma; fin.mean_s = s.bnp + 2 * C; fin.rstd_s = s.bnp + 3 * C; fin.coef_s = n->small + 6 * 256;
        fin.dgamma_s = ps.dgamma; fin.dbeta_s = ps.dbeta;
    }
    return fin;
}

// maps up to this many pixels take the one-launch BN backward (bn_bwd_fused_small_kernel)
static long long bn_fused_small_max() {
    static const long long v = getenv("HD_BN_FUSED_SMALL_MAX") ? atoll(getenv("HD_BN_FUSED_SMALL_MAX")) : 32768;
    return v;
}

// presummed: the producer of `dout` (hd_conv2d_igemm_bwdstat) has already left this BN's coefficients and dgamma / dbeta
// in the lane's scratch - only the apply pass remains (plain conv + BN + ReLU units)
static void bn_bwd_unit(hd_net* n, int ui, const bf16* dout, const bf16* out, const bf16* y, bf16* dy, int us,
                        const bf16* ys, bf16* dys, bf16* gout, const uint8_t* mask = nullptr,
                        const uint8_t* pool_idx = nullptr, int pB = 0, int pH = 0, int pW = 0, bool presummed = false) {
    Unit& u = n->units[ui];
    const int C = u.cout;
    float* sums = n->small;            // [3][C]
    float* coef = n->small + 3 * 256;  // [3][C]
    float* coef_s = n->small + 6 * 256;
    const Unit* s = us >= 0 ? &n->units[us] : nullptr;
    // out == nullptr: plain conv+BN+ReLU unit, the mask is recomputed from y and this unit's scale/shift.
    // The reduction's last block also produces the coefficients and dgamma / dbeta (no separate finalize launch).
    hd_bn_bwd_fuse fin = bn_bwd_fin(n, ui, us);
    if (presummed) {
        RUN(hd_bn_bwd_apply(dout, nullptr, u.bnp, u.bnp + C, nullptr, nullptr, y, coef, dy, nullptr, nullptr, nullptr, nullptr,
                            u.npix, C, n->stream));
        return;
    }
    // two-branch tail (skip conv + BN): the ReLU mask is rebuilt from the two conv outputs the kernels read anyway
    // instead of reading the stored block output (saves a 537 MB read per kernel at 256x256)
    const float* sc_s = s ? s->bnp : nullptr;
    const float* sh_s = s ? s->bnp + C : nullptr;
    if (pool_idx) {
        if (!s) { if (n->rc == 0) n->rc = fail(HD_ERR_INVALID, "net: pooled BN backward needs a two-branch tail"); return; }
        RUN(hd_bn_bwd_reduce_pool_fin(dout, pool_idx, u.bnp, u.bnp + C, sc_s, sh_s, y, ys, sums, pB, pH, pW, C, &fin, n->stream));
        RUN(hd_bn_bwd_apply_pool(dout, pool_idx, u.bnp, u.bnp + C, sc_s, sh_s, y, ys, coef, coef_s, dy, dys, pB, pH, pW, C,
                                 n->stream));
        return;
    }
    // small maps (the deep hourglass levels): one launch instead of two, see bn_bwd_fused_small_kernel
    const long long fused_max = bn_fused_small_max();
    unsigned int* epoch = reinterpret_cast<unsigned int*>(n->small + 9 * 256) + 1;
    if (!s && u.npix <= fused_max && (mask || out == nullptr)) {
        RUN(hd_bn_bwd_fused_small(dout, mask, mask ? nullptr : u.bnp, mask ? nullptr : u.bnp + C, y, sums, dy,
                                  mask ? gout : nullptr, u.npix, C, &fin, epoch, n->stream));
        if (mask || gout == nullptr) return;
    }
    if (mask && !s) {     // single-BN residual tail with stored ReLU mask bits: `out` is not read at all
        RUN(hd_bn_bwd_reduce_fin_mask(dout, mask, y, sums, u.npix, C, &fin, n->stream));
        RUN(hd_bn_bwd_apply_mask(dout, mask, y, coef, dy, gout, u.npix, C, n->stream));
        return;
    }
    const bf16* mask_src = s ? nullptr : out;
    RUN(hd_bn_bwd_reduce_fin(dout, mask_src, u.bnp, u.bnp + C, sc_s, sh_s, y, ys, sums, u.npix, C, &fin, n->stream));
    RUN(hd_bn_bwd_apply(dout, mask_src, u.bnp, u.bnp + C, sc_s, sh_s, y, coef, dy, ys, s ? coef_s : nullptr, dys, gout,
                        u.npix, C, n->stream));
}

static void residual_bwd(hd_net* n, int ri, const bf16* dOut, bf16* dX, int B, const uint8_t* pool_idx = nullptr) {
    ResSaved& r = n->res[ri];
    const int H = r.H, W = r.W;
    const size_t mark = n->bw.off;
    const size_t bytes_o = act_bytes(B, H, W, r.cout), bytes_i = act_bytes(B, H, W, r.cin);
    // dY buffers are wgrad operands read asynchronously by the side stream: bump-only region, never reused
    bf16* dY2 = reinterpret_cast<bf16*>(n->wg.alloc(bytes_o));
    bf16* dYs = r.us >= 0 ? reinterpret_cast<bf16*>(n->wg.alloc(bytes_o)) : nullptr;
    bf16* dY1 = reinterpret_cast<bf16*>(n->wg.alloc(bytes_o));
    bf16* G = r.us >= 0 ? nullptr : reinterpret_cast<bf16*>(n->bw.alloc(bytes_o));
    bn_bwd_unit(n, r.u2, dOut, r.Out, r.Y2, dY2, r.us, r.Ys, dYs, G, r.Mask, pool_idx, B, H, W);
    bf16* dZ1 = reinterpret_cast<bf16*>(n->bw.alloc(bytes_o));
    // HD_DGRAD_BNSTAT=1 (opt-in, read per call so that one process can compare both): conv2's dgrad also reduces the
    // statistics of conv1's BN backward in its epilogue when it runs on the halo kernel (maps >= 16x16 with enough tiles)
    // and the map is too large for the one-launch BN backward; the reduction pass over dZ1 and Y1 disappears (-1.6 GB of
    // DRAM traffic per step, -0.32 ms of serialised kernel time). MEASURED NEUTRAL on the step (same box, graph replay:
    // 11.97 / 11.82 ms without, 11.76 / 11.90 with): the dgrad gets 25 us longer at 128x128 (112 -> 137 us; 427 -> 516 at
    // 256x256) and, unlike the HBM-bound reduction it replaces, a persistent tensor-core kernel cannot share the SMs with
    // the weight-gradient stream - the work moved from a kernel that overlapped into one that does not.
    const char* bnstat_env = getenv("HD_DGRAD_BNSTAT");
    Unit& uc2 = n->units[r.u2];
    const bool bnstat = bnstat_env != nullptr && bnstat_env[0] == '1' && uc2.cin == 128 && uc2.cout == 128 && n->units[r.u1].npix > bn_fused_small_max() &&
                        hd_conv2d_igemm_halo_eligible(B, H, W, 128, uc2.k) == 1;
    if (bnstat) {
        Unit& uc1 = n->units[r.u1];
        hd_bn_bwd_fuse fin = bn_bwd_fin(n, r.u1, -1);
        RUN(hd_conv2d_igemm_bwdstat(dY2, uc2.wpd, dZ1, B, H, W, pad64(uc2.cout), uc2.cin, uc2.k, r.Y1, uc1.bnp, uc1.bnp + uc1.cout,
                                    n->small, &fin, n->stream));
    } else {
        dgrad_unit(n, r.u2, dY2, dZ1, B, H, W, nullptr);
    }
    // The readiness event is recorded AFTER the dgrad launch on purpose: dgrad and wgrad are both persistent
    // tensor-core kernels that cannot share an SM, so the wgrad should start when the dgrad ends - exactly when the
    // HBM-bound BN-backward kernels below start on the main stream and can co-run with it.
    cudaEvent_t e2 = mark_ready(n);
    wgrad_unit(n, r.u2, r.Z1, dY2, B, H, W, e2);
    if (r.us >= 0) wgrad_unit(n, r.us, r.X, dYs, B, H, W, e2);
    bn_bwd_unit(n, r.u1, dZ1, nullptr, r.Y1, dY1, -1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, bnstat);
    if (r.us < 0) {
        dgrad_unit(n, r.u1, dY1, dX, B, H, W, G);
    } else {
        // dX = dgrad3x3(dY1) + dgrad1x1(dYs): one launch when the 64-output-channel halo kernel applies (the 256x256
        // level, where this pair was 0.68 ms of the step), else two with the first result as the second's addend
        Unit &u1 = n->units[r.u1], &us = n->units[r.us];
        static const bool no_dual = getenv("HD_NO_DUAL_DGRAD") != nullptr;
        const long long tiles16 = static_cast<long long>((W + 15) / 16) * ((H + 15) / 16) * B;
        if (!no_dual && u1.cin == 64 && us.cin == 64 && H >= 16 && W >= 16 && tiles16 >= hd_sm_count()) {
            // this 0.45 ms persistent kernel owns every SM it runs on (all registers): leaving a few SMs out of its grid
            // lets the weight-gradient stream's small kernels (the split-K reduction queued right now) get through
            static const int n64_reserve = getenv("HD_N64_RESERVE") ? atoi(getenv("HD_N64_RESERVE")) : 0;
            SmReserve room(n64_reserve);
            RUN(hd_conv2d_igemm_dual(dY1, u1.wpd, dYs, us.wpd, dX, nullptr, B, H, W, pad64(u1.cout), pad64(us.cout), u1.cin,
                                     64, 3, u1.cin, n->stream));
        } else {
            bf16* dXa = reinterpret_cast<bf16*>(n->bw.alloc(bytes_i));
            dgrad_unit(n, r.u1, dY1, dXa, B, H, W, nullptr);
            dgrad_unit(n, r.us, dYs, dX, B, H, W, dXa);
        }
    }
    cudaEvent_t e1 = mark_ready(n);
    wgrad_unit(n, r.u1, r.X, dY1, B, H, W, e1);         // overlaps the next block's BN backward
    n->bw.off = mark;
}

static void hourglass_bwd(hd_net* n, int hi, const bf16* dOut, bf16* dX, const bf16* extra_add, int B) {
    HgSaved& h = n->hgs[hi];
    const int H = h.H, W = h.W, C = n->in_ch;
    const size_t mark = n->bw.off;
    const size_t half = act_bytes(B, H / 2, W / 2, C), full = act_bytes(B, H, W, C);
    // up1's backward (needs only dOut) on the other lane; its result d_xa is consumed by this lane's max-pool backward
    bf16* d_xa = reinterpret_cast<bf16*>(n->bw.alloc(full));
    cudaEvent_t dout_ready = mark_ready(n);
    swap_lane(n);
    wait_on(n, n->stream, dout_ready);
    residual_bwd(n, h.up1, dOut, d_xa, B);
    cudaEvent_t up1_done = mark_ready(n);
    swap_lane(n);
    bf16* d_low3 = reinterpret_cast<bf16*>(n->bw.alloc(half));
    RUN(hd_sum2x2(dOut, d_low3, B, H, W, C, n->stream));
    bf16* d_low2 = reinterpret_cast<bf16*>(n->bw.alloc(half));
    residual_bwd(n, h.low3, d_low3, d_low2, B);
    bf16* d_low1 = d_low3;  // d_low3 is dead
    if (h.low2_hg >= 0) hourglass_bwd(n, h.low2_hg, d_low2, d_low1, nullptr, B);
    else residual_bwd(n, h.low2_res, d_low2, d_low1, B);
    bf16* d_pool = d_low2;  // dead as well (a nested level has joined its up1 lane before returning)
    residual_bwd(n, h.low1, d_low1, d_pool, B);
    wait_on(n, n->stream, up1_done);
    RUN(hd_maxpool2_bwd(h.x, d_pool, d_xa, extra_add, dX, B, H, W, C, n->stream));
    n->bw.off = mark;
}

// stages: bit 0 = the stacks (head / neck / hourglass of every stack, last to first) and the two 128x128 Residuals of
// PreLayer, bit 1 = the 256x256 level of PreLayer (pool backward, Residual(64,128), stem) + stream join.
// Every parameter gradient of stage 1 - 96 % of the flat gradient buffer for one stack - is complete (on the side stream
// for the weights, on the caller's stream for BN / bias gradients) when stage 1 has been enqueued, ~3 ms before the
// backward pass ends: that is where the data-parallel exchange of that bucket starts (parallel.py).
static void backward_impl(hd_net* n, const float* dlogits, int stages = 3) {
    const int B = n->B, C = n->in_ch;
    const int H2 = n->H / 2, W2 = n->W / 2, H4 = n->H / 4, W4 = n->W / 4;
    const size_t full4 = act_bytes(B, H4, W4, C);
    const long long hw4 = static_cast<long long>(H4) * W4;
    const long long npix4 = static_cast<long long>(B) * hw4;
    bf16* dXn = nullptr;  // gradient w.r.t. the input of stack i+1
    if (stages & 1) {
    phase_mark(n, "(loss)bwd:start");
    for (int i = n->S - 1; i >= 0; --i) {
        hd_net::StackSaved& s = n->stacks[i];
        const bool merge = i < n->S - 1;
        bf16* dpred = nullptr;
        bf16* dF2m = nullptr;
        if (merge) {
            const hd_unit_ptrs& pmf = UP(n, s.u_mf);
            const hd_unit_ptrs& pmp = UP(n, s.u_mp);
            RUN(hd_colsum(dXn, pmp.db, npix4, C, C, n->stream));
            RUN(hd_colsum(dXn, pmf.db, npix4, C, C, n->stream));
            cudaEvent_t em = mark_ready(n);
            wgrad_unit(n, s.u_mp, s.pred64, dXn, B, H4, W4, em);
            wgrad_unit(n, s.u_mf, s.F2, dXn, B, H4, W4, em);
            dpred = reinterpret_cast<bf16*>(n->bw.alloc(act_bytes(B, H4, W4, 16)));
            Unit& ump = n->units[s.u_mp];
            RUN(hd_conv2d_igemm(dXn, ump.wpd, dpred, nullptr, nullptr, nullptr, nullptr, nullptr, B, H4, W4, C,
                                n->out_ch, 16, 1, 0, 16, 0, 0, 1, n->stream));
        }
        // head
        bf16* dF2 = reinterpret_cast<bf16*>(n->bw.alloc(full4));
        const hd_unit_ptrs& ph = UP(n, s.u_head);
        RUN(hd_head_backward(dlogits + static_cast<long long>(i) * n->out_ch * hw4,
                             static_cast<long long>(n->S) * n->out_ch * hw4, dpred, 16, s.F2, n->units[s.u_head].wp,
                             dF2, ph.dw, ph.db, B, H4, W4, n->out_ch, n->stream));
        if (merge) {
            dF2m = reinterpret_cast<bf16*>(n->bw.alloc(full4));
            dgrad_unit(n, s.u_mf, dXn, dF2m, B, H4, W4, dF2);
            dF2 = dF2m;
        }
        // neck: Residual, then conv1x1 + bias + BN + ReLU
        bf16* dF1 = reinterpret_cast<bf16*>(n->bw.alloc(full4));
        residual_bwd(n, s.neck_res, dF2, dF1, B);
        bf16* dYn = reinterpret_cast<bf16*>(n->bw.alloc(full4));
        bn_bwd_unit(n, s.u_neck, dF1, nullptr, s.Yn, dYn, -1, nullptr, nullptr, nullptr);
        // The neck (and stem) conv bias is followed by a train-mode BatchNorm: its gradient is the
        // column sum of dYn, which BN backward makes exactly zero (sum of a*g + b*y + c over the batch vanishes). The
        // reference's fp32 autograd returns ~1e-7 there; a bf16 column sum would return ~1e-3 of noise, so the
        // caller-zeroed gradient is left untouched (HD_BIAS_COLSUM=1 restores the explicit reduction).
        static const bool bias_colsum = getenv("HD_BIAS_COLSUM") != nullptr;
        if (bias_colsum) RUN(hd_colsum(dYn, UP(n, s.u_neck).db, npix4, C, C, n->stream));
        bf16* dHg = dF1;  // dead
        dgrad_unit(n, s.u_neck, dYn, dHg, B, H4, W4, nullptr);
        cudaEvent_t en = mark_ready(n);
        wgrad_unit(n, s.u_neck, s.hg_out, dYn, B, H4, W4, en);
        bf16* dXi = reinterpret_cast<bf16*>(n->bw.alloc(full4));
        phase_mark(n, "head+neck bwd");
        {
            SmReserve lanes(hg_sm_reserve());
            hourglass_bwd(n, s.hg_root, dHg, dXi, merge ? dXn : nullptr, B);
        }
        phase_mark(n, "hourglass bwd");
        dXn = dXi;
    }
    // PreLayer, 128x128 part (the two Residuals behind the pool): still stage 1, so that the early gradient bucket covers
    // everything but the 256x256 level (stem + Residual(64,128): 0.2 M of the 4.98 M parameters)
    bf16* dR3 = reinterpret_cast<bf16*>(n->bw.alloc(full4));
    residual_bwd(n, n->r_pre4, dXn, dR3, B);
    bf16* dP1 = reinterpret_cast<bf16*>(n->bw.alloc(full4));
    residual_bwd(n, n->r_pre3, dR3, dP1, B);
    phase_mark(n, "pre4,3 bwd");
    n->dX_pre = dP1;
    }   // stage 1
    if (!(stages & 2)) return;
    bf16* dP = n->dX_pre;
    // PreLayer, 256x256 part
    // The pool's backward is folded into the BN backward of the block before it when that block's tail was fused with the
    // pool in the forward pass: the routed gradient (537 MB at 256x256, B = 32) is then neither written nor read
    // (HD_NO_POOL_BWD_FUSE=1: the separate hd_maxpool2_bwd_idx launch of the first version)
    static const bool no_pool_bwd_fuse = getenv("HD_NO_POOL_BWD_FUSE") != nullptr;
    const bool fold_pool = n->R1idx != nullptr && !no_pool_bwd_fuse;
    bf16* dR1 = nullptr;
    if (!fold_pool) {
        dR1 = reinterpret_cast<bf16*>(n->bw.alloc(act_bytes(B, H2, W2, 128)));
        if (n->R1idx) RUN(hd_maxpool2_bwd_idx(n->R1idx, dP, nullptr, nullptr, dR1, B, H2, W2, 128, n->stream));
        else RUN(hd_maxpool2_bwd(n->res[n->r_pre1].Out, dP, nullptr, nullptr, dR1, B, H2, W2, 128, n->stream));
    }
    phase_mark(n, "pool bwd");
    bf16* dZ0 = reinterpret_cast<bf16*>(n->bw.alloc(act_bytes(B, H2, W2, 64)));
    residual_bwd(n, n->r_pre1, fold_pool ? dP : dR1, dZ0, B, fold_pool ? n->R1idx : nullptr);
    phase_mark(n, "pre1@256 bwd");
    bf16* dY0 = reinterpret_cast<bf16*>(n->bw.alloc(act_bytes(B, H2, W2, 64)));
    bn_bwd_unit(n, 0, dZ0, nullptr, n->Y0, dY0, -1, nullptr, nullptr, nullptr);
    cudaEvent_t e0 = mark_ready(n);
    static const bool bias_colsum0 = getenv("HD_BIAS_COLSUM") != nullptr;     // see the neck above
    if (bias_colsum0) RUN(hd_colsum(dY0, UP(n, 0).db, static_cast<long long>(B) * H2 * W2, 64, 64, n->stream));
    wgrad_unit(n, 0, n->patches, dY0, B, H2, W2, e0);
    // join: everything the side stream produced (all weight gradients) is ordered before whatever follows on `stream`
    if (!n->dry && n->rc == 0) {
        cudaEvent_t done = next_event(n);
        phase_mark(n, "stem bwd");
        if (!done || cudaEventRecord(done, n->side) != cudaSuccess || cudaStreamWaitEvent(n->stream, done, 0) != cudaSuccess)
            n->rc = fail(HD_ERR_CUDA, "net_backward: stream join failed");
        phase_mark(n, "wgrad tail");
        phase_report(n);
    }
}

// ------------------------------------------------------------------------------------------------ planning + entry points
static size_t max_wgrad_ws(int B, int H, int W) {
    size_t m = 0;
    const int H2 = H / 2, W2 = W / 2;
    size_t v = hd_conv2d_wgrad_workspace_bytes(B, H2, W2, 192, 1); if (v > m) m = v;
    v = hd_conv2d_wgrad_workspace_bytes(B, H2, W2, 64, 1); if (v > m) m = v;      // space-to-depth stem: 4 taps x 64
    v = hd_conv2d_wgrad_workspace_bytes(B, H2, W2, 128, 3); if (v > m) m = v;
    for (int h = H / 4, w = W / 4; h >= 1 && w >= 1; h /= 2, w /= 2) {
        v = hd_conv2d_wgrad_workspace_bytes(B, h, w, 128, 3);
        if (v > m) m = v;
    }
    return m;
}

static void plan(hd_net* n, uint8_t* base, size_t cap, int B, int H, int W, bool with_backward, size_t* total) {
    // persistent | wgrad scratch | forward | backward
    n->persist = Arena(); n->fw = Arena(); n->bw = Arena();
    n->persist.base = base; n->persist.cap = cap;
    plan_persistent(n);
    size_t off = (n->persist.peak + 255) & ~size_t(255);
    n->wgrad_ws_bytes = with_backward ? max_wgrad_ws(B, H, W) : 0;
    n->wgrad_ws = base ? base + off : nullptr;
    off += (n->wgrad_ws_bytes + 255) & ~size_t(255);
    n->fw.base = base ? base + off : nullptr;
    n->fw.cap = cap > off ? cap - off : 0;
    *total = off;
}

extern "C" size_t hd_net_workspace_bytes(hd_net* n, int B, int H, int W, int with_backward) {
    if (B <= 0 || H <= 0 || W <= 0 || H % 64 || W % 64) return 0;
    size_t head = 0;
    n->dry = true; n->rc = 0; n->up = nullptr;
    plan(n, nullptr, 0, B, H, W, with_backward != 0, &head);
    n->B = B; n->H = H; n->W = W;
    forward_impl(n, nullptr, nullptr, B, H, W, 1);
    size_t total = head + ((n->fw.peak + 255) & ~size_t(255));
    if (with_backward) {
        n->bw = Arena();
        n->alt.bw = Arena();
        n->wg = Arena();
        backward_impl(n, nullptr);
        total += ((n->bw.peak + 255) & ~size_t(255)) + ((n->alt.bw.peak + 255) & ~size_t(255)) +
                 ((n->wg.peak + 255) & ~size_t(255));
    }
    n->dry = false;
    n->trained_fwd = false;
    return total + 4096;
}

extern "C" int hd_net_forward(hd_net* n, const hd_unit_ptrs* units, int n_units, const float* x, float* logits,
                              void* workspace, size_t workspace_bytes, int B, int H, int W, int training,
                              hd_stream_t stream) {
    HD_REQUIRE(n && units && x && logits && workspace, "net_forward: null argument");
    HD_REQUIRE(n_units == static_cast<int>(n->units.size()), "net_forward: %d units given, the network has %d", n_units,
               static_cast<int>(n->units.size()));
    HD_REQUIRE(B > 0 && H > 0 && W > 0 && H % 64 == 0 && W % 64 == 0,
               "net_forward: input (%d,3,%d,%d): H and W must be positive multiples of 64", B, H, W);
    HD_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "net_forward: workspace must be 256-byte aligned");
    size_t head = 0;
    static const int train_pdl = getenv("HD_TRAIN_PDL") ? atoi(getenv("HD_TRAIN_PDL")) : 0;
    // see hd_common.h: all launches in eval; training: off (HD_TRAIN_PDL=2 enables it for small grids only - measured neutral)
    PdlScope pdl(training == 0 ? 1 : train_pdl);
    n->dry = false; n->rc = 0; n->up = units; n->stream = stream;
    {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &st);
        if (st == cudaStreamCaptureStatusNone) {      // resources are created by eager calls only
            if (!n->pinned && cudaHostAlloc(reinterpret_cast<void**>(&n->pinned), kPinnedSlots * kPinnedSlotBytes,
                                            cudaHostAllocDefault) != cudaSuccess)
                return fail(HD_ERR_CUDA, "net_forward: cannot allocate the pinned job-table slots");
            if (n->events.empty()) next_event(n);
        }
    }
    if (!n->alt_stream && cudaStreamCreateWithFlags(&n->alt_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(HD_ERR_CUDA, "net_forward: cannot create the second lane's stream");
    static const bool single_lane = getenv("HD_SINGLE_LANE") != nullptr;   // debug knob: everything on one stream
    n->alt.stream = single_lane ? stream : n->alt_stream;
    if (workspace != n->last_ws || workspace_bytes != n->last_ws_bytes) {     // another arena: nothing is known about its contents
        n->uploaded.clear();
        n->last_ws = workspace; n->last_ws_bytes = workspace_bytes;
    }
    plan(n, reinterpret_cast<uint8_t*>(workspace), workspace_bytes, B, H, W, training != 0, &head);
    n->B = B; n->H = H; n->W = W;
    // capacity check with a dry pass first (cheap: pointer arithmetic only)
    {
        n->dry = true;
        Arena keep = n->fw;
        forward_impl(n, nullptr, nullptr, B, H, W, training);
        const size_t need = head + n->fw.peak;
        n->fw = keep;
        n->dry = false;
        HD_REQUIRE(need <= workspace_bytes, "net_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    }
    forward_impl(n, x, logits, B, H, W, training);
    n->trained_fwd = training != 0 && n->rc == 0;
    // backward region starts after the forward region
    const size_t fw_end = head + ((n->fw.peak + 255) & ~size_t(255));
    n->wg = Arena();
    n->wg.base = reinterpret_cast<uint8_t*>(workspace) + fw_end;
    n->wg.cap = workspace_bytes > fw_end ? workspace_bytes - fw_end : 0;
    return n->rc;
}

static int backward_begin(hd_net* n, const hd_unit_ptrs* units, int n_units, const float* dlogits, void* workspace,
                          size_t workspace_bytes, hd_stream_t stream) {
    HD_REQUIRE(n && units && dlogits && workspace, "net_backward: null argument");
    HD_REQUIRE(n_units == static_cast<int>(n->units.size()), "net_backward: unit table size mismatch");
    HD_REQUIRE(n->trained_fwd, "net_backward: no training-mode forward pass is pending on this network");
    HD_REQUIRE(n->bwd_stage == 0, "net_backward: a staged backward pass is already in flight (call stage 2 first)");
    HD_REQUIRE(reinterpret_cast<uint8_t*>(workspace) == n->persist.base, "net_backward: workspace moved since the forward pass");
    n->up = units; n->stream = stream; n->rc = 0;
    if (!n->side) {
        // The weight-gradient stream gets the highest priority: its kernels (wgrad, and above all the tiny split-K
        // reduce that follows each) share the machine with the main stream's HBM-bound kernels, whose thousands of
        // queued CTAs otherwise starve them of SM slots (trace: a 12 us wgrad_reduce took 280-470 us at 256x256).
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        static const bool flat = getenv("HD_FLAT_PRIORITY") != nullptr;
        if (cudaStreamCreateWithPriority(&n->side, cudaStreamNonBlocking, flat ? lo : hi) != cudaSuccess)
            return fail(HD_ERR_CUDA, "net_backward: cannot create the weight-gradient stream");
    }
    static const bool serial = getenv("HD_SERIAL_WGRAD") != nullptr;   // debug knob: wgrad on the main stream
    n->side_keep = n->side;
    if (serial) n->side = stream;
    if (!n->alt_stream && cudaStreamCreateWithFlags(&n->alt_stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(HD_ERR_CUDA, "net_backward: cannot create the second lane's stream");
    static const bool single_lane = getenv("HD_SINGLE_LANE") != nullptr;
    n->alt.stream = single_lane ? stream : n->alt_stream;
    // capacity check (dry) then run: [wg: bump-only dY operands][bw: this lane's stack][alt.bw: the up1 lane's stack]
    uint8_t* region = n->wg.base;
    const size_t region_cap = n->wg.cap;
    n->dry = true;
    n->wg = Arena(); n->bw = Arena(); n->alt.bw = Arena();
    backward_impl(n, nullptr);
    const size_t wg_need = (n->wg.peak + 255) & ~size_t(255), bw_need = (n->bw.peak + 255) & ~size_t(255),
                 alt_need = n->alt.bw.peak;
    n->dry = false;
    n->wg = Arena(); n->wg.base = region; n->wg.cap = region_cap;
    n->bw = Arena(); n->bw.base = region + wg_need; n->bw.cap = bw_need;
    n->alt.bw = Arena(); n->alt.bw.base = region + wg_need + bw_need; n->alt.bw.cap = alt_need;
    if (wg_need + bw_need + alt_need > region_cap) {
        n->side = n->side_keep;
        return fail(HD_ERR_INVALID, "net_backward: workspace too small for the backward pass (%zu < %zu bytes)", region_cap,
                    wg_need + bw_need + alt_need);
    }
    (void)workspace_bytes;
    return HD_OK;
}

static void backward_end(hd_net* n) {
    n->side = n->side_keep;
    n->trained_fwd = false;
    n->bwd_stage = 0;
}

extern "C" int hd_net_backward(hd_net* n, const hd_unit_ptrs* units, int n_units, const float* dlogits,
                               void* workspace, size_t workspace_bytes, hd_stream_t stream) {
    static const int train_pdl = getenv("HD_TRAIN_PDL") ? atoi(getenv("HD_TRAIN_PDL")) : 0;
    PdlScope pdl(train_pdl);
    int rc = backward_begin(n, units, n_units, dlogits, workspace, workspace_bytes, stream);
    if (rc) return rc;
    backward_impl(n, dlogits, 3);
    backward_end(n);
    return n->rc;
}

// The same pass in two calls, for callers that start the data-parallel exchange of the stacks' gradients while the
// PreLayer backward (the 256x256 level: ~3 ms of a 12 ms step) is still running. Stage 1 enqueues the stacks and, when
// `comm_stream` is given, makes it wait for everything that produced their parameter gradients (the caller's stream and
// the internal weight-gradient stream) - the caller then enqueues its collective on `comm_stream`. Stage 2 enqueues the
// PreLayer and joins the internal streams into `stream`.
extern "C" int hd_net_backward_stage(hd_net* n, const hd_unit_ptrs* units, int n_units, const float* dlogits,
                                     void* workspace, size_t workspace_bytes, hd_stream_t stream, int stage,
                                     hd_stream_t comm_stream) {
    static const int train_pdl = getenv("HD_TRAIN_PDL") ? atoi(getenv("HD_TRAIN_PDL")) : 0;
    PdlScope pdl(train_pdl);
    HD_REQUIRE(stage == 1 || stage == 2, "net_backward_stage: stage=%d", stage);
    if (stage == 1) {
        int rc = backward_begin(n, units, n_units, dlogits, workspace, workspace_bytes, stream);
        if (rc) return rc;
        backward_impl(n, dlogits, 1);
        if (n->rc == 0 && comm_stream) {
            cudaEvent_t em = next_event(n), es = next_event(n);
            if (!em || !es || cudaEventRecord(em, n->stream) != cudaSuccess || cudaEventRecord(es, n->side) != cudaSuccess ||
                cudaStreamWaitEvent(comm_stream, em, 0) != cudaSuccess || cudaStreamWaitEvent(comm_stream, es, 0) != cudaSuccess)
                n->rc = fail(HD_ERR_CUDA, "net_backward_stage: cannot order the communication stream");
        }
        if (n->rc != 0) { backward_end(n); return n->rc; }
        n->bwd_stage = 1;
        n->comm_overlap = comm_stream != nullptr;
        return HD_OK;
    }
    HD_REQUIRE(n && n->bwd_stage == 1, "net_backward_stage: stage 2 without a pending stage 1");
    HD_REQUIRE(stream == n->stream && units == n->up, "net_backward_stage: stage 2 must use the stream / unit table of stage 1");
    n->rc = 0;
    {
        // The collective of the stacks' bucket runs while this stage executes. Its kernel needs a few SMs; the persistent
        // convolution / weight-gradient grids own every SM they run on, so without room the collective only advances
        // between them (measured: overlap == no overlap at N = 8). HD_COMM_RESERVE leaves that many SMs out of their grids.
        // Measured at N = 8 on two boxes (profiles/r02_allreduce_timing.txt): box A flat 13.89 / overlap 13.76 / + 8 SMs 13.54 /
        // + 16 SMs 13.60 ms; box B (early bucket widened to 96 %) flat 13.63 / overlap 13.56 / + 8 SMs 13.75 - no robust
        // gain from the reservation, so the default is 0.
        static const int comm_reserve = getenv("HD_COMM_RESERVE") ? atoi(getenv("HD_COMM_RESERVE")) : 0;
        SmReserve room(n->comm_overlap ? comm_reserve : 0);
        backward_impl(n, dlogits, 2);
    }
    n->comm_overlap = false;
    backward_end(n);
    return n->rc;
}
