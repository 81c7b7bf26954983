/* hd_b200.h — C ABI of libhd_b200.so, the B200-native (sm_100a) hot path of
 * tyui592/Real_Time_Helmet_Detection.
 *
 * The reference has no FFI / plugin layer of its own: its boundary is a set of Python call signatures
 * (hourglass.py:198-237 StackedHourglass, loss.py:6-40 LossCalculator, transform.py:73 hm2box,
 * evaluate.py:114-182 Prediction). This header is the native boundary underneath our drop-in Python
 * mirrors of those signatures; every entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - activations: NHWC bf16 ("nhwc"), channel count a multiple of 8; images / logits / targets: NCHW fp32;
 *   - every call only ENQUEUES work on `stream`: no allocation, no synchronisation, no host read-back;
 *     outputs and workspaces are allocated by the caller (torch caching allocator in the Python host);
 *   - return value 0 on success, negative errno-style code otherwise (-22 invalid argument, -5 CUDA error,
 *     -38 unsupported); hd_last_error() returns a thread-local message for the last failure;
 *   - there is no CPU fallback: without a CUDA device the compute entry points fail with -5.
 */
#ifndef HD_B200_H
#define HD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* hd_stream_t; /* == cudaStream_t */

const char* hd_last_error(void);
int hd_version(void);
/* Kernels launched by this library in this process so far (instrumentation for bench.py's gpu_launches). */
long long hd_launch_count(void);
/* Programmatic dependent launch (every kernel is launched with the programmatic-stream-serialization attribute and
 * starts with griddepcontrol.launch_dependents / griddepcontrol.wait): 1 = on (default), 0 = off. HD_NO_PDL=1 in the
 * environment also turns it off. */
void hd_set_pdl(int on);
/* Profiling aid: with HD_TRACE=1 in the environment every launch is bracketed by timing events; this prints one
 * "[hd_trace] s<stream> start_ms end_ms duration kernel" row per launch since the last dump to stderr (syncs the device). */
void hd_trace_dump(void);

/* ------------------------------------------------------------------ convolutions (hourglass.py:94-108) */

/* Implicit-GEMM conv, stride 1, "same" padding, on tcgen05 tensor cores. Replaces nn.Conv2d (hourglass.py:100)
 * forward for the 3x3 / 1x1 convs and, with mode-1 packed weights, its input gradient (dgrad).
 *   x        nhwc [N,H,W,cin], cin % 64 == 0
 *   w_packed bf16 [ksize*ksize][block_n][cin] from hd_pack_conv_weight / hd_stem_pack_weight
 *   out      out_mode 0: nhwc bf16, channel stride out_cs ; out_mode 1 (block_n 16 only): fp32 NCHW slice
 *            [:, stack_idx] of a (N, num_stack, cout, H, W) logits tensor (hourglass.py:237 torch.stack)
 *   out2     optional (block_n 16): second nhwc bf16 copy, channel stride out2_cs, channels [cout,16) zeroed
 *   bias     optional fp32 [cout] ; addend optional nhwc bf16 like `out` (fused residual / merge add, :235)
 *   stat_sum / stat_sqsum  optional fp32 [cout]: per-channel sum / sum of squares of the fp32 result are
 *            ACCUMULATED (train-mode BatchNorm statistics, hourglass.py:103)
 *   block_n  128, 64 or 16 (>= cout) */
int hd_conv2d_igemm(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                    const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin, int cout,
                    int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx, int num_stack,
                    hd_stream_t stream);

/* Optional fusion of the train-mode BatchNorm finalize (hourglass.py:103) into the convolution: the last CTA to
 * flush its statistics computes out[0..4*cout) = scale | shift | mean | rstd, updates the running statistics
 * (momentum, unbiased variance) and num_batches_tracked. `counter` is a zero-initialised device word the kernel
 * leaves at zero again. */
typedef struct hd_bn_fuse {
    const float* gamma; const float* beta;
    float* running_mean; float* running_var; long long* num_batches_tracked;
    float momentum, eps;
    float* out;
    unsigned int* counter;
} hd_bn_fuse;
int hd_conv2d_igemm_bn(const void* x, const void* w_packed, void* out, void* out2, const float* bias,
                       const void* addend, float* stat_sum, float* stat_sqsum, int N, int H, int W, int cin, int cout,
                       int block_n, int ksize, int out_mode, int out_cs, int out2_cs, int stack_idx, int num_stack,
                       const hd_bn_fuse* bn, hd_stream_t stream);

/* Same convolution with a per-channel affine (+ residual addend) (+ ReLU) epilogue - the eval-mode `Convolution`
 * (hourglass.py:105-108 with BatchNorm on running statistics) and the `Residual` tail (hourglass.py:123-127) in ONE
 * launch:  out = relu?( (conv + bias) * scale[c] + shift[c] + addend ),  NHWC bf16 output with channel stride out_cs.
 * scale / shift: fp32 [cout] or both NULL; block_n 128 or 64. */
int hd_conv2d_igemm_affine(const void* x, const void* w_packed, void* out, const float* bias, const void* addend,
                           const float* scale, const float* shift, int relu, int N, int H, int W, int cin, int cout,
                           int block_n, int ksize, int out_cs, hd_stream_t stream);
/* Eval-mode BatchNorm folded to scale / shift for a whole table of layers in one launch. jobs_host: njobs records
 * {const float* gamma, beta, running_mean, running_var; float* out; int channels; float eps} (out[0..C) = scale =
 * gamma * rsqrt(var + eps), out[C..2C) = shift = beta - mean * scale); jobs_dev: device scratch of the same size
 * (jobs_host may be NULL when the table is already in jobs_dev). */
int hd_bn_fold_all(const void* jobs_host, int njobs, void* jobs_dev, hd_stream_t stream);

/* Tuning / test knob for hd_conv2d_igemm: 0 = automatic choice (default), 1 = generic kernel only, 2 = use the
 * halo kernel (3x3, 128 output channels, map >= 16x16) whenever the shape is eligible. Applies to the CALLING THREAD
 * only (like hd_last_error). */
void hd_set_conv_variant(int variant);
/* Profiling only (results are wrong when non-zero): 1 = epilogue drains TMEM but skips math/stores, 2 = MMA issue
 * skipped (halo kernel). */
void hd_set_conv_debug(int mode);

/* Weight gradient of the same convs (autograd of hourglass.py:100): grad_w (OIHW fp32 [cout][cin_real][k][k])
 * = (accumulate ? grad_w : 0) + sum_pixels dy[p, co] * x[p + tap, ci].  x nhwc [.., cin], dy nhwc [.., cout],
 * cout in {64,128}, cin in {64,128,192(k=1)}. workspace: hd_conv2d_wgrad_workspace_bytes() bytes.
 * stem_perm: K index is the stem im2col order, grad_w is [cout][3][7][7] (hourglass.py:163). */
int hd_conv2d_wgrad(const void* x, const void* dy, float* grad_w, void* workspace, int N, int H, int W, int cin,
                    int cin_real, int cout, int ksize, int accumulate, int stem_perm, hd_stream_t stream);
/* The same launch with caller-owned barrier words (two unsigned ints, ZERO on entry, left zero on exit) instead of the
 * per-call memset of the workspace tail: the split-K partials are reduced INSIDE the kernel after a grid-wide barrier, so a
 * sequence of weight-gradient launches on one stream needs neither a second kernel nor a memset in between. */
int hd_conv2d_wgrad_sync(const void* x, const void* dy, float* grad_w, void* workspace, int N, int H, int W, int cin,
                         int cin_real, int cout, int ksize, int accumulate, int stem_perm, unsigned int* sync_words,
                         hd_stream_t stream);
int hd_conv2d_wgrad_ksplit(int N, int H, int W, int ksize);
size_t hd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int cin, int ksize);

/* OIHW fp32 -> packed bf16 UMMA operand. mode 0: forward [tap][co][ci]; mode 1: dgrad [tap'][ci][co] (rotated). */
int hd_pack_conv_weight(const float* w_oihw, void* out, int cout, int cin, int ksize, int rows_pad, int k_pad,
                        int mode, hd_stream_t stream);
/* All weights of a network in one launch; `jobs` is a DEVICE table of {const float* w; bf16* out; int cout, cin, taps,
 * rows_pad, k_pad, mode (0 fwd, 1 dgrad, 2 stem); long long start} with ascending `start` (flat output offset). */
int hd_pack_all_weights(const void* jobs, int njobs, long long total, hd_stream_t stream);
int hd_nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int c_pad, hd_stream_t stream);
int hd_nhwc_bf16_to_nchw_f32(const void* x, float* y, int N, int C, int H, int W, int c_stride, hd_stream_t stream);

/* 7x7 stride-2 stem (hourglass.py:163): image (N,3,H,W) fp32 -> patch matrix nhwc [N,H/2,W/2,192] (K = 147 padded),
 * then hd_conv2d_igemm(cin=192, ksize=1) with weights from hd_stem_pack_weight ([1][64][192]). */
int hd_stem_im2col(const float* x, void* patches, int N, int H, int W, hd_stream_t stream);
int hd_stem_pack_weight(const float* w, void* out, int cout, hd_stream_t stream);
/* Space-to-depth formulation of the same stem (csrc/stem.cu): unfolded [N][H/2][W/2][64] bf16 with
 * unfolded[.., dx*12 + (c*2+sy)*2 + sx] = img[c][2Y+sy][2(X+dx-2)+sx] (48 of 64 channels used, zero outside the image):
 * the stem is then four vertical taps over it - hd_conv2d_igemm_vtaps with packed weights [4][64][64] (pack mode 3 of the
 * executor) - and its weight gradient hd_conv2d_wgrad(.., cin 64, cin_real 48, cout 64, ksize 1, stem_perm 2). */
int hd_stem_unfold(const float* x_nchw, void* unfolded, int N, int H, int W, hd_stream_t stream);
/* Convolution with a COLUMN of `vtaps` vertical taps (input rows y - pad_top .. y - pad_top + vtaps - 1, same column):
 * w_packed [vtaps][block_n][cin] bf16. Train mode: stat_sum / stat_sqsum (+ optional fused BN finalize `bn`); eval mode:
 * optional per-channel scale / shift (+ ReLU) epilogue. Generic kernel only. */
int hd_conv2d_igemm_vtaps(const void* x, const void* w_packed, void* out, const float* bias, float* stat_sum,
                          float* stat_sqsum, int N, int H, int W, int cin, int cout, int block_n, int vtaps, int pad_top,
                          int out_cs, const hd_bn_fuse* bn, const float* scale, const float* shift, int relu,
                          hd_stream_t stream);

/* Two convolutions summed in one pass: out = conv_{ksize x ksize}(x; w_packed) + conv_{1x1}(x2; w2_packed) (+ addend), all
 * NHWC bf16, 64 output channels on a map with >= #SM 16x16 tiles. In the backward pass of `Residual(64, 128)`
 * (hourglass.py:111-127, PreLayer's 256x256 level) this is dX = dgrad(conv1) + dgrad(skip): the 1x1 skip-branch gradient
 * enters the 3x3 dgrad as one extra tap, so the intermediate tensor and the second launch disappear. */
int hd_conv2d_igemm_dual(const void* x, const void* w_packed, const void* x2, const void* w2_packed, void* out,
                         const void* addend, int N, int H, int W, int cin, int cin2, int cout, int block_n, int ksize,
                         int out_cs, hd_stream_t stream);

/* Backward of the 1x1 prediction head (hourglass.py:189-195). dw/dbias are accumulated into. */
int hd_head_backward(const float* dlogits, long long batch_stride, const void* extra, int extra_cs, const void* feat,
                     const void* w_packed, void* dfeat, float* dw, float* dbias, int N, int H, int W, int cout,
                     hd_stream_t stream);

/* ------------------------------------------------------------------ BatchNorm / activation / pooling */

/* nn.BatchNorm2d (hourglass.py:103) finalize: statistics -> scale/shift (+ running-stat update in training). */
int hd_bn_finalize(const float* sum, const float* sqsum, float count, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                   int training, float* scale, float* shift, float* save_mean, float* save_rstd, int C,
                   hd_stream_t stream);
/* z = act(y*scale + shift)  (Convolution.forward, hourglass.py:107-108) */
int hd_bn_act(const void* y, const float* scale, const float* shift, void* z, long long npix, int C, int relu,
              hd_stream_t stream);
/* out = relu(y2*s2 + b2 + skip) with skip = x or ys*ss + bs  (Residual.forward, hourglass.py:125-127) */
int hd_bn_add_relu(const void* y2, const float* s2, const float* b2, const void* skip, const float* ss,
                   const float* bs, void* out, long long npix, int C, hd_stream_t stream);
/* Same, additionally storing the ReLU mask of the (bf16-rounded) output as bits: mask_out [npix][C/8] bytes, bit j of
 * byte v = (out[8v + j] > 0). hd_bn_bwd_reduce_fin_mask / hd_bn_bwd_apply_mask then take that mask instead of re-reading
 * the activated tensor in the BatchNorm backward of a single-BN residual tail (1 byte instead of 16 per 8 channels). */
int hd_bn_add_relu_mask(const void* y2, const float* s2, const float* b2, const void* skip, const float* ss,
                        const float* bs, void* out, void* mask_out, long long npix, int C, hd_stream_t stream);
int hd_maxpool2(const void* x, void* y, int N, int H, int W, int C, hd_stream_t stream);           /* :72  */
int hd_upsample2_add(const void* up1, const void* low, void* out, int N, int H, int W, int C,
                     hd_stream_t stream);                                                         /* :147,:156 */
/* backward counterparts */
/* g = dout * relu_mask; mask = (out > 0), or, when out == NULL, (y*act_scale + act_shift > 0) recomputed from y.
 * sums[0] += sum g, sums[1] += sum g*y, sums[2] += sum g*ys (RAW moments; hd_bn_bwd_finalize applies mean / rstd). */
int hd_bn_bwd_reduce(const void* dout, const void* out, const float* act_scale, const float* act_shift,
                     const void* y, const float* mean, const float* rstd,
                     const void* ys, const float* mean_s, const float* rstd_s, float* sums, long long npix, int C,
                     hd_stream_t stream);
/* out == NULL: the ReLU mask is rebuilt instead of read - from y*act_scale+act_shift (conv+BN+ReLU unit), or, when ys
 * is given, from y*act_scale+act_shift + ys*act_scale_s+act_shift_s (two-branch residual tail, hourglass.py:125-127);
 * act_scale_s / act_shift_s are only needed in that last case (hd_bn_bwd_reduce_fin and hd_bn_bwd_apply).
 * Same reduction with the finalize fused in: the last block to add its partial sums also computes coef (and coef_s),
 * dgamma / dbeta (and the skip branch's) exactly like hd_bn_bwd_finalize, then re-zeroes `sums`.
 * `counter`: zero-initialised device word (left at zero). */
typedef struct hd_bn_bwd_fuse {
    const float* gamma; const float* mean; const float* rstd; float* coef; float* dgamma; float* dbeta;
    const float* gamma_s; const float* mean_s; const float* rstd_s; float* coef_s; float* dgamma_s; float* dbeta_s;
    float count;
    unsigned int* counter;
} hd_bn_bwd_fuse;
int hd_bn_bwd_reduce_fin(const void* dout, const void* out, const float* act_scale, const float* act_shift,
                         const float* act_scale_s, const float* act_shift_s, const void* y, const void* ys,
                         float* sums, long long npix, int C, const hd_bn_bwd_fuse* fin, hd_stream_t stream);
int hd_bn_bwd_reduce_fin_mask(const void* dout, const void* mask, const void* y, float* sums, long long npix, int C,
                              const hd_bn_bwd_fuse* fin, hd_stream_t stream);
int hd_bn_bwd_apply_mask(const void* dout, const void* mask, const void* y, const float* coef, void* dy, void* gout,
                         long long npix, int C, hd_stream_t stream);
/* Reduce + coefficients + apply of ONE single-BN backward in one launch, for small maps (the latency-bound deep
 * hourglass levels): a grid of <= 64 co-resident CTAs, the last one to arrive builds the coefficients and raises
 * `epoch` (a device word the caller zero-initialises once; it only ever grows), all CTAs then apply them. mask: stored
 * ReLU bits (then gout = dout * mask may be requested) or NULL (mask = y * act_scale + act_shift > 0). */
int hd_bn_bwd_fused_small(const void* dout, const void* mask, const float* act_scale, const float* act_shift, const void* y,
                          float* sums, void* dy, void* gout, long long npix, int C, const hd_bn_bwd_fuse* fin,
                          unsigned int* epoch, hd_stream_t stream);
int hd_bn_bwd_finalize(const float* s0, const float* s1, float count, const float* gamma, const float* mean,
                       const float* rstd, float* coef, float* dgamma, float* dbeta, int accumulate, int C,
                       hd_stream_t stream);
int hd_bn_bwd_apply(const void* dout, const void* out, const float* act_scale, const float* act_shift,
                    const float* act_scale_s, const float* act_shift_s, const void* y, const float* coef, void* dy,
                    const void* ys, const float* coef_s, void* dys, void* gout, long long npix, int C,
                    hd_stream_t stream);
int hd_maxpool2_bwd(const void* x, const void* dpool, const void* add1, const void* add2, void* dx, int N, int H,
                    int W, int C, hd_stream_t stream);
/* `Residual` tail with a BN'd skip branch fused with the 2x2 max pool that follows it in the PreLayer
 * (hourglass.py:165-166, 125-127): pooled [N][H/2][W/2][C] bf16 = maxpool2(relu(y2*s2+b2 + ys*ss+bs)) (values rounded
 * to bf16 before the comparison, like the stored tensor would be) and idx [N][H/2][W/2][C] uint8 = window position
 * (0..3, row-major) of the first maximum; the un-pooled block output is never written. hd_maxpool2_bwd_idx is the
 * pool's backward from those indices (same semantics as hd_maxpool2_bwd, which recomputes them from x). */
int hd_bn_add_relu_pool2(const void* y2, const float* s2, const float* b2, const void* ys, const float* ss,
                         const float* bs, void* pooled, void* idx, int N, int H, int W, int C, hd_stream_t stream);
int hd_maxpool2_bwd_idx(const void* idx, const void* dpool, const void* add1, const void* add2, void* dx, int N, int H,
                        int W, int C, hd_stream_t stream);
/* Backward of that fused tail WITHOUT materialising the routed, masked gradient g = route(dpool, idx) * (pre > 0)
 * at the un-pooled size (autograd of hourglass.py:125-127 + :166): both BatchNorms' reduction (+ fused finalize, as
 * hd_bn_bwd_reduce_fin) and their apply (dy = a*g + b*y2 + c, dys = as*g + bs*ys + cs) read dpool [N][H/2][W/2][C] and
 * idx directly. H, W are the UN-pooled sizes; sc/sh, sc_s/sh_s the two BN scale/shift vectors of the forward pass. */
int hd_bn_bwd_reduce_pool_fin(const void* dpool, const void* idx, const float* sc, const float* sh, const float* sc_s,
                              const float* sh_s, const void* y2, const void* ys, float* sums, int N, int H, int W, int C,
                              const hd_bn_bwd_fuse* fin, hd_stream_t stream);
int hd_bn_bwd_apply_pool(const void* dpool, const void* idx, const float* sc, const float* sh, const float* sc_s,
                         const float* sh_s, const void* y2, const void* ys, const float* coef, const float* coef_s,
                         void* dy, void* dys, int N, int H, int W, int C, hd_stream_t stream);
/* Data-gradient convolution that ALSO produces the BatchNorm-backward statistics of its consumer (autograd of
 * hourglass.py:120-123: conv2's dgrad writes dZ1, the gradient behind conv1's BN + ReLU): out = conv(x) as
 * hd_conv2d_igemm (NHWC bf16, cout == 128 dense), and with g = out * (y * act_scale + act_shift > 0) the epilogue
 * accumulates sums[0][c] += sum g, sums[1][c] += sum g * y; the last CTA writes the coefficients / dgamma / dbeta of `fin`
 * and re-zeroes `sums` - exactly what hd_bn_bwd_reduce_fin(out, NULL, act_scale, act_shift, ..., y, ...) would leave, so
 * hd_bn_bwd_apply can follow directly and the reduction pass over out and y disappears. Only shapes that run on the
 * transposed halo kernel (hd_conv2d_igemm_halo_eligible() == 1); y: NHWC bf16 of the output's shape. */
int hd_conv2d_igemm_halo_eligible(int N, int H, int W, int cout, int ksize);
int hd_conv2d_igemm_bwdstat(const void* x, const void* w_packed, void* out, int N, int H, int W, int cin, int cout,
                            int ksize, const void* y, const float* act_scale, const float* act_shift, float* sums,
                            const hd_bn_bwd_fuse* fin, hd_stream_t stream);
int hd_sum2x2(const void* dout, void* dlow, int N, int H, int W, int C, hd_stream_t stream);
int hd_add(const void* a, const void* b, const void* c, void* out, long long nelem, hd_stream_t stream);
int hd_colsum(const void* x, float* out, long long npix, int C, int cs, hd_stream_t stream);

/* ------------------------------------------------------------------ losses (loss.py:6-69, train.py:105-120) */

/* Fused focal + masked-L1 loss. hm/off/size: fp32 (B,C|2|2,H,W) views with dense planes and batch strides *_bs.
 * out[0..4] = hm, offset, size, total, 1/(B*num_pos). from_logits: sigmoid(hm) inside; sigmoid_reg: sigmoid(off/size). */
int hd_loss_forward(const float* hm, long long hm_bs, const float* off, long long off_bs, const float* size,
                    long long size_bs, const float* ghm, const float* goff, const float* gsize, const float* gmask,
                    int B, int C, int H, int W, float alpha, float beta, float w_hm, float w_off, float w_size,
                    int from_logits, int sigmoid_reg, float* sums, float* out, hd_stream_t stream);
int hd_loss_backward(const float* hm, long long hm_bs, const float* off, long long off_bs, const float* size,
                     long long size_bs, const float* ghm, const float* goff, const float* gsize, const float* gmask,
                     int B, int C, int H, int W, float alpha, float beta, float w_hm, float w_off, float w_size,
                     int from_logits, int sigmoid_reg, const float* fwd_out, const float* grad_out, float* d_hm,
                     long long d_hm_bs, float* d_off, long long d_off_bs, float* d_size, long long d_size_bs,
                     hd_stream_t stream);

/* ------------------------------------------------------------------ decode (transform.py:73-110, evaluate.py:126-182) */

size_t hd_decode_scratch_bytes(int B, int S, int C, int H, int W);
/* Zero the candidate counters at the head of a scratch buffer. Needed ONCE per buffer: hd_decode_nms expects them zero
 * on entry and leaves them zero on exit (its second kernel cleans up after itself), so steady-state calls need no memset. */
int hd_decode_scratch_init(void* scratch, int B, int S, hd_stream_t stream);
/* Two launches for the whole batch: (1) grid-wide sigmoid (apply_sigmoid) + 3x3 peak test + candidate compaction - with
 * conf_th > 0 only elements whose logit can reach the threshold take the nine-sigmoid test -, (2) one CTA per image: per
 * stack exact joint top-k + gather + boxes + threshold, then (do_nms) class-agnostic hard NMS over the concatenated
 * stacks (IoU bit matrix with one thread per pair, serial sweep, parallel write-out). heat/off/wh: fp32 planes with
 * batch/stack strides. scratch: hd_decode_scratch_bytes(), initialised once with hd_decode_scratch_init().
 * Outputs: boxes [B][S*topk][4] fp32, cls [B][S*topk] int64, scores [B][S*topk] fp32, count [B] int32. */
int hd_decode_nms(const float* heat, long long bs_heat, long long ss_heat, const float* off, long long bs_off,
                  long long ss_off, const float* wh, long long bs_wh, long long ss_wh, int B, int S, int C, int H,
                  int W, int topk, float scale_factor, float conf_th, float nms_th, int normalized,
                  int apply_sigmoid, int do_nms, void* scratch, float* out_boxes, long long* out_cls,
                  float* out_scores, int* out_count, hd_stream_t stream);

/* ------------------------------------------------------------------ GT encoder + input normalisation (8(f)-2) */

/* box2hm + draw_gaussian (transform.py:4-70) for a whole batch, as data.py:108-115 calls it from collate_fn, on the
 * device: boxes [B][nmax][4] fp32 (xmin, ymin, xmax, ymax in input pixels), labels [B][nmax] int32 (< 0: empty slot,
 * the reference's `box is None`), nmax <= 128. Outputs (fp32, fully written): heat [B][num_cls][h][w],
 * offset / size [B][2][h][w], mask [B][1][h][w]; h = imsize_y / scale_factor, w = imsize_x / scale_factor.
 * Boxes are applied in list order (the last box owning a centre cell wins, heat is the running maximum). A box whose
 * centre cell falls outside the map (IndexError / negative-index wrap-around in the reference) or whose label is
 * >= num_cls is skipped and counted in *err_count (optional device int, caller-zeroed). */
int hd_encode_targets(const float* boxes, const int* labels, int B, int nmax, int h, int w, int num_cls,
                      int scale_factor, int normalized, float* heat, float* offset, float* size, float* mask,
                      int* err_count, hd_stream_t stream);
/* TF.to_tensor + torchvision Normalize (data.py:118, utils.py:55-68): uint8 [B][H][W][3] -> fp32 [B][3][H][W],
 * ((u8 / 255) - mean[c]) / std[c] in fp32 with IEEE division. mean3 / std3 are HOST arrays of 3 floats. */
int hd_normalize_u8(const void* img_nhwc_u8, float* out_nchw, int B, int H, int W, const float* mean3,
                    const float* std3, hd_stream_t stream);

/* ------------------------------------------------------------------ optimizer (optim.py:3-12, train.py:128-139) */

/* Fused multi-tensor Adam: one launch over all tensors. jobs_host: njobs records {float* p; const float* g; float* m;
 * float* v; long long n; long long chunk_start} (chunks of 1024 elements, ascending); jobs_dev: device scratch of the
 * same size; step_dev: device float = steps taken so far (advanced here); grad_scale / found_inf: optional device
 * floats of torch.amp.GradScaler (gradients are multiplied by 1/grad_scale; nothing happens when found_inf != 0). */
int hd_adam_step(const void* jobs_host, int njobs, void* jobs_dev, long long nchunks, float lr, float beta1,
                 float beta2, float eps, float* step_dev, const float* grad_scale, const float* found_inf,
                 hd_stream_t stream);

/* ------------------------------------------------------------------ whole-network executor (hourglass.py:198-237) */

/* Pointers of one `Convolution` module (hourglass.py:94-108): parameters, buffers and gradient destinations.
 * Unused members are NULL. Gradient destinations must be zero-initialised by the caller before hd_net_backward. */
typedef struct hd_unit_ptrs {
    float* w;      float* b;                       /* convolution.weight (OIHW fp32), convolution.bias */
    float* gamma;  float* beta;                    /* bn.weight, bn.bias */
    float* running_mean; float* running_var; long long* num_batches_tracked;
    float* dw;     float* db;   float* dgamma;  float* dbeta;
} hd_unit_ptrs;

typedef struct hd_net hd_net;
/* num_stack, in_ch (128) and out_ch (= num_cls + 4) of StackedHourglass.__init__ (hourglass.py:199). */
int hd_net_create(int num_stack, int in_ch, int out_ch, hd_net** net);
void hd_net_destroy(hd_net* net);
int hd_net_num_units(const hd_net* net);
/* Inference with frozen parameters: when on, eval-mode forwards after the next one reuse the packed bf16 weights and
 * the folded BatchNorm constants in the workspace instead of rebuilding them from the fp32 parameters on every call
 * (same workspace pointer required; any training-mode forward or a new call of this function invalidates them). */
void hd_net_set_static_weights(hd_net* net, int on);
/* Bytes of workspace hd_net_forward (+ hd_net_backward when with_backward) need for a (B,3,H,W) input. */
size_t hd_net_workspace_bytes(hd_net* net, int B, int H, int W, int with_backward);
/* StackedHourglass.forward (hourglass.py:223-237): x (B,3,H,W) fp32 -> logits (B,S,out_ch,H/4,W/4) fp32.
 * training: batch statistics + running-stat update and activations kept for hd_net_backward. */
int hd_net_forward(hd_net* net, const hd_unit_ptrs* units, int n_units, const float* x, float* logits,
                   void* workspace, size_t workspace_bytes, int B, int H, int W, int training, hd_stream_t stream);
/* Autograd of the above: dlogits (B,S,out_ch,H/4,W/4) fp32 -> all parameter gradients (units[i].d*). */
int hd_net_backward(hd_net* net, const hd_unit_ptrs* units, int n_units, const float* dlogits, void* workspace,
                    size_t workspace_bytes, hd_stream_t stream);
/* The same pass in two calls - the hook for DistributedDataParallel's overlap of the gradient exchange with backward
 * (train.py:174-175; the reducer all-reduces finished buckets while autograd still runs): stage 1 enqueues the backward
 * of the stacks (head, neck, hourglass) and, if comm_stream != NULL, makes comm_stream wait until every parameter
 * gradient of those units is complete, so the caller can enqueue its collective there; stage 2 (same arguments)
 * enqueues the PreLayer backward and joins the library's internal streams into `stream`. */
int hd_net_backward_stage(hd_net* net, const hd_unit_ptrs* units, int n_units, const float* dlogits, void* workspace,
                          size_t workspace_bytes, hd_stream_t stream, int stage, hd_stream_t comm_stream);

#ifdef __cplusplus
}
#endif
#endif /* HD_B200_H */
