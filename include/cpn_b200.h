/*
 * cpn_b200.h -- C ABI of the B200-native Contour Proposal Network inference path.
 *
 * The reference (FZJ-INM1-BDA/celldetection) is pure Python and has no FFI layer; its boundary for this path is the
 * public Python API.  This header is the C boundary a host in any language binds instead: plain device
 * pointers + sizes + a cudaStream_t (passed as void*), int status return (0 = ok) and cpn_last_error().  No torch
 * types appear in any signature.  Each entry point cites the reference code it replaces (paths relative to
 * /root/reference/celldetection unless noted).  The ctypes binding the reference side would add is shown in
 * INTEGRATION.md; celldetection_b200/_lib.py is that binding for this repository.
 *
 * All pointers are DEVICE pointers unless a parameter name ends in _host.  All kernels are enqueued on `stream` and
 * return without synchronising unless stated.
 */
#ifndef CPN_B200_H
#define CPN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPN_B200_ABI_VERSION 5

/* ---------------------------------------------------------------------------------------------------------------- */
/* status / diagnostics                                                                                             */
/* ---------------------------------------------------------------------------------------------------------------- */
int cpn_abi_version(void);
/* Thread-local message of the last failing call (never NULL). */
const char* cpn_last_error(void);
/* Number of kernels this library has launched since load (all streams); feeds bench.py's "gpu_launches". */
int64_t cpn_launch_count(void);
/* Fills name[0..n) with the device name and returns the SM count of the current device (or <0 on error). */
int cpn_device_info(char* name_host, int n, int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---------------------------------------------------------------------------------------------------------------- */
/* layer plan: backbone + heads (replaces CPNCore.forward, models/cpn.py:238-283, and everything below it:           */
/* models/unet.py:29-58,178-249, models/resnet.py:56-290, models/fpn.py:50-134, models/commons.py:120-149,461-511,  */
/* 686-700, i.e. the nn.Conv2d/BatchNorm2d/MaxPool2d/F.interpolate calls into cuDNN/ATen)                           */
/* ---------------------------------------------------------------------------------------------------------------- */

enum {
  CPN_DT_F32 = 0,
  CPN_DT_F16 = 1,
  CPN_DT_U8 = 2,
  CPN_DT_U16 = 5,   /* raw 16-bit image data (cpn_histogram / cpn_apply_lut only) */
  CPN_DT_F16X2 = 3, /* split fp16 pair: value = hi + lo, hi = fp16(v), lo = fp16(v - hi); channel c of a view lives at
                      element c (hi) and c + lo_delta (lo) of the pixel.  Used by the 3-pass tensor-core engine
                      (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo, fp32 accumulate) that reaches fp32-level accuracy. */
  CPN_DT_F16F8 = 4 /* fp16 value + two e4m3 correction operands (the 2-pass tensor-core engine: one kind::f16 pass
                      A_hi*W_hi and ONE kind::f8f6f4 pass over the K-concatenated (A_lo | A_hi) x (W_hi ; W_lo) e4m3
                      operands, both into the same fp32 accumulator).  Channel c of a view lives at element c (hi =
                      fp16(v)); the pixel's 8-bit block starts lo_delta elements after the view's first element and
                      holds, per 32-channel chunk j, 64 bytes: 32 x e4m3((v - hi) * 2^(8+fp8_exp)) followed by
                      32 x e4m3(hi * 2^(-2+fp8_exp)) (round to nearest even, saturating at +-448), so a 64-channel
                      K block is one 128-byte row for either instruction kind.  Channel offsets of views must be
                      multiples of 32.  Same 4 bytes per element as CPN_DT_F16X2. */
};

/* NHWC view into the activation arena: element (n,y,x,c) lives at
 *   arena + offset + ((n*h + y)*w + x) * pitch * sizeof(dtype) + c * sizeof(dtype).
 * `pitch` (in elements) >= c lets a tensor be a channel slice of a wider buffer, which is how torch.cat
 * (models/unet.py:219-224) is realised without a copy. */
typedef struct {
  int64_t offset; /* bytes from arena base (for INPUT/OUTPUT bindings: bytes from the bound pointer) */
  int32_t n, h, w, c;
  int32_t pitch;
  int32_t dtype;  /* CPN_DT_* */
  int32_t lo_delta; /* CPN_DT_F16X2: elements from a channel's hi half to its lo half; CPN_DT_F16F8: elements from the
                       view's first hi element to its 8-bit block (0 otherwise) */
  int32_t fp8_exp;  /* CPN_DT_F16F8: power-of-two exponent of the tensor's 8-bit operand scales (0 suits O(1e-3..1e3)
                       activations); all views of one buffer carry the same value */
} cpn_view_t;

enum {
  CPN_OP_PREP = 0,      /* network input -> NHWC activations; fuses Normalize's range check (commons.py:694-700),
                           uint8 -> float / 255 (lightning_base.py:774-780) and the layout change.  With r > 0 it
                           writes the im2col operand of a k=r, stride, pad stem convolution instead:
                           dst[n,oy,ox,(r*k+s)*C+c] = in[n,c,oy*stride-pad+r,ox*stride-pad+s], zero-padded channels */
  CPN_OP_CONV = 1,      /* conv (+ folded BN) (+ residual) (+ ReLU)  */
  CPN_OP_MAXPOOL = 2,   /* nn.MaxPool2d(k, stride, pad), -inf padding */
  CPN_OP_UPSAMPLE = 3,  /* F.interpolate(mode='nearest'): src = floor(dst * in / out) */
  CPN_OP_BILINEAR = 4,  /* F.interpolate(mode='bilinear', align_corners=False) (models/cpn.py:109-115) */
  CPN_OP_PROJ = 5       /* ReadOut's final 1x1 conv (+ ScaledTanh) to fp32 head tensors (commons.py:494-511) */
};

enum { CPN_IN_F32_NCHW = 0, CPN_IN_U8_NCHW = 1, CPN_IN_U8_NHWC = 2 };
enum { CPN_ACT_NONE = 0, CPN_ACT_RELU = 1, CPN_ACT_SCALED_TANH = 2,
       CPN_ACT_SIGMOID = 3 /* ReadOut(final_activation='sigmoid') of the uncertainty head, models/cpn.py:208-219 */ };
/* cpn_op_t::flags */
enum { CPN_CONV_UP2 = 1 /* F.interpolate(scale 2, 'nearest') followed by this 3x3 / stride 1 / pad 1 convolution
                           (models/unet.py:213-217 + the bridge block :95-98), computed WITHOUT materialising the
                           up-sampled tensor: src is the low-resolution input [n,h,w,cin], dst the high-resolution output
                           [n,2h,2w,cout], and the weights are the four phase kernels of the composition,
                           [3*3][4*cout][kslab] with output channel (2a+b)*cout + c = output pixel (2y+a, 2x+b), channel
                           c (each phase kernel sums the taps of the original kernel that read the same source pixel;
                           zero padding of the up-sampled image equals zero padding of the source).  The bias holds
                           4*cout values.  Needs 4*cout % 256 == 0. */ };
/* conv engines */
enum { CPN_ENGINE_SIMT = 0, CPN_ENGINE_TCGEN05 = 1 };

typedef struct {
  int32_t kind;          /* CPN_OP_* */
  int32_t engine;        /* CONV only: CPN_ENGINE_* */
  cpn_view_t src;        /* PREP: describes the *logical* input (n,h,w,c); the pointer is bound at forward time */
  cpn_view_t dst;
  cpn_view_t res;        /* CONV: optional residual (n == 0 -> none).  If res.h/res.w differ from dst.h/dst.w the
                            residual is read through nearest up-sampling (torchvision FPN top-down path). */
  int64_t w_offset;      /* CONV/PROJ: bytes into the weight blob.
                            SIMT   : float  [R*S][kslab][cout]
                            TCGEN05: __half [R*S][cout][kslab]   (K-major, TMA box {64, BN, 1}); for CPN_DT_F16X2
                                     activations [R*S][cout][3*kslab] = (W_hi | W_hi | W_lo) along K; for CPN_DT_F16F8
                                     [R*S][cout][2*kslab] half slots = (W8 | fp16(S * W_hi)) along K, W8 holding per
                                     32-channel chunk 32 x e4m3(W_hi * 2^a) then 32 x e4m3(W_lo * 2^(a+10)) (pairing
                                     with the activations' (lo | hi) bytes), S = 2^(8 + src.fp8_exp + a) = 1/acc_scale
                            PROJ   : float  [cout][cin_slice]    */
  int64_t b_offset;      /* CONV/PROJ: bytes into the weight blob of float bias[cout] (-1 -> no bias) */
  int32_t r, s, stride, pad;
  int32_t kslab;         /* input channels each 64-wide output-channel slab contracts over: cin for dense convs,
                            max(64, cin/groups) for grouped convs (block-diagonal weights are expanded to the slab) */
  int32_t slab_mode;     /* 0: dense (slab base 0); 1: grouped (slab base = (n0 / kslab) * kslab) */
  int32_t act;           /* CPN_ACT_* */
  float act_scale;       /* ScaledTanh factor (models/commons.py:167-168) */
  int32_t proj_cin_off;  /* PROJ: first input channel of the slice this projection contracts over */
  int32_t proj_cin;      /* PROJ: number of input channels */
  int32_t out_binding;   /* PROJ/any: -1 -> dst is in the arena; >= 0 -> dst is the caller's output pointer #k */
  int32_t fuse_next;     /* CONV (TCGEN05): number of immediately following PROJ ops computed inside this
                            convolution's epilogue (one per output-channel tile, in tile order); they are skipped by
                            cpn_plan_forward and the convolution's own dst is then not written */
  float acc_scale;       /* CONV (TCGEN05): the fp32 accumulator is multiplied by this before bias / residual / activation
                            (1/S of the CPN_DT_F16F8 weight packing; 0 is read as 1) */
  int32_t flags;         /* CONV (TCGEN05): CPN_CONV_UP2 */
} cpn_op_t;

typedef struct cpn_plan cpn_plan_t;

/* Validates the op list, builds every TMA tensor map / launch configuration once, and keeps device pointers to the
 * caller-owned weight blob and activation arena (both must outlive the plan; the library never allocates device
 * memory).  `flags_dev` must point to >= 4 int32 of device memory (bit 0 of flags_dev[0] is set by PREP when an
 * input value is outside [0, 1] -- the reference raises AssertionError there, commons.py:695-697). */
int cpn_plan_create(const cpn_op_t* ops_host, int n_ops, const void* weights, size_t weights_bytes, void* arena,
                    size_t arena_bytes, int32_t* flags_dev, cpn_plan_t** plan_out);
/* Runs the whole list on `stream`.  `input` is the network input in `input_format`; outputs_host[k] is the device
 * pointer bound to ops with out_binding == k. */
int cpn_plan_forward(cpn_plan_t* plan, const void* input, int input_format, void* const* outputs_host, int n_outputs,
                     void* stream);
/* Runs ops [first_op, end_op) only (fused projections travel with their convolution and must not be split from it).  The
 * host uses it to enqueue the score head first, start the read-back of the proposal count, and enqueue the rest of the plan
 * (the full-resolution branch and the refinement head) while that read-back is in flight. */
int cpn_plan_forward_range(cpn_plan_t* plan, int first_op, int end_op, const void* input, int input_format,
                           void* const* outputs_host, int n_outputs, void* stream);
/* Kernel launches one forward of this plan issues. */
int cpn_plan_num_launches(const cpn_plan_t* plan);
/* Sparse-heads plans (one 1x1 layer over a [1, rows/16, 16] row matrix): the following forwards compute only the 128-row
 * blocks that hold the first `rows` rows (the proposals actually gathered); rows < 0 restores the whole matrix. */
int cpn_plan_set_active_rows(cpn_plan_t* plan, int64_t rows);
/* Debug/profiling: run only op `index` (same bindings as forward). */
int cpn_plan_run_op(cpn_plan_t* plan, int index, const void* input, int input_format, void* const* outputs_host,
                    int n_outputs, void* stream);
void cpn_plan_destroy(cpn_plan_t* plan);

/* Stand-alone convolution (same kernels as the plan; used by the per-layer parity tests).  Views are relative to
 * the given base pointers instead of an arena. */
int cpn_conv2d(const cpn_op_t* op_host, const void* src_base, void* dst_base, const void* res_base,
               const void* weights, void* stream);

/* ---------------------------------------------------------------------------------------------------------------- */
/* post-head chain (replaces models/cpn.py:575-734 and ops/cpn.py:15-165)                                           */
/* ---------------------------------------------------------------------------------------------------------------- */

/* Bytes of scratch cpn_select_* needs for a score map with `pixels` = N*h*w entries. */
size_t cpn_select_workspace_bytes(int64_t pixels);

/* Pass 1 of proposal selection (models/cpn.py:578-579, 616-620): sigmoid(logit) [min upper] [max lower] > thresh.
 * logits [N,h,w] fp32; lower/upper: optional [N,h,w] fp32 bounds already at head resolution (NULL -> none).
 * Writes the total number of proposals to total_dev[0] (int64). */
int cpn_select_count(const float* logits, const float* lower, const float* upper, int64_t pixels, float thresh,
                     void* workspace, int64_t* total_dev, void* stream);
/* Pass 2: writes, in torch.where order (n, y, x ascending), the flat pixel index (int32, n*h*w + y*w + x) and the
 * post-bound score of every proposal, plus seg_offsets[N+1] (int32 row offsets of each image's proposals).
 * `capacity` rows are available in idx/score; must be >= the total from pass 1 (same workspace, same inputs). */
int cpn_select_write(const float* logits, const float* lower, const float* upper, int n_images, int64_t hw,
                     float thresh, void* workspace, int32_t* idx, float* score, int64_t capacity,
                     int32_t* seg_offsets, void* stream);

/* General form of the two selection passes for every scoring variant of models/cpn.py:575-587 and the certainty filter
 * of :616-618.  `channels` = score_channels of the model:
 *   1   scores = sigmoid(logit) [min upper] [max lower]; class = scores > thresh; foreground = class > 0
 *   > 2 scores = softmax over the channels, every channel [min upper] [max lower]; class = argmax (first maximum);
 *       foreground = class > 0 (thresh unused); the selected score is scores[class]
 * logits [N,h,w,channels] fp32 (channel-last records, as the plan's score head writes them); lower/upper [N,h,w] or
 * NULL.  uncertainty: optional [N,h,w,4] fp32 (sigmoid outputs of the uncertainty head); when use_certainty != 0 a
 * pixel is foreground only if mean(uncertainty[pixel, :]) < certainty_limit (= 1 - certainty_thresh, rounded to fp32
 * by the host exactly like the reference's tensor-scalar comparison). */
typedef struct {
  const float* logits;
  const float* lower;
  const float* upper;
  const float* uncertainty;
  int32_t channels;
  int32_t use_certainty;
  float thresh;
  float certainty_limit;
} cpn_select_params_t;
int cpn_select_count_ex(const cpn_select_params_t* params_host, int64_t pixels, void* workspace, int64_t* total_dev,
                        void* stream);
/* classes (optional, may be NULL) receives the class of every proposal as int64 (1 for channels == 1). */
int cpn_select_write_ex(const cpn_select_params_t* params_host, int n_images, int64_t hw, void* workspace, int32_t* idx,
                        float* score, int64_t* classes, int64_t capacity, int32_t* seg_offsets, void* stream);

/* F.interpolate(mode='bilinear', align_corners=False) of a stand-alone fp32 NHWC tensor [n,h,w,c] -> [n,ho,wo,c]: the
 * resize of scores_lower_bound / scores_upper_bound to the head resolution (_apply_score_bounds / _equal_size,
 * models/cpn.py:109-123; with c == 1 the reference's [N,1,h,w] layout is the same memory). */
int cpn_resize_bilinear(const float* src, int n, int h, int w, int c, float* dst, int ho, int wo, void* stream);

/* NMS weights of models/cpn.py:723-726 (uncertainty_nms): out[i] = scores[i] * (1 - mean(uncertainty[idx[i], 0:4])). */
int cpn_nms_weights(const float* scores, const float* uncertainty, const int32_t* idx, int64_t P, float* out,
                    void* stream);

/* Decode + rescale + local refinement + clamp + boxes + offsets for P selected pixels in one kernel
 * (ops/cpn.py:15-41 rel->abs, :44-95 inverse DFT, :98-165 rescale; models/cpn.py:63-85 refinement, :661-670 clamp
 * and boxes, :695-702 offsets).
 *   idx        [P] int32 flat pixel indices into the head maps
 *   locfou     [N,h,w,2+4*order_core] fp32 records: (loc_x, loc_y, fourier[order_core][4]) per pixel
 *   trig       [2][order][samples] fp32: cos then sin of 2*pi*k*t_s (host computes it exactly as ops/cpn.py:74-81)
 *   refinement [N,H,W,2] fp32 (x then y displacement, already 3*tanh) or NULL
 *   offsets    [N,2] fp32 (x, y) or NULL
 * outputs (any may be NULL): contours [P,S,2], proposals [P,S,2], boxes [P,4], locations [P,2], fourier [P,order,4] */
int cpn_decode_refine(const int32_t* idx, int64_t P, const float* locfou, int order_core, int order, int n_images,
                      int h, int w, int H, int W, const float* trig, int samples, const float* refinement, int iters,
                      const float* offsets, float* contours, float* proposals, float* boxes, float* locations,
                      float* fourier_out, void* stream);

/* cpn_decode_refine with bucketed local refinement (refinement_buckets > 1; models/cpn.py:73-82, ops/cpn.py:238-255):
 * refinement is [N,H,W,2*buckets] (bucket j: x at channel 2j, y at 2j+1); per contour sample s the three neighbouring
 * buckets and their weights are read from bucket_idx [samples][3] int32 / bucket_w [samples][3] fp32, which the host
 * computes from the sampling exactly like resolve_refinement_buckets; the displacement is
 * ((r[b0]*w0) + (r[b1]*w1)) + (r[b2]*w2), the reference's summation order.  buckets == 1 ignores the tables. */
int cpn_decode_refine_buckets(const int32_t* idx, int64_t P, const float* locfou, int order_core, int order,
                              int n_images, int h, int w, int H, int W, const float* trig, int samples,
                              const float* refinement, int iters, int buckets, const int32_t* bucket_idx,
                              const float* bucket_w, const float* offsets, float* contours, float* proposals,
                              float* boxes, float* locations, float* fourier_out, void* stream);

/* cpn_decode_refine_buckets whose locfou holds one record per PROPOSAL (row i belongs to idx[i]) when records_by_row != 0
 * -- the output of the sparse location / fourier heads (cpn_gather_patches + a 1x1 plan) -- instead of one per pixel. */
int cpn_decode_refine_rows(const int32_t* idx, int64_t P, const float* locfou, int records_by_row, int order_core,
                           int order, int n_images, int h, int w, int H, int W, const float* trig, int samples,
                           const float* refinement, int iters, int buckets, const int32_t* bucket_idx,
                           const float* bucket_w, const float* offsets, float* contours, float* proposals, float* boxes,
                           float* locations, float* fourier_out, void* stream);

/* Sparse evaluation of the ReadOut heads that are only read at proposals (location, fourier: models/cpn.py:253-263 compute
 * them on every pixel, :620-623 gather them at the selected pixels).  For the P flat pixel indices idx (n*h*w + y*w + x,
 * as written by cpn_select_write) of the head feature map `src` (view: fp16-family NHWC, c % 64 == 0) writes the k x k x c
 * neighbourhood of every pixel as row i of `dst` (view with c_dst = k*k*c channels and at least P pixels, same dtype):
 * zero outside the image, channel order [c/64 blocks][k*k taps][64 channels] -- the order in which the dense k x k
 * convolution contracts.  A plan holding ONE 1x1 convolution over dst (weights in the same K order) plus the fused
 * projections then yields the head records of the P proposals only.  View offsets are relative to src / dst. */
int cpn_gather_patches(const void* src, const cpn_view_t* src_view_host, const int32_t* idx, int64_t P, int k, void* dst,
                       const cpn_view_t* dst_view_host, void* stream);

/* ops.cpn.fouriers2contours (ops/cpn.py:44-95): fourier [P,order,4], locations [P,2] -> out [P,samples,2].
 * trig as above; or sampling [P,samples] (per-proposal t, :67-71) with trig == NULL. */
int cpn_fouriers2contours(const float* fourier, const float* locations, int64_t P, int order, int samples,
                          const float* trig, const float* sampling, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------- */
/* NMS (replaces torch.ops.torchvision.nms as called by ops/cpn.py:189-227 and cpn_inference.py:405-408)            */
/* ---------------------------------------------------------------------------------------------------------------- */
size_t cpn_nms_workspace_bytes(int64_t n_boxes, int n_segments);
/* Greedy NMS per segment with torchvision semantics: stable descending score order, suppress when IoU > thr (strict,
 * NaN never suppresses), areas without +1.  seg_offsets [n_segments+1] int32 (device).  Segments larger than
 * `chunk` follow ops/cpn.py:213-224 (NMS per chunk, then NMS over the survivors); chunk <= 0 disables the rule.
 * keep [n_boxes] int32 receives, per segment and packed at the segment's offset, the kept GLOBAL row indices in
 * descending score order; keep_counts [n_segments] int32 the number kept. */
int cpn_nms_segments(const float* boxes, const float* scores, const int32_t* seg_offsets, int n_segments,
                     int64_t n_boxes, float iou_threshold, int chunk, void* workspace, int32_t* keep,
                     int32_t* keep_counts, void* stream);

/* Exact greedy NMS of ONE large segment (the global stitch over all tiles, cpn_inference.py:405-408: 1e5-1e6 boxes)
 * by parallel fixed-point rounds over a uniform spatial grid; same semantics and result as cpn_nms_segments with one
 * segment and chunk <= 0.  Synchronises the stream internally (the round loop polls a device counter).
 * rounds_host (optional) receives the number of rounds executed. */
size_t cpn_nms_grid_workspace_bytes(int64_t n_boxes);
int cpn_nms_grid(const float* boxes, const float* scores, int64_t n_boxes, float iou_threshold, void* workspace,
                 int32_t* keep, int32_t* keep_count, int* rounds_host, void* stream);

/* Box voting of model ensembles (cd.ops.filter_by_box_voting / get_iou_voting, ops/boxes.py:53-83, called by
 * cpn_inference.py:419-424): votes[i] = sum over all boxes j (including i) of iou(i,j) * (iou(i,j) > iou_threshold),
 * torchvision box_iou arithmetic.  Sparse: only boxes in neighbouring cells of a uniform grid are visited (all other
 * pairs contribute exact zeros), O(n) memory.  The host keeps rows with votes >= min_vote. */
size_t cpn_box_votes_workspace_bytes(int64_t n_boxes);
int cpn_box_votes(const float* boxes, int64_t n_boxes, float iou_threshold, void* workspace, float* votes, void* stream);

/* remove_border_contours (ops/cpn.py:258-290) as called by cpn_inference.py:375-380: keep[i] = 1 iff every vertex of
 * contour i satisfies the enabled side tests in tile-local coordinates (contours + (-offset)).
 * contours [K,S,2]; tile_of_row [K] int32 -> row of tile_meta; tile_meta [T,CPN_TILE_META] float:
 * (off_x, off_y, h, w, top, right, bottom, left, ex_br, stop_x, stop_y, 0).  With ex_br != 0 the 'ex_br' stitching rule
 * (filter_contours_by_stitching_rule, ops/cpn.py:293-325) additionally drops contours whose every vertex has
 * x >= stop_x or y >= stop_y (stop = tile size - right/bottom overlap), in the same tile-local coordinates. */
#define CPN_TILE_META 12
int cpn_border_filter(const float* contours, const int32_t* tile_of_row, const float* tile_meta, int64_t K, int samples,
                      float padding, uint8_t* keep, void* stream);

/* ---------------------------------------------------------------------------------------------------------------- */
/* label rasterisation (replaces cd.data.contours2labels, data/cpn.py:292-358, incl. render_contour :246-256 and the  */
/* third-party cv2.drawContours(thickness=-1) polygon fill it calls; used by cpn_inference.py:809-813)               */
/* ---------------------------------------------------------------------------------------------------------------- */
size_t cpn_contours2labels_workspace_bytes(int64_t n_contours, int samples);
/* contours [K,S,2] fp32 (x, y) -> labels [H,W,channels] int32 (zeroed by the call): contour k is rounded (half to even,
 * `rounded`), clipped to the image (`clip`), filled exactly like OpenCV's integer scan-line polygon fill and written as
 * label k+1 into the first channel whose bounding-box region dilated by `gap` is still empty -- the reference's
 * sequential rule, evaluated as an ordered dependency schedule (one warp per contour).  info_dev [4] int32 receives
 * [0] channels used, [1] channels needed (> channels: the result is incomplete, call again with more channels; the
 * reference appends channels on demand), [2] non-zero on an internal scheduling time-out.  channels <= 64. */
int cpn_contours2labels(const float* contours, int64_t n_contours, int samples, int H, int W, int rounded, int clip,
                        int gap, int32_t* labels, int channels, void* workspace, int32_t* info_dev, void* stream);

/* cd.data.resolve_label_channels (data/cpn.py:361-398, method='dilation', kernel (3,3) = cv2 MORPH_CROSS): labels
 * [H,W,channels] int32 -> flat [H,W] int32.  Pixels of exactly one object keep its label, pixels of several are filled
 * by Jacobi sweeps of a 5-point grey dilation until nothing changes or max_iter sweeps ran (reference default 999).
 * Synchronises the stream (one flag read per sweep).  sweeps_host (optional) receives the number of sweeps. */
size_t cpn_resolve_label_channels_workspace_bytes(int H, int W);
int cpn_resolve_label_channels(const int32_t* labels, int H, int W, int channels, int max_iter, int32_t* flat,
                               void* workspace, int* sweeps_host, void* stream);

/* Region statistics of a label image for the csv output (cd.data.labels2property_table, data/misc.py:320-345 -> third-party
 * skimage.measure.regionprops_table; cpn_inference.py:824-837).  labels [H, W, channels] int32 (<= 0 = background; values >
 * max_label set bit 0 of flags[0] and are ignored).  Slot = channel * (max_label + 1) + label: area[slot] pixel count,
 * bbox[slot] = (min_row, min_col, max_row + 1, max_col + 1), sums[slot] = (sum of rows, sum of columns).  All outputs are
 * initialised by the call. */
int cpn_label_props(const int32_t* labels, int H, int W, int channels, int max_label, uint32_t* area, int32_t* bbox,
                    uint64_t* sums, int32_t* flags, void* stream);

/* ---------------------------------------------------------------------------------------------------------------- */
/* input-side preprocessing of a slide (celldetection_scripts/cpn_inference.py:196-222 `preprocess`,                 */
/* cd.data.normalize_percentile data/misc.py:156-161)                                                               */
/* ---------------------------------------------------------------------------------------------------------------- */
/* Exact histogram of n uint8 (CPN_DT_U8, 256 bins) or uint16 (CPN_DT_U16, 65536 bins) values into hist (uint32, zeroed by
 * the call).  The host derives np.percentile's order statistics and the image mean from it. */
int cpn_histogram(const void* data, int dtype, int64_t n, uint32_t* hist, void* stream);
/* dst[i] = lut[src[i]] (lut: 256 or 65536 uint8 entries on the device): percentile normalisation to uint8, gamma and
 * brightness / contrast adjustment composed into one table by the host. */
int cpn_apply_lut(const void* src, int dtype, int64_t n, const uint8_t* lut, uint8_t* dst, void* stream);
/* `grayscale` branch of preprocess (cpn_inference.py:203-213, third-party cv2.cvtColor COLOR_RGB2GRAY / COLOR_RGBA2GRAY, 8-bit
 * fixed point (9798 R + 19235 G + 3735 B + 2^14) >> 15): src [n_px, channels] interleaved (channels 3 or 4; uint16 sources
 * need `lut`), each element first mapped through `lut` when given (nullable), dst [n_px] uint8. */
int cpn_rgb2gray(const void* src, int dtype, int64_t n_px, int channels, const uint8_t* lut, uint8_t* dst, void* stream);

/* Phase-decomposed refinement head (models/cpn.py:274-279: F.interpolate(x2, bilinear) followed by the ReadOut's k x k
 * convolution, run as four phase convolutions on the low-resolution map).  cpn_unshuffle2: phase-packed records
 * rec [N, h, w, 4 * c] (phase (a, b) at channels (2a + b) * c ...) -> out [N, 2h, 2w, c] with out[n, 2y + a, 2x + b] = that
 * phase's record.  cpn_copy_window: copies the hh x ww pixel window at (ys, xs) of every image of src [N, Hs, Ws] to
 * (yd, xd) of dst [N, Hd, Wd]; pixels are bytes_per_px bytes (multiple of 4) in both -- crops the border strips of a
 * feature map into the small tensors the plain path recomputes them on, and pastes their results back. */
int cpn_unshuffle2(const float* rec, int N, int h, int w, int c, float* out, void* stream);
int cpn_copy_window(const void* src, void* dst, int N, int Hs, int Ws, int Hd, int Wd, int bytes_per_px, int ys, int xs,
                    int yd, int xd, int hh, int ww, void* stream);

/* dst[i, :] = src[index[i], :] for rows of row_bytes (multiple of 4) bytes (resolve_keep_indices, cpn.py:53-60). */
int cpn_gather_rows(const void* src, int64_t row_bytes, const int32_t* index, int64_t n_rows, void* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CPN_B200_H */
