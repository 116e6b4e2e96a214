/* frcnn_b200.h -- C ABI of the B200-native Faster R-CNN detection hot path.
 *
 * Drop-in boundary for andreaskoepf/faster-rcnn.torch.  The reference has no FFI of its own (pure Lua on top
 * of Torch7 nn/cunn); each entry point below names the reference interface (file:line) it replaces.  The Lua
 * glue (faster-rcnn.torch_b200/lua/) binds exactly these symbols through LuaJIT FFI; the
 * Python host mirror binds the same symbols through ctypes.
 *
 * The declarations are cdef-clean: after removing lines that start with '#' and the extern "C" braces the file
 * can be passed verbatim to LuaJIT ffi.cdef / Python cffi.
 *
 * Conventions
 *   - every function returns an int status (FRCNN_OK == 0); the message of the last failure is available from
 *     frcnn_last_error(ctx) (ctx may be NULL for failures of frcnn_create).  No C++ exception and no exit()
 *     crosses this boundary.
 *   - "_dev" pointers are CUDA device pointers borrowed for the duration of the call (Torch / the caller owns
 *     every tensor); "_host" pointers are ordinary host memory.  Plain pointers and sizes only.
 *   - all work is enqueued on the stream given to frcnn_create (NULL = legacy default stream, which keeps the
 *     ordering with surrounding cutorch work).  Entry points that return host data synchronise that stream.
 *   - a ctx is bound to one device and is not re-entrant.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with FRCNN_E_CUDA.
 */
#ifndef FRCNN_B200_H
#define FRCNN_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct frcnn_ctx frcnn_ctx;

enum {
  FRCNN_OK = 0,
  FRCNN_E_INVALID = 1,   /* bad argument */
  FRCNN_E_CUDA = 2,      /* CUDA runtime / driver failure (including: no device) */
  FRCNN_E_NOMEM = 3,
  FRCNN_E_STATE = 4,     /* call order violated (e.g. forward before bind_params) */
  FRCNN_E_ROI_EMPTY = 5, /* an ROI clipped to max == 0: the reference raises here (objective.lua:11) */
  FRCNN_E_OVERFLOW = 6,  /* more candidates than the capacity given */
  FRCNN_E_NCCL = 7
};

/* nms.lua:37-43 -- what the `scores` argument selects.  A tensor argument falls through to Y2 (sic). */
enum { FRCNN_NMS_ORDER_Y2 = 0, FRCNN_NMS_ORDER_AREA = 1, FRCNN_NMS_ORDER_COLUMN = 2 };

/* models/vgg_small.lua:5-10 `layers` entries */
typedef struct frcnn_block_desc {
  int filters, kW, kH, padW, padH, conv_steps;
  float dropout;
} frcnn_block_desc;
/* models/vgg_small.lua:12-17 `anchor_nets` entries (input is 1-based, as in Lua) */
typedef struct frcnn_head_desc {
  int kW, n, input;
} frcnn_head_desc;
/* models/vgg_small.lua:19-22 `class_layers` entries */
typedef struct frcnn_fc_desc {
  int n;
  float dropout;
  int batch_norm;
} frcnn_fc_desc;

/* One entry of the match list built by Detector.lua:36-66 ({p, a, r, l}) */
typedef struct frcnn_candidate {
  double r[4];   /* decoded rect, Lua doubles {minX, minY, maxX, maxY} (Anchors.lua:245-252) */
  float box[4];  /* r:totensor(), fp32 (Rect.lua:143-145) */
  float logp;    /* c[1], foreground log-probability (Detector.lua:52) */
  int layer, aspect, y, x; /* 1-based anchor coordinates (Anchors.lua:60-67) */
  int pad_;
} frcnn_candidate;

/* One winner of Detector.lua:140 ({p, a, r, l, r2, class, confidence}) */
typedef struct frcnn_detection {
  double r[4];    /* proposal rect */
  double r2[4];   /* refined rect (Detector.lua:107) */
  float p;        /* RPN foreground log-probability */
  float confidence; /* class log-probability (Detector.lua:112) */
  int cls;        /* 1-based class index (Detector.lua:111) */
  int layer, aspect, y, x; /* anchor a */
  int image;      /* 0-based index of the frame inside the batch */
} frcnn_detection;

/* One training example of objective.lua:91-140: a positive {anchor, roi} or a negative {anchor} as produced by
 * BatchIterator:nextTraining (from Anchors:findPositive / sampleNegative: frcnn_find_positive / frcnn_sample_negative). */
typedef struct frcnn_example {
  double anchor[4];     /* anchor rect {minX, minY, maxX, maxY} (Anchors:get) */
  double roi[4];        /* ground-truth rect roi.rect (positives only) */
  float reg_target[4];  /* Anchors.inputToAnchor(anchor, roi.rect), the fp32 FloatTensor of objective.lua:110 */
  int layer, aspect, y, x; /* 1-based anchor index {layer, aspect, index[2], index[3]} */
  int class_index;      /* roi.class_index, 1-based (positives only) */
  int pad_;
} frcnn_example;

/* ---- lifetime -------------------------------------------------------------------------------------------- */
int frcnn_version(void);
/* replaces cutorch.setDevice + :cuda() setup (main.lua:52, Detector.lua:13-14).  stream: cudaStream_t or NULL. */
int frcnn_create(frcnn_ctx** out, int device, void* stream);
int frcnn_destroy(frcnn_ctx* ctx);
const char* frcnn_last_error(const frcnn_ctx* ctx);
int frcnn_device_info(frcnn_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched on ctx so far (bench.py's gpu_launches) */
int64_t frcnn_launch_count(const frcnn_ctx* ctx);

/* ---- model plan: replaces create_model (models/model_utilities.lua:126-136) -------------------------------- */
/* dropout_eval_scale < 0 selects the Torch7 SpatialDropout v1 default (1 - p). */
int frcnn_model_plan(frcnn_ctx* ctx, const frcnn_block_desc* blocks, int n_blocks, const frcnn_head_desc* heads,
                     int n_heads, const frcnn_fc_desc* fcs, int n_fcs, int class_count, int roi_kh, int roi_kw,
                     const double* scales, int n_scales, float dropout_eval_scale);
/* Parameter tensors in bind order.  Shapes are Torch's: conv [Cout][Cin][kH][kW], linear [out][in]. */
int frcnn_param_count(const frcnn_ctx* ctx);
int frcnn_param_info(const frcnn_ctx* ctx, int index, char* name, int name_cap, int64_t* numel);
/* Borrows device pointers into Torch's flat weight buffer (utilities.lua:136-147); fp32. */
int frcnn_bind_params(frcnn_ctx* ctx, const float* const* params_dev, int n);
/* Re-packs the bound fp32 weights into the bf16 tensor-core layouts; call after every optimiser step
 * (main.lua:133) or restore (main.lua:97). */
int frcnn_pack_weights(frcnn_ctx* ctx);

/* ---- geometry: Localizer.lua / Anchors.lua (host-side, exact double arithmetic) ---------------------------- */
/* which: 0..n_heads-1 = Localizer.new(pnet.outnode.children[which+1]) (Anchors.lua:11); n_heads = the ROI
 * localizer (Detector.lua:12, objective.lua:22).  layers6: rows {kW,kH,dW,dH,padW,padH} (Localizer.lua:28-36). */
int frcnn_localizer_layers(const frcnn_ctx* ctx, int which, int* layers6, int cap_layers, int* n_layers);
/* Localizer:inputToFeatureRect (Localizer.lua:41-67) */
int frcnn_input_to_feature_rect(const frcnn_ctx* ctx, int which, const double rect[4], double out[4]);
/* Localizer:featureToInputRect (Localizer.lua:69-79) */
int frcnn_feature_to_input_rect(const frcnn_ctx* ctx, int which, const double rect[4], double out[4]);
/* The same with the Lua methods' optional `layer_index` argument (Localizer.lua:41-42,69-70): only the first
 * layer_index layers of the localizer take part; 0 (or the layer count) = all. */
int frcnn_input_to_feature_rect_upto(const frcnn_ctx* ctx, int which, int layer_index, const double rect[4], double out[4]);
int frcnn_feature_to_input_rect_upto(const frcnn_ctx* ctx, int which, int layer_index, const double rect[4], double out[4]);
/* Anchors.__init LUTs (Anchors.lua:15-57): w_lut/h_lut are [n_scales][3][200][2] fp32 */
int frcnn_anchors_build(frcnn_ctx* ctx, float* w_lut_host, float* h_lut_host);

/* ---- pnet: replaces pnet:forward (Detector.lua:33, objective.lua:71) --------------------------------------- */
/* img_dev: [n][3][h][w] fp32.  out_dev[0..n_heads-1]: [n][18][hi][wi] fp32; out_dev[n_heads]: [n][C][hf][wf] fp32
 * (Torch layouts).  Any out_dev entry may be NULL to skip that export.  Evaluate mode. */
int frcnn_pnet_forward(frcnn_ctx* ctx, const float* img_dev, int n, int h, int w, float* const* out_dev);
/* Shapes of the outputs above for an h x w input: dims[i] = {channels, height, width}, i = 0..n_heads */
int frcnn_pnet_output_dims(const frcnn_ctx* ctx, int h, int w, int* dims3);

/* ---- training: replaces pnet:training() / pnet:forward / pnet:backward (objective.lua:47,71,189) ------------- */
/* Gradient views in frcnn_param_info order (the flat gradient CudaTensor of utilities.lua:136-147).  Gradients are
 * ACCUMULATED (the reference zeroes the flat buffer once per batch, objective.lua:49).  Entries of non-learnable
 * parameters (BatchNorm running statistics) are never written. */
int frcnn_bind_grads(frcnn_ctx* ctx, float* const* grads_dev, int n);
/* Channels of every nn.SpatialDropout layer in trunk order (model_utilities.lua:10-12). */
int frcnn_dropout_layers(const frcnn_ctx* ctx, int* channels, int cap, int* n_layers);
/* Training-mode forward: SpatialDropout draws one Bernoulli(1 - p) mask per (image, channel) without rescale
 * (SURVEY Q5); masks_dev[i] ([n][channels_i] fp32 of 0 / 1) injects the masks of dropout layer i (NULL = draw from
 * seed).  Keeps what frcnn_pnet_backward needs (activations, pooling winners, split-K slices of the heads). */
int frcnn_pnet_forward_train(frcnn_ctx* ctx, const float* img_dev, int n, int h, int w, float* const* out_dev,
                             const float* const* masks_dev, uint64_t seed);
/* pnet:backward(img, delta_outputs): d_out_dev[0..n_heads-1]: [n][18][hi][wi] fp32; d_out_dev[n_heads]: [n][C][hf][wf]
 * fp32 (ROI-pool gradients, objective.lua:184); NULL entries are treated as zero.  dgrad / wgrad run on the
 * tcgen05 conv kernel with bf16 gradient maps and fp32 accumulation.  PReLU slopes must be > 0. */
int frcnn_pnet_backward(frcnn_ctx* ctx, const float* const* d_out_dev);

/* One iteration of the per-image loop of lossAndGradient (objective.lua:65-198) for ONE frame img_dev [3][h][w]:
 * pnet forward (training) -> RPN criteria on the listed anchors (CrossEntropy on the fg/bg pair, 10 * SmoothL1) ->
 * ROI pooling of the ground-truth rects (positives) / anchor rects (negatives) -> cnet forward (training: BatchNorm
 * batch statistics, Dropout v2) -> 10 * SmoothL1 on the positives' bbox outputs + mean ClassNLL -> cnet backward ->
 * ROI-pool backward -> pnet backward.  Parameter gradients are ACCUMULATED into the views given to
 * frcnn_bind_grads; the caller zeroes them per batch and divides by cls_count (objective.lua:49,200).
 * Examples must already be cleaned (cleanAnchors, objective.lua:32-43).  losses_host: {sum CE, 10 * sum SmoothL1,
 * 10 * SmoothL1 sum of the detection stage, mean NLL} of this frame.
 * pnet_masks_dev: per SpatialDropout layer [1][C] or NULL; cnet_masks_dev: per class layer [n_pos + n_neg][n] of
 * 0 / 1 or NULL (= drawn from seed). */
int frcnn_train_image(frcnn_ctx* ctx, const float* img_dev, int h, int w, const frcnn_example* pos_host, int n_pos,
                      const frcnn_example* neg_host, int n_neg, const float* const* pnet_masks_dev,
                      const float* const* cnet_masks_dev, uint64_t seed, float losses_host[4]);
/* The same for n frames of ONE size in one call (img_dev [n][3][h][w]): pnet forward (training) and pnet:backward run
 * once over the whole batch, the per-image stages (criteria, ROI pooling, cnet with its per-image BatchNorm statistics,
 * objective.lua:91-185) frame by frame in between; one host synchronisation.  pos_host / neg_host: n pointers to the
 * frames' example lists, n_pos / n_neg their lengths, seeds one per frame, losses_host [n][4].  Gradients accumulate
 * exactly as n calls of frcnn_train_image would (same sums; the tensor-core reductions differ in order).
 * Both calls return as soon as the losses are on the host (they are final before pnet:backward starts); the gradient
 * accumulation may still be running on the context's stream.  Work queued on that stream, or on the legacy default
 * stream (Torch's; it synchronises with the context's blocking stream), is ordered behind it as in any CUDA program;
 * a host-side reader of the gradient calls frcnn_synchronize first.  FRCNN_TRAIN_SYNC=1 waits for the whole step. */
int frcnn_train_batch(frcnn_ctx* ctx, const float* img_dev, int n, int h, int w, const frcnn_example* const* pos_host,
                      const int* n_pos, const frcnn_example* const* neg_host, const int* n_neg,
                      const float* const* pnet_masks_dev, const uint64_t* seeds, float* losses_host);
/* Blocks until everything queued on the context's stream has finished (cutorch.synchronize() for this context). */
int frcnn_synchronize(frcnn_ctx* ctx);

/* cnet:forward(cinput) in training mode + the detection-stage criteria + cnet:backward (objective.lua:164-179).
 * x_dev: [R][kh*kw*C] fp32 (reference ordering), the first n_pos rows are positives; crtarget_dev [R][4];
 * cctarget_dev [R] 0-based class targets (background = class_count).  dx_dev (optional) receives post_roi_delta
 * [R][kh*kw*C] fp32; parameter gradients are accumulated; the BatchNorm running statistics are updated (momentum
 * 0.1).  losses_host: {10 * SmoothL1 sum over the positives' bbox outputs, mean ClassNLL}. */
int frcnn_cnet_train_step(frcnn_ctx* ctx, const float* x_dev, int R, int n_pos, const float* crtarget_dev,
                          const int32_t* cctarget_dev, const float* const* masks_dev, uint64_t seed, float* dx_dev,
                          float losses_host[2]);

/* cnet:forward(cinput) in TRAINING mode (objective.lua:164: BatchNormalization batch statistics + running-statistics update,
 * Dropout v2 masks drawn from `seed` or injected through masks_dev[layer] [R][n]) and, separately, cnet:backward(cinput,
 * {crdelta, ccdelta}) (objective.lua:179) with the caller's gradients wrt the two outputs -- d_reg_dev [R][4] and d_cls_dev
 * [R][class_count+1] (wrt the log-softmax output) -- for hosts that keep their own criteria (the unmodified objective.lua).
 * The backward pass uses the state the forward call left in the context; dx_dev (optional) receives post_roi_delta
 * [R][kh*kw*C]; parameter gradients are accumulated. */
int frcnn_cnet_forward_train(frcnn_ctx* ctx, const float* x_dev, int R, const float* const* masks_dev, uint64_t seed,
                             float* reg_dev, float* cls_dev);
int frcnn_cnet_backward(frcnn_ctx* ctx, const float* d_reg_dev, const float* d_cls_dev, float* dx_dev);

/* ---- RPN decode: replaces the per-anchor Lua loop Detector.lua:36-66 ------------------------------------- */
/* heads_dev[i]: [18][hi][wi] fp32 of ONE image.  Writes the ordered match list (layer, y, x, aspect order) to
 * cand_host and its length to n_cand.  threshold = 0.95 in the reference (Detector.lua:54). */
int frcnn_rpn_decode(frcnn_ctx* ctx, const float* const* heads_dev, int h, int w, double threshold,
                     frcnn_candidate* cand_host, int cap, int* n_cand);

/* ---- NMS: replaces the global nms(boxes, overlap, scores) (nms.lua:23-102) ---------------------------------- */
/* boxes_host: n rows of row_stride floats {minX,minY,maxX,maxY,...}.  pick_host receives 0-based indices in pick
 * order (the Lua shim adds 1), n_pick their number.  order_col is 0-based. */
int frcnn_nms(frcnn_ctx* ctx, const float* boxes_host, int64_t n, int64_t row_stride, float overlap, int order_mode,
              int order_col, int64_t* pick_host, int64_t* n_pick);
/* Same on device buffers (boxes_dev fp32 row-major, pick_dev int64[n], n_pick_dev int64[1]). Asynchronous. */
int frcnn_nms_dev(frcnn_ctx* ctx, const float* boxes_dev, int64_t n, int64_t row_stride, float overlap,
                  int order_mode, int order_col, int64_t* pick_dev, int64_t* n_pick_dev);
/* Per-class NMS of Detector.lua:125-136 in one call: segment s is rows seg_offsets[s]..seg_offsets[s+1]-1.
 * pick (segment-local 0-based indices, pick order) is written at pick[seg_offsets[s]...]; counts[s] = #picks. */
int frcnn_nms_segmented(frcnn_ctx* ctx, const float* boxes_host, int64_t row_stride, const int64_t* seg_offsets_host,
                        int n_seg, float overlap, int order_mode, int order_col, int64_t* pick_host,
                        int64_t* counts_host);
int frcnn_nms_segmented_dev(frcnn_ctx* ctx, const float* boxes_dev, int64_t n_total, int64_t row_stride,
                            const int64_t* seg_offsets_host, int n_seg, float overlap, int order_mode, int order_col,
                            int64_t* pick_dev, int64_t* counts_dev);

/* ---- ROI pooling: replaces extract_roi_pooling_input + nn.SpatialAdaptiveMaxPooling (objective.lua:5-13,30;
 *      Detector.lua:96-97) ------------------------------------------------------------------------------------ */
/* fmap_dev: [C][H][W] fp32 (Torch layout); rects_host: R input-space rects (doubles).  out_dev: [R][C*kh*kw] fp32,
 * element c*kh*kw + by*kw + bx; argmax_dev (optional): flat index y*W + x into the feature plane (0-based). */
int frcnn_roi_pool_forward(frcnn_ctx* ctx, const float* fmap_dev, int C, int H, int W, const double* rects_host, int R,
                           float* out_dev, int32_t* argmax_dev);

/* The `amp` module itself, nn.SpatialAdaptiveMaxPooling(kw, kh) (objective.lua:30,118-119,138-139,183-184;
 * Detector.lua:14,97), for callers that keep the reference's per-ROI loop: forward on a [C][h][w] view with element strides
 * (stride_c, stride_h, stride_w) -- the crop extract_roi_pooling_input returns is a non-contiguous view of the feature map --
 * writes out_dev [C][kh][kw] and, when idx_dev is non-NULL, the module's `indices` [C][kh][kw]: the winner's position
 * y*w + x inside the view, stored as float (objective.lua only clones the field and puts it back).  Backward zero-fills
 * dx_dev [C][h][w] (contiguous, the module's gradInput) and adds every output gradient at its winner. */
int frcnn_adaptive_maxpool_forward(frcnn_ctx* ctx, const float* x_dev, int C, int h, int w, int64_t stride_c, int64_t stride_h,
                                   int64_t stride_w, int kh, int kw, float* out_dev, float* idx_dev);
int frcnn_adaptive_maxpool_backward(frcnn_ctx* ctx, const float* dout_dev, const float* idx_dev, int C, int h, int w, int kh,
                                    int kw, float* dx_dev);

/* ---- cnet: replaces cnet:forward (Detector.lua:101, objective.lua:164), evaluate mode ---------------------- */
/* x_dev: [R][kh*kw*C] fp32 (reference ordering).  reg_dev: [R][4]; cls_dev: [R][class_count+1] log-softmax. */
int frcnn_cnet_forward(frcnn_ctx* ctx, const float* x_dev, int R, float* reg_dev, float* cls_dev);

/* ---- whole pipeline: replaces Detector:detect (Detector.lua:17-141) ------------------------------------- */
/* img_host: [n][3][h][w] fp32 in host memory (copied to the device inside the call).  det_host receives the
 * winners of all frames grouped by (image, class ascending) and, inside a class, in pick order. */
int frcnn_detect(frcnn_ctx* ctx, const float* img_host, int n, int h, int w, frcnn_detection* det_host, int cap,
                 int* n_det);
/* Same with the frames already resident in device memory. */
int frcnn_detect_dev(frcnn_ctx* ctx, const float* img_dev, int n, int h, int w, frcnn_detection* det_host, int cap,
                     int* n_det);
/* Stage statistics of the last detect call: {matches, candidates after NMS, classified (non-bg, conf>0.2),
 * winners} summed over the batch. */
/* The same call in two halves, for a host loop that keeps several frames in flight (the loop of main.lua:198-206
 * over a stream of frames).  frcnn_detect_begin copies the frames (img: host memory, or device memory with
 * img_on_device != 0) into the context's staging buffer and enqueues the whole launch sequence on the context's
 * stream without waiting; frcnn_detect_end waits for it and returns the winners exactly as frcnn_detect does.
 * One detection may be in flight per context (FRCNN_E_STATE otherwise); contexts are independent (own stream,
 * workspaces, CUDA graph), so K contexts holding the same parameters keep K frames in flight: the few-CTA stages of
 * one frame (NMS, finalize) then overlap the convolutions of the next.  Results are identical to frcnn_detect. */
int frcnn_detect_begin(frcnn_ctx* ctx, const float* img, int img_on_device, int n, int h, int w);
int frcnn_detect_end(frcnn_ctx* ctx, frcnn_detection* det_host, int cap, int* n_det);
int frcnn_detect_stats(const frcnn_ctx* ctx, int64_t stats[4]);
/* Thresholds of Detector.lua:54,81,115,133; defaults 0.95, 0.25, 0.2, 0.1. */
int frcnn_set_detect_thresholds(frcnn_ctx* ctx, double fg_prob, float nms_proposals, double class_prob,
                                float nms_classes);
/* Device-side time of the kernel groups of the last detect call when profiling was enabled with
 * frcnn_set_profiling(ctx, 1): ms[0]=trunk+heads convs, [1]=decode+nms, [2]=roi pool, [3]=cnet, [4]=final nms,
 * [5]=total. */
int frcnn_set_profiling(frcnn_ctx* ctx, int enable);
/* frcnn_detect / frcnn_detect_dev replay their fixed kernel sequence from a CUDA graph from the third call with the
 * same image pointer, shape and thresholds on (default on; profiling mode always runs eagerly). */
int frcnn_set_graph_replay(frcnn_ctx* ctx, int enable);
/* Launch schedule of pnet:forward in evaluate mode (frcnn_pnet_forward, frcnn_detect*).  FRCNN_SCHED_LATENCY
 * (default): every stage is cut so that ONE frame batch fills the machine -- the anchor-head convolutions run as
 * split-K units on all SMs followed by a tail kernel (fastest single synchronous call).  FRCNN_SCHED_THROUGHPUT: the
 * least SM time per frame -- the four AnchorNetworks (model_utilities.lua:29-35) run as unsplit units with bias +
 * PReLU + the 1x1 convolution fused into the conv epilogue (no slice workspace, no tail launch); meant for several
 * frames in flight (frcnn_detect_begin / _end on several contexts), where the other frames fill the machine.  Both
 * compute the same fp32 sums in a different, fixed order (results agree to fp32 rounding). */
enum { FRCNN_SCHED_LATENCY = 0, FRCNN_SCHED_THROUGHPUT = 1 };
int frcnn_set_schedule(frcnn_ctx* ctx, int schedule);
/* 16-bit operand format of the tensor-core convolutions / Linear layers in EVALUATE mode (frcnn_pnet_forward,
 * frcnn_cnet_forward, frcnn_detect*).  The reference computes in fp32 (cunn); FRCNN_PREC_FP16 (default) rounds operands
 * to 11 significand bits -- the precision of tf32 at the bf16 tensor-core rate; outputs saturate at +-65504 -- and
 * reproduces the fp32 path's discrete decisions (matches / candidates / winners) almost everywhere; FRCNN_PREC_BF16
 * (8 bits) is kept for comparison.  Accumulation is fp32 either way.  Training entry points always use bf16 operands:
 * gradient maps need its exponent range.  Measured agreement with the fp32 path: DESIGN.md 4, profiles/r2_precision*. */
enum { FRCNN_PREC_BF16 = 0, FRCNN_PREC_FP16 = 1 };
int frcnn_set_eval_precision(frcnn_ctx* ctx, int precision);
int frcnn_last_timings(const frcnn_ctx* ctx, float ms[6]);
/* Profiling mode also brackets every launch of the tcgen05 conv/GEMM kernel with a CUDA event pair on the ctx
 * stream: summed device time, summed algorithmic FLOPs (2*M*N*K of the un-padded problems) and launch count of
 * the last detect call (bench.py's roofline figure). */
int frcnn_last_conv_profile(const frcnn_ctx* ctx, float* ms, double* flops, int* launches);

/* ---- frame normalisation: replaces the tail of load_image / BatchIterator:processImage (utilities.lua:211-212,
 *      BatchIterator.lua:86,146-161) -- the step before pnet:forward, SURVEY 8f row 2 ------------------------------ */
/* In place on img_dev [3][h][w] fp32 (already resized): optional image.rgb2yuv; per-channel centering (x - mean);
 * per-channel scaling (x / std, unbiased, skipped when std <= 1e-8); nn.SpatialContrastiveNormalization(1,
 * image.gaussian1D(contrastive_width)) on channel 1 (contrastive_width = 0: none; config/duplo.lua:6 uses 7).
 * `image` / `nn` are un-vendored: algorithms restated in oracle/preprocess.py (tolerance 1e-5, parity unpinned). */
int frcnn_normalize_frame(frcnn_ctx* ctx, float* img_dev, int h, int w, int rgb2yuv, int centering, int scaling,
                          int contrastive_width);

/* find_target_size (utilities.lua:188-204): the frame size BatchIterator:processImage / main.lua:199 resize to.  Host only. */
int frcnn_find_target_size(int orig_w, int orig_h, double target_smaller_side, double max_pixel_size, int* w, int* h);
/* image.scale(img, dst_w, dst_h), default 'bilinear' mode (BatchIterator.lua:49-52, main.lua:200): src_dev [c][src_h][src_w]
 * fp32 -> dst_dev [c][dst_h][dst_w].  Rows first, then columns; enlarging an axis interpolates linearly with scale
 * (src - 1) / (dst - 1), shrinking averages the source interval with fractional end weights (torch/image's
 * Main_scaleLinear_rowcol, un-vendored: restated in oracle/preprocess.py, bit-exact against that restatement, parity
 * unpinned).  src_dev may be the frame still in page-locked host memory mapped into the device (the copy then rides in
 * the first pass). */
int frcnn_scale_frame(frcnn_ctx* ctx, const float* src_dev, int c, int src_h, int src_w, float* dst_dev, int dst_h, int dst_w);

/* ---- anchor labelling: replaces Anchors:findPositive / Anchors:sampleNegative (Anchors.lua:147-235; called per
 *      training image by BatchIterator.lua:200-225) -- SURVEY 8f row 1 ------------------------------------------ */
typedef struct frcnn_anchor_ref {
  int layer, aspect, y, x; /* the arguments of Anchors:get (Anchors.lua:60-67), 1-based */
} frcnn_anchor_ref;
/* rois_host: n_rois ground-truth rects {minX, minY, maxX, maxY} (doubles); clip_host: 4 doubles or NULL.  Writes the
 * match list in the reference's order -- ROI by ROI; inside a ROI the anchors with IoU > pos_threshold in
 * enumeration order (scale, aspect, y, x), or, when there is none and include_best, the best set of
 * Anchors.lua:171-186 -- as (anchor, roi index 0-based) pairs.  Bit-exact vs the reference's loops: float32 LUT
 * entries read as doubles, Rect.IoU in double. */
int frcnn_find_positive(frcnn_ctx* ctx, const double* rois_host, int n_rois, const double* clip_host,
                        double pos_threshold, double neg_threshold, int include_best, frcnn_anchor_ref* out_host,
                        int* out_roi_host, int cap, int* n_out);
/* The nearby-aversion list of BatchIterator.lua:206-217 (before shuffle_n): for every positive anchor p (in list order)
 * Anchors:findNearby(p:center()) (Anchors.lua:69-84: the anchors whose cell centre falls into p's 16-pixel bin on both axes,
 * in the order of the Lua bin tables = scale, aspect, y cell, x cell ascending), kept when Rect.IoU(p, a) < neg_threshold.
 * out_pos_host[k] = 0-based index of the positive entry k belongs to. */
int frcnn_find_nearby_negative(frcnn_ctx* ctx, const frcnn_anchor_ref* pos_host, int n_pos, double neg_threshold,
                               frcnn_anchor_ref* out_host, int* out_pos_host, int cap, int* n_out);
/* Anchors:sampleNegative(image_rect, roi_list, neg_threshold, count).  The reference draws three torch.random()
 * values per trial (range, x, y); the caller supplies that stream: rnd_host holds 3 * n_trials uint32 values.  Returns
 * the accepted anchors in order, how many trials the loop consumed (so the caller can keep its generator in step) and
 * whether the loop's own stopping rule fired (count reached, or 500 consecutive rejections) before the supplied
 * stream ran out (*finished == 0: call again with more random numbers, count reduced by *n_out and retry_in set to
 * *retry_out, the run of rejections the loop was in; retry_in = 0 for a fresh loop; retry_out may be NULL). */
int frcnn_sample_negative(frcnn_ctx* ctx, const double image_rect[4], const double* rois_host, int n_rois,
                          double neg_threshold, int count, const uint32_t* rnd_host, int n_trials, int retry_in,
                          frcnn_anchor_ref* out_host, int cap, int* n_out, int* trials_consumed, int* finished,
                          int* retry_out);

/* ---- optimiser step on the flat buffers: replaces gradient:div(cls_count) (objective.lua:200) + optim.rmsprop
 *      (main.lua:122,133) -- SURVEY 8f row 3 ------------------------------------------------------------------ */
/* One fused pass over the three flat device buffers of n floats (16-byte aligned): g /= grad_div (1 = no division);
 * [g += weight_decay * w]; m = alpha * m + (1 - alpha) * g * g; w += -lr * g / (sqrt(m) + epsilon) -- the statement
 * sequence of optim.rmsprop (an un-vendored dependency; restated, see oracle/optim.py), every operation rounded to
 * fp32 on its own.  state_m_dev is rmsprop_state.m (zero before the first step).  Asynchronous on the ctx stream;
 * call frcnn_pack_weights afterwards.
 * The scalars are doubles as in Lua and are rounded to fp32 where TH does: -lr, alpha, (1.0 - alpha) [evaluated in
 * double first], epsilon, weight_decay, grad_div. */
int frcnn_rmsprop_step(frcnn_ctx* ctx, float* weights_dev, float* gradient_dev, float* state_m_dev, int64_t n,
                       double grad_div, double lr, double alpha, double epsilon, double weight_decay);

/* ---- data-parallel training: the gradient all-reduce of SURVEY 8e (objective.lua:189,200 sum the per-image gradients
 *      into the flat `gradient` and divide once by the example count; with the frames of a batch sharded over GPUs the sum
 *      runs over the ranks) ---------------------------------------------------------------------------------------------
 * NCCL is loaded at run time (libnccl.so.2, or the path in FRCNN_NCCL_LIB); failures return FRCNN_E_NCCL.  One
 * communicator rank per context.  Two ways to form the communicator:
 *   one process per GPU (torchrun, MPI): rank 0 calls frcnn_dp_unique_id and distributes the 128 bytes by any means; every
 *     rank calls frcnn_dp_init_rank on its context;
 *   one process, several GPUs (a LuaJIT host is single-threaded, main.lua:52): frcnn_dp_init_all over one context per
 *     device (ncclCommInitAll).
 * frcnn_dp_allreduce sums IN PLACE, across the ranks, every gradient view bound with frcnn_bind_grads (adjacent views --
 * nn.Module.flatten's layout -- go out as single calls) and `n_counters` floats of counters_dev[i] (the example / loss
 * counters of objective.lua:196-200; NULL / 0 = none), for the `n` contexts this thread drives (n = 1 with one process per
 * GPU), inside one NCCL group.  The work runs on a side stream ordered after the context's stream; the context's stream
 * waits for the result, so the optimiser step may simply be enqueued next.  The caller then divides by the summed counter.
 * With frcnn_dp_set_overlap(ctx, 1) (one process per GPU only) frcnn_train_batch / frcnn_pnet_backward send each bucket
 * -- cnet, anchor networks, conv block 4 ... 1, the order in which the backward pass finishes them -- as soon as it is
 * final, so that only the last bucket's transfer is exposed; the caller promises that the call is the step's last
 * accumulation into the gradient.  frcnn_dp_allreduce then sends what is left and joins the streams. */
int frcnn_dp_unique_id(char id_out[128]);
int frcnn_dp_init_rank(frcnn_ctx* ctx, const char id[128], int rank, int nranks);
int frcnn_dp_init_all(frcnn_ctx* const* ctxs, int n);
int frcnn_dp_set_overlap(frcnn_ctx* ctx, int enable);
int frcnn_dp_allreduce(frcnn_ctx* const* ctxs, int n, float* const* counters_dev, int n_counters);
/* rank / size of the context's communicator (nranks = 0: none), the loaded NCCL's version, bytes all-reduced so far */
int frcnn_dp_info(const frcnn_ctx* ctx, int* rank, int* nranks, int* nccl_version, int64_t* bytes_reduced);

/* Diagnostic (training parity tests): the pooled output of conv block `block` (1-based) of the LAST pnet forward on this
 * context as fp32 [n][C][h][w] (Torch layout); dims3 receives {C, h, w}.  out_dev may be NULL to query the shape. */
int frcnn_block_output(frcnn_ctx* ctx, int block, float* out_dev, int* dims3);

/* ---- low-level conv / GEMM entry (tests, roofline measurement) ---------------------------------------- */
/* y = prelu(conv(x, w) + bias) * scale on NHWC bf16 activations (passed as uint16 bit patterns).
 * x_dev: [n][h][w][cin]; w_dev: fp32 Torch layout [cout][cin][k][k]; out_dev: [n][ho][wo][cout] bf16, or with
 * pool != 0 the 2x2 stride-2 ceil-mode max-pooled map [n][ceil(ho/2)][ceil(wo/2)][cout] (model_utilities.lua:23).
 * splits > 1 exercises the split-K fp32-atomic path; bn in {0 (auto), 64, 128, 192, 256}; mt in {0 (auto), 1, 2} =
 * 128-row sub-tiles per CTA of the tap-per-box kernel, 11 / 12 = the halo-tile kernel with 1 / 2 sub-tiles, 21 / 22 = the
 * halo-tile kernel on CTA pairs (cta_group::2, M = 256 MMAs) with 1 / 2 sub-tiles per CTA.  elapsed_ms (optional) receives the device time of `iters` back-to-back launches of
 * the conv kernel alone. */
int frcnn_conv_bf16(frcnn_ctx* ctx, const uint16_t* x_dev, const float* w_dev, const float* bias_dev,
                    const float* prelu_dev, float scale, int n, int h, int w, int cin, int cout, int k, int pad,
                    int splits, int bn, int mt, int pool, uint16_t* out_dev, int iters, float* elapsed_ms);
/* Backward primitives of nn.SpatialConvolution (pnet:backward, objective.lua:189), stride 1, NHWC bf16 tensors:
 * dgrad: dx[n][h][w][cin] = conv_transpose(dy[n][ho][wo][cout], w); wgrad: dw (fp32, Torch layout [cout][cin][k][k])
 * += sum over pixels of dy (x) x.  Both run on the tcgen05 conv kernel (dgrad: flipped filters, padding k-1-pad;
 * wgrad: K = output pixels over planar copies of x and dy, one filter tap per work unit, TMA reduce-add). */
int frcnn_conv_dgrad_bf16(frcnn_ctx* ctx, const uint16_t* dy_dev, const float* w_dev, int n, int h, int w, int cin,
                          int cout, int k, int pad, uint16_t* dx_dev);
int frcnn_conv_wgrad_bf16(frcnn_ctx* ctx, const uint16_t* x_dev, const uint16_t* dy_dev, int n, int h, int w, int cin,
                          int cout, int k, int pad, float* dw_dev);
/* The fused first layer: img_dev [n][3][h][w] fp32 (Torch layout), w_dev [cout = 64][3][3][3] fp32; output as above. */
int frcnn_conv_first(frcnn_ctx* ctx, const float* img_dev, const float* w_dev, const float* bias_dev,
                     const float* prelu_dev, float scale, int n, int h, int w, int cout, int pad, int pool,
                     uint16_t* out_dev, int iters, float* elapsed_ms);
/* Its weight gradient (accGradParameters of the first nn.SpatialConvolution inside pnet:backward, objective.lua:189):
 * dw_dev [64][3][3][3] fp32 += sum over pixels of dy_dev [n][h][w][64] (bf16, NHWC) x the padded frame (pad = 1: the
 * output has the frame's size, as in both reference models; anything else is rejected).  Warp-level
 * tensor-core kernel with the frame split hi/lo, so the frame enters at 16 mantissa bits. */
int frcnn_conv_first_wgrad(frcnn_ctx* ctx, const uint16_t* dy_dev, const float* img_dev, int n, int h, int w, int pad,
                           float* dw_dev, int iters, float* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* FRCNN_B200_H */
