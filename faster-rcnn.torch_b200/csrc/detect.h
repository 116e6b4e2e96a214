// Internal interface of the detector-side kernels (detect_kernels.cu): RPN decode, ROI pooling, candidate
// refinement / class grouping and winner assembly.
#pragma once
#include "common.h"
#include "nms.h"

namespace frcnn {

static constexpr int MAX_LOC_LAYERS = 24;
static constexpr int MAX_HEADS = 4;  // Detector.lua:38 and Anchors.lua:108 hard-code 4 anchor layers x 3 aspects
static constexpr int LUT_EXTENT = 200;  // Anchors.lua:15

struct LocalizerDev {
  int n;
  int l[MAX_LOC_LAYERS][6];  // kW, kH, dW, dH, padW, padH (Localizer.lua:28-36)
};

struct DecodeParams {
  const float* head[MAX_HEADS];  // [N][18][hh][hw] fp32
  int hh[MAX_HEADS], hw[MAX_HEADS];
  int offs[MAX_HEADS + 1];       // anchor index prefix per layer (layer-major, then y, x, aspect)
  int total;                     // anchors per image
  const float* w_lut;            // [4][3][200][2]
  const float* h_lut;
  double img_w, img_h, threshold;
  int cap;                       // candidate capacity per image
  double* cand_r;                // [N][cap][4]
  float4* cand_box;              // [N][cap]
  float* cand_logp;              // [N][cap]
  int4* cand_anchor;             // [N][cap] {layer, aspect, y, x} 1-based
  int* cand_count;               // [N]   (clamped to cap)
  int* cand_overflow;            // [1]   set when an image produced more than cap matches
  // chained-scan state
  unsigned long long* ticket;    // [N] 64-bit running block tickets
  unsigned long long* status;    // [N][nblocks]
  int nblocks;
};
void launch_rpn_decode(const DecodeParams& p, int N, cudaStream_t st);

struct RoiParams {
  const bf16* fmap;  // [N][FH][FW][C] 16-bit (bf16, or fp16 with f16 != 0)
  int f16;
  int FH, FW, C, kh, kw;
  LocalizerDev loc;
  const double* cand_r;    // [N][cap][4]
  const int* pick;         // [N][cap] candidate indices in pick order
  const int* pick_count;   // [N]
  const int* roi_base;     // [N] exclusive prefix of pick_count
  int cap;
  bf16* out;               // [R_total][kh*kw][C]
  int* roi_img;            // [R_total]
  int* roi_cand;           // [R_total]
  int* status;             // [1] number of degenerate ROIs
  int* roi_base_out;       // [N] written by the prepare kernel (exclusive prefix of pick_count)
  int* roi_total;          // [1] written by the prepare kernel (clamped to total_cap)
  int total_cap;
  int4* roi_rect;          // [R_total] crop {y0, y1, x0, x1} in feature cells, y0 < 0 = degenerate
};
// prepare: ROI row table (image, candidate, crop rect) for every NMS survivor, one thread per survivor;
// pool: persistent CTAs over the rows
void launch_roi_pool_nhwc(const RoiParams& p, int N, int num_sms, cudaStream_t st);
void launch_roi_pool_chw(const float* fmap, int C, int H, int W, const LocalizerDev& loc, const double* rects_dev, int R,
                         int kh, int kw, float* out, int32_t* argmax, int* status, cudaStream_t st);

// nn.SpatialAdaptiveMaxPooling on a strided [C][h][w] view (the `amp` module of objective.lua:30 / Detector.lua:14)
void launch_adaptive_maxpool_fwd(const float* x, int C, int h, int w, long sc, long sh, long sw, int kh, int kw, float* out,
                                 float* idx, cudaStream_t st);
void launch_adaptive_maxpool_bwd(const float* dout, const float* idx, int C, int h, int w, int kh, int kw, float* dx, cudaStream_t st);

struct FinalizeParams {
  const double* cand_r;   // [N][cap][4]
  const float* cand_logp;
  const int4* cand_anchor;
  int cap;
  const int* roi_img;     // [R]
  const int* roi_cand;    // [R]
  const int* roi_total;   // [1]
  const float* reg;       // [R][4]
  const float* cls;       // [R][ncls]
  int ncls;               // class_count + 1, background = ncls (1-based)
  double class_prob;
  double* fin_r2;         // [R][4]
  float4* fin_box;        // [R]
  int* fin_cls;           // [R] 1-based class, 0 when rejected
  float* fin_conf;        // [R]
};
void launch_finalize(const FinalizeParams& p, int R_cap, cudaStream_t st);

struct GroupParams {
  const int* roi_base;    // [N]
  const int* pick_count;  // [N]
  const int* fin_cls;
  const float4* fin_box;
  int cap, n_classes;     // n_classes = class_count (foreground classes)
  float4* gbox;           // [N][cap]
  int* grow;              // [N][cap] roi row
  int* n_pass;            // [N]
  int* overflow;          // set when an image has more NMS survivors than the in-CTA sort holds (8192)
};
void launch_group_by_class(const GroupParams& p, NmsWorkspace* ws, int N, cudaStream_t st);

struct AssembleParams {
  const int* grow;
  const double* cand_r;
  const float* cand_logp;
  const int4* cand_anchor;
  const int* roi_img;
  const int* roi_cand;
  const double* fin_r2;
  const int* fin_cls;
  const float* fin_conf;
  int cap, n_classes;
  frcnn_detection* det;   // [det_cap]
  int det_cap;
  int* n_det;             // [1] total winners (may exceed det_cap: only det_cap are written)
};
void launch_assemble(const AssembleParams& p, NmsWorkspace* ws, int n_seg, cudaStream_t st);

}  // namespace frcnn
