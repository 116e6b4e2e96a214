// Internal interface of the training-side kernels (train_kernels.cu).
#pragma once
#include <algorithm>

#include "common.h"

namespace frcnn {

void launch_dropout_mask(float* mask, int n, float p, uint64_t seed, uint32_t layer, cudaStream_t st);

// Row ranges of the frames of one training batch: frame f owns rows [off[f], off[f] + R[f]) of every per-row buffer, its
// first n_pos[f] rows are the positives.  Passed by value to the per-frame stages (criteria, ROI pooling, cnet's
// BatchNorm chains), which run ONE launch over all frames with the frame on blockIdx.y or looked up from the row.
constexpr int MAX_TRAIN_FRAMES = 64;
struct FrameList {
  int nf;
  int off[MAX_TRAIN_FRAMES], R[MAX_TRAIN_FRAMES], n_pos[MAX_TRAIN_FRAMES];
  int max_R() const { int m = 0; for (int f = 0; f < nf; ++f) m = R[f] > m ? R[f] : m; return m; }
  int rows() const { return nf > 0 ? off[nf - 1] + R[nf - 1] : 0; }
};
// (an empty frame shares its offset with the next one and is stepped over)
__device__ __forceinline__ int frame_of_row(const FrameList& fl, int r) {
  int f = 0;
  while (f + 1 < fl.nf && r >= fl.off[f + 1]) ++f;
  return f;
}
struct FrameSeeds { uint64_t seed[MAX_TRAIN_FRAMES]; };
// masks of all frames: frame f draws R[f] * per_row values at mask + off[f] * per_row from seeds.seed[f]
void launch_dropout_mask_frames(float* mask, const FrameList& fl, int per_row, float p, const FrameSeeds& seeds, uint32_t layer,
                                cudaStream_t st);
void launch_unpool_prelu_bwd(const float* g, const uint8_t* arg, const bf16* yp, const float* slope, const float* mask, bf16* dpre,
                             float* dbias, float* dslope, int N, int H, int W, int C, int num_sms, cudaStream_t st);
void launch_prelu_bwd(const bf16* d, bf16* out, const bf16* y, const float* slope, const float* mask, float* dbias, float* dslope, int N, int H, int W,
                      int C, int num_sms, cudaStream_t st);

struct HeadTailBwd {
  const float* d_out;   // [N][18][HW] fp32 (Torch layout)
  const float* ws;      // [splits][npix][256] forward split-K slices of the k x k conv
  int splits;
  long slice_stride, npix;
  int HW;
  const float *bias, *prelu, *w2;
  bf16* dpre;           // [npix][256] gradient wrt the k x k conv's pre-activation output
  float *dw2, *db2, *db1, *dslope;
  const int* list = nullptr;   // sparse form: the M pixels (flat n * HW + y * hw + x, unique) that can carry gradient;
  int M = 0;                   // dpre is then the COMPACT [M][256] matrix, row r = pixel list[r]
  int ws_compact = 0;          // sparse forward: ws is the compact [M][256] conv output of the listed pixels (one slice)
};
void launch_head_tail_bwd(const HeadTailBwd& H, int num_sms, cudaStream_t st);
// Sparse anchor-head backward (lossAndGradient lists <= 256 anchors per frame, objective.lua:91-140, so delta_outputs is
// zero at all but M pixels of a head): the k x k conv's data and weight gradients become two small GEMMs over the M
// listed pixels instead of two dense convolutions over the map.
//   rows   [M][k*k*Cin] bf16 = the input windows of the listed pixels (im2col rows, tap-major then channel)
//   w_rows [k*k*Cin][Cout] bf16 = the filter as the GEMM operand of  G[M][k*k*Cin] = dpre[M][Cout] x W
//   scatter: dx[n][y + ty][x + tx][ci] += G[r][(ty * k + tx) * Cin + ci]
// forward tail of the listed pixels: out[n][o][y][x] = b2[o] + sum_c w2[o][c] * PReLU(hpre[r][c] + bias[c]), written at
// the listed pixels only (the criteria read nothing else, objective.lua:96-131)
struct HeadTailList {
  const float* hpre;   // [M][256] conv output of the listed pixels (no bias)
  const int* list;
  int M, HW;
  const float *bias, *prelu, *w2, *b2;
  float* out;          // [N][18][HW]
};
void launch_head_tail_list(const HeadTailList& H, cudaStream_t st);
void launch_pack_head_weight_rows(const float* w, bf16* out, int Cout, int Cin, int K, cudaStream_t st);
void launch_head_gather_rows(const bf16* x, const int* list, int M, int hh, int hw, int Hin, int Win, int Cin, int K, bf16* rows,
                             cudaStream_t st);
void launch_head_scatter_rows(const float* g, const int* list, int M, int hh, int hw, int Hin, int Win, int Cin, int K, float* dx,
                              cudaStream_t st);
void launch_first_wgrad(const bf16* dpre, const float* img, float* dw, int N, int H, int W, int pad, int num_sms, cudaStream_t st);
void launch_add_chw_to_nhwc(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);
void launch_check_slopes(const float* const* slopes_dev, int n, int* flag, cudaStream_t st);

}  // namespace frcnn

// ---- objective.lua:91-186: RPN losses, training ROI pooling, cnet training forward / backward ------------------
#include "detect.h"
namespace frcnn {

struct ExampleDev {          // device copy of frcnn_example
  double anchor[4], roi[4];
  float reg_target[4];
  int layer, aspect, y, x, class_index, pad_;
};

struct RpnLossParams {       // every pointer is the batch base: rows by FrameList offsets, maps by frame index
  const ExampleDev* ex;      // per frame: positives first, then negatives
  const float* out[MAX_HEADS];   // [N][18][hh][hw] fp32
  float* d_out[MAX_HEADS];       // zeroed delta_outputs, same shape
  int hh[MAX_HEADS], hw[MAX_HEADS];
  float* crtarget;           // [R][4] regression targets of the detection stage (zeros for negatives)
  int* cctarget;             // [R] 0-based class targets (background = class_count)
  int bg_class;              // 0-based background index
  double* rects;             // [R][4] rect that is ROI-pooled: ground truth (positives) / anchor (negatives)
  float* losses;             // frame f: [8 f + 0] += sum CE, [8 f + 1] += 10 * sum SmoothL1
  int* status;               // set when an example indexes outside its map (cleanAnchors was not applied)
};
void launch_rpn_loss(const RpnLossParams& p, const FrameList& fl, cudaStream_t st);

// ROI pooling of explicit rects with winners (flat feature-plane positions) for the backward scatter; row r pools the
// feature map of its frame (fmap + frame * FH * FW * C)
void launch_roi_pool_train(const bf16* fmap, int FH, int FW, int C, int kh, int kw, const LocalizerDev& loc, const double* rects_dev,
                           const FrameList& fl, bf16* out, int* argmax, int* status, cudaStream_t st);
// dfeat (fp32 NHWC [N][FH][FW][C]) += scatter of d_rows ([rows][bins][C] fp32) to the winners, frame by frame
void launch_roi_pool_bwd(const float* d_rows, const int* argmax, const FrameList& fl, int bins, int C, float* dfeat, long fmap_elems,
                         cudaStream_t st);

struct FcTrainFwd {          // Linear bias (+ BatchNorm, training statistics) + PReLU + Dropout v2 on the GEMM output
                             // (pointers are batch bases; the kernel offsets them by the frame's rows)
  const float* acc;          // [R][n] fp32 GEMM result
  const float *bias, *bn_w, *bn_b, *prelu;
  float *bn_mean, *bn_var;   // running statistics, updated with momentum 0.1 (null: no BatchNorm)
  const float* mask;         // [R][n] 0/1 dropout mask
  float keep_scale;          // 1 / (1 - p)
  float* pre;                // [R][n] value entering the PReLU (after BN)
  float* xhat;               // [R][n] normalised BN input (BatchNorm only)
  float* rstd;               // [nf][n]
  float* stat;               // [nf][2][n] scratch: batch mean / unbiased variance of every frame (BatchNorm, nf > 1)
  bf16* out_bf16;            // [R][n] next GEMM operand (after dropout), or null
  float* out_f32;            // [R][n] same in fp32, or null
  int n;
};
void launch_fc_train_fwd(const FcTrainFwd& p, const FrameList& fl, cudaStream_t st);

struct FcTrainBwd {          // backward of the same chain: d (wrt the dropout output) -> d wrt the Linear output
  const float* d_in;         // [R][n]
  const float *pre, *xhat, *rstd, *bn_w, *prelu, *mask;
  float keep_scale;
  bf16* d_out_bf16;          // [R][n] gradient wrt the Linear output (GEMM operand)
  float *g_bias, *g_bn_w, *g_bn_b, *g_prelu;
  int n;
};
void launch_fc_train_bwd(const FcTrainBwd& p, const FrameList& fl, cudaStream_t st);

struct CnetLossParams {      // objective.lua:166-177 on the two cnet outputs + backward through the output Linear layers
  const float* hidden;       // [R][nin] fp32 input of both branches
  const float *w_reg, *b_reg, *w_cls, *b_cls;
  const float* crtarget;     // [R][4]
  const int* cctarget;       // [R]
  int nin, ncls;
  float* d_hidden;           // [R][nin]
  float* dz;                 // [R][ncls + 4] scratch: gradient wrt the branch outputs (4 reg, then ncls logits)
  float *g_w_reg, *g_b_reg, *g_w_cls, *g_b_cls;
  float* losses;             // frame f: [loss_stride f + 2] += 10 * SmoothL1 sum, [loss_stride f + 3] += mean NLL
  int loss_stride;
  const float* ext_dreg;     // optional [R][4]: gradient wrt the bbox output supplied by the caller (cnet:backward, objective.lua:179)
  const float* ext_dcls;     // optional [R][ncls]: gradient wrt the LOG-SOFTMAX output; with both set no criterion is evaluated
};
void launch_cnet_loss_bwd(const CnetLossParams& p, const FrameList& fl, cudaStream_t st);

// fc weight for the dgrad GEMM: out[k'][o] = w[o][src(k')] bf16, k' = b*C + c <- c*bins + b when permute
void launch_pack_fc_weight_dgrad(const float* w, bf16* out, int nout, int C, int bins, int permute, cudaStream_t st);
// grad[o][src(k')] += dw[o][k'] (inverse of the ROI column permutation)
void launch_wgrad_finish_fc(const float* dw, float* grad, int nout, int C, int bins, int permute, cudaStream_t st);

}  // namespace frcnn
