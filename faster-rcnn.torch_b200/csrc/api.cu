// C ABI of the library (include/frcnn_b200.h): context, model plan, weight packing, pnet / cnet forward and the
// fused Detector:detect pipeline.  Host orchestration only -- every arithmetic step is a kernel of this library.
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <memory>

#include "common.h"
#include "detect.h"
#include "nms.h"
#include "train.h"
#include "label.h"

namespace frcnn {

static std::string g_last_error;
void set_global_error(const std::string& msg) { g_last_error = msg; }

struct ParamInfo {
  std::string name;
  int64_t numel;
};

struct ConvLayer {
  int cin, cout, k, pad;
  bool first;       // 3-channel first layer: fused in-kernel im2col (conv_first_kernel)
  bool pooled;      // last conv of a block: the 2x2 ceil max pool is fused into its epilogue
  float scale;      // SpatialDropout evaluate-mode factor (1 if no dropout follows)
  int p_w, p_b, p_prelu;
  bf16* w_packed = nullptr;
  int block;        // trunk block index, or -1 for an anchor-head conv
  // per-shape state
  int hin = 0, win = 0, hout = 0, wout = 0;
  bf16* in = nullptr;
  bf16* out = nullptr;
  ConvLaunch launch;
  // training
  float dropout = 0.f;          // SpatialDropout p applied to this conv's output (first conv of a block), else 0
  float* mask = nullptr;        // [N][cout] Bernoulli(1 - p) mask of the last training forward
  bf16* w_dgrad = nullptr;      // [cin][k][k][cout] flipped filters (transposed convolution)
  float* dw_taps = nullptr;     // [cout][k*k][cin] fp32 wgrad accumulator
  long dgrad_gen = -1;          // weights generation w_dgrad was packed from
  ConvLaunch dgrad, wgrad;
};

struct Head {
  int kW, n, input;
  ConvLayer conv;
  int p_w2, p_b2;
  float* acc = nullptr;      // [splits * N][hh][hw][n] fp32 split-K slices
  float* out = nullptr;      // [N][18][hh][hw] fp32
  ConvLaunch fused;          // throughput schedule: k x k conv + tail in one unsplit halo-kernel unit (EPI_HEAD)
  int hh = 0, hw = 0;
  bf16* dpre = nullptr;      // training: [N][hh][hw][n] gradient wrt the k x k conv's pre-activation output
  bf16* w_rows = nullptr;    // sparse backward: [k * k * cin][n] filter as the operand of the listed-pixel GEMM
  long rows_gen = -1;        // weights generation w_rows was packed from
};

struct FcLayer {
  int nin, nout;
  bool bn;
  int p_w, p_b, p_bn_w, p_bn_b, p_bn_mean, p_bn_var, p_prelu;
  bf16* w_packed = nullptr;
  float* acc = nullptr;      // [R_cap][nout] fp32
  bf16* out_bf16 = nullptr;  // [R_cap][nout]
  float* out_f32 = nullptr;  // last layer only
  ConvLaunch launch;
  float dropout = 0.f;
  // training buffers for up to cw_rows example rows
  float *t_acc = nullptr, *t_pre = nullptr, *t_xhat = nullptr, *t_rstd = nullptr, *t_stat = nullptr, *t_mask = nullptr, *t_din = nullptr, *t_out32 = nullptr;
  bf16 *t_out = nullptr, *t_dy = nullptr, *w_dgrad = nullptr;
  float* dw_taps = nullptr;
  long dgrad_gen = -1;
};

}  // namespace frcnn

using namespace frcnn;

struct frcnn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  std::string err;
  int64_t launches = 0;

  // plan
  bool planned = false;
  std::vector<frcnn_block_desc> blocks;
  std::vector<frcnn_head_desc> head_desc;
  std::vector<frcnn_fc_desc> fc_desc;
  int class_count = 0, roi_kh = 6, roi_kw = 6;
  std::vector<double> scales;
  std::vector<ParamInfo> params;
  std::vector<const float*> bound;
  bool packed = false;
  std::vector<ConvLayer> trunk;
  std::vector<Head> heads;
  std::vector<FcLayer> fcs;
  int p_reg_w = -1, p_reg_b = -1, p_cls_w = -1, p_cls_b = -1;
  int feat_c = 0;
  std::vector<std::vector<std::array<int, 6>>> loc;  // heads..., ROI
  std::vector<float> w_lut, h_lut;
  std::vector<double> cen_x, cen_y;   // [scale][200] cell centres (Anchors.lua:40-41,49-50): the findNearby bins
  float* d_w_lut = nullptr;
  float* d_h_lut = nullptr;
  double* d_cen = nullptr;            // cen_x | cen_y
  LocalizerDev roi_loc;

  // training (pnet:backward): gradient views, per-block winners / gradient accumulators, two scratch gradient maps
  std::vector<float*> grads;
  bool train_ready = false;      // the last pnet forward ran in training mode on the current workspace
  int tws_n = 0, tws_h = 0, tws_w = 0;
  std::vector<void*> tws_allocs;
  std::vector<uint8_t*> pool_arg;  // per block [N][Hp][Wp][C]
  std::vector<float*> dblock;      // per block fp32 [N][Hp][Wp][C]: gradient wrt the block's (pooled) output
  bf16* gscratch[2] = {nullptr, nullptr};
  const float* train_img = nullptr;
  // objective.lua stage buffers (frcnn_train_image), sized for cw_rows examples
  int cnet_train_rows = 0;         // rows of the last frcnn_cnet_forward_train (0: none pending for frcnn_cnet_backward)
  int cw_rows = 0;
  int cw_n = 0;                    // frames the per-frame objective buffers (head_dout, losses_dev) are sized for
  long cw_gen = -1;                // pnet workspace generation they were sized against
  float* losses_cur = nullptr;     // the 8-float loss slot of the frame being processed
  std::vector<ExampleDev> ex_host; // host staging of a batch's example records
  // sparse anchor-head backward of lossAndGradient: per head the unique pixels its listed anchors sit on
  std::vector<int> hl_host;        // the lists back to back (staging, alive until the next call)
  int* hl_dev = nullptr;           // [rows capacity]
  int hl_off[MAX_HEADS] = {0}, hl_count[MAX_HEADS] = {0};
  bf16* hs_d = nullptr;            // [rows][256] compact pre-activation gradients of one head
  bf16* hs_x = nullptr;            // [rows][max k*k*cin] gathered input windows, the heads' row blocks back to back
  float* hs_h = nullptr;           // [rows][256] conv outputs of the listed pixels (forward), heads back to back
  bool hs_forward = false;         // this step's anchor-network forward ran on the listed pixels (hs_x / hs_h are valid)
  float* hs_g = nullptr;           // [rows][max k*k*cin] listed-pixel data gradients before the scatter
  std::vector<void*> cw_allocs;
  ExampleDev* ex_dev = nullptr;
  double* ex_rects = nullptr;
  float *crtarget = nullptr, *losses_dev = nullptr, *t_dhidden = nullptr, *t_dz = nullptr, *t_dx = nullptr;
  int *cctarget = nullptr, *t_status = nullptr, *t_argmax = nullptr;
  bf16* t_rows = nullptr;
  std::vector<float*> head_dout;   // per head [18][hh][hw] fp32 (delta_outputs of one frame)
  // thresholds (Detector.lua:54,81,115,133)
  double thr_fg = 0.95, thr_class = 0.2;
  float thr_nms1 = 0.25f, thr_nms2 = 0.1f;

  // activation workspace for (N, H, W)
  int ws_n = 0, ws_h = 0, ws_w = 0, ws_sched = -1;
  std::vector<void*> ws_allocs;
  std::vector<bf16*> pool_out;   // per block
  std::vector<int> pool_h, pool_w;
  HeadPlan head_plan;            // conv_head_kernel: the anchor networks of a frame batch in one launch (evaluate mode)
  int feat_h = 0, feat_w = 0;

  // detector workspace
  int cand_cap = 4096;
  int det_n = 0;             // batch size the detector buffers were sized for
  int roi_cap = 0;           // total ROI rows
  std::vector<void*> det_allocs;
  double* cand_r = nullptr;
  float4* cand_box = nullptr;
  float* cand_logp = nullptr;
  int4* cand_anchor = nullptr;
  int* cand_count = nullptr;
  int* flags = nullptr;      // [0] candidate overflow, [1] degenerate ROIs, [2] roi_total, [3] n_det
  unsigned long long* ticket = nullptr;
  unsigned long long* status = nullptr;
  int status_blocks = 0;
  int decode_nblocks = 0;      // block count the ticket / scan-state words are currently valid for
  bool own_stream = false;
  // CUDA graph of the detect pipeline (replayed while image pointer / shape / thresholds stay the same)
  bool graph_enabled = true;
  int schedule = FRCNN_SCHED_LATENCY;  // frcnn_set_schedule
  int eval_f16 = 1;            // frcnn_set_eval_precision: evaluate-mode operands fp16 (default) or bf16
  int act_f16 = 0;             // 16-bit format of the activations the last pnet forward left in the workspace
  cudaGraphExec_t graph_exec = nullptr;
  struct GraphKey { const float* img; int N, H, W; double thr_fg, thr_class; float thr_nms1, thr_nms2; long gen; } graph_key = {};
  GraphKey eager_key = {};     // key of the last eager run (a config is captured on its second use)
  void* nms_stage = nullptr;   // device staging of the host-side nms entry points (boxes | picks | counts)
  size_t nms_stage_bytes = 0;
  bool det_pending = false;    // frcnn_detect_begin without its frcnn_detect_end yet
  int det_pending_n = 0, det_pending_h = 0, det_pending_w = 0;
  const float* det_pending_img = nullptr;  // device frames of the detection in flight (re-run if the match list outgrew cand_cap)
  long weights_gen = 0;        // bumped by frcnn_pack_weights: invalidates the cached dgrad weight layouts
  long ws_gen = 0;             // bumped whenever a workspace pointer baked into the graph may have changed
  int64_t launches_per_detect = 0;
  int spec_det = 256;          // winners copied to the host speculatively together with the counters
  int* pick1 = nullptr;
  int* count1 = nullptr;
  int* roi_base = nullptr;
  int4* roi_rect = nullptr;
  bf16* roi_out = nullptr;
  int* roi_img = nullptr;
  int* roi_cand = nullptr;
  float* reg_out = nullptr;
  float* cls_out = nullptr;
  double* fin_r2 = nullptr;
  float4* fin_box = nullptr;
  int* fin_cls = nullptr;
  float* fin_conf = nullptr;
  float4* gbox = nullptr;
  int* grow = nullptr;
  int* n_pass = nullptr;
  frcnn_detection* det_dev = nullptr;
  int det_cap = 0;
  void* nms_mem = nullptr;
  size_t nms_bytes = 0;
  NmsWorkspace nms;
  int nms_cap_total = 0, nms_cap_seg = 0;
  // pinned host staging
  float* h_loss = nullptr;         // page-locked [64 * 8 losses | 4 status ints]: lossAndGradient's early read-back
  cudaEvent_t loss_ev = nullptr;   // recorded behind that read-back
  int* h_ints = nullptr;           // [16]
  frcnn_detection* h_det = nullptr;
  int h_det_cap = 0;
  float* h_img = nullptr;
  size_t h_img_bytes = 0;
  float* d_img = nullptr;
  size_t d_img_bytes = 0;
  int64_t stats[4] = {0, 0, 0, 0};
  bool profiling = false;
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float timings[6] = {0, 0, 0, 0, 0, 0};
  // per-launch timing of the tcgen05 conv/GEMM kernel (profiling mode): event pairs around every launch
  std::vector<cudaEvent_t> conv_ev;
  int conv_ev_used = 0;
  double conv_flops = 0.0;
  float conv_ms = 0.f;
  int conv_launches = 0;
  long prof_rows = 0;  // ROI rows of the previous profiled detect call (FLOP accounting of the cnet GEMMs)
  // scratch for API-level calls
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // data-parallel training: one NCCL communicator rank per context, gradient buckets all-reduced on a
  // side stream as pnet:backward finishes them
  void* dp_comm = nullptr;           // ncclComm_t
  int dp_rank = 0, dp_nranks = 1;
  cudaStream_t dp_stream = nullptr;
  cudaEvent_t dp_ready = nullptr, dp_done = nullptr;
  bool dp_overlap = false;           // frcnn_dp_set_overlap: buckets go out from inside frcnn_train_batch / frcnn_pnet_backward
  std::vector<char> dp_bucket_sent;  // per bucket: already all-reduced in this step
  int64_t dp_bytes = 0;              // bytes all-reduced so far (bench accounting)
};

namespace frcnn {

static void* dev_alloc(std::vector<void*>& list, size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) bytes = 256;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) throw Error{FRCNN_E_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e)};
  list.push_back(p);
  return p;
}
static void free_all(std::vector<void*>& list) {
  for (void* p : list) cudaFree(p);
  list.clear();
}
static void* ensure_scratch(frcnn_ctx* c, size_t bytes) {
  if (bytes > c->scratch_bytes) {
    if (c->scratch) {
      FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
      cudaFree(c->scratch);
      c->scratch = nullptr;
      c->scratch_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&c->scratch, bytes);
    if (e != cudaSuccess) throw Error{FRCNN_E_NOMEM, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e)};
    c->scratch_bytes = bytes;
  }
  return c->scratch;
}

// small utility kernels --------------------------------------------------------------------------------------
__global__ void set_int_kernel(int* p, int v) { *p = v; }
// fp32 [R][C*bins] in the reference's flatten order (c*bins + b) -> bf16 [R][bins][C]
__global__ void pack_roi_rows_kernel(const float* __restrict__ x, bf16* __restrict__ out, long R, int C, int bins, int f16) {
  long total = R * C * bins;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / ((long)C * bins);
    int k = i - r * (long)C * bins;
    int b = k / C, c = k - b * C;
    const float v = x[r * (long)C * bins + (long)c * bins + b];
    if (f16) reinterpret_cast<__half*>(out)[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else out[i] = __float2bfloat16_rn(v);
  }
}
// split-K sums -> prelu(acc + bias) * scale -> bf16 (test entry frcnn_conv_bf16 with splits > 1)
__global__ void acc_tail_kernel(const float* __restrict__ acc, const float* __restrict__ bias, const float* __restrict__ prelu,
                                float scale, bf16* __restrict__ out, long total, int n) {
  const bool hp = prelu != nullptr;
  const float slope = hp ? prelu[0] : 1.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float x = acc[i] + (bias ? bias[i % n] : 0.f);
    if (hp) x = x > 0.f ? x : x * slope;
    out[i] = __float2bfloat16_rn(x * scale);
  }
}

// geometry (host) ------------------------------------------------------------------------------------------
static double lua_mod_h(double a, double b) { return a - floor(a / b) * b; }

// Localizer:inputToFeatureRect (Localizer.lua:41-67)
static void input_to_feature(const std::vector<std::array<int, 6>>& L, const double in[4], double out[4]) {
  double minX = in[0], minY = in[1], maxX = in[2], maxY = in[3];
  for (const auto& l : L) {
    const double kW = l[0], kH = l[1], dW = l[2], dH = l[3], pW = l[4], pH = l[5];
    if (dW < kW) {
      minX -= (kW - dW); maxX += (kW - dW);
      minY -= (kH - dH); maxY += (kH - dH);
    }
    minX += pW; maxX += pW; minY += pH; maxY += pH;
    minX = minX / dH;
    minY = minY / dH;
    if (lua_mod_h(maxX - kW, dW) == 0.0) maxX = std::max((maxX - kW) / dW + 1.0, minX + 1.0);
    else maxX = std::max(ceil((maxX - kW) / dW) + 1.0, minX + 1.0);
    if (lua_mod_h(maxY - kH, dH) == 0.0) maxY = std::max((maxY - kH) / dW + 1.0, minY + 1.0);
    else maxY = std::max(ceil((maxY - kH) / dH) + 1.0, minY + 1.0);
  }
  out[0] = floor(minX); out[1] = floor(minY); out[2] = ceil(maxX); out[3] = ceil(maxY);
}
// Localizer:featureToInputRect (Localizer.lua:69-79)
static void feature_to_input(const std::vector<std::array<int, 6>>& L, const double in[4], double out[4]) {
  double minX = in[0], minY = in[1], maxX = in[2], maxY = in[3];
  for (int i = (int)L.size() - 1; i >= 0; --i) {
    const auto& l = L[i];
    minX = minX * l[2] - l[4];
    minY = minY * l[3] - l[4];               // sic: padW (Localizer.lua:74)
    maxX = maxX * l[2] - l[5] + l[0] - l[2]; // sic: padH (Localizer.lua:75)
    maxY = maxY * l[3] - l[5] + l[1] - l[3];
  }
  out[0] = minX; out[1] = minY; out[2] = maxX; out[3] = maxY;
}
// Anchors.__init (Anchors.lua:15-57)
static void build_luts(frcnn_ctx* c) {
  const int S = (int)c->scales.size();
  c->w_lut.assign((size_t)S * 3 * LUT_EXTENT * 2, 0.f);
  c->h_lut.assign((size_t)S * 3 * LUT_EXTENT * 2, 0.f);
  c->cen_x.assign((size_t)S * LUT_EXTENT, 0.0);
  c->cen_y.assign((size_t)S * LUT_EXTENT, 0.0);
  for (int i = 0; i < S; ++i) {
    const double s = c->scales[i];
    const double a = s / sqrt(2.0);
    const double asp[3][2] = {{s, s}, {2 * a, a}, {a, 2 * a}};
    for (int j = 0; j < 3; ++j) {
      for (int y = 1; y <= LUT_EXTENT; ++y) {
        double in[4] = {0, (double)(y - 1), 0, (double)y}, r[4];
        feature_to_input(c->loc[i], in, r);
        const double cy = (r[1] + r[3]) / 2;
        c->cen_y[(size_t)i * LUT_EXTENT + (y - 1)] = cy;
        const double mn = cy - asp[j][1] * 0.5;  // Rect.fromCenterWidthHeight (Rect.lua:30-36)
        c->h_lut[(((size_t)i * 3 + j) * LUT_EXTENT + (y - 1)) * 2 + 0] = (float)mn;
        c->h_lut[(((size_t)i * 3 + j) * LUT_EXTENT + (y - 1)) * 2 + 1] = (float)(mn + asp[j][1]);
      }
      for (int x = 1; x <= LUT_EXTENT; ++x) {
        double in[4] = {(double)(x - 1), 0, (double)x, 0}, r[4];
        feature_to_input(c->loc[i], in, r);
        const double cx = (r[0] + r[2]) / 2;
        c->cen_x[(size_t)i * LUT_EXTENT + (x - 1)] = cx;
        const double mn = cx - asp[j][0] * 0.5;
        c->w_lut[(((size_t)i * 3 + j) * LUT_EXTENT + (x - 1)) * 2 + 0] = (float)mn;
        c->w_lut[(((size_t)i * 3 + j) * LUT_EXTENT + (x - 1)) * 2 + 1] = (float)(mn + asp[j][0]);
      }
    }
  }
}

static int add_param(frcnn_ctx* c, const std::string& name, int64_t numel) {
  c->params.push_back({name, numel});
  return (int)c->params.size() - 1;
}

static void do_plan(frcnn_ctx* c, const frcnn_block_desc* blocks, int n_blocks, const frcnn_head_desc* heads, int n_heads,
                    const frcnn_fc_desc* fcs, int n_fcs, int class_count, int roi_kh, int roi_kw, const double* scales,
                    int n_scales, float dropout_eval_scale) {
  FRCNN_REQUIRE(!c->planned, FRCNN_E_STATE, "model plan already set on this ctx");
  FRCNN_REQUIRE(n_blocks >= 1 && n_blocks <= 8 && n_fcs >= 1 && class_count >= 1, FRCNN_E_INVALID, "bad model description");
  FRCNN_REQUIRE(n_heads == MAX_HEADS && n_scales == MAX_HEADS, FRCNN_E_INVALID,
                "the reference hard-codes 4 anchor layers (Detector.lua:38, Anchors.lua:108)");
  c->blocks.assign(blocks, blocks + n_blocks);
  c->head_desc.assign(heads, heads + n_heads);
  c->fc_desc.assign(fcs, fcs + n_fcs);
  c->class_count = class_count;
  c->roi_kh = roi_kh;
  c->roi_kw = roi_kw;
  c->scales.assign(scales, scales + n_scales);
  int cin = 3;
  for (int b = 0; b < n_blocks; ++b) {
    const auto& l = blocks[b];
    FRCNN_REQUIRE(l.kW == l.kH && l.padW == l.padH && l.conv_steps >= 1, FRCNN_E_INVALID, "square kernels / symmetric pads only");
    FRCNN_REQUIRE(l.filters % 64 == 0, FRCNN_E_INVALID, "block filters must be a multiple of 64");
    for (int s = 0; s < l.conv_steps; ++s) {
      ConvLayer cv;
      cv.cin = cin; cv.cout = l.filters; cv.k = l.kW; cv.pad = l.padW;
      cv.first = (cin == 3);
      cv.block = b;
      cv.scale = 1.0f;
      if (s == 0 && l.dropout > 0.f) {
        cv.scale = dropout_eval_scale < 0.f ? 1.0f - l.dropout : dropout_eval_scale;  // Q5
        cv.dropout = l.dropout;
      }
      std::string n = "b" + std::to_string(b + 1) + "_c" + std::to_string(s + 1);
      cv.p_w = add_param(c, n + ".weight", (int64_t)l.filters * cin * l.kH * l.kW);
      cv.p_b = add_param(c, n + ".bias", l.filters);
      cv.p_prelu = add_param(c, n + ".prelu", 1);
      c->trunk.push_back(cv);
      cin = l.filters;
    }
  }
  c->feat_c = cin;
  for (int h = 0; h < n_heads; ++h) {
    const auto& a = heads[h];
    FRCNN_REQUIRE(a.input >= 1 && a.input <= n_blocks, FRCNN_E_INVALID, "anchor net input out of range");
    FRCNN_REQUIRE(a.n % 128 == 0, FRCNN_E_INVALID, "anchor net width must be a multiple of 128");
    Head hd;
    hd.kW = a.kW; hd.n = a.n; hd.input = a.input;
    hd.conv.cin = blocks[a.input - 1].filters; hd.conv.cout = a.n; hd.conv.k = a.kW; hd.conv.pad = 0;
    hd.conv.first = false; hd.conv.block = -1; hd.conv.scale = 1.f;
    std::string n = "h" + std::to_string(h + 1);
    hd.conv.p_w = add_param(c, n + "_conv.weight", (int64_t)a.n * hd.conv.cin * a.kW * a.kW);
    hd.conv.p_b = add_param(c, n + "_conv.bias", a.n);
    hd.conv.p_prelu = add_param(c, n + "_conv.prelu", 1);
    hd.p_w2 = add_param(c, n + "_out.weight", 18 * a.n);
    hd.p_b2 = add_param(c, n + "_out.bias", 18);
    c->heads.push_back(hd);
  }
  int fin = roi_kh * roi_kw * c->feat_c;
  for (int i = 0; i < n_fcs; ++i) {
    FRCNN_REQUIRE(fcs[i].n % 64 == 0, FRCNN_E_INVALID, "class layer width must be a multiple of 64");
    FcLayer f;
    f.nin = fin; f.nout = fcs[i].n; f.bn = fcs[i].batch_norm != 0;
    f.dropout = fcs[i].dropout;
    std::string n = "fc" + std::to_string(i + 1);
    f.p_w = add_param(c, n + ".weight", (int64_t)f.nout * fin);
    f.p_b = add_param(c, n + ".bias", f.nout);
    f.p_bn_w = f.p_bn_b = f.p_bn_mean = f.p_bn_var = -1;
    if (f.bn) {
      f.p_bn_w = add_param(c, n + ".bn_weight", f.nout);
      f.p_bn_b = add_param(c, n + ".bn_bias", f.nout);
      f.p_bn_mean = add_param(c, n + ".bn_mean", f.nout);
      f.p_bn_var = add_param(c, n + ".bn_var", f.nout);
    }
    f.p_prelu = add_param(c, n + ".prelu", 1);
    c->fcs.push_back(f);
    fin = f.nout;
  }
  c->p_reg_w = add_param(c, "reg.weight", 4 * (int64_t)fin);
  c->p_reg_b = add_param(c, "reg.bias", 4);
  c->p_cls_w = add_param(c, "cls.weight", (int64_t)(class_count + 1) * fin);
  c->p_cls_b = add_param(c, "cls.bias", class_count + 1);

  // Localizer layer lists (Localizer.lua:8-38): convs and max-pools in forward order
  auto trunk_layers = [&](int upto) {
    std::vector<std::array<int, 6>> L;
    for (int b = 0; b < upto; ++b) {
      for (int s = 0; s < blocks[b].conv_steps; ++s)
        L.push_back({blocks[b].kW, blocks[b].kH, 1, 1, blocks[b].padW, blocks[b].padH});
      L.push_back({2, 2, 2, 2, 0, 0});
    }
    return L;
  };
  c->loc.clear();
  for (int h = 0; h < n_heads; ++h) {
    auto L = trunk_layers(heads[h].input);
    L.push_back({heads[h].kW, heads[h].kW, 1, 1, 0, 0});
    L.push_back({1, 1, 1, 1, 0, 0});
    c->loc.push_back(L);
  }
  c->loc.push_back(trunk_layers(n_blocks));
  FRCNN_REQUIRE((int)c->loc.back().size() <= MAX_LOC_LAYERS, FRCNN_E_INVALID, "too many layers for the ROI localizer");
  c->roi_loc.n = (int)c->loc.back().size();
  for (int i = 0; i < c->roi_loc.n; ++i)
    for (int e = 0; e < 6; ++e) c->roi_loc.l[i][e] = c->loc.back()[i][e];
  build_luts(c);
  c->bound.assign(c->params.size(), nullptr);
  c->grads.assign(c->params.size(), nullptr);
  c->planned = true;
}

static const float* P(frcnn_ctx* c, int idx) { return idx >= 0 ? c->bound[idx] : nullptr; }

#define REQUIRE_DEVICE(c) FRCNN_REQUIRE((c)->device >= 0, FRCNN_E_CUDA, "host-only context: no CUDA device (there is no CPU fallback)")

// the anchor LUTs (Anchors.lua:18-19) in device memory; the host vectors are built at plan time
static void ensure_luts_dev(frcnn_ctx* c) {
  if (c->d_w_lut) return;
  FRCNN_REQUIRE(!c->w_lut.empty(), FRCNN_E_STATE, "frcnn_model_plan must be called first");
  FRCNN_CUDA_TRY(cudaMalloc(&c->d_w_lut, c->w_lut.size() * sizeof(float)));
  FRCNN_CUDA_TRY(cudaMalloc(&c->d_h_lut, c->h_lut.size() * sizeof(float)));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_w_lut, c->w_lut.data(), c->w_lut.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_h_lut, c->h_lut.data(), c->h_lut.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMalloc(&c->d_cen, (c->cen_x.size() + c->cen_y.size()) * sizeof(double)));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_cen, c->cen_x.data(), c->cen_x.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_cen + c->cen_x.size(), c->cen_y.data(), c->cen_y.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
}

static void do_pack(frcnn_ctx* c) {
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(c->planned, FRCNN_E_STATE, "frcnn_model_plan must be called first");
  for (auto p : c->bound) FRCNN_REQUIRE(p != nullptr, FRCNN_E_STATE, "frcnn_bind_params must be called before packing");
  auto alloc_w = [&](size_t elems) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, elems * sizeof(bf16));
    if (e != cudaSuccess) throw Error{FRCNN_E_NOMEM, "cudaMalloc(packed weights) failed"};
    return (bf16*)p;
  };
  for (auto& cv : c->trunk) {
    if (cv.first) {
      FRCNN_REQUIRE(cv.cin * cv.k * cv.k <= 32, FRCNN_E_INVALID, "first-layer im2col K exceeds 32");
      // every forward weight is packed twice: [bf16 copy | fp16 copy] (training / evaluate operands, ConvParams::f16)
      if (!cv.w_packed) cv.w_packed = alloc_w(2 * (size_t)cv.cout * 32);
      launch_pack_first_conv_weight(P(c, cv.p_w), cv.w_packed, cv.cout, cv.cin, cv.k, cv.k, c->stream);
    } else {
      if (!cv.w_packed) cv.w_packed = alloc_w(2 * (size_t)cv.cout * cv.cin * cv.k * cv.k);
      launch_pack_conv_weight(P(c, cv.p_w), cv.w_packed, cv.cout, cv.cin, cv.k, cv.k, c->stream, 2);
    }
    ++c->launches;
  }
  for (auto& h : c->heads) {
    auto& cv = h.conv;
    if (!cv.w_packed) cv.w_packed = alloc_w(2 * (size_t)cv.cout * cv.cin * cv.k * cv.k);
    launch_pack_conv_weight(P(c, cv.p_w), cv.w_packed, cv.cout, cv.cin, cv.k, cv.k, c->stream, 2);
    ++c->launches;
  }
  for (size_t i = 0; i < c->fcs.size(); ++i) {
    auto& f = c->fcs[i];
    if (!f.w_packed) f.w_packed = alloc_w(2 * (size_t)f.nout * f.nin);
    if (i == 0) launch_pack_fc_weight(P(c, f.p_w), f.w_packed, f.nout, c->feat_c, c->roi_kh * c->roi_kw, 1, c->stream, 2);
    else launch_pack_fc_weight(P(c, f.p_w), f.w_packed, f.nout, f.nin, 1, 0, c->stream, 2);
    ++c->launches;
  }
  ensure_luts_dev(c);
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));  // the LUT host vectors / caller buffers may change afterwards
  ++c->weights_gen;
  c->packed = true;
}

// (re)builds the activation workspace and the prepared conv launches for an (N, H, W) input
static void ensure_pnet_workspace(frcnn_ctx* c, int N, int H, int W) {
  if (c->ws_n == N && c->ws_h == H && c->ws_w == W && c->ws_sched == c->schedule) return;
  c->ws_sched = c->schedule;   // the throughput schedule picks the trunk kernels by SM-time (conv_prepare, force_mt = -1)
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  free_all(c->ws_allocs);
  free_all(c->tws_allocs);
  c->tws_n = c->tws_h = c->tws_w = 0;
  c->train_ready = false;
  ++c->ws_gen;
  c->ws_n = c->ws_h = c->ws_w = 0;
  const int nb = (int)c->blocks.size();
  c->pool_out.assign(nb, nullptr);
  c->pool_h.assign(nb, 0);
  c->pool_w.assign(nb, 0);
  int h = H, w = W;
  bf16* cur = nullptr;
  size_t li = 0;
  for (int b = 0; b < nb; ++b) {
    for (int s = 0; s < c->blocks[b].conv_steps; ++s, ++li) {
      ConvLayer& cv = c->trunk[li];
      cv.pooled = (s + 1 == c->blocks[b].conv_steps);  // model_utilities.lua:23: the pool follows the block's last conv
      cv.hin = h; cv.win = w;
      cv.hout = h + 2 * cv.pad - cv.k + 1;
      cv.wout = w + 2 * cv.pad - cv.k + 1;
      const int oh = cv.pooled ? (cv.hout + 1) / 2 : cv.hout, ow = cv.pooled ? (cv.wout + 1) / 2 : cv.wout;  // :ceil()
      cv.out = (bf16*)dev_alloc(c->ws_allocs, (size_t)N * oh * ow * cv.cout * sizeof(bf16));
      const int mode = cv.pooled ? EPI_POOL : EPI_STORE;
      if (cv.first) {
        cv.in = nullptr;
        conv_first_prepare(&cv.launch, cv.w_packed, N, h, w, cv.cin, cv.cout, cv.k, cv.k, cv.pad, cv.pad, mode, cv.out, c->sm_count);
      } else {
        cv.in = cur;
        conv_prepare(&cv.launch, cur, cv.w_packed, N, h, w, cv.cin, cv.cout, cv.k, cv.k, cv.pad, cv.pad, mode, cv.out,
                     c->sm_count, 0, 0, c->schedule == FRCNN_SCHED_THROUGHPUT ? -1 : 0, 2);
      }
      cv.launch.p.scale = cv.scale;
      cur = cv.out;
      h = oh; w = ow;
    }
    c->pool_out[b] = cur;
    c->pool_h[b] = h; c->pool_w[b] = w;
  }
  c->feat_h = h; c->feat_w = w;
  // anchor heads: ONE grouped conv launch with deterministic split-K slices, sized so that all work units of all
  // heads have similar K length and there are about two units per SM
  long total_k = 0;
  for (auto& hd : c->heads) {
    const int ih = c->pool_h[hd.input - 1], iw = c->pool_w[hd.input - 1];
    FRCNN_REQUIRE(ih >= hd.kW && iw >= hd.kW, FRCNN_E_INVALID, "image too small for the anchor networks");
    hd.hh = ih - hd.kW + 1; hd.hw = iw - hd.kW + 1;
    FRCNN_REQUIRE(hd.hh <= LUT_EXTENT && hd.hw <= LUT_EXTENT, FRCNN_E_INVALID,
                  "feature map exceeds the 200-cell anchor LUT (Anchors.lua:15)");
    FRCNN_REQUIRE(hd.n == 256, FRCNN_E_INVALID, "anchor net width must be 256 (models/vgg_*.lua)");
    const long tiles = (long)N * ((hd.hh * hd.hw + 127) / 128);
    total_k += tiles * hd.kW * hd.kW * (hd.conv.cin / 64);
  }
  const long k_target = std::max<long>(16, (total_k + 2L * c->sm_count - 1) / (2L * c->sm_count));
  for (auto& hd : c->heads) {
    const int ih = c->pool_h[hd.input - 1], iw = c->pool_w[hd.input - 1];
    const int k_iters = hd.kW * hd.kW * (hd.conv.cin / 64);
    int splits = (int)((k_iters + k_target / 2) / k_target);
    splits = std::max(1, std::min(splits, std::max(1, k_iters / 8)));
    hd.conv.hin = ih; hd.conv.win = iw; hd.conv.hout = hd.hh; hd.conv.wout = hd.hw;
    conv_prepare(&hd.conv.launch, c->pool_out[hd.input - 1], hd.conv.w_packed, N, ih, iw, hd.conv.cin, hd.conv.cout, hd.kW,
                 hd.kW, 0, 0, EPI_F32_SLICES, nullptr, c->sm_count, splits, 256, 1, 2);
    const size_t slices = (size_t)hd.conv.launch.p.splits * N;
    hd.acc = (float*)dev_alloc(c->ws_allocs, slices * hd.hh * hd.hw * hd.n * sizeof(float));
    hd.out = (float*)dev_alloc(c->ws_allocs, (size_t)N * 18 * hd.hh * hd.hw * sizeof(float));
    conv_set_f32_output(&hd.conv.launch, hd.acc);
    conv_prepare_head(&hd.fused, c->pool_out[hd.input - 1], hd.conv.w_packed, N, ih, iw, hd.conv.cin, hd.kW, c->sm_count);
  }
  // the fused anchor-network kernel (evaluate mode): unit schedule, slices of the split heads, arrival counters
  {
    HeadDesc hdesc[MAX_GROUP];
    const int nh = (int)c->heads.size();
    FRCNN_REQUIRE(nh <= MAX_GROUP, FRCNN_E_INVALID, "at most 4 anchor networks (Detector.lua:38)");
    for (int i = 0; i < nh; ++i) {
      Head& hd = c->heads[i];
      hdesc[i] = HeadDesc{c->pool_out[hd.input - 1], hd.conv.w_packed, c->pool_h[hd.input - 1], c->pool_w[hd.input - 1], hd.conv.cin, hd.kW};
    }
    std::vector<int4> units;
    std::vector<int> cta_off;
    size_t slice_floats[MAX_GROUP];
    int counter_ints[MAX_GROUP];
    conv_head_plan(&c->head_plan, hdesc, nh, N, c->sm_count, &units, &cta_off, slice_floats, counter_ints);
    int4* d_units = (int4*)dev_alloc(c->ws_allocs, units.size() * sizeof(int4));
    int* d_off = (int*)dev_alloc(c->ws_allocs, cta_off.size() * sizeof(int));
    FRCNN_CUDA_TRY(cudaMemcpyAsync(d_units, units.data(), units.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
    FRCNN_CUDA_TRY(cudaMemcpyAsync(d_off, cta_off.data(), cta_off.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    c->head_plan.sched.units = d_units;
    c->head_plan.sched.cta_off = d_off;
    for (int i = 0; i < MAX_GROUP; ++i) {
      c->head_plan.sched.slices[i] = slice_floats[i] ? (float*)dev_alloc(c->ws_allocs, slice_floats[i] * sizeof(float)) : nullptr;
      c->head_plan.sched.counters[i] = nullptr;
      if (counter_ints[i]) {
        c->head_plan.sched.counters[i] = (int*)dev_alloc(c->ws_allocs, counter_ints[i] * sizeof(int));
        FRCNN_CUDA_TRY(cudaMemsetAsync(c->head_plan.sched.counters[i], 0, counter_ints[i] * sizeof(int), c->stream));
      }
    }
    c->head_plan.sched.trace = nullptr;
    if (getenv("FRCNN_HEAD_TRACE")) {   // measurement only: per-unit globaltimer stamps, dumped after every launch
      const size_t bytes = (size_t)c->head_plan.grid * 64 * sizeof(unsigned long long);
      c->head_plan.sched.trace = (unsigned long long*)dev_alloc(c->ws_allocs, bytes);
      FRCNN_CUDA_TRY(cudaMemsetAsync(c->head_plan.sched.trace, 0, bytes, c->stream));
    }
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));   // `units` / `cta_off` are stack vectors
  }
  c->ws_n = N; c->ws_h = H; c->ws_w = W;
}

// Launches the conv/GEMM kernel; in profiling mode brackets it with an event pair and adds its algorithmic FLOPs
// (2 * M * N * K of the un-padded problem; `rows` overrides M for the m_limit-ed cnet GEMMs).
static void conv_launch_timed(frcnn_ctx* c, const ConvLaunch& L, long rows = -1) {
  if (c->profiling) {
    if ((int)c->conv_ev.size() < c->conv_ev_used + 2) {
      cudaEvent_t a, b;
      FRCNN_CUDA_TRY(cudaEventCreate(&a));
      FRCNN_CUDA_TRY(cudaEventCreate(&b));
      c->conv_ev.push_back(a);
      c->conv_ev.push_back(b);
    }
    cudaEventRecord(c->conv_ev[c->conv_ev_used], c->stream);
  }
  conv_launch(L, c->stream);
  ++c->launches;
  if (c->profiling) {
    cudaEventRecord(c->conv_ev[c->conv_ev_used + 1], c->stream);
    c->conv_ev_used += 2;
    const double M = rows >= 0 ? (double)rows : (double)L.p.N * L.p.Hout * L.p.Wout;
    c->conv_flops += 2.0 * M * L.p.Cout * (double)L.p.KH * L.p.KW * (L.first ? L.p.Cimg : L.p.Cin);
  }
}
static void conv_group_launch_timed(frcnn_ctx* c, const ConvLaunch* const* Ls, int n) {
  const bool prof = c->profiling;
  if (prof) {
    if ((int)c->conv_ev.size() < c->conv_ev_used + 2) {
      cudaEvent_t a, b;
      FRCNN_CUDA_TRY(cudaEventCreate(&a));
      FRCNN_CUDA_TRY(cudaEventCreate(&b));
      c->conv_ev.push_back(a);
      c->conv_ev.push_back(b);
    }
    cudaEventRecord(c->conv_ev[c->conv_ev_used], c->stream);
  }
  conv_launch_group(Ls, n, c->sm_count, c->stream);
  ++c->launches;
  if (prof) {
    cudaEventRecord(c->conv_ev[c->conv_ev_used + 1], c->stream);
    c->conv_ev_used += 2;
    for (int i = 0; i < n; ++i) {
      const ConvParams& p = Ls[i]->p;
      c->conv_flops += 2.0 * (double)p.N * p.Hout * p.Wout * p.Cout * (double)p.KH * p.KW * p.Cin;
    }
  }
}
static void conv_profile_begin(frcnn_ctx* c) {
  c->conv_ev_used = 0;
  c->conv_flops = 0.0;
}
// after a stream synchronisation
static void conv_profile_end(frcnn_ctx* c) {
  c->conv_ms = 0.f;
  c->conv_launches = c->conv_ev_used / 2;
  for (int i = 0; i + 1 < c->conv_ev_used; i += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->conv_ev[i], c->conv_ev[i + 1]);
    c->conv_ms += ms;
  }
}

static void run_conv(frcnn_ctx* c, ConvLayer& cv) {
  cv.launch.p.bias = P(c, cv.p_b);
  cv.launch.p.prelu = P(c, cv.p_prelu);
  const float keep = cv.launch.p.scale;  // set by the caller (evaluate / training)
  conv_launch_timed(c, cv.launch);
  (void)keep;
}

static void ensure_train_workspace(frcnn_ctx* c, int N, int H, int W);
static int detect_stop_after();

// skip_heads: lossAndGradient evaluates the anchor networks on the listed pixels only (sparse_head_forward)
static void do_pnet_forward(frcnn_ctx* c, const float* img_dev, int N, int H, int W, bool train = false, bool skip_heads = false) {
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called before the forward pass");
  FRCNN_REQUIRE(N >= 1 && H >= 16 && W >= 16, FRCNN_E_INVALID, "bad input size");
  ensure_pnet_workspace(c, N, H, W);
  if (train) ensure_train_workspace(c, N, H, W);
  c->train_ready = train;
  c->train_img = train ? img_dev : nullptr;
  // operand format: training stays bf16 (gradient maps need the exponent range); evaluate mode uses fp16 by default --
  // three more significand bits at the same tensor-core rate (precision contract)
  const int f16 = (!train && c->eval_f16) ? 1 : 0;
  c->act_f16 = f16;
  size_t li = 0;
  for (size_t b = 0; b < c->blocks.size(); ++b) {
    for (int s = 0; s < c->blocks[b].conv_steps; ++s, ++li) {
      ConvLayer& cv = c->trunk[li];
      if (cv.first) cv.launch.p.img = img_dev;
      // evaluate: SpatialDropout multiplies by (1 - p) (Q5); training: per-(image, channel) Bernoulli mask, no rescale,
      // and the pooled convs remember the winner of every window for the backward pass
      const bool masked = train && cv.dropout > 0.f;
      cv.launch.p.scale = masked ? 1.0f : (train ? 1.0f : cv.scale);
      cv.launch.p.chan_scale = masked ? cv.mask : nullptr;
      cv.launch.p.pool_arg = (train && cv.pooled) ? c->pool_arg[b] : nullptr;
      cv.launch.p.f16 = f16;
      run_conv(c, cv);
    }
  }
  // anchor heads (model_utilities.lua:29-35,51-54).  Throughput schedule (several frames in flight, evaluate mode):
  // ONE launch of unsplit units, k x k conv and tail fused -- the least SM time per frame, the machine is filled by
  // the other frames.  Latency schedule: one grouped split-K conv launch (every SM busy on this frame), heaviest
  // units first, then one grouped tail launch.
  if (detect_stop_after() == 1 && !train) return;
  if (skip_heads) {
    FRCNN_CUDA_TRY(cudaGetLastError());
    return;
  }
  static const int old_heads = getenv("FRCNN_HEADS_OLD") ? atoi(getenv("FRCNN_HEADS_OLD")) : 0;   // A/B measurements only
  if (!train && !old_heads) {
    // evaluate mode, both schedules: conv_head_kernel (linear tiles, reduction split by filter rows with in-kernel fix-up)
    HeadPlan& hp = c->head_plan;
    for (size_t i = 0; i < c->heads.size(); ++i) {
      Head& hd = c->heads[i];
      ConvParams& p = hp.grp.p[i];
      p.bias = P(c, hd.conv.p_b); p.prelu = P(c, hd.conv.p_prelu); p.w2 = P(c, hd.p_w2); p.b2 = P(c, hd.p_b2);
      p.out = hd.out;
      p.f16 = f16;
    }
    for (size_t i = c->heads.size(); i < (size_t)MAX_GROUP; ++i) hp.grp.p[i] = hp.grp.p[c->heads.size() - 1];
    const bool prof = c->profiling;
    if (prof) {
      if ((int)c->conv_ev.size() < c->conv_ev_used + 2) {
        cudaEvent_t a, b;
        FRCNN_CUDA_TRY(cudaEventCreate(&a));
        FRCNN_CUDA_TRY(cudaEventCreate(&b));
        c->conv_ev.push_back(a);
        c->conv_ev.push_back(b);
      }
      cudaEventRecord(c->conv_ev[c->conv_ev_used], c->stream);
    }
    conv_launch_heads(hp, c->stream);
    c->launches += hp.fix.n > 0 ? 2 : 1;   // + head_fixup_kernel when a head's reduction is split
    if (hp.sched.trace) {
      std::vector<unsigned long long> t((size_t)hp.grid * 64);
      FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
      FRCNN_CUDA_TRY(cudaMemcpy(t.data(), hp.sched.trace, t.size() * 8, cudaMemcpyDeviceToHost));
      FRCNN_CUDA_TRY(cudaMemset(hp.sched.trace, 0, t.size() * 8));
      if (FILE* f = fopen(getenv("FRCNN_HEAD_TRACE"), "wb")) {
        fwrite(t.data(), 8, t.size(), f);
        fclose(f);
      }
    }
    if (prof) {
      cudaEventRecord(c->conv_ev[c->conv_ev_used + 1], c->stream);
      c->conv_ev_used += 2;
      c->conv_flops += hp.flops;
    }
    FRCNN_CUDA_TRY(cudaGetLastError());
    return;
  }
  if (c->schedule == FRCNN_SCHED_THROUGHPUT && !train) {
    std::vector<const ConvLaunch*> order;
    for (auto& hd : c->heads) {
      ConvParams& p = hd.fused.p;
      p.bias = P(c, hd.conv.p_b); p.prelu = P(c, hd.conv.p_prelu); p.w2 = P(c, hd.p_w2); p.b2 = P(c, hd.p_b2);
      p.out = hd.out;
      p.f16 = f16;
      order.push_back(&hd.fused);
    }
    std::stable_sort(order.begin(), order.end(), [](const ConvLaunch* a, const ConvLaunch* b) { return a->p.k_iters > b->p.k_iters; });
    const bool prof = c->profiling;
    if (prof) {
      if ((int)c->conv_ev.size() < c->conv_ev_used + 2) {
        cudaEvent_t a, b;
        FRCNN_CUDA_TRY(cudaEventCreate(&a));
        FRCNN_CUDA_TRY(cudaEventCreate(&b));
        c->conv_ev.push_back(a);
        c->conv_ev.push_back(b);
      }
      cudaEventRecord(c->conv_ev[c->conv_ev_used], c->stream);
    }
    conv_launch_head_group(order.data(), (int)order.size(), c->sm_count, c->stream);
    ++c->launches;
    if (prof) {
      cudaEventRecord(c->conv_ev[c->conv_ev_used + 1], c->stream);
      c->conv_ev_used += 2;
      for (auto* L : order) c->conv_flops += 2.0 * (double)L->p.N * L->p.Hout * L->p.Wout * L->p.Cout * (double)L->p.KH * L->p.KW * L->p.Cin;
    }
  } else {
    std::vector<const ConvLaunch*> order;
    for (auto& hd : c->heads) {
      hd.conv.launch.p.bias = nullptr;
      hd.conv.launch.p.prelu = nullptr;
      hd.conv.launch.p.f16 = f16;
      order.push_back(&hd.conv.launch);
    }
    std::stable_sort(order.begin(), order.end(),
                     [](const ConvLaunch* a, const ConvLaunch* b) { return a->p.k_per_split > b->p.k_per_split; });
    conv_group_launch_timed(c, order.data(), (int)order.size());
    HeadTailGroup tg;
    tg.n = (int)c->heads.size();
    for (int i = 0; i < tg.n; ++i) {
      Head& hd = c->heads[i];
      HeadTail& t = tg.h[i];
      t.ws = hd.acc;
      t.splits = hd.conv.launch.p.splits;
      t.npix = (long)N * hd.hh * hd.hw;
      t.slice_stride = t.npix * hd.n;
      t.HW = hd.hh * hd.hw;
      t.bias = P(c, hd.conv.p_b); t.prelu = P(c, hd.conv.p_prelu); t.w2 = P(c, hd.p_w2); t.b2 = P(c, hd.p_b2);
      t.out = hd.out;
      t.block_end = 0;
    }
    launch_head_tail_group(tg, c->sm_count, c->stream);
    ++c->launches;
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------- training workspace
static float* G(frcnn_ctx* c, int idx) { return idx >= 0 ? c->grads[idx] : nullptr; }

// buffers and prepared dgrad / wgrad launches of pnet:backward for the (N, H, W) activation workspace
static void ensure_train_workspace(frcnn_ctx* c, int N, int H, int W) {
  if (c->tws_n == N && c->tws_h == H && c->tws_w == W) return;
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  free_all(c->tws_allocs);
  c->tws_n = c->tws_h = c->tws_w = 0;
  auto& A = c->tws_allocs;
  const int nb = (int)c->blocks.size();
  c->pool_arg.assign(nb, nullptr);
  c->dblock.assign(nb, nullptr);
  size_t max_map = 0;
  size_t li = 0;
  for (int b = 0; b < nb; ++b) {
    const int C = c->blocks[b].filters;
    c->pool_arg[b] = (uint8_t*)dev_alloc(A, (size_t)N * c->pool_h[b] * c->pool_w[b] * C);
    c->dblock[b] = (float*)dev_alloc(A, (size_t)N * c->pool_h[b] * c->pool_w[b] * C * sizeof(float));
    for (int s = 0; s < c->blocks[b].conv_steps; ++s, ++li) {
      ConvLayer& cv = c->trunk[li];
      max_map = std::max(max_map, (size_t)N * cv.hout * cv.wout * cv.cout);
      max_map = std::max(max_map, (size_t)N * cv.hin * cv.win * (size_t)std::max(cv.cin, 64));
      if (cv.dropout > 0.f) cv.mask = (float*)dev_alloc(A, (size_t)N * cv.cout * sizeof(float));
      if (!cv.first) {
        cv.w_dgrad = (bf16*)dev_alloc(A, (size_t)cv.cout * cv.cin * cv.k * cv.k * sizeof(bf16));
        cv.dw_taps = (float*)dev_alloc(A, (size_t)cv.cout * cv.cin * cv.k * cv.k * sizeof(float));
      }
    }
  }
  for (auto& cv : c->trunk) cv.dgrad_gen = -1;
  for (auto& hd : c->heads) hd.conv.dgrad_gen = -1;
  c->gscratch[0] = (bf16*)dev_alloc(A, max_map * sizeof(bf16));
  c->gscratch[1] = (bf16*)dev_alloc(A, max_map * sizeof(bf16));
  // trunk launches: dy always lives in gscratch[0] ("cur"), the bf16 dgrad output in gscratch[1]
  li = 0;
  for (int b = 0; b < nb; ++b) {
    for (int s = 0; s < c->blocks[b].conv_steps; ++s, ++li) {
      ConvLayer& cv = c->trunk[li];
      if (cv.first) continue;
      conv_wgrad_prepare(&cv.wgrad, c->gscratch[0], cv.in, cv.dw_taps, N, cv.hin, cv.win, cv.cin, cv.cout, cv.k, cv.k, cv.pad, cv.pad,
                         c->sm_count);
      const int padT = cv.k - 1 - cv.pad;
      if (s > 0) {  // the producer of this conv's input is a plain conv of the same block: bf16 gradient map
        conv_prepare(&cv.dgrad, c->gscratch[0], cv.w_dgrad, N, cv.hout, cv.wout, cv.cout, cv.cin, cv.k, cv.k, padT, padT, EPI_STORE,
                     c->gscratch[1], c->sm_count, 0, 0, 0);
      } else {      // the input is the pooled output of the previous block: accumulate into its fp32 gradient
        conv_prepare(&cv.dgrad, c->gscratch[0], cv.w_dgrad, N, cv.hout, cv.wout, cv.cout, cv.cin, cv.k, cv.k, padT, padT,
                     EPI_F32_REDUCE, nullptr, c->sm_count, 1, 0, 0);
        conv_set_f32_output(&cv.dgrad, c->dblock[b - 1]);
      }
    }
  }
  for (auto& hd : c->heads) {
    ConvLayer& cv = hd.conv;
    hd.dpre = (bf16*)dev_alloc(A, (size_t)N * hd.hh * hd.hw * hd.n * sizeof(bf16));
    cv.w_dgrad = (bf16*)dev_alloc(A, (size_t)cv.cout * cv.cin * cv.k * cv.k * sizeof(bf16));
    cv.dw_taps = (float*)dev_alloc(A, (size_t)cv.cout * cv.cin * cv.k * cv.k * sizeof(float));
    conv_wgrad_prepare(&cv.wgrad, hd.dpre, c->pool_out[hd.input - 1], cv.dw_taps, N, cv.hin, cv.win, cv.cin, cv.cout, cv.k, cv.k, 0, 0,
                       c->sm_count);
    conv_prepare(&cv.dgrad, hd.dpre, cv.w_dgrad, N, hd.hh, hd.hw, cv.cout, cv.cin, cv.k, cv.k, cv.k - 1, cv.k - 1, EPI_F32_REDUCE,
                 nullptr, c->sm_count, 1, 0, 0);
    conv_set_f32_output(&cv.dgrad, c->dblock[hd.input - 1]);
  }
  c->tws_n = N; c->tws_h = H; c->tws_w = W;
}

static void dp_bucket_ready(frcnn_ctx* c, int bucket);

// weight gradient of one conv: taps buffer zeroed, tensor-core wgrad, accumulated into the bound Torch-layout gradient
static void run_wgrad(frcnn_ctx* c, ConvLayer& cv) {
  FRCNN_CUDA_TRY(cudaMemsetAsync(cv.dw_taps, 0, (size_t)cv.cout * cv.cin * cv.k * cv.k * sizeof(float), c->stream));
  conv_launch(cv.wgrad, c->stream);
  launch_wgrad_finish(cv.dw_taps, G(c, cv.p_w), cv.cout, cv.cin, cv.k, cv.k, c->stream);
  c->launches += 2;
}
static void run_dgrad(frcnn_ctx* c, ConvLayer& cv) {
  // the flipped filters are re-packed once per frcnn_pack_weights (= once per optimiser step), not per frame
  if (cv.dgrad_gen != c->weights_gen) {
    launch_pack_conv_weight_dgrad(P(c, cv.p_w), cv.w_dgrad, cv.cout, cv.cin, cv.k, cv.k, c->stream);
    cv.dgrad_gen = c->weights_gen;
    ++c->launches;
  }
  cv.dgrad.p.bias = nullptr;
  cv.dgrad.p.prelu = nullptr;
  cv.dgrad.p.scale = 1.f;
  conv_launch(cv.dgrad, c->stream);
  ++c->launches;
}

// pnet:backward(img, delta_outputs) (objective.lua:189): parameter gradients are ACCUMULATED into the bound gradient
// views (the reference zeroes them once per batch, objective.lua:49); the input gradient is not computed (unused).
static void zero_block_grads(frcnn_ctx* c) {
  const int N = c->ws_n, nb = (int)c->blocks.size();
  for (int b = 0; b < nb; ++b)
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->dblock[b], 0, (size_t)N * c->pool_h[b] * c->pool_w[b] * c->blocks[b].filters * sizeof(float),
                                   c->stream));
}

// keep_block_grads: the caller has already zeroed the per-block gradient maps and added the ROI-pool gradients
static void gemm_rows(frcnn_ctx* c, const bf16* a, const bf16* w, int R, int nin, int nout, float* out);
// element offset of head i's block of gathered windows in c->hs_x (the heads' [M][k*k*cin] blocks back to back)
static size_t head_rows_offset(const frcnn_ctx* c, int i) {
  size_t o = 0;
  for (int j = 0; j < i; ++j) o += (size_t)c->hl_count[j] * c->heads[j].conv.k * c->heads[j].conv.k * c->heads[j].conv.cin;
  return o;
}
// The anchor networks on the listed pixels only (lossAndGradient reads their outputs nowhere else, objective.lua:96-131):
// gather the k x k windows, one tensor-core GEMM per head against the packed forward filters, tail on the M rows.
static void sparse_head_forward(frcnn_ctx* c) {
  cudaStream_t st = c->stream;
  for (size_t i = 0; i < c->heads.size(); ++i) {
    const int M = c->hl_count[i];
    if (M == 0) continue;
    Head& hd = c->heads[i];
    ConvLayer& cv = hd.conv;
    const int* list = c->hl_dev + c->hl_off[i];
    const int kkc = cv.k * cv.k * cv.cin;
    bf16* windows = c->hs_x + head_rows_offset(c, (int)i);
    float* hpre = c->hs_h + (size_t)c->hl_off[i] * 256;
    launch_head_gather_rows(c->pool_out[hd.input - 1], list, M, hd.hh, hd.hw, cv.hin, cv.win, cv.cin, cv.k, windows, st);
    gemm_rows(c, windows, cv.w_packed, M, kkc, cv.cout, hpre);
    HeadTailList t;
    t.hpre = hpre; t.list = list; t.M = M; t.HW = hd.hh * hd.hw;
    t.bias = P(c, cv.p_b); t.prelu = P(c, cv.p_prelu); t.w2 = P(c, hd.p_w2); t.b2 = P(c, hd.p_b2);
    t.out = hd.out;
    launch_head_tail_list(t, st);
    c->launches += 2;
  }
  c->hs_forward = true;
}
static void wgrad_rows(frcnn_ctx* c, const bf16* dy, const bf16* x, int R, int nin, int nout, float* dw, bool zero);
// sparse_heads: the caller (lossAndGradient) has listed, per head, the pixels delta_outputs can be non-zero at
// (c->hl_dev / hl_off / hl_count): the head convolutions' backward runs on those pixels only
static void do_pnet_backward(frcnn_ctx* c, const float* const* d_out, bool keep_block_grads = false, bool sparse_heads = false) {
  FRCNN_REQUIRE(c->train_ready, FRCNN_E_STATE, "pnet:backward needs a preceding training-mode forward on this context");
  for (auto g : c->grads) FRCNN_REQUIRE(g != nullptr, FRCNN_E_STATE, "frcnn_bind_grads must be called before the backward pass");
  const int N = c->ws_n, nb = (int)c->blocks.size();
  cudaStream_t st = c->stream;
  if (!keep_block_grads) zero_block_grads(c);
  // delta_outputs[5]: ROI-pool gradients on the last conv block (objective.lua:184)
  const int nh = (int)c->heads.size();
  if (d_out[nh]) {
    launch_add_chw_to_nhwc(d_out[nh], c->dblock[nb - 1], N, c->feat_h, c->feat_w, c->feat_c, st);
    ++c->launches;
  }
  // anchor heads
  for (int i = 0; i < nh; ++i) {
    if (!d_out[i]) continue;
    Head& hd = c->heads[i];
    HeadTailBwd hb;
    hb.d_out = d_out[i];
    hb.ws = hd.acc;
    hb.splits = hd.conv.launch.p.splits;
    hb.npix = (long)N * hd.hh * hd.hw;
    hb.slice_stride = hb.npix * hd.n;
    hb.HW = hd.hh * hd.hw;
    hb.bias = P(c, hd.conv.p_b); hb.prelu = P(c, hd.conv.p_prelu); hb.w2 = P(c, hd.p_w2);
    hb.dpre = hd.dpre;
    hb.dw2 = G(c, hd.p_w2); hb.db2 = G(c, hd.p_b2); hb.db1 = G(c, hd.conv.p_b); hb.dslope = G(c, hd.conv.p_prelu);
    if (sparse_heads) {
      const int M = c->hl_count[i];
      if (M == 0) continue;               // no listed anchor on this head: all its gradients are zero
      ConvLayer& cv = hd.conv;
      const int* list = c->hl_dev + c->hl_off[i];
      const int kkc = cv.k * cv.k * cv.cin;
      hb.list = list; hb.M = M; hb.dpre = c->hs_d;
      const bf16* windows = c->hs_x;
      if (c->hs_forward) {   // the forward left this head's windows and conv outputs behind
        windows = c->hs_x + head_rows_offset(c, i);
        hb.ws = c->hs_h + (size_t)c->hl_off[i] * 256;
        hb.ws_compact = 1;
      }
      launch_head_tail_bwd(hb, c->sm_count, st);
      // weight gradient: dW[co][tap][ci] = dpre[M][co]^T x windows[M][tap][ci]
      if (!c->hs_forward)
        launch_head_gather_rows(c->pool_out[hd.input - 1], list, M, hd.hh, hd.hw, cv.hin, cv.win, cv.cin, cv.k, c->hs_x, st);
      wgrad_rows(c, c->hs_d, windows, M, kkc, cv.cout, cv.dw_taps, true);
      launch_wgrad_finish(cv.dw_taps, G(c, cv.p_w), cv.cout, cv.cin, cv.k, cv.k, st);
      // data gradient: G[M][tap][ci] = dpre[M][co] x W, scattered onto the windows
      if (hd.rows_gen != c->weights_gen) {
        launch_pack_head_weight_rows(P(c, cv.p_w), hd.w_rows, cv.cout, cv.cin, cv.k, st);
        hd.rows_gen = c->weights_gen;
        ++c->launches;
      }
      gemm_rows(c, c->hs_d, hd.w_rows, M, cv.cout, kkc, c->hs_g);
      launch_head_scatter_rows(c->hs_g, list, M, hd.hh, hd.hw, cv.hin, cv.win, cv.cin, cv.k, c->dblock[hd.input - 1], st);
      c->launches += 4;
      continue;
    }
    launch_head_tail_bwd(hb, c->sm_count, st);
    ++c->launches;
    run_wgrad(c, hd.conv);
    run_dgrad(c, hd.conv);
  }
  dp_bucket_ready(c, 1);   // the anchor networks' gradients are final
  // trunk, last block first
  size_t li_end = c->trunk.size();
  for (int b = nb - 1; b >= 0; --b) {
    const size_t li_first = li_end - c->blocks[b].conv_steps;
    ConvLayer& last = c->trunk[li_end - 1];
    launch_unpool_prelu_bwd(c->dblock[b], c->pool_arg[b], c->pool_out[b], P(c, last.p_prelu), last.dropout > 0.f ? last.mask : nullptr,
                            c->gscratch[0], G(c, last.p_b), G(c, last.p_prelu), N, last.hout, last.wout, last.cout, c->sm_count, st);
    ++c->launches;
    for (size_t li = li_end; li-- > li_first;) {
      ConvLayer& cv = c->trunk[li];
      if (cv.first) {
        FRCNN_REQUIRE(cv.k == 3 && cv.pad == 1 && cv.cout == 64, FRCNN_E_INVALID,
                      "training: the first convolution must be 3 -> 64 channels, 3x3, padding 1 (both reference models)");
        launch_first_wgrad(c->gscratch[0], c->train_img, G(c, cv.p_w), N, cv.hin, cv.win, cv.pad, c->sm_count, st);
        ++c->launches;
        break;
      }
      run_wgrad(c, cv);
      run_dgrad(c, cv);
      if (li > li_first) {
        // gradient wrt the previous conv's output (bf16, gscratch[1]) -> its pre-activation gradient, which becomes
        // the next iteration's dy in gscratch[0]
        ConvLayer& prev = c->trunk[li - 1];
        launch_prelu_bwd(c->gscratch[1], c->gscratch[0], prev.out, P(c, prev.p_prelu), prev.dropout > 0.f ? prev.mask : nullptr,
                         G(c, prev.p_b), G(c, prev.p_prelu), N, prev.hout, prev.wout, prev.cout, c->sm_count, st);
        ++c->launches;
      }
    }
    li_end = li_first;
    dp_bucket_ready(c, 2 + (nb - 1 - b));   // this block's gradients are final
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------- objective.lua stage
static void ensure_objective_workspace(frcnn_ctx* c, int rows) {
  const int NF = std::max(1, c->ws_n);
  if (rows <= c->cw_rows && !c->head_dout.empty() && c->cw_n >= NF && c->cw_gen == c->ws_gen) return;
  rows = std::max(rows, c->cw_rows);
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  free_all(c->cw_allocs);
  c->cw_rows = 0;
  auto& A = c->cw_allocs;
  const int R = std::max(64, (rows + 63) / 64 * 64);
  const int feat = c->roi_kh * c->roi_kw * c->feat_c;
  c->ex_dev = (ExampleDev*)dev_alloc(A, (size_t)R * sizeof(ExampleDev));
  c->ex_rects = (double*)dev_alloc(A, (size_t)R * 4 * sizeof(double));
  c->crtarget = (float*)dev_alloc(A, (size_t)R * 4 * sizeof(float));
  c->cctarget = (int*)dev_alloc(A, (size_t)R * sizeof(int));
  c->losses_dev = (float*)dev_alloc(A, (size_t)NF * 8 * sizeof(float));
  c->losses_cur = c->losses_dev;
  c->t_status = (int*)dev_alloc(A, 4 * sizeof(int));
  c->t_rows = (bf16*)dev_alloc(A, (size_t)R * feat * sizeof(bf16));
  c->t_argmax = (int*)dev_alloc(A, (size_t)R * feat * sizeof(int));
  c->t_dx = (float*)dev_alloc(A, (size_t)R * feat * sizeof(float));
  const int last_n = c->fcs.back().nout;
  c->t_dhidden = (float*)dev_alloc(A, (size_t)R * last_n * sizeof(float));
  c->t_dz = (float*)dev_alloc(A, (size_t)R * (c->class_count + 5) * sizeof(float));
  for (auto& f : c->fcs) {
    const size_t rn = (size_t)R * f.nout;
    f.t_acc = (float*)dev_alloc(A, rn * sizeof(float));
    f.t_pre = (float*)dev_alloc(A, rn * sizeof(float));
    f.t_xhat = f.bn ? (float*)dev_alloc(A, rn * sizeof(float)) : nullptr;
    f.t_rstd = (float*)dev_alloc(A, (size_t)NF * f.nout * sizeof(float));   // BatchNorm 1/std of every frame's ROI batch
    f.t_stat = (float*)dev_alloc(A, (size_t)NF * 2 * f.nout * sizeof(float));   // and its mean / unbiased variance
    f.t_mask = (float*)dev_alloc(A, rn * sizeof(float));
    f.t_din = (float*)dev_alloc(A, rn * sizeof(float));
    f.t_out32 = (float*)dev_alloc(A, rn * sizeof(float));
    f.t_out = (bf16*)dev_alloc(A, rn * sizeof(bf16));
    f.t_dy = (bf16*)dev_alloc(A, rn * sizeof(bf16));
    f.w_dgrad = (bf16*)dev_alloc(A, (size_t)f.nin * f.nout * sizeof(bf16));
    f.dgrad_gen = -1;
    f.dw_taps = (float*)dev_alloc(A, (size_t)f.nin * f.nout * sizeof(float));
  }
  c->head_dout.clear();
  for (auto& hd : c->heads) c->head_dout.push_back((float*)dev_alloc(A, (size_t)NF * 18 * hd.hh * hd.hw * sizeof(float)));
  size_t max_kkc = 0;
  for (auto& hd : c->heads) {
    const size_t kkc = (size_t)hd.conv.k * hd.conv.k * hd.conv.cin;
    max_kkc = std::max(max_kkc, kkc);
    hd.w_rows = (bf16*)dev_alloc(A, kkc * hd.n * sizeof(bf16));
    hd.rows_gen = -1;
  }
  c->hl_dev = (int*)dev_alloc(A, (size_t)R * sizeof(int));
  c->hs_d = (bf16*)dev_alloc(A, (size_t)R * 256 * sizeof(bf16));
  c->hs_x = (bf16*)dev_alloc(A, (size_t)R * max_kkc * sizeof(bf16));
  c->hs_h = (float*)dev_alloc(A, (size_t)R * 256 * sizeof(float));
  c->hs_g = (float*)dev_alloc(A, (size_t)R * max_kkc * sizeof(float));
  c->cw_rows = R;
  c->cw_n = NF;
  c->cw_gen = c->ws_gen;
}

// [R][nin] bf16 rows x [nout][nin] bf16 weights -> fp32 [R][nout].  Few output tiles with a long K (fc1: 8 tiles,
// K = 13824) are split along K and summed by TMA reduce-add into the zeroed output; otherwise plain stores.
static void gemm_rows(frcnn_ctx* c, const bf16* a, const bf16* w, int R, int nin, int nout, float* out) {
  ConvLaunch L;
  const int tiles = ((R + 127) / 128) * ((nout + 255) / 256);
  const int k_iters = nin / 64;
  int splits = std::min(std::max(1, c->sm_count / std::max(tiles, 1)), std::max(1, k_iters / 8));
  if (splits > 1) {
    FRCNN_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)R * nout * sizeof(float), c->stream));
    conv_prepare(&L, a, w, 1, 1, R, nin, nout, 1, 1, 0, 0, EPI_F32_REDUCE, nullptr, c->sm_count, splits, 0, 1);
  } else {
    conv_prepare(&L, a, w, 1, 1, R, nin, nout, 1, 1, 0, 0, EPI_F32_SLICES, nullptr, c->sm_count, 1, 0, 1);
  }
  conv_set_f32_output(&L, out);
  conv_launch(L, c->stream);
  ++c->launches;
}
// dW[nout][nin] (fp32 taps buffer, zeroed here) = dy[R][nout]^T x[R][nin]
static void wgrad_rows(frcnn_ctx* c, const bf16* dy, const bf16* x, int R, int nin, int nout, float* dw, bool zero) {
  if (zero) FRCNN_CUDA_TRY(cudaMemsetAsync(dw, 0, (size_t)nin * nout * sizeof(float), c->stream));
  ConvLaunch L;
  conv_wgrad_prepare(&L, dy, x, dw, 1, 1, R, nin, nout, 1, 1, 0, 0, c->sm_count);
  conv_launch(L, c->stream);
  ++c->launches;
}

// cnet:forward (training) -> detection-stage criteria -> cnet:backward (objective.lua:164-179) for the example rows of
// nf frames stored back to back in c->t_rows ([rows][bins][C] bf16; frame f = rows [off[f], off[f] + R[f])) with targets
// c->crtarget / c->cctarget; leaves d(loss)/d(rows) in c->t_dx and adds frame f's losses to c->losses_dev[8 f + 2..3].
// Stage-wise over the frames: every Linear layer is ONE tensor-core GEMM over all rows (forward, data gradient, weight
// gradient), and BatchNormalization / PReLU / Dropout and the criteria are ONE launch each with the frame on blockIdx.y
// (or looked up from the row) -- in the reference cnet sees the ROI batch of one image at a time (objective.lua:164
// inside the per-image loop), so the batch statistics, the running-statistics updates (in frame order) and the mean
// over the frame's rows stay per frame inside those launches.
static FrameList make_frames(int nf, const int* off, const int* R, const int* n_pos) {
  FrameList fl;
  memset(&fl, 0, sizeof(fl));
  fl.nf = nf;
  for (int f = 0; f < nf; ++f) { fl.off[f] = off[f]; fl.R[f] = R[f]; fl.n_pos[f] = n_pos ? n_pos[f] : 0; }
  return fl;
}
static void cnet_train_forward(frcnn_ctx* c, const FrameList& fl, const float* const* cnet_masks, const uint64_t* seeds) {
  cudaStream_t st = c->stream;
  const int rows = fl.rows();
  if (rows <= 0) return;
  FrameSeeds fs = {};
  for (int f = 0; f < fl.nf; ++f) fs.seed[f] = seeds[f];
  // ---- cnet forward, training mode (objective.lua:164)
  const bf16* in = c->t_rows;
  for (size_t i = 0; i < c->fcs.size(); ++i) {
    FcLayer& f = c->fcs[i];
    gemm_rows(c, in, f.w_packed, rows, f.nin, f.nout, f.t_acc);
    if (cnet_masks && cnet_masks[i]) {   // explicit masks: single-frame (test) facility
      FRCNN_CUDA_TRY(cudaMemcpyAsync(f.t_mask, cnet_masks[i], (size_t)rows * f.nout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      launch_dropout_mask_frames(f.t_mask, fl, f.nout, f.dropout, fs, 100u + (uint32_t)i, st);
    }
    FcTrainFwd ff;
    ff.acc = f.t_acc; ff.bias = P(c, f.p_b); ff.bn_w = P(c, f.p_bn_w); ff.bn_b = P(c, f.p_bn_b); ff.prelu = P(c, f.p_prelu);
    ff.bn_mean = const_cast<float*>(P(c, f.p_bn_mean)); ff.bn_var = const_cast<float*>(P(c, f.p_bn_var));
    ff.mask = f.t_mask; ff.keep_scale = f.dropout > 0.f ? 1.f / (1.f - f.dropout) : 1.f;
    ff.pre = f.t_pre; ff.xhat = f.t_xhat; ff.rstd = f.t_rstd; ff.stat = f.t_stat;
    ff.out_bf16 = f.t_out; ff.out_f32 = f.t_out32; ff.n = f.nout;
    launch_fc_train_fwd(ff, fl, st);
    c->launches += 2 + ((f.bn && fl.nf > 1) ? 1 : 0);
    in = f.t_out;
  }
}
// ext_dreg / ext_dcls: gradients wrt cnet's two outputs supplied by the caller (cnet:backward(cinput, {crdelta, ccdelta}),
// objective.lua:179) instead of the built-in criteria
static void cnet_train_loss_bwd(frcnn_ctx* c, const FrameList& fl, const float* ext_dreg = nullptr, const float* ext_dcls = nullptr) {
  if (fl.rows() <= 0) return;
  // ---- detection-stage criteria + backward through the two output branches (objective.lua:166-179): one launch over
  // the rows of all frames, each row averaged over ITS frame's ROI batch
  FcLayer& last = c->fcs.back();
  CnetLossParams cl;
  cl.hidden = last.t_out32;
  cl.w_reg = P(c, c->p_reg_w); cl.b_reg = P(c, c->p_reg_b); cl.w_cls = P(c, c->p_cls_w); cl.b_cls = P(c, c->p_cls_b);
  cl.crtarget = c->crtarget; cl.cctarget = c->cctarget;
  cl.nin = last.nout; cl.ncls = c->class_count + 1;
  cl.d_hidden = c->t_dhidden; cl.dz = c->t_dz;
  cl.g_w_reg = G(c, c->p_reg_w); cl.g_b_reg = G(c, c->p_reg_b); cl.g_w_cls = G(c, c->p_cls_w); cl.g_b_cls = G(c, c->p_cls_b);
  cl.losses = fl.nf == 1 ? c->losses_cur : c->losses_dev;
  cl.loss_stride = fl.nf == 1 ? 0 : 8;
  cl.ext_dreg = ext_dreg;
  cl.ext_dcls = ext_dcls;
  launch_cnet_loss_bwd(cl, fl, c->stream);
  c->launches += 2;
}
static void cnet_train_backward(frcnn_ctx* c, const FrameList& fl) {
  cudaStream_t st = c->stream;
  const int bins = c->roi_kh * c->roi_kw;
  const int rows = fl.rows();
  if (rows <= 0) return;
  // ---- cnet backward (objective.lua:179)
  const float* d_in = c->t_dhidden;
  for (int i = (int)c->fcs.size() - 1; i >= 0; --i) {
    FcLayer& f = c->fcs[i];
    FcTrainBwd fb;
    fb.d_in = d_in; fb.pre = f.t_pre; fb.xhat = f.t_xhat; fb.rstd = f.t_rstd;
    fb.bn_w = P(c, f.p_bn_w); fb.prelu = P(c, f.p_prelu);
    fb.mask = f.t_mask; fb.keep_scale = f.dropout > 0.f ? 1.f / (1.f - f.dropout) : 1.f;
    fb.d_out_bf16 = f.t_dy;
    fb.g_bias = G(c, f.p_b); fb.g_bn_w = G(c, f.p_bn_w); fb.g_bn_b = G(c, f.p_bn_b); fb.g_prelu = G(c, f.p_prelu);
    fb.n = f.nout;
    launch_fc_train_bwd(fb, fl, st);
    ++c->launches;
    const bf16* x_in = i == 0 ? c->t_rows : c->fcs[i - 1].t_out;
    // weight gradient over ALL rows: one fp32 TMA reduce-add GEMM into dw_taps, transposed into Torch's layout
    wgrad_rows(c, f.t_dy, x_in, rows, f.nin, f.nout, f.dw_taps, true);
    const bool perm = i == 0;
    launch_wgrad_finish_fc(f.dw_taps, G(c, f.p_w), f.nout, perm ? c->feat_c : f.nin, perm ? bins : 1, perm ? 1 : 0, st);
    ++c->launches;
    if (f.dgrad_gen != c->weights_gen) {  // the transposed bf16 weights of the data gradient: once per weight update
      launch_pack_fc_weight_dgrad(P(c, f.p_w), f.w_dgrad, f.nout, perm ? c->feat_c : f.nin, perm ? bins : 1, perm ? 1 : 0, st);
      f.dgrad_gen = c->weights_gen;
      ++c->launches;
    }
    float* d_prev = i == 0 ? c->t_dx : c->fcs[i - 1].t_din;
    gemm_rows(c, f.t_dy, f.w_dgrad, rows, f.nout, f.nin, d_prev);
    d_in = d_prev;
  }
}
static void run_cnet_train_frames(frcnn_ctx* c, const FrameList& fl, const float* const* cnet_masks, const uint64_t* seeds) {
  cnet_train_forward(c, fl, cnet_masks, seeds);
  cnet_train_loss_bwd(c, fl);
  cnet_train_backward(c, fl);
}

static void run_cnet_train(frcnn_ctx* c, int R, int n_pos, const float* const* cnet_masks, uint64_t seed) {
  const int off = 0;
  run_cnet_train_frames(c, make_frames(1, &off, &R, &n_pos), cnet_masks, &seed);
}

// The per-image loop of lossAndGradient (objective.lua:65-198) for N frames of one size: pnet forward (training) and
// pnet:backward run ONCE over the whole batch (the conv kernels then work at batch efficiency and every elementwise
// backward kernel is launched once instead of N times); the criteria, the ROI pooling and the cnet stage -- whose
// BatchNormalization sees the ROI batch of ONE image (objective.lua:164 sits inside the per-image loop) -- run frame
// by frame in between, each writing its own slice of delta_outputs.  One host synchronisation per batch.
static void do_train_batch(frcnn_ctx* c, const float* img_dev, int N, int H, int W, const frcnn_example* const* pos, const int* n_pos,
                           const frcnn_example* const* neg, const int* n_neg, const float* const* pnet_masks,
                           const float* const* cnet_masks, const uint64_t* seeds, float* losses_host /* [N][4] */) {
  static_assert(sizeof(frcnn_example) == sizeof(ExampleDev), "example layouts must match");
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called first");
  for (auto g : c->grads) FRCNN_REQUIRE(g != nullptr, FRCNN_E_STATE, "frcnn_bind_grads must be called first");
  FRCNN_REQUIRE(N >= 1 && N <= 64, FRCNN_E_INVALID, "train: 1..64 frames per call");
  FRCNN_REQUIRE(cnet_masks == nullptr || N == 1, FRCNN_E_INVALID, "explicit cnet masks are a single-frame (test) facility");
  std::vector<int> off(N), Rn(N);
  int rows = 0;
  for (int n = 0; n < N; ++n) {
    off[n] = rows;
    Rn[n] = n_pos[n] + n_neg[n];
    rows += Rn[n];
  }
  cudaStream_t st = c->stream;
  ensure_pnet_workspace(c, N, H, W);
  ensure_train_workspace(c, N, H, W);
  ensure_objective_workspace(c, rows);
  const FrameList fl = make_frames(N, off.data(), Rn.data(), n_pos);
  // FRCNN_HEAD_SPARSE=0: dense anchor-head backward (the path pnet:backward with caller-supplied deltas always takes)
  const char* hs_env = getenv("FRCNN_HEAD_SPARSE");
  const bool sparse_heads = !(hs_env && atoi(hs_env) == 0);
  for (int l = 0; l < MAX_HEADS; ++l) c->hl_count[l] = 0;
  FrameList per_frame;   // "one row per frame": the SpatialDropout masks
  memset(&per_frame, 0, sizeof(per_frame));
  per_frame.nf = N;
  FrameSeeds fs = {};
  for (int n = 0; n < N; ++n) { per_frame.off[n] = n; per_frame.R[n] = 1; fs.seed[n] = seeds[n]; }
  // ---- pnet forward, training mode (objective.lua:60,71)
  int mi = 0;
  for (auto& cv : c->trunk) {
    if (cv.dropout <= 0.f) continue;
    if (pnet_masks && pnet_masks[mi]) {
      FRCNN_CUDA_TRY(cudaMemcpyAsync(cv.mask, pnet_masks[mi], (size_t)N * cv.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      launch_dropout_mask_frames(cv.mask, per_frame, cv.cout, cv.dropout, fs, (uint32_t)mi, st);   // one draw per (frame, channel)
    }
    ++mi;
  }
  c->hs_forward = false;
  do_pnet_forward(c, img_dev, N, H, W, true, sparse_heads);
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->losses_dev, 0, (size_t)N * 8 * sizeof(float), st));
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->t_status, 0, 4 * sizeof(int), st));
  for (size_t i = 0; i < c->heads.size(); ++i)
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->head_dout[i], 0, (size_t)N * 18 * c->heads[i].hh * c->heads[i].hw * sizeof(float), st));
  zero_block_grads(c);
  const int bins = c->roi_kh * c->roi_kw;
  const size_t fmap_elems = (size_t)c->feat_h * c->feat_w * c->feat_c;
  if (rows > 0) {
    // all example records of the batch in one upload (frame f = rows [off[f], off[f] + R[f]): positives, then negatives)
    std::vector<ExampleDev>& ex = c->ex_host;   // staging owned by the context: alive until the next call
    ex.resize((size_t)rows);
    for (int n = 0; n < N; ++n) {
      if (n_pos[n]) memcpy(&ex[off[n]], pos[n], (size_t)n_pos[n] * sizeof(ExampleDev));
      if (n_neg[n]) memcpy(&ex[off[n] + n_pos[n]], neg[n], (size_t)n_neg[n] * sizeof(ExampleDev));
    }
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->ex_dev, ex.data(), (size_t)rows * sizeof(ExampleDev), cudaMemcpyHostToDevice, st));
    if (sparse_heads) {
      // the pixels of every head that carry a listed anchor (several aspects / examples may share one): the only places
      // delta_outputs is written (objective.lua:102-103,131); records that index outside their map are skipped here and
      // reported by rpn_loss_kernel's status flag
      std::vector<int> per_head[MAX_HEADS];
      const int nh = (int)c->heads.size();
      for (int n = 0; n < N; ++n)
        for (int e = off[n]; e < off[n] + Rn[n]; ++e) {
          const ExampleDev& x = ex[e];
          const int l = x.layer - 1, a = x.aspect - 1, yy = x.y - 1, xx = x.x - 1;
          if (l < 0 || l >= nh || a < 0 || a >= 3) continue;
          const Head& hd = c->heads[l];
          if (yy < 0 || yy >= hd.hh || xx < 0 || xx >= hd.hw) continue;
          per_head[l].push_back((n * hd.hh + yy) * hd.hw + xx);
        }
      c->hl_host.clear();
      for (int l = 0; l < MAX_HEADS; ++l) {
        auto& v = per_head[l];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        c->hl_off[l] = (int)c->hl_host.size();
        c->hl_count[l] = (int)v.size();
        c->hl_host.insert(c->hl_host.end(), v.begin(), v.end());
      }
      if (!c->hl_host.empty())
        FRCNN_CUDA_TRY(cudaMemcpyAsync(c->hl_dev, c->hl_host.data(), c->hl_host.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      sparse_head_forward(c);
    }
    // ---- RPN criteria on the listed anchors (objective.lua:91-140), all frames in one launch
    RpnLossParams lp;
    lp.ex = c->ex_dev;
    for (int i = 0; i < MAX_HEADS; ++i) {
      lp.out[i] = c->heads[i].out; lp.d_out[i] = c->head_dout[i]; lp.hh[i] = c->heads[i].hh; lp.hw[i] = c->heads[i].hw;
    }
    lp.crtarget = c->crtarget; lp.cctarget = c->cctarget; lp.bg_class = c->class_count;
    lp.rects = c->ex_rects;
    lp.losses = c->losses_dev; lp.status = c->t_status;
    launch_rpn_loss(lp, fl, st);
    // ---- ROI pooling of ground-truth rects / negative anchors (objective.lua:117-119,137-139)
    launch_roi_pool_train(c->pool_out.back(), c->feat_h, c->feat_w, c->feat_c, c->roi_kh, c->roi_kw, c->roi_loc, c->ex_rects, fl,
                          c->t_rows, c->t_argmax, c->t_status + 1, st);
    c->launches += 2;
    c->losses_cur = c->losses_dev;
    run_cnet_train_frames(c, fl, cnet_masks, seeds);
    dp_bucket_ready(c, 0);   // cnet's gradients are final: their all-reduce overlaps pnet:backward
    // ---- ROI-pool backward into delta_outputs[5] (objective.lua:182-185), kept as the fp32 NHWC block gradient
    launch_roi_pool_bwd(c->t_dx, c->t_argmax, fl, bins, c->feat_c, c->dblock.back(), (long)fmap_elems, st);
    ++c->launches;
  }
  c->losses_cur = c->losses_dev;
  // The losses and the status flags are final here, before pnet:backward: they go to page-locked memory behind an event,
  // and the call returns on THAT event -- the caller's host work between two steps (cleanAnchors, marshalling) then
  // overlaps the backward pass instead of leaving the GPU idle.  The gradient accumulation is still ordered on the
  // context's stream (FRCNN_TRAIN_SYNC=1: wait for the whole step, the round-1 behaviour).
  if (!c->h_loss) {
    FRCNN_CUDA_TRY(cudaMallocHost(&c->h_loss, (size_t)(MAX_TRAIN_FRAMES * 8 + 4) * sizeof(float)));
    FRCNN_CUDA_TRY(cudaEventCreateWithFlags(&c->loss_ev, cudaEventDisableTiming));
  }
  float* lh = c->h_loss;
  int* sh = reinterpret_cast<int*>(c->h_loss + MAX_TRAIN_FRAMES * 8);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(lh, c->losses_dev, (size_t)N * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(sh, c->t_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  FRCNN_CUDA_TRY(cudaEventRecord(c->loss_ev, st));
  // ---- pnet backward (objective.lua:189), all frames at once
  std::vector<const float*> d_out(c->heads.size() + 1, nullptr);
  for (size_t i = 0; i < c->heads.size(); ++i) d_out[i] = c->head_dout[i];
  do_pnet_backward(c, d_out.data(), true, sparse_heads);
  if (sparse_heads) c->train_ready = false;   // the dense head activations of this forward do not exist: a later
                                              // pnet:backward must follow its own pnet:forward
  const char* ts_env = getenv("FRCNN_TRAIN_SYNC");
  if (ts_env && atoi(ts_env) != 0) FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
  else FRCNN_CUDA_TRY(cudaEventSynchronize(c->loss_ev));
  for (int n = 0; n < N; ++n)
    for (int i = 0; i < 4; ++i) losses_host[4 * n + i] = lh[8 * n + i];
  FRCNN_REQUIRE(sh[0] == 0, FRCNN_E_INVALID, "an example indexes outside its anchor map: apply cleanAnchors first (objective.lua:32-43)");
  FRCNN_REQUIRE(sh[1] == 0, FRCNN_E_ROI_EMPTY, "an ROI clipped to max == 0; the reference raises an index error here (objective.lua:11)");
}

static void do_train_image(frcnn_ctx* c, const float* img_dev, int H, int W, const frcnn_example* pos, int n_pos,
                           const frcnn_example* neg, int n_neg, const float* const* pnet_masks, const float* const* cnet_masks,
                           uint64_t seed, float losses_host[4]) {
  do_train_batch(c, img_dev, 1, H, W, &pos, &n_pos, &neg, &n_neg, pnet_masks, cnet_masks, &seed, losses_host);
}

static int cnet_ctas(const frcnn_ctx* c) {
  if (const char* e = getenv("FRCNN_CNET_CTAS")) return atoi(e);
  return c->schedule == FRCNN_SCHED_THROUGHPUT ? std::max(1, c->sm_count / 4) : c->sm_count;
}

// detector / cnet workspace for a batch of N images
static void ensure_det_workspace(frcnn_ctx* c, int N, int min_rows) {
  int want_rows = std::max(N * c->cand_cap, min_rows);
  if (c->det_n >= N && c->roi_cap >= want_rows) return;
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  free_all(c->det_allocs);
  if (c->nms_mem) { cudaFree(c->nms_mem); c->nms_mem = nullptr; }
  c->det_n = 0;
  auto& A = c->det_allocs;
  const int cap = c->cand_cap;
  const size_t NC = (size_t)N * cap;
  c->cand_r = (double*)dev_alloc(A, NC * 4 * sizeof(double));
  c->cand_box = (float4*)dev_alloc(A, NC * sizeof(float4));
  c->cand_logp = (float*)dev_alloc(A, NC * sizeof(float));
  c->cand_anchor = (int4*)dev_alloc(A, NC * sizeof(int4));
  c->flags = (int*)dev_alloc(A, (16 + 2 * (size_t)N) * sizeof(int));  // flags | match counts | accepted counts
  c->cand_count = c->flags + 16;
  c->n_pass = c->flags + 16 + N;
  c->ticket = (unsigned long long*)dev_alloc(A, N * sizeof(unsigned long long));
  c->status_blocks = 4096;
  c->status = (unsigned long long*)dev_alloc(A, (size_t)N * c->status_blocks * sizeof(unsigned long long));
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->ticket, 0, N * sizeof(unsigned long long), c->stream));
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->status, 0, (size_t)N * c->status_blocks * sizeof(unsigned long long), c->stream));
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->flags, 0, (16 + 2 * (size_t)N) * sizeof(int), c->stream));
  c->decode_nblocks = 0;
  c->pick1 = (int*)dev_alloc(A, NC * sizeof(int));
  c->count1 = (int*)dev_alloc(A, N * sizeof(int));
  c->roi_base = (int*)dev_alloc(A, N * sizeof(int));
  const int R = want_rows;
  const int bins = c->roi_kh * c->roi_kw;
  c->roi_out = (bf16*)dev_alloc(A, (size_t)R * bins * c->feat_c * sizeof(bf16));
  c->roi_img = (int*)dev_alloc(A, (size_t)R * sizeof(int));
  c->roi_rect = (int4*)dev_alloc(A, (size_t)R * sizeof(int4));
  c->roi_cand = (int*)dev_alloc(A, (size_t)R * sizeof(int));
  const int ncls = c->class_count + 1;
  c->reg_out = (float*)dev_alloc(A, (size_t)R * 4 * sizeof(float));
  c->cls_out = (float*)dev_alloc(A, (size_t)R * ncls * sizeof(float));
  c->fin_r2 = (double*)dev_alloc(A, (size_t)R * 4 * sizeof(double));
  c->fin_box = (float4*)dev_alloc(A, (size_t)R * sizeof(float4));
  c->fin_cls = (int*)dev_alloc(A, (size_t)R * sizeof(int));
  c->fin_conf = (float*)dev_alloc(A, (size_t)R * sizeof(float));
  c->gbox = (float4*)dev_alloc(A, NC * sizeof(float4));
  c->grow = (int*)dev_alloc(A, NC * sizeof(int));
  c->det_cap = (int)NC;
  c->det_dev = (frcnn_detection*)dev_alloc(A, (size_t)c->det_cap * sizeof(frcnn_detection));
  // cnet layers: GEMM rows = R
  const bf16* in = c->roi_out;
  for (size_t i = 0; i < c->fcs.size(); ++i) {
    FcLayer& f = c->fcs[i];
    const bool last = i + 1 == c->fcs.size();
    f.out_bf16 = last ? nullptr : (bf16*)dev_alloc(A, (size_t)R * f.nout * sizeof(bf16));
    f.out_f32 = last ? (float*)dev_alloc(A, (size_t)R * f.nout * sizeof(float)) : nullptr;
    // split-K sized for the detector's typical few hundred ROIs: 8 K-iterations per split
    int k_iters = f.nin / 64;
    int splits = std::max(1, k_iters / 8);
    // deterministic split-K: every split writes its own fp32 slice (no reduce-add, no memset); fc_tail sums the slices in
    // ascending order.  Tile-major slices bound the workspace by max(M tiles, dyn_ctas / N tiles) tiles of 128 rows.
    conv_prepare(&f.launch, in, f.w_packed, 1, 1, R, f.nin, f.nout, 1, 1, 0, 0, EPI_F32_SLICES, nullptr, c->sm_count, splits, 0, 0, 2);
    const ConvParams& fp = f.launch.p;
    const size_t ws_tiles = (size_t)std::max(fp.n_tiles_m, c->sm_count / std::max(1, fp.n_tiles_n) + 1) + 1;
    f.acc = (float*)dev_alloc(A, ws_tiles * 128 * f.nout * sizeof(float));
    conv_set_f32_output(&f.launch, f.acc);
    f.launch.p.slice_tile_major = 1;
    f.launch.p.m_limit = c->flags + 2;  // roi_total
    // split-K factor chosen on the device from the live row count so that the units fill about dyn_ctas CTAs: the whole
    // machine on the latency schedule, a quarter of it on the throughput schedule (measured: 37 CTAs cost a single
    // frame nothing and leave the other SMs to the frames in flight; the 148-way TMA reduce-add contention goes away)
    f.launch.p.dyn_ctas = cnet_ctas(c);
    f.launch.grid = std::min(f.launch.grid, f.launch.p.dyn_ctas);   // no idle CTAs: each would still claim a whole SM (TMEM, 200 KB)
    in = f.out_bf16;
  }
  // NMS workspace: segments = max(N images, N * classes)
  c->nms_cap_total = (int)NC;
  c->nms_cap_seg = N * std::max(1, c->class_count);
  c->nms_bytes = nms_workspace_bytes(c->nms_cap_total, c->nms_cap_seg);
  FRCNN_CUDA_TRY(cudaMalloc(&c->nms_mem, c->nms_bytes));
  nms_workspace_init(&c->nms, c->nms_mem, c->nms_bytes, c->nms_cap_total, c->nms_cap_seg);
  if (c->h_ints) cudaFreeHost(c->h_ints);
  FRCNN_CUDA_TRY(cudaMallocHost(&c->h_ints, (96 + 2 * (size_t)N) * sizeof(int)));
  ++c->ws_gen;
  if (c->h_det_cap < c->det_cap) {
    if (c->h_det) cudaFreeHost(c->h_det);
    FRCNN_CUDA_TRY(cudaMallocHost(&c->h_det, (size_t)c->det_cap * sizeof(frcnn_detection)));
    c->h_det_cap = c->det_cap;
  }
  c->det_n = N;
  c->roi_cap = R;
}

// grid sizing of the small cnet kernels: the typical live row count (the grid-stride loops cover more); FRCNN_CNET_GRID_ROWS
// overrides it for measurements
static int env_rows() {
  static const int v = getenv("FRCNN_CNET_GRID_ROWS") ? atoi(getenv("FRCNN_CNET_GRID_ROWS")) : 128;
  return v;
}

// cnet on rows [0, *roi_total) of roi_out (bf16, [bins][C] order)
static void run_cnet(frcnn_ctx* c, int rows_max) {
  for (size_t i = 0; i < c->fcs.size(); ++i) {
    FcLayer& f = c->fcs[i];
    f.launch.p.f16 = c->eval_f16;   // cnet:forward here is evaluate mode (Detector.lua:101): the rows are in that format
    conv_launch_timed(c, f.launch, c->profiling ? c->prof_rows : -1);
    const ConvParams& fp = f.launch.p;
    const FcSlices sl = {fp.k_iters, fp.splits, fp.n_tiles_n, fp.dyn_ctas};
    launch_fc_tail(f.acc, P(c, f.p_b), P(c, f.p_bn_w), P(c, f.p_bn_b), P(c, f.p_bn_mean), P(c, f.p_bn_var), P(c, f.p_prelu),
                   f.out_bf16, f.out_f32, rows_max, c->flags + 2, f.nout, c->stream, &sl, env_rows(), c->eval_f16);
    ++c->launches;
  }
  const FcLayer& last = c->fcs.back();
  launch_cnet_out(last.out_f32, P(c, c->p_reg_w), P(c, c->p_reg_b), P(c, c->p_cls_w), P(c, c->p_cls_b), c->reg_out, c->cls_out,
                  rows_max, c->flags + 2, last.nout, c->class_count + 1, c->stream);
  ++c->launches;
  FRCNN_CUDA_TRY(cudaGetLastError());
}

static void run_decode(frcnn_ctx* c, const float* const* heads_dev, int N, int H, int W, double threshold) {
  DecodeParams p;
  int off = 0;
  for (int i = 0; i < MAX_HEADS; ++i) {
    p.head[i] = heads_dev[i];
    p.hh[i] = c->heads[i].hh;
    p.hw[i] = c->heads[i].hw;
    p.offs[i] = off;
    off += p.hh[i] * p.hw[i] * 3;
  }
  p.offs[MAX_HEADS] = off;
  p.total = off;
  p.w_lut = c->d_w_lut;
  p.h_lut = c->d_h_lut;
  p.img_w = W; p.img_h = H; p.threshold = threshold;
  p.cap = c->cand_cap;
  p.cand_r = c->cand_r; p.cand_box = c->cand_box; p.cand_logp = c->cand_logp; p.cand_anchor = c->cand_anchor;
  p.cand_count = c->cand_count;
  p.cand_overflow = c->flags + 0;
  p.ticket = c->ticket;
  p.status = c->status;
  p.nblocks = (off + 255) / 256;
  FRCNN_REQUIRE(p.nblocks <= c->status_blocks, FRCNN_E_INVALID, "too many anchors for the decode scan state");
  // status words are laid out with the per-image stride nblocks and tagged with the launch number derived from the
  // running ticket counter; a different block count restarts the numbering
  if (c->decode_nblocks != p.nblocks) {
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->ticket, 0, c->det_n * sizeof(unsigned long long), c->stream));
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->status, 0, (size_t)c->det_n * c->status_blocks * sizeof(unsigned long long), c->stream));
    c->decode_nblocks = p.nblocks;
    ++c->ws_gen;
  }
  launch_rpn_decode(p, N, c->stream);
  ++c->launches;
}

// Enqueues the whole Detector:detect pipeline (Detector.lua:31-136) plus the result copies on the ctx stream.
static int detect_stop_after() {
  // FRCNN_DETECT_STOP (measurement only: tools/stage_costs.sh): 1 = trunk, 2 = + anchor heads, 3 = + decode / NMS,
  // 4 = + ROI pooling, 5 = + cnet; 0 / unset = the whole pipeline.  Winners are meaningless when set.
  static const int v = getenv("FRCNN_DETECT_STOP") ? atoi(getenv("FRCNN_DETECT_STOP")) : 0;
  return v;
}

static void enqueue_detect(frcnn_ctx* c, const float* img_dev, int N, int H, int W) {
  cudaStream_t st = c->stream;
  const bool prof = c->profiling;
  const int stop = detect_stop_after();
  auto copy_back = [&]() {
    if (prof) for (int i = 1; i <= 5; ++i) cudaEventRecord(c->ev[i], st);
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_ints, c->flags, (16 + 2 * (size_t)c->det_n) * sizeof(int), cudaMemcpyDeviceToHost, st));
  };
  if (prof) {
    conv_profile_begin(c);
    cudaEventRecord(c->ev[0], st);
  }
  do_pnet_forward(c, img_dev, N, H, W);
  if (prof) cudaEventRecord(c->ev[1], st);
  if (stop == 1 || stop == 2) return copy_back();
  // --- Detector.lua:36-66
  const float* heads_dev[MAX_HEADS];
  for (int i = 0; i < MAX_HEADS; ++i) heads_dev[i] = c->heads[i].out;
  run_decode(c, heads_dev, N, H, W, c->thr_fg);
  // --- Detector.lua:68-85: nms(bb, 0.25, score) -- the score tensor is ignored, order key = y2 (nms.lua:41-42)
  if (c->cand_cap <= NMS_CTA_MAX_SEG) {
    nms_set_segments_from_counts(&c->nms, c->cand_count, N, c->cand_cap, st);
    c->nms.fused = false;  // up to cand_cap matches per image: sort / matrix / resolve as three launches
    c->launches += nms_run(&c->nms, reinterpret_cast<const float*>(c->cand_box), 4, N, N * c->cand_cap, c->cand_cap, c->thr_nms1,
                           FRCNN_NMS_ORDER_Y2, 0, st, nullptr);
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->pick1, c->nms.st.pick, (size_t)N * c->cand_cap * sizeof(int), cudaMemcpyDeviceToDevice, st));
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->count1, c->nms.st.counts, N * sizeof(int), cudaMemcpyDeviceToDevice, st));
  } else {
    // The candidate capacity has grown past the CTA-level NMS (a frame with more than 8192 matches above the
    // foreground threshold: rare, the match list is unbounded in the reference, Detector.lua:59).  The radix-sort path
    // takes its segment table from the host: read the match counts back and run it image by image (eager only).
    int* hs = c->h_ints + 16 + 2 * c->det_n;  // pinned scratch behind the counters: [0..1] segment table, [8..24) counts
    FRCNN_CUDA_TRY(cudaMemcpyAsync(hs + 8, c->cand_count, std::min(N, 16) * sizeof(int), cudaMemcpyDeviceToHost, st));
    FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
    FRCNN_REQUIRE(N <= 16, FRCNN_E_OVERFLOW, "more than 8192 matches per image is supported for batches of at most 16 frames");
    int counts[16];
    for (int i = 0; i < N; ++i) counts[i] = hs[8 + i];
    for (int i = 0; i < N; ++i) {
      const int n = counts[i];
      if (n == 0) {
        FRCNN_CUDA_TRY(cudaMemsetAsync(c->count1 + i, 0, sizeof(int), st));
        continue;
      }
      hs[0] = 0; hs[1] = n;
      FRCNN_CUDA_TRY(cudaMemcpyAsync(c->nms.st.seg_beg, hs, sizeof(int), cudaMemcpyHostToDevice, st));
      FRCNN_CUDA_TRY(cudaMemcpyAsync(c->nms.st.seg_len, hs + 1, sizeof(int), cudaMemcpyHostToDevice, st));
      FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
      c->nms.fused = false;
      c->nms.seg_counts = nullptr;
      c->launches += nms_run(&c->nms, reinterpret_cast<const float*>(c->cand_box + (size_t)i * c->cand_cap), 4, 1, n, n, c->thr_nms1,
                             FRCNN_NMS_ORDER_Y2, 0, st, nullptr);
      FRCNN_CUDA_TRY(cudaMemcpyAsync(c->pick1 + (size_t)i * c->cand_cap, c->nms.st.pick, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st));
      FRCNN_CUDA_TRY(cudaMemcpyAsync(c->count1 + i, c->nms.st.counts, sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
  }
  if (prof) cudaEventRecord(c->ev[2], st);
  if (stop == 3) return copy_back();
  // --- Detector.lua:91-98: ROI pooling of every candidate
  RoiParams rp;
  rp.f16 = c->act_f16;
  rp.fmap = c->pool_out.back(); rp.FH = c->feat_h; rp.FW = c->feat_w; rp.C = c->feat_c; rp.kh = c->roi_kh; rp.kw = c->roi_kw;
  rp.loc = c->roi_loc;
  rp.cand_r = c->cand_r; rp.pick = c->pick1; rp.pick_count = c->count1; rp.roi_base = c->roi_base; rp.cap = c->cand_cap;
  rp.out = c->roi_out; rp.roi_img = c->roi_img; rp.roi_cand = c->roi_cand; rp.status = c->flags + 1;
  rp.roi_base_out = c->roi_base; rp.roi_total = c->flags + 2; rp.total_cap = c->roi_cap; rp.roi_rect = c->roi_rect;
  launch_roi_pool_nhwc(rp, N, c->sm_count, st);
  c->launches += 2;
  if (prof) cudaEventRecord(c->ev[3], st);
  if (stop == 4) return copy_back();
  // --- Detector.lua:101: cnet
  run_cnet(c, c->roi_cap);
  if (prof) cudaEventRecord(c->ev[4], st);
  if (stop == 5) return copy_back();
  // --- Detector.lua:106-122
  FinalizeParams fp;
  fp.cand_r = c->cand_r; fp.cand_logp = c->cand_logp; fp.cand_anchor = c->cand_anchor; fp.cap = c->cand_cap;
  fp.roi_img = c->roi_img; fp.roi_cand = c->roi_cand; fp.roi_total = c->flags + 2; fp.reg = c->reg_out; fp.cls = c->cls_out;
  fp.ncls = c->class_count + 1; fp.class_prob = c->thr_class;
  fp.fin_r2 = c->fin_r2; fp.fin_box = c->fin_box; fp.fin_cls = c->fin_cls; fp.fin_conf = c->fin_conf;
  launch_finalize(fp, c->roi_cap, st);
  GroupParams gp;
  gp.roi_base = c->roi_base; gp.pick_count = c->count1; gp.fin_cls = c->fin_cls; gp.fin_box = c->fin_box;
  gp.cap = c->cand_cap; gp.n_classes = c->class_count; gp.gbox = c->gbox; gp.grow = c->grow; gp.n_pass = c->n_pass;
  gp.overflow = c->flags + 4;
  launch_group_by_class(gp, &c->nms, N, st);
  c->launches += 2;
  // --- Detector.lua:125-136: per-class nms(bb, 0.1, bb[{{},5}]) -- order key again y2
  const int n_seg = N * c->class_count;
  c->nms.fused = true;   // per-class segments are small: one fused launch
  c->launches += nms_run(&c->nms, reinterpret_cast<const float*>(c->gbox), 4, n_seg, N * c->cand_cap,
                         std::min(c->cand_cap, NMS_CTA_MAX_SEG), c->thr_nms2, FRCNN_NMS_ORDER_Y2, 0, st, nullptr);
  AssembleParams ap;
  ap.grow = c->grow; ap.cand_r = c->cand_r; ap.cand_logp = c->cand_logp; ap.cand_anchor = c->cand_anchor;
  ap.roi_img = c->roi_img; ap.roi_cand = c->roi_cand; ap.fin_r2 = c->fin_r2; ap.fin_cls = c->fin_cls; ap.fin_conf = c->fin_conf;
  ap.cap = c->cand_cap; ap.n_classes = c->class_count; ap.det = c->det_dev; ap.det_cap = c->det_cap; ap.n_det = c->flags + 3;
  launch_assemble(ap, &c->nms, n_seg, st);
  ++c->launches;
  if (prof) cudaEventRecord(c->ev[5], st);
  // --- results to the host in ONE batch: counters (flags | per-image match counts | per-image accepted counts are
  // one allocation) and, speculatively, the first spec_det winners
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_ints, c->flags, (16 + 2 * (size_t)c->det_n) * sizeof(int), cudaMemcpyDeviceToHost, st));
  const int spec = std::min(c->spec_det, c->det_cap);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_det, c->det_dev, (size_t)spec * sizeof(frcnn_detection), cudaMemcpyDeviceToHost, st));
}

// First half of a detection: the whole launch sequence (graph replay from the third use of a configuration) goes onto
// the context's stream; nothing waits.  do_detect_finish is the second half.
static void do_detect_enqueue(frcnn_ctx* c, const float* img_dev, int N, int H, int W) {
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called before detect");
  FRCNN_REQUIRE(!c->det_pending, FRCNN_E_STATE, "a detection is already in flight on this context (frcnn_detect_end it first)");
  ensure_pnet_workspace(c, N, H, W);
  ensure_det_workspace(c, N, 0);
  cudaStream_t st = c->stream;
  const bool prof = c->profiling;
  if (prof) for (int i = 0; i < 7; ++i) if (!c->ev[i]) FRCNN_CUDA_TRY(cudaEventCreate(&c->ev[i]));
  // The pipeline is a fixed launch sequence without host decisions: the second call with the same image pointer,
  // shape and thresholds captures it into a CUDA graph, later calls replay the graph (one host call per step).
  const frcnn_ctx::GraphKey key = {img_dev, N, H, W, c->thr_fg, c->thr_class, c->thr_nms1, c->thr_nms2, c->ws_gen};
  auto same = [](const frcnn_ctx::GraphKey& a, const frcnn_ctx::GraphKey& b) {
    return a.img == b.img && a.N == b.N && a.H == b.H && a.W == b.W && a.thr_fg == b.thr_fg && a.thr_class == b.thr_class &&
           a.thr_nms1 == b.thr_nms1 && a.thr_nms2 == b.thr_nms2 && a.gen == b.gen;
  };
  const bool want_graph = c->graph_enabled && !prof && c->decode_nblocks != 0 && c->cand_cap <= NMS_CTA_MAX_SEG;
  if (want_graph && c->graph_exec && same(key, c->graph_key)) {
    FRCNN_CUDA_TRY(cudaGraphLaunch(c->graph_exec, st));
    c->launches += c->launches_per_detect;
  } else if (want_graph && same(key, c->eager_key)) {
    if (c->graph_exec) {
      cudaGraphExecDestroy(c->graph_exec);
      c->graph_exec = nullptr;
    }
    const int64_t l0 = c->launches;
    cudaGraph_t graph = nullptr;
    FRCNN_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_detect(c, img_dev, N, H, W);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      c->graph_enabled = false;
      throw;
    }
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (e == cudaSuccess) e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {  // capture not possible in this environment: stay eager
      cudaGetLastError();
      c->graph_exec = nullptr;
      c->graph_enabled = false;
      c->launches = l0;
      enqueue_detect(c, img_dev, N, H, W);
    } else {
      c->launches_per_detect = c->launches - l0;
      c->graph_key = key;
      FRCNN_CUDA_TRY(cudaGraphLaunch(c->graph_exec, st));
    }
  } else {
    enqueue_detect(c, img_dev, N, H, W);
    c->eager_key = {img_dev, N, H, W, c->thr_fg, c->thr_class, c->thr_nms1, c->thr_nms2, c->ws_gen};
  }
  c->det_pending = true;
  c->det_pending_n = N; c->det_pending_h = H; c->det_pending_w = W;
  c->det_pending_img = img_dev;
}

static void do_detect_finish(frcnn_ctx* c, frcnn_detection* det_host, int cap, int* n_det) {
  FRCNN_REQUIRE(det_host != nullptr && n_det != nullptr && cap >= 0, FRCNN_E_INVALID, "bad output buffer");
  FRCNN_REQUIRE(c->det_pending, FRCNN_E_STATE, "no detection in flight on this context");
  cudaStream_t st = c->stream;
  const bool prof = c->profiling;
  const int N = c->det_pending_n;
  c->det_pending = false;
  FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
  // The reference's match list has no bound (Detector.lua:59): when a frame produced more matches than the candidate
  // buffers hold, grow them (doubling, up to the number of anchors) and run the frames again.
  while (c->h_ints[0]) {
    long total_anchors = 0;
    for (auto& hd : c->heads) total_anchors += 3L * hd.hh * hd.hw;
    FRCNN_REQUIRE(c->cand_cap < total_anchors, FRCNN_E_OVERFLOW, "candidate overflow with a capacity of every anchor");
    c->cand_cap = (int)std::min<long>(2L * c->cand_cap, (total_anchors + 255) / 256 * 256);
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->flags, 0, 2 * sizeof(int), st));
    if (c->graph_exec) {
      cudaGraphExecDestroy(c->graph_exec);
      c->graph_exec = nullptr;
    }
    c->eager_key = {};
    c->det_n = 0;  // forces ensure_det_workspace to rebuild for the new capacity
    do_detect_enqueue(c, c->det_pending_img, N, c->det_pending_h, c->det_pending_w);
    c->det_pending = false;
    FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
  }
  const int overflow = c->h_ints[0], degenerate = c->h_ints[1], roi_total = c->h_ints[2], ndet = c->h_ints[3];
  if (overflow || degenerate) FRCNN_CUDA_TRY(cudaMemsetAsync(c->flags, 0, 2 * sizeof(int), st));
  c->stats[0] = c->stats[2] = 0;
  for (int i = 0; i < N; ++i) {
    c->stats[0] += c->h_ints[16 + i];
    c->stats[2] += c->h_ints[16 + c->det_n + i];
  }
  c->stats[1] = roi_total;
  c->stats[3] = ndet;
  if (prof) {
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&c->timings[i], c->ev[i], c->ev[i + 1]);
    cudaEventElapsedTime(&c->timings[5], c->ev[0], c->ev[5]);
    conv_profile_end(c);
    c->prof_rows = roi_total;
  }
  FRCNN_REQUIRE(!overflow, FRCNN_E_OVERFLOW, "more RPN matches than the candidate capacity (" + std::to_string(c->cand_cap) + " per image)");
  if (c->h_ints[4]) {
    FRCNN_CUDA_TRY(cudaMemsetAsync(c->flags + 4, 0, sizeof(int), st));
    FRCNN_REQUIRE(false, FRCNN_E_OVERFLOW, "more than 8192 candidates of one image survived nms(bb, 0.25) (Detector.lua:82)");
  }
  FRCNN_REQUIRE(!degenerate, FRCNN_E_ROI_EMPTY,
                "an ROI clipped to max == 0; the reference raises an index error here (objective.lua:11)");
  const int ncopy = std::min(ndet, std::min(cap, c->det_cap));
  const int spec = std::min(c->spec_det, c->det_cap);
  if (ncopy > spec) {  // more winners than the speculative copy brought over
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_det + spec, c->det_dev + spec, (size_t)(ncopy - spec) * sizeof(frcnn_detection),
                                   cudaMemcpyDeviceToHost, st));
    FRCNN_CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (ncopy > 0) memcpy(det_host, c->h_det, (size_t)ncopy * sizeof(frcnn_detection));
  *n_det = ncopy;
  FRCNN_REQUIRE(ndet <= cap, FRCNN_E_OVERFLOW, "more winners than the output capacity");
}

static void do_detect(frcnn_ctx* c, const float* img_dev, int N, int H, int W, frcnn_detection* det_host, int cap, int* n_det) {
  FRCNN_REQUIRE(det_host != nullptr && n_det != nullptr && cap >= 0, FRCNN_E_INVALID, "bad output buffer");
  do_detect_enqueue(c, img_dev, N, H, W);
  do_detect_finish(c, det_host, cap, n_det);
}

// Input:cuda() (Detector.lua:32) into the context's staging frame buffer, asynchronously on the context's stream.
// A page-locked caller buffer is DMA'd directly; pageable memory is staged through the context's pinned buffer so
// that the copy is a true asynchronous DMA either way; a device buffer is copied device to device (the staging
// buffer keeps the captured graph's input pointer constant whatever frame the caller passes).
static const float* stage_frames(frcnn_ctx* c, const float* img, bool on_device, int n, int h, int w) {
  const size_t bytes = (size_t)n * 3 * h * w * sizeof(float);
  if (bytes > c->d_img_bytes) {
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->d_img) cudaFree(c->d_img);
    if (c->h_img) cudaFreeHost(c->h_img);
    c->d_img = nullptr; c->h_img = nullptr; c->d_img_bytes = 0;
    FRCNN_CUDA_TRY(cudaMalloc(&c->d_img, bytes));
    FRCNN_CUDA_TRY(cudaMallocHost(&c->h_img, bytes));
    c->d_img_bytes = bytes;
  }
  if (on_device) {
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_img, img, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return c->d_img;
  }
  cudaPointerAttributes attr;
  const bool pinned = cudaPointerGetAttributes(&attr, img) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (pinned) {
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_img, img, bytes, cudaMemcpyHostToDevice, c->stream));
  } else {
    memcpy(c->h_img, img, bytes);
    FRCNN_CUDA_TRY(cudaMemcpyAsync(c->d_img, c->h_img, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  return c->d_img;
}

// ---------------------------------------------------------------------------------------------- data-parallel training
// The one collective of the path (objective.lua:189,200 sum the per-image gradients and divide once): an
// all-reduce (sum) of the flat gradient over the ranks.  NCCL is loaded at run time (dlopen: the library itself links
// nothing but the CUDA runtime); one communicator rank per context.  The gradient is reduced IN PLACE in buckets, in the
// order pnet:backward finishes them -- cnet, anchor networks, conv block 4 ... 1 -- on a side stream, so that all but the
// last bucket travel over NVLink while the remaining weight gradients are still being computed.
typedef struct { char internal[128]; } NcclUniqueId;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {getenv("FRCNN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !n[0]) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  FRCNN_REQUIRE(api.handle != nullptr, FRCNN_E_NCCL, std::string("cannot load NCCL (libnccl.so.2; set FRCNN_NCCL_LIB): ") + (dlerror() ? dlerror() : ""));
  auto sym = [&](const char* name) {
    void* p = dlsym(api.handle, name);
    FRCNN_REQUIRE(p != nullptr, FRCNN_E_NCCL, std::string("NCCL symbol missing: ") + name);
    return p;
  };
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclUniqueId, int)>(sym("ncclCommInitRank"));
  api.CommInitAll = reinterpret_cast<int (*)(void**, int, const int*)>(sym("ncclCommInitAll"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclAllReduce"));
  api.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  return api;
}
#define FRCNN_NCCL_TRY(expr)                                                                                            \
  do {                                                                                                                  \
    int _r = (expr);                                                                                                    \
    if (_r != 0) throw ::frcnn::Error{FRCNN_E_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r)};             \
  } while (0)
static constexpr int NCCL_FLOAT = 7, NCCL_SUM = 0;

// Buckets in completion order: 0 = cnet, 1 = anchor networks, 2.. = conv blocks, last block first.  [lo, hi] param indices.
static int dp_bucket_count(const frcnn_ctx* c) { return 2 + (int)c->blocks.size(); }
static void dp_bucket_range(const frcnn_ctx* c, int bucket, int* lo, int* hi) {
  if (bucket == 0) {
    *lo = c->fcs.front().p_w;
    *hi = c->p_cls_b;
  } else if (bucket == 1) {
    *lo = c->heads.front().conv.p_w;
    *hi = c->heads.back().p_b2;
  } else {
    const int b = (int)c->blocks.size() - 1 - (bucket - 2);
    size_t first = 0;
    for (int i = 0; i < b; ++i) first += c->blocks[i].conv_steps;
    *lo = c->trunk[first].p_w;
    *hi = c->trunk[first + c->blocks[b].conv_steps - 1].p_prelu;
  }
}
static void dp_setup_streams(frcnn_ctx* c) {
  if (c->dp_stream) return;
  FRCNN_CUDA_TRY(cudaStreamCreateWithFlags(&c->dp_stream, cudaStreamNonBlocking));
  FRCNN_CUDA_TRY(cudaEventCreateWithFlags(&c->dp_ready, cudaEventDisableTiming));
  FRCNN_CUDA_TRY(cudaEventCreateWithFlags(&c->dp_done, cudaEventDisableTiming));
  c->dp_bucket_sent.assign(dp_bucket_count(c), 0);
}
// Enqueues the all-reduce of one bucket on the side stream, ordered after everything enqueued so far on the context's
// stream.  Gradient views that are adjacent in memory (nn.Module.flatten lays them out back to back) go out as one call.
static void dp_send_bucket(frcnn_ctx* c, int bucket) {
  int lo, hi;
  dp_bucket_range(c, bucket, &lo, &hi);
  FRCNN_CUDA_TRY(cudaEventRecord(c->dp_ready, c->stream));
  FRCNN_CUDA_TRY(cudaStreamWaitEvent(c->dp_stream, c->dp_ready, 0));
  int i = lo;
  while (i <= hi) {
    float* base = c->grads[i];
    int64_t n = c->params[i].numel;
    int j = i + 1;
    while (j <= hi && c->grads[j] == base + n) {
      n += c->params[j].numel;
      ++j;
    }
    FRCNN_NCCL_TRY(nccl().AllReduce(base, base, (size_t)n, NCCL_FLOAT, NCCL_SUM, c->dp_comm, c->dp_stream));
    c->dp_bytes += n * 4;
    i = j;
  }
  c->dp_bucket_sent[bucket] = 1;
}
static void dp_bucket_ready(frcnn_ctx* c, int bucket) {
  if (!c->dp_comm || !c->dp_overlap || c->dp_nranks <= 1) return;
  if (bucket >= (int)c->dp_bucket_sent.size() || c->dp_bucket_sent[bucket]) return;
  dp_send_bucket(c, bucket);
}

}  // namespace frcnn

// =============================================================================================== C ABI
#define API_BEGIN(ctx)                                                     \
  if (!(ctx)) {                                                            \
    frcnn::set_global_error("null ctx");                                   \
    return FRCNN_E_INVALID;                                                \
  }                                                                        \
  try {                                                                    \
    if ((ctx)->device >= 0) {                                              \
      cudaError_t _sd = cudaSetDevice((ctx)->device);                      \
      if (_sd != cudaSuccess) throw frcnn::Error{FRCNN_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_sd)}; \
    }

#define API_END(ctx)                                  \
    return FRCNN_OK;                                  \
  } catch (const frcnn::Error& e) {                   \
    (ctx)->err = e.msg;                               \
    return e.code;                                    \
  } catch (const std::exception& e) {                 \
    (ctx)->err = e.what();                            \
    return FRCNN_E_INVALID;                           \
  } catch (...) {                                     \
    (ctx)->err = "unknown error";                     \
    return FRCNN_E_INVALID;                           \
  }

#pragma GCC visibility push(default)
extern "C" {

int frcnn_version(void) { return 100; }

int frcnn_create(frcnn_ctx** out, int device, void* stream) {
  if (!out) return FRCNN_E_INVALID;
  *out = nullptr;
  if (device == -1) {  // host-only context: model plan + Localizer / Anchors geometry, no compute entry point works
    frcnn_ctx* c = new frcnn_ctx();
    c->device = -1;
    *out = c;
    return FRCNN_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    frcnn::set_global_error(std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                            " (this library has no CPU fallback)");
    cudaGetLastError();
    return FRCNN_E_CUDA;
  }
  if (device < 0 || device >= count) {
    frcnn::set_global_error("device index out of range");
    return FRCNN_E_INVALID;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    frcnn::set_global_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return FRCNN_E_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    frcnn::set_global_error(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    return FRCNN_E_CUDA;
  }
  if (prop.major != 10) {
    frcnn::set_global_error("this library is built for sm_100a (B200) only; found compute capability " + std::to_string(prop.major) +
                            "." + std::to_string(prop.minor));
    return FRCNN_E_CUDA;
  }
  frcnn_ctx* c = new frcnn_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  if (c->stream == nullptr) {
    // The legacy default stream cannot be captured into a CUDA graph.  A BLOCKING stream (plain cudaStreamCreate)
    // synchronises implicitly with the legacy default stream in both directions, so the ordering with surrounding
    // cutorch work on stream 0 (main.lua uses no other stream) is exactly what enqueueing on stream 0 would give.
    e = cudaStreamCreate(&c->stream);
    if (e != cudaSuccess) {
      frcnn::set_global_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
      delete c;
      return FRCNN_E_CUDA;
    }
    c->own_stream = true;
  }
  if (const char* ng = getenv("FRCNN_NO_GRAPH")) c->graph_enabled = !(ng[0] == '1');  // profiling under ncu
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  *out = c;
  return FRCNN_OK;
}

int frcnn_destroy(frcnn_ctx* c) {
  if (!c) return FRCNN_E_INVALID;
  if (c->device < 0) {
    delete c;
    return FRCNN_OK;
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  frcnn::free_all(c->ws_allocs);
  frcnn::free_all(c->det_allocs);
  frcnn::free_all(c->tws_allocs);
  frcnn::free_all(c->cw_allocs);
  if (c->nms_mem) cudaFree(c->nms_mem);
  for (auto& cv : c->trunk) if (cv.w_packed) cudaFree(cv.w_packed);
  for (auto& h : c->heads) if (h.conv.w_packed) cudaFree(h.conv.w_packed);
  for (auto& f : c->fcs) if (f.w_packed) cudaFree(f.w_packed);
  if (c->d_w_lut) cudaFree(c->d_w_lut);
  if (c->d_h_lut) cudaFree(c->d_h_lut);
  if (c->d_cen) cudaFree(c->d_cen);
  if (c->scratch) cudaFree(c->scratch);
  if (c->nms_stage) cudaFree(c->nms_stage);
  if (c->d_img) cudaFree(c->d_img);
  if (c->h_ints) cudaFreeHost(c->h_ints);
  if (c->h_det) cudaFreeHost(c->h_det);
  if (c->h_img) cudaFreeHost(c->h_img);
  if (c->h_loss) cudaFreeHost(c->h_loss);
  if (c->loss_ev) cudaEventDestroy(c->loss_ev);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->conv_ev) cudaEventDestroy(e);
  if (c->dp_comm) {
    cudaStreamSynchronize(c->dp_stream);
    try {
      frcnn::nccl().CommDestroy(c->dp_comm);
    } catch (...) {
    }
  }
  if (c->dp_ready) cudaEventDestroy(c->dp_ready);
  if (c->dp_done) cudaEventDestroy(c->dp_done);
  if (c->dp_stream) cudaStreamDestroy(c->dp_stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return FRCNN_OK;
}

const char* frcnn_last_error(const frcnn_ctx* c) { return c ? c->err.c_str() : frcnn::g_last_error.c_str(); }

int frcnn_device_info(frcnn_ctx* c, int* sm_count, int* cc_major, int* cc_minor) {
  API_BEGIN(c)
  if (sm_count) *sm_count = c->sm_count;
  if (cc_major) *cc_major = c->cc_major;
  if (cc_minor) *cc_minor = c->cc_minor;
  API_END(c)
}

int64_t frcnn_launch_count(const frcnn_ctx* c) { return c ? c->launches : 0; }

int frcnn_model_plan(frcnn_ctx* c, const frcnn_block_desc* blocks, int n_blocks, const frcnn_head_desc* heads, int n_heads,
                     const frcnn_fc_desc* fcs, int n_fcs, int class_count, int roi_kh, int roi_kw, const double* scales,
                     int n_scales, float dropout_eval_scale) {
  API_BEGIN(c)
  FRCNN_REQUIRE(blocks && heads && fcs && scales, FRCNN_E_INVALID, "null model description");
  frcnn::do_plan(c, blocks, n_blocks, heads, n_heads, fcs, n_fcs, class_count, roi_kh, roi_kw, scales, n_scales, dropout_eval_scale);
  API_END(c)
}

int frcnn_param_count(const frcnn_ctx* c) { return c ? (int)c->params.size() : 0; }

int frcnn_param_info(const frcnn_ctx* c, int index, char* name, int name_cap, int64_t* numel) {
  if (!c || index < 0 || index >= (int)c->params.size()) return FRCNN_E_INVALID;
  if (name && name_cap > 0) {
    strncpy(name, c->params[index].name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (numel) *numel = c->params[index].numel;
  return FRCNN_OK;
}

int frcnn_bind_params(frcnn_ctx* c, const float* const* params_dev, int n) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->planned, FRCNN_E_STATE, "frcnn_model_plan must be called first");
  FRCNN_REQUIRE(params_dev && n == (int)c->params.size(), FRCNN_E_INVALID,
                "expected " + std::to_string(c->params.size()) + " parameter pointers");
  for (int i = 0; i < n; ++i) {
    FRCNN_REQUIRE(params_dev[i] != nullptr, FRCNN_E_INVALID, "null parameter pointer: " + c->params[i].name);
    c->bound[i] = params_dev[i];
  }
  ++c->ws_gen;
  c->packed = false;
  API_END(c)
}

int frcnn_pack_weights(frcnn_ctx* c) {
  API_BEGIN(c)
  frcnn::do_pack(c);
  API_END(c)
}

int frcnn_localizer_layers(const frcnn_ctx* cc, int which, int* layers6, int cap_layers, int* n_layers) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c) return FRCNN_E_INVALID;
  try {
    FRCNN_REQUIRE(c->planned && which >= 0 && which < (int)c->loc.size(), FRCNN_E_INVALID, "bad localizer index");
    const auto& L = c->loc[which];
    FRCNN_REQUIRE(n_layers != nullptr, FRCNN_E_INVALID, "null n_layers");
    *n_layers = (int)L.size();
    if (layers6) {
      FRCNN_REQUIRE(cap_layers >= (int)L.size(), FRCNN_E_INVALID, "layer buffer too small");
      for (size_t i = 0; i < L.size(); ++i)
        for (int e = 0; e < 6; ++e) layers6[i * 6 + e] = L[i][e];
    }
    return FRCNN_OK;
  } catch (const frcnn::Error& e) {
    c->err = e.msg;
    return e.code;
  }
}

int frcnn_input_to_feature_rect(const frcnn_ctx* cc, int which, const double rect[4], double out[4]) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c || !c->planned || which < 0 || which >= (int)c->loc.size() || !rect || !out) return FRCNN_E_INVALID;
  frcnn::input_to_feature(c->loc[which], rect, out);
  return FRCNN_OK;
}

int frcnn_feature_to_input_rect(const frcnn_ctx* cc, int which, const double rect[4], double out[4]) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c || !c->planned || which < 0 || which >= (int)c->loc.size() || !rect || !out) return FRCNN_E_INVALID;
  frcnn::feature_to_input(c->loc[which], rect, out);
  return FRCNN_OK;
}

// the same with the optional layer_index argument of the Lua methods (Localizer.lua:41-42,69-70): only the first
// `layer_index` layers take part (0 or #layers = all of them)
int frcnn_input_to_feature_rect_upto(const frcnn_ctx* cc, int which, int layer_index, const double rect[4], double out[4]) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c || !c->planned || which < 0 || which >= (int)c->loc.size() || !rect || !out) return FRCNN_E_INVALID;
  const auto& L = c->loc[which];
  if (layer_index < 0 || layer_index > (int)L.size()) return FRCNN_E_INVALID;
  if (layer_index == 0 || layer_index == (int)L.size()) {
    frcnn::input_to_feature(L, rect, out);
  } else {
    std::vector<std::array<int, 6>> part(L.begin(), L.begin() + layer_index);
    frcnn::input_to_feature(part, rect, out);
  }
  return FRCNN_OK;
}

int frcnn_feature_to_input_rect_upto(const frcnn_ctx* cc, int which, int layer_index, const double rect[4], double out[4]) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c || !c->planned || which < 0 || which >= (int)c->loc.size() || !rect || !out) return FRCNN_E_INVALID;
  const auto& L = c->loc[which];
  if (layer_index < 0 || layer_index > (int)L.size()) return FRCNN_E_INVALID;
  if (layer_index == 0 || layer_index == (int)L.size()) {
    frcnn::feature_to_input(L, rect, out);
  } else {
    std::vector<std::array<int, 6>> part(L.begin(), L.begin() + layer_index);
    frcnn::feature_to_input(part, rect, out);
  }
  return FRCNN_OK;
}

int frcnn_anchors_build(frcnn_ctx* c, float* w_lut_host, float* h_lut_host) {
  if (!c || !c->planned || !w_lut_host || !h_lut_host) return FRCNN_E_INVALID;
  memcpy(w_lut_host, c->w_lut.data(), c->w_lut.size() * sizeof(float));
  memcpy(h_lut_host, c->h_lut.data(), c->h_lut.size() * sizeof(float));
  return FRCNN_OK;
}

int frcnn_pnet_output_dims(const frcnn_ctx* cc, int h, int w, int* dims3) {
  frcnn_ctx* c = const_cast<frcnn_ctx*>(cc);
  if (!c || !c->planned || !dims3) return FRCNN_E_INVALID;
  std::vector<int> ph, pw;
  int ch = h, cw = w;
  for (auto& b : c->blocks) {
    for (int s = 0; s < b.conv_steps; ++s) {
      ch = ch + 2 * b.padH - b.kH + 1;
      cw = cw + 2 * b.padW - b.kW + 1;
    }
    ch = (ch + 1) / 2;
    cw = (cw + 1) / 2;
    ph.push_back(ch);
    pw.push_back(cw);
  }
  for (size_t i = 0; i < c->heads.size(); ++i) {
    dims3[i * 3 + 0] = 18;
    dims3[i * 3 + 1] = ph[c->heads[i].input - 1] - c->heads[i].kW + 1;
    dims3[i * 3 + 2] = pw[c->heads[i].input - 1] - c->heads[i].kW + 1;
  }
  dims3[c->heads.size() * 3 + 0] = c->feat_c;
  dims3[c->heads.size() * 3 + 1] = ch;
  dims3[c->heads.size() * 3 + 2] = cw;
  return FRCNN_OK;
}

int frcnn_pnet_forward(frcnn_ctx* c, const float* img_dev, int n, int h, int w, float* const* out_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_dev != nullptr, FRCNN_E_INVALID, "null image");
  frcnn::conv_profile_begin(c);
  frcnn::do_pnet_forward(c, img_dev, n, h, w);
  if (out_dev) {
    for (size_t i = 0; i < c->heads.size(); ++i)
      if (out_dev[i])
        FRCNN_CUDA_TRY(cudaMemcpyAsync(out_dev[i], c->heads[i].out, (size_t)n * 18 * c->heads[i].hh * c->heads[i].hw * sizeof(float),
                                       cudaMemcpyDeviceToDevice, c->stream));
    if (out_dev[c->heads.size()]) {
      frcnn::launch_nhwc_bf16_to_chw_f32(c->pool_out.back(), out_dev[c->heads.size()], n, c->feat_h, c->feat_w, c->feat_c, c->stream,
                                         c->act_f16);
      ++c->launches;
    }
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  API_END(c)
}

int frcnn_bind_grads(frcnn_ctx* c, float* const* grads_dev, int n) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->planned, FRCNN_E_STATE, "frcnn_model_plan must be called first");
  FRCNN_REQUIRE(grads_dev && n == (int)c->params.size(), FRCNN_E_INVALID,
                "expected " + std::to_string(c->params.size()) + " gradient pointers");
  for (int i = 0; i < n; ++i) {
    FRCNN_REQUIRE(grads_dev[i] != nullptr, FRCNN_E_INVALID, "null gradient pointer: " + c->params[i].name);
    c->grads[i] = grads_dev[i];
  }
  API_END(c)
}

int frcnn_dropout_layers(const frcnn_ctx* c, int* channels, int cap, int* n_layers) {
  if (!c || !c->planned || !n_layers) return FRCNN_E_INVALID;
  int n = 0;
  for (const auto& cv : c->trunk)
    if (cv.dropout > 0.f) {
      if (channels && n < cap) channels[n] = cv.cout;
      ++n;
    }
  *n_layers = n;
  return FRCNN_OK;
}

int frcnn_pnet_forward_train(frcnn_ctx* c, const float* img_dev, int n, int h, int w, float* const* out_dev,
                             const float* const* masks_dev, uint64_t seed) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_dev != nullptr, FRCNN_E_INVALID, "null image");
  frcnn::ensure_pnet_workspace(c, n, h, w);
  frcnn::ensure_train_workspace(c, n, h, w);
  // SpatialDropout masks: injected by the caller (parity tests) or drawn from the seed
  int mi = 0;
  for (auto& cv : c->trunk) {
    if (cv.dropout <= 0.f) continue;
    if (masks_dev && masks_dev[mi]) {
      FRCNN_CUDA_TRY(cudaMemcpyAsync(cv.mask, masks_dev[mi], (size_t)n * cv.cout * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    } else {
      frcnn::launch_dropout_mask(cv.mask, n * cv.cout, cv.dropout, seed, (uint32_t)mi, c->stream);
      ++c->launches;
    }
    ++mi;
  }
  frcnn::conv_profile_begin(c);
  frcnn::do_pnet_forward(c, img_dev, n, h, w, true);
  if (out_dev) {
    for (size_t i = 0; i < c->heads.size(); ++i)
      if (out_dev[i])
        FRCNN_CUDA_TRY(cudaMemcpyAsync(out_dev[i], c->heads[i].out, (size_t)n * 18 * c->heads[i].hh * c->heads[i].hw * sizeof(float),
                                       cudaMemcpyDeviceToDevice, c->stream));
    if (out_dev[c->heads.size()]) {
      frcnn::launch_nhwc_bf16_to_chw_f32(c->pool_out.back(), out_dev[c->heads.size()], n, c->feat_h, c->feat_w, c->feat_c, c->stream,
                                         c->act_f16);
      ++c->launches;
    }
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  API_END(c)
}

int frcnn_pnet_backward(frcnn_ctx* c, const float* const* d_out_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(d_out_dev != nullptr, FRCNN_E_INVALID, "null delta_outputs");
  frcnn::do_pnet_backward(c, d_out_dev);
  API_END(c)
}

int frcnn_train_image(frcnn_ctx* c, const float* img_dev, int h, int w, const frcnn_example* pos_host, int n_pos,
                      const frcnn_example* neg_host, int n_neg, const float* const* pnet_masks_dev, const float* const* cnet_masks_dev,
                      uint64_t seed, float losses_host[4]) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_dev && losses_host && n_pos >= 0 && n_neg >= 0 && (n_pos == 0 || pos_host) && (n_neg == 0 || neg_host),
                FRCNN_E_INVALID, "bad argument");
  frcnn::do_train_image(c, img_dev, h, w, pos_host, n_pos, neg_host, n_neg, pnet_masks_dev, cnet_masks_dev, seed, losses_host);
  API_END(c)
}

// reference order (c*bins + b) fp32 <-> [row][bin][C]
__global__ void unpack_roi_rows_kernel(const float* __restrict__ x, float* __restrict__ out, long R, int C, int bins) {
  long total = R * C * bins;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / ((long)C * bins);
    int k = i - r * (long)C * bins;
    int b = k / C, cc = k - b * C;
    out[r * (long)C * bins + (long)cc * bins + b] = x[i];
  }
}

int frcnn_synchronize(frcnn_ctx* c) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int frcnn_train_batch(frcnn_ctx* c, const float* img_dev, int n, int h, int w, const frcnn_example* const* pos_host, const int* n_pos,
                      const frcnn_example* const* neg_host, const int* n_neg, const float* const* pnet_masks_dev, const uint64_t* seeds,
                      float* losses_host) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_dev && losses_host && pos_host && neg_host && n_pos && n_neg && seeds && n >= 1, FRCNN_E_INVALID, "bad argument");
  for (int i = 0; i < n; ++i)
    FRCNN_REQUIRE(n_pos[i] >= 0 && n_neg[i] >= 0 && (n_pos[i] == 0 || pos_host[i]) && (n_neg[i] == 0 || neg_host[i]), FRCNN_E_INVALID,
                  "bad example list");
  frcnn::do_train_batch(c, img_dev, n, h, w, pos_host, n_pos, neg_host, n_neg, pnet_masks_dev, nullptr, seeds, losses_host);
  API_END(c)
}

int frcnn_cnet_train_step(frcnn_ctx* c, const float* x_dev, int R, int n_pos, const float* crtarget_dev, const int32_t* cctarget_dev,
                          const float* const* masks_dev, uint64_t seed, float* dx_dev, float losses_host[2]) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called first");
  for (auto g : c->grads) FRCNN_REQUIRE(g != nullptr, FRCNN_E_STATE, "frcnn_bind_grads must be called first");
  FRCNN_REQUIRE(x_dev && crtarget_dev && cctarget_dev && losses_host && R >= 1 && n_pos >= 0 && n_pos <= R, FRCNN_E_INVALID, "bad argument");
  frcnn::ensure_objective_workspace(c, R);
  const int bins = c->roi_kh * c->roi_kw;
  const long total = (long)R * bins * c->feat_c;
  const int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
  frcnn::pack_roi_rows_kernel<<<blocks, 256, 0, c->stream>>>(x_dev, c->t_rows, R, c->feat_c, bins, 0);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->crtarget, crtarget_dev, (size_t)R * 4 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->cctarget, cctarget_dev, (size_t)R * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemsetAsync(c->losses_dev, 0, 8 * sizeof(float), c->stream));
  c->losses_cur = c->losses_dev;
  frcnn::run_cnet_train(c, R, n_pos, masks_dev, seed);
  if (dx_dev) unpack_roi_rows_kernel<<<blocks, 256, 0, c->stream>>>(c->t_dx, dx_dev, R, c->feat_c, bins);
  c->launches += 2;
  float lh[8];
  FRCNN_CUDA_TRY(cudaMemcpyAsync(lh, c->losses_dev, 8 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  losses_host[0] = lh[2];
  losses_host[1] = lh[3];
  API_END(c)
}

int frcnn_cnet_forward_train(frcnn_ctx* c, const float* x_dev, int R, const float* const* masks_dev, uint64_t seed, float* reg_dev,
                             float* cls_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called first");
  FRCNN_REQUIRE(x_dev && reg_dev && cls_dev && R >= 1, FRCNN_E_INVALID, "bad argument");
  frcnn::ensure_objective_workspace(c, R);
  const int bins = c->roi_kh * c->roi_kw;
  const long total = (long)R * bins * c->feat_c;
  const int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
  frcnn::pack_roi_rows_kernel<<<blocks, 256, 0, c->stream>>>(x_dev, c->t_rows, R, c->feat_c, bins, 0);
  ++c->launches;
  const int off = 0;
  frcnn::cnet_train_forward(c, frcnn::make_frames(1, &off, &R, nullptr), masks_dev, &seed);
  // the two output branches in fp32 on the stored hidden activations (model_utilities.lua:96-105)
  const frcnn::FcLayer& last = c->fcs.back();
  frcnn::launch_cnet_out(last.t_out32, frcnn::P(c, c->p_reg_w), frcnn::P(c, c->p_reg_b), frcnn::P(c, c->p_cls_w), frcnn::P(c, c->p_cls_b), reg_dev,
                         cls_dev, R, nullptr, last.nout, c->class_count + 1, c->stream);
  ++c->launches;
  FRCNN_CUDA_TRY(cudaGetLastError());
  c->cnet_train_rows = R;
  API_END(c)
}

int frcnn_cnet_backward(frcnn_ctx* c, const float* d_reg_dev, const float* d_cls_dev, float* dx_dev) {
  API_BEGIN(c)
  for (auto g : c->grads) FRCNN_REQUIRE(g != nullptr, FRCNN_E_STATE, "frcnn_bind_grads must be called first");
  FRCNN_REQUIRE(c->cnet_train_rows > 0, FRCNN_E_STATE, "cnet:backward needs a preceding frcnn_cnet_forward_train on this context");
  FRCNN_REQUIRE(d_reg_dev && d_cls_dev, FRCNN_E_INVALID, "null gradient");
  const int R = c->cnet_train_rows, off = 0, n_pos = R;
  const frcnn::FrameList fl = frcnn::make_frames(1, &off, &R, &n_pos);
  frcnn::cnet_train_loss_bwd(c, fl, d_reg_dev, d_cls_dev);
  frcnn::cnet_train_backward(c, fl);
  if (dx_dev) {
    const int bins = c->roi_kh * c->roi_kw;
    const long total = (long)R * bins * c->feat_c;
    const int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
    unpack_roi_rows_kernel<<<blocks, 256, 0, c->stream>>>(c->t_dx, dx_dev, R, c->feat_c, bins);
    ++c->launches;
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  c->cnet_train_rows = 0;
  API_END(c)
}

int frcnn_rpn_decode(frcnn_ctx* c, const float* const* heads_dev, int h, int w, double threshold, frcnn_candidate* cand_host, int cap,
                     int* n_cand) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called first (uploads the anchor LUTs)");
  FRCNN_REQUIRE(heads_dev && cand_host && n_cand, FRCNN_E_INVALID, "null argument");
  // head sizes follow from the image size
  std::vector<int> dims((c->heads.size() + 1) * 3);
  frcnn_pnet_output_dims(c, h, w, dims.data());
  for (size_t i = 0; i < c->heads.size(); ++i) {
    c->heads[i].hh = dims[i * 3 + 1];
    c->heads[i].hw = dims[i * 3 + 2];
    FRCNN_REQUIRE(c->heads[i].hh <= frcnn::LUT_EXTENT && c->heads[i].hw <= frcnn::LUT_EXTENT, FRCNN_E_INVALID,
                  "feature map exceeds the 200-cell anchor LUT (Anchors.lua:15)");
  }
  if (c->ws_h != h || c->ws_w != w) c->ws_n = 0;  // head sizes were overwritten: force a workspace rebuild later
  frcnn::ensure_det_workspace(c, 1, 0);
  frcnn::run_decode(c, heads_dev, 1, h, w, threshold);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_ints, c->flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(c->h_ints + 4, c->cand_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const int overflow = c->h_ints[0];
  const int n = c->h_ints[4];
  if (overflow) FRCNN_CUDA_TRY(cudaMemsetAsync(c->flags, 0, sizeof(int), c->stream));
  FRCNN_REQUIRE(!overflow, FRCNN_E_OVERFLOW, "more RPN matches than the candidate capacity");
  FRCNN_REQUIRE(n <= cap, FRCNN_E_OVERFLOW, "more RPN matches than the output capacity");
  std::vector<double> r((size_t)n * 4);
  std::vector<float> box((size_t)n * 4), lp(n);
  std::vector<int> an((size_t)n * 4);
  if (n > 0) {
    FRCNN_CUDA_TRY(cudaMemcpy(r.data(), c->cand_r, r.size() * sizeof(double), cudaMemcpyDeviceToHost));
    FRCNN_CUDA_TRY(cudaMemcpy(box.data(), c->cand_box, box.size() * sizeof(float), cudaMemcpyDeviceToHost));
    FRCNN_CUDA_TRY(cudaMemcpy(lp.data(), c->cand_logp, lp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    FRCNN_CUDA_TRY(cudaMemcpy(an.data(), c->cand_anchor, an.size() * sizeof(int), cudaMemcpyDeviceToHost));
  }
  for (int i = 0; i < n; ++i) {
    frcnn_candidate& o = cand_host[i];
    for (int e = 0; e < 4; ++e) { o.r[e] = r[i * 4 + e]; o.box[e] = box[i * 4 + e]; }
    o.logp = lp[i];
    o.layer = an[i * 4]; o.aspect = an[i * 4 + 1]; o.y = an[i * 4 + 2]; o.x = an[i * 4 + 3];
    o.pad_ = 0;
  }
  *n_cand = n;
  API_END(c)
}

static void nms_dev_impl(frcnn_ctx* c, const float* boxes_dev, int64_t n_total, int64_t row_stride, const int64_t* seg_offsets_host,
                         int n_seg, float overlap, int order_mode, int order_col, int64_t* pick_dev, int64_t* counts_dev) {
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(n_total >= 0 && n_total < (1ll << 31) && n_seg >= 1 && row_stride >= 4, FRCNN_E_INVALID, "bad NMS sizes");
  FRCNN_REQUIRE(order_mode >= 0 && order_mode <= 2 && (order_mode != FRCNN_NMS_ORDER_COLUMN || (order_col >= 0 && order_col < row_stride)),
                FRCNN_E_INVALID, "bad NMS order mode");
  // a private workspace sized for this call (kept until a larger one is needed)
  std::vector<int> beg(n_seg), len(n_seg);
  int max_len = 0;
  for (int s = 0; s < n_seg; ++s) {
    int64_t a = seg_offsets_host ? seg_offsets_host[s] : 0, b = seg_offsets_host ? seg_offsets_host[s + 1] : n_total;
    FRCNN_REQUIRE(a >= 0 && b >= a && b <= n_total, FRCNN_E_INVALID, "bad segment offsets");
    beg[s] = (int)a;
    len[s] = (int)(b - a);
    max_len = std::max(max_len, len[s]);
  }
  FRCNN_REQUIRE(!seg_offsets_host || (seg_offsets_host[0] == 0 && seg_offsets_host[n_seg] == n_total), FRCNN_E_INVALID,
                "segment offsets must partition [0, n_total)");
  if (n_total == 0 || max_len == 0) {
    FRCNN_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, n_seg * sizeof(int64_t), c->stream));
    return;
  }
  const int cap_total = (int)n_total, cap_seg = n_seg;
  size_t need = frcnn::nms_workspace_bytes(cap_total, cap_seg);
  void* mem = frcnn::ensure_scratch(c, need);
  frcnn::NmsWorkspace ws;
  frcnn::nms_workspace_init(&ws, mem, need, cap_total, cap_seg);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(ws.st.seg_beg, beg.data(), n_seg * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(ws.st.seg_len, len.data(), n_seg * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));  // beg/len are stack vectors
  if (!c->h_ints) FRCNN_CUDA_TRY(cudaMallocHost(&c->h_ints, 64 * sizeof(int)));
  int* h_rem = max_len > 8192 ? c->h_ints + 40 : nullptr;
  ws.fused = max_len <= 1024;
  c->launches += frcnn::nms_run(&ws, boxes_dev, (int)row_stride, n_seg, cap_total, max_len, overlap, order_mode, order_col, c->stream, h_rem);
  frcnn::nms_export(&ws, n_seg, max_len, pick_dev, counts_dev, c->stream);
  ++c->launches;
  FRCNN_CUDA_TRY(cudaGetLastError());
}

int frcnn_nms_segmented_dev(frcnn_ctx* c, const float* boxes_dev, int64_t n_total, int64_t row_stride, const int64_t* seg_offsets_host,
                            int n_seg, float overlap, int order_mode, int order_col, int64_t* pick_dev, int64_t* counts_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(boxes_dev && seg_offsets_host && pick_dev && counts_dev, FRCNN_E_INVALID, "null argument");
  nms_dev_impl(c, boxes_dev, n_total, row_stride, seg_offsets_host, n_seg, overlap, order_mode, order_col, pick_dev, counts_dev);
  API_END(c)
}

int frcnn_nms_dev(frcnn_ctx* c, const float* boxes_dev, int64_t n, int64_t row_stride, float overlap, int order_mode, int order_col,
                  int64_t* pick_dev, int64_t* n_pick_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(pick_dev && n_pick_dev && (boxes_dev || n == 0), FRCNN_E_INVALID, "null argument");
  nms_dev_impl(c, boxes_dev, n, row_stride, nullptr, 1, overlap, order_mode, order_col, pick_dev, n_pick_dev);
  API_END(c)
}

int frcnn_nms_segmented(frcnn_ctx* c, const float* boxes_host, int64_t row_stride, const int64_t* seg_offsets_host, int n_seg,
                        float overlap, int order_mode, int order_col, int64_t* pick_host, int64_t* counts_host) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(seg_offsets_host && pick_host && counts_host && n_seg >= 1, FRCNN_E_INVALID, "null argument");
  const int64_t n = seg_offsets_host[n_seg];
  FRCNN_REQUIRE(n >= 0 && (boxes_host || n == 0), FRCNN_E_INVALID, "null boxes");
  if (n == 0) {
    for (int s = 0; s < n_seg; ++s) counts_host[s] = 0;
    return FRCNN_OK;
  }
  // staging buffers: boxes | picks | counts -- one allocation kept by the context and grown on demand (a cudaMalloc /
  // cudaFree pair per call costs more than the whole NMS of a few thousand boxes)
  const size_t b_boxes = ((size_t)n * row_stride * sizeof(float) + 255) & ~size_t(255);
  const size_t b_pick = ((size_t)n * sizeof(int64_t) + 255) & ~size_t(255);
  const size_t b_counts = ((size_t)n_seg * sizeof(int64_t) + 255) & ~size_t(255);
  if (b_boxes + b_pick + b_counts > c->nms_stage_bytes) {
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->nms_stage) cudaFree(c->nms_stage);
    c->nms_stage = nullptr;
    c->nms_stage_bytes = 0;
    FRCNN_CUDA_TRY(cudaMalloc(&c->nms_stage, b_boxes + b_pick + b_counts));
    c->nms_stage_bytes = b_boxes + b_pick + b_counts;
  }
  float* d_boxes = (float*)c->nms_stage;
  int64_t* d_pick = (int64_t*)((uint8_t*)c->nms_stage + b_boxes);
  int64_t* d_counts = (int64_t*)((uint8_t*)c->nms_stage + b_boxes + b_pick);
  FRCNN_CUDA_TRY(cudaMemcpyAsync(d_boxes, boxes_host, (size_t)n * row_stride * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  nms_dev_impl(c, d_boxes, n, row_stride, seg_offsets_host, n_seg, overlap, order_mode, order_col, d_pick, d_counts);
  // results: the counts first, then only the picked prefix of every segment (the picks are a small fraction of the
  // boxes: copying all n int64 slots back cost more than the NMS itself)
  FRCNN_CUDA_TRY(cudaMemcpyAsync(counts_host, d_counts, n_seg * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  if ((size_t)n * sizeof(int64_t) <= (256u << 10)) {   // small problem: one copy and one synchronisation beat per-segment copies
    FRCNN_CUDA_TRY(cudaMemcpyAsync(pick_host, d_pick, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRCNN_OK;
  }
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int sgm = 0; sgm < n_seg; ++sgm) {
    const int64_t a = seg_offsets_host[sgm], cnt = counts_host[sgm];
    if (cnt > 0) FRCNN_CUDA_TRY(cudaMemcpyAsync(pick_host + a, d_pick + a, (size_t)cnt * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  }
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int frcnn_nms(frcnn_ctx* c, const float* boxes_host, int64_t n, int64_t row_stride, float overlap, int order_mode, int order_col,
              int64_t* pick_host, int64_t* n_pick) {
  if (!n_pick) return FRCNN_E_INVALID;
  if (n == 0) {  // nms.lua:26-28
    *n_pick = 0;
    return FRCNN_OK;
  }
  int64_t seg[2] = {0, n};
  return frcnn_nms_segmented(c, boxes_host, row_stride, seg, 1, overlap, order_mode, order_col, pick_host, n_pick);
}

int frcnn_roi_pool_forward(frcnn_ctx* c, const float* fmap_dev, int C, int H, int W, const double* rects_host, int R, float* out_dev,
                           int32_t* argmax_dev) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(c->planned, FRCNN_E_STATE, "frcnn_model_plan must be called first");
  FRCNN_REQUIRE(fmap_dev && rects_host && out_dev && R >= 0, FRCNN_E_INVALID, "null argument");
  if (R == 0) return FRCNN_OK;
  uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, (size_t)R * 4 * sizeof(double) + 256);
  int* status = (int*)mem;
  double* rects_dev = (double*)(mem + 256);
  FRCNN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(rects_dev, rects_host, (size_t)R * 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  frcnn::launch_roi_pool_chw(fmap_dev, C, H, W, c->roi_loc, rects_dev, R, c->roi_kh, c->roi_kw, out_dev, argmax_dev, status, c->stream);
  ++c->launches;
  int h_status = 0;
  FRCNN_CUDA_TRY(cudaMemcpyAsync(&h_status, status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  FRCNN_REQUIRE(h_status == 0, FRCNN_E_ROI_EMPTY, "an ROI clipped to max == 0; the reference raises an index error here (objective.lua:11)");
  API_END(c)
}

int frcnn_adaptive_maxpool_forward(frcnn_ctx* c, const float* x_dev, int C, int h, int w, int64_t stride_c, int64_t stride_h,
                                   int64_t stride_w, int kh, int kw, float* out_dev, float* idx_dev) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(x_dev && out_dev && C >= 1 && kh >= 1 && kw >= 1, FRCNN_E_INVALID, "bad argument");
  // an empty view is what extract_roi_pooling_input yields for a rect clipped to nothing; cunn raises on it
  FRCNN_REQUIRE(h >= 1 && w >= 1, FRCNN_E_ROI_EMPTY, "adaptive max pooling of an empty view (objective.lua:11)");
  frcnn::launch_adaptive_maxpool_fwd(x_dev, C, h, w, (long)stride_c, (long)stride_h, (long)stride_w, kh, kw, out_dev, idx_dev, c->stream);
  ++c->launches;
  FRCNN_CUDA_TRY(cudaGetLastError());
  API_END(c)
}

int frcnn_adaptive_maxpool_backward(frcnn_ctx* c, const float* dout_dev, const float* idx_dev, int C, int h, int w, int kh, int kw,
                                    float* dx_dev) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(dout_dev && idx_dev && dx_dev && C >= 1 && h >= 1 && w >= 1 && kh >= 1 && kw >= 1, FRCNN_E_INVALID, "bad argument");
  frcnn::launch_adaptive_maxpool_bwd(dout_dev, idx_dev, C, h, w, kh, kw, dx_dev, c->stream);
  ++c->launches;
  FRCNN_CUDA_TRY(cudaGetLastError());
  API_END(c)
}

int frcnn_cnet_forward(frcnn_ctx* c, const float* x_dev, int R, float* reg_dev, float* cls_dev) {
  API_BEGIN(c)
  FRCNN_REQUIRE(c->packed, FRCNN_E_STATE, "frcnn_pack_weights must be called first");
  FRCNN_REQUIRE(x_dev && reg_dev && cls_dev && R >= 0, FRCNN_E_INVALID, "null argument");
  if (R == 0) return FRCNN_OK;
  frcnn::conv_profile_begin(c);
  frcnn::ensure_det_workspace(c, 1, R);
  const int bins = c->roi_kh * c->roi_kw;
  long total = (long)R * bins * c->feat_c;
  int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
  frcnn::pack_roi_rows_kernel<<<blocks, 256, 0, c->stream>>>(x_dev, c->roi_out, R, c->feat_c, bins, c->eval_f16);
  frcnn::set_int_kernel<<<1, 1, 0, c->stream>>>(c->flags + 2, R);
  c->launches += 2;
  frcnn::run_cnet(c, R);
  const int ncls = c->class_count + 1;
  FRCNN_CUDA_TRY(cudaMemcpyAsync(reg_dev, c->reg_out, (size_t)R * 4 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  FRCNN_CUDA_TRY(cudaMemcpyAsync(cls_dev, c->cls_out, (size_t)R * ncls * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  API_END(c)
}

int frcnn_detect_dev(frcnn_ctx* c, const float* img_dev, int n, int h, int w, frcnn_detection* det_host, int cap, int* n_det) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_dev != nullptr, FRCNN_E_INVALID, "null image");
  frcnn::do_detect(c, img_dev, n, h, w, det_host, cap, n_det);
  API_END(c)
}

int frcnn_detect(frcnn_ctx* c, const float* img_host, int n, int h, int w, frcnn_detection* det_host, int cap, int* n_det) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img_host != nullptr && n >= 1 && h >= 1 && w >= 1, FRCNN_E_INVALID, "null image");
  const float* img_dev = frcnn::stage_frames(c, img_host, false, n, h, w);
  frcnn::do_detect(c, img_dev, n, h, w, det_host, cap, n_det);
  API_END(c)
}

int frcnn_detect_begin(frcnn_ctx* c, const float* img, int img_on_device, int n, int h, int w) {
  API_BEGIN(c)
  FRCNN_REQUIRE(img != nullptr && n >= 1 && h >= 1 && w >= 1, FRCNN_E_INVALID, "null image");
  FRCNN_REQUIRE(!c->det_pending, FRCNN_E_STATE, "a detection is already in flight on this context (frcnn_detect_end it first)");
  const float* img_dev = frcnn::stage_frames(c, img, img_on_device != 0, n, h, w);
  frcnn::do_detect_enqueue(c, img_dev, n, h, w);
  API_END(c)
}

int frcnn_detect_end(frcnn_ctx* c, frcnn_detection* det_host, int cap, int* n_det) {
  API_BEGIN(c)
  frcnn::do_detect_finish(c, det_host, cap, n_det);
  API_END(c)
}

int frcnn_detect_stats(const frcnn_ctx* c, int64_t stats[4]) {
  if (!c || !stats) return FRCNN_E_INVALID;
  for (int i = 0; i < 4; ++i) stats[i] = c->stats[i];
  return FRCNN_OK;
}

int frcnn_set_detect_thresholds(frcnn_ctx* c, double fg_prob, float nms_proposals, double class_prob, float nms_classes) {
  if (!c) return FRCNN_E_INVALID;
  c->thr_fg = fg_prob; c->thr_nms1 = nms_proposals; c->thr_class = class_prob; c->thr_nms2 = nms_classes;
  return FRCNN_OK;
}

int frcnn_block_output(frcnn_ctx* c, int block, float* out_dev, int* dims3) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(c->planned && c->ws_n > 0, FRCNN_E_STATE, "no pnet forward has run on this context");
  FRCNN_REQUIRE(block >= 1 && block <= (int)c->blocks.size() && dims3, FRCNN_E_INVALID, "bad block index");
  const int C = c->blocks[block - 1].filters, h = c->pool_h[block - 1], w = c->pool_w[block - 1];
  dims3[0] = C; dims3[1] = h; dims3[2] = w;
  if (out_dev) {
    // training forwards store bf16 activations; evaluate-mode forwards the context's operand format
    frcnn::launch_nhwc_bf16_to_chw_f32(c->pool_out[block - 1], out_dev, c->ws_n, h, w, C, c->stream, c->act_f16);
    ++c->launches;
    FRCNN_CUDA_TRY(cudaGetLastError());
  }
  API_END(c)
}

int frcnn_find_target_size(int orig_w, int orig_h, double target_smaller_side, double max_pixel_size, int* w_out, int* h_out) {
  if (orig_w < 1 || orig_h < 1 || !w_out || !h_out) return FRCNN_E_INVALID;
  double w, h;   // Lua numbers are doubles; math.floor(x + 0.5)
  if (orig_h < orig_w) {
    w = std::min((double)orig_w * target_smaller_side / (double)orig_h, max_pixel_size);
    h = floor((double)orig_h * w / (double)orig_w + 0.5);
    w = floor(w + 0.5);
  } else {
    h = std::min((double)orig_h * target_smaller_side / (double)orig_w, max_pixel_size);
    w = floor((double)orig_w * h / (double)orig_h + 0.5);
    h = floor(h + 0.5);
  }
  if (!(w >= 1 && h >= 1)) return FRCNN_E_INVALID;   // utilities.lua:201 asserts
  *w_out = (int)w;
  *h_out = (int)h;
  return FRCNN_OK;
}

int frcnn_scale_frame(frcnn_ctx* c, const float* src_dev, int ch, int src_h, int src_w, float* dst_dev, int dst_h, int dst_w) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(src_dev && dst_dev && ch >= 1 && src_h >= 1 && src_w >= 1 && dst_h >= 1 && dst_w >= 1, FRCNN_E_INVALID, "bad argument");
  float* tmp = (float*)frcnn::ensure_scratch(c, (size_t)ch * src_h * dst_w * sizeof(float) + 256);
  frcnn::launch_scale_image(src_dev, ch, src_h, src_w, tmp, dst_dev, dst_h, dst_w, c->stream);
  c->launches += 2;
  FRCNN_CUDA_TRY(cudaGetLastError());
  API_END(c)
}

int frcnn_normalize_frame(frcnn_ctx* c, float* img_dev, int h, int w, int rgb2yuv, int centering, int scaling, int contrastive_width) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(img_dev && h >= 1 && w >= 1 && contrastive_width >= 0, FRCNN_E_INVALID, "bad argument");
  FRCNN_REQUIRE(contrastive_width == 0 || (contrastive_width % 2 == 1 && contrastive_width <= 17), FRCNN_E_INVALID,
                "contrastive normalisation: odd kernel width <= 17");
  const size_t b_stat = 4096, b_k = 256;
  uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, b_stat + b_k + (size_t)h * w * sizeof(float) + 256);
  float kh[32];
  if (contrastive_width > 0) {
    // image.gaussian1D(size): sigma 0.25, amplitude 1, mean 0.5, not normalised: exp(-((i - center) / (sigma * size))^2 / 2)
    // with center = mean * size + 0.5 (1-based i); nn.SpatialSubtractiveNormalization then divides by kernel:sum()
    const int n = contrastive_width;
    double sum = 0.0;
    float g[32];
    for (int i = 1; i <= n; ++i) {
      const double center = 0.5 * n + 0.5;
      const double t = ((double)i - center) / (0.25 * n);
      g[i - 1] = (float)exp(-(t * t) / 2.0);
      sum += (double)g[i - 1];
    }
    for (int i = 0; i < n; ++i) kh[i] = g[i] / (float)sum;   // kernel:div(kernel:sum() * sqrt(nInputPlane)), nInputPlane = 1
    FRCNN_CUDA_TRY(cudaMemcpyAsync(mem + b_stat, kh, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  frcnn::launch_normalize_frame(img_dev, h, w, rgb2yuv, centering, scaling, (const float*)(mem + b_stat), contrastive_width, 1e-4f,
                                (double*)mem, (float*)(mem + b_stat + b_k), c->stream);
  c->launches += (rgb2yuv ? 1 : 0) + (centering ? 3 : 0) + (scaling ? 5 : 0) + (contrastive_width ? 2 : 0);
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));   // kh lives on this stack frame
  API_END(c)
}

int frcnn_find_positive(frcnn_ctx* c, const double* rois_host, int n_rois, const double* clip_host, double pos_threshold,
                        double neg_threshold, int include_best, frcnn_anchor_ref* out_host, int* out_roi_host, int cap, int* n_out) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(n_out != nullptr && cap >= 0 && n_rois >= 0 && (n_rois == 0 || rois_host), FRCNN_E_INVALID, "bad argument");
  *n_out = 0;
  if (n_rois > 0) {
    frcnn::ensure_luts_dev(c);
    const int n_scales = (int)c->w_lut.size() / (3 * frcnn::LUT_CELLS * 2);
    FRCNN_REQUIRE(n_scales * 3 <= frcnn::MAX_LABEL_IJ, FRCNN_E_INVALID, "find_positive: at most 4 scales");
    const int cap_roi = 16384;
    const size_t b_rois = ((size_t)n_rois * 4 * sizeof(double) + 255) & ~size_t(255);
    const size_t b_out = (size_t)n_rois * cap_roi * sizeof(frcnn_anchor_ref);
    const size_t b_cnt = (((size_t)n_rois + 1) * sizeof(int) + 255) & ~size_t(255);
    uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, b_rois + 2 * b_out + b_cnt);
    frcnn::FindPositiveParams p;
    p.w_lut = c->d_w_lut; p.h_lut = c->d_h_lut; p.n_scales = n_scales;
    p.rois = (const double*)mem;
    p.out = (frcnn_anchor_ref*)(mem + b_rois);
    p.best_scratch = (frcnn_anchor_ref*)(mem + b_rois + b_out);
    p.n_out = (int*)(mem + b_rois + 2 * b_out);
    p.status = p.n_out + n_rois;
    p.cap_per_roi = cap_roi;
    p.has_clip = clip_host != nullptr;
    for (int k = 0; k < 4; ++k) p.clip[k] = clip_host ? clip_host[k] : 0.0;
    p.pos_threshold = pos_threshold; p.neg_threshold = neg_threshold; p.include_best = include_best ? 1 : 0;
    FRCNN_CUDA_TRY(cudaMemcpyAsync(mem, rois_host, (size_t)n_rois * 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    FRCNN_CUDA_TRY(cudaMemsetAsync(p.status, 0, sizeof(int), c->stream));
    frcnn::launch_find_positive(p, n_rois, c->stream);
    ++c->launches;
    std::vector<int> counts(n_rois + 1);
    FRCNN_CUDA_TRY(cudaMemcpyAsync(counts.data(), p.n_out, ((size_t)n_rois + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    FRCNN_REQUIRE(counts[n_rois] == 0, FRCNN_E_OVERFLOW, "find_positive: more than 16384 matches for one ROI");
    long total = 0;
    for (int r = 0; r < n_rois; ++r) total += counts[r];
    FRCNN_REQUIRE(total <= cap, FRCNN_E_OVERFLOW, "find_positive: more matches than the output capacity");
    FRCNN_REQUIRE(total == 0 || (out_host && out_roi_host), FRCNN_E_INVALID, "null output");
    int at = 0;
    for (int r = 0; r < n_rois; ++r) {
      if (!counts[r]) continue;
      FRCNN_CUDA_TRY(cudaMemcpyAsync(out_host + at, p.out + (size_t)r * cap_roi, (size_t)counts[r] * sizeof(frcnn_anchor_ref),
                                     cudaMemcpyDeviceToHost, c->stream));
      for (int i = 0; i < counts[r]; ++i) out_roi_host[at + i] = r;
      at += counts[r];
    }
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_out = at;
  }
  API_END(c)
}

int frcnn_find_nearby_negative(frcnn_ctx* c, const frcnn_anchor_ref* pos_host, int n_pos, double neg_threshold,
                               frcnn_anchor_ref* out_host, int* out_pos_host, int cap, int* n_out) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(n_out && cap >= 0 && n_pos >= 0 && (n_pos == 0 || pos_host), FRCNN_E_INVALID, "bad argument");
  *n_out = 0;
  if (n_pos > 0) {
    frcnn::ensure_luts_dev(c);
    const int n_scales = (int)c->w_lut.size() / (3 * frcnn::LUT_CELLS * 2);
    for (int i = 0; i < n_pos; ++i)
      FRCNN_REQUIRE(pos_host[i].layer >= 1 && pos_host[i].layer <= n_scales && pos_host[i].aspect >= 1 && pos_host[i].aspect <= 3 &&
                        pos_host[i].y >= 1 && pos_host[i].y <= frcnn::LUT_CELLS && pos_host[i].x >= 1 && pos_host[i].x <= frcnn::LUT_CELLS,
                    FRCNN_E_INVALID, "find_nearby: anchor index out of the LUT range");
    const size_t b_pos = ((size_t)n_pos * sizeof(frcnn_anchor_ref) + 255) & ~size_t(255);
    const size_t b_out = ((size_t)cap * sizeof(frcnn_anchor_ref) + 255) & ~size_t(255);
    const size_t b_idx = ((size_t)cap * sizeof(int) + 255) & ~size_t(255);
    uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, b_pos + b_out + b_idx + 256);
    frcnn::FindNearbyParams p;
    p.w_lut = c->d_w_lut; p.h_lut = c->d_h_lut; p.n_scales = n_scales;
    p.cen_x = c->d_cen; p.cen_y = c->d_cen + c->cen_x.size();
    p.pos = (const frcnn_anchor_ref*)mem;
    p.n_pos = n_pos;
    p.neg_threshold = neg_threshold;
    p.out = (frcnn_anchor_ref*)(mem + b_pos);
    p.out_pos = (int*)(mem + b_pos + b_out);
    p.cap = cap;
    p.result = (int*)(mem + b_pos + b_out + b_idx);
    FRCNN_CUDA_TRY(cudaMemcpyAsync(mem, pos_host, (size_t)n_pos * sizeof(frcnn_anchor_ref), cudaMemcpyHostToDevice, c->stream));
    frcnn::launch_find_nearby(p, c->stream);
    ++c->launches;
    int total = 0;
    FRCNN_CUDA_TRY(cudaMemcpyAsync(&total, p.result, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    FRCNN_REQUIRE(total <= cap, FRCNN_E_OVERFLOW, "find_nearby: more entries than the output capacity");
    FRCNN_REQUIRE(total == 0 || (out_host && out_pos_host), FRCNN_E_INVALID, "null output");
    if (total > 0) {
      FRCNN_CUDA_TRY(cudaMemcpyAsync(out_host, p.out, (size_t)total * sizeof(frcnn_anchor_ref), cudaMemcpyDeviceToHost, c->stream));
      FRCNN_CUDA_TRY(cudaMemcpyAsync(out_pos_host, p.out_pos, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    *n_out = total;
  }
  API_END(c)
}

int frcnn_sample_negative(frcnn_ctx* c, const double image_rect[4], const double* rois_host, int n_rois, double neg_threshold, int count,
                          const uint32_t* rnd_host, int n_trials, int retry_in, frcnn_anchor_ref* out_host, int cap, int* n_out,
                          int* trials_consumed, int* finished, int* retry_out) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(image_rect && n_out && trials_consumed && finished && cap >= 0 && n_rois >= 0 && n_trials >= 0 &&
                    (n_rois == 0 || rois_host) && (n_trials == 0 || rnd_host),
                FRCNN_E_INVALID, "bad argument");
  frcnn::ensure_luts_dev(c);
  const int n_scales = (int)c->w_lut.size() / (3 * frcnn::LUT_CELLS * 2);
  FRCNN_REQUIRE(n_scales * 3 <= frcnn::MAX_LABEL_IJ, FRCNN_E_INVALID, "sample_negative: at most 4 scales");
  const size_t b_rois = ((size_t)std::max(n_rois, 1) * 4 * sizeof(double) + 255) & ~size_t(255);
  const size_t b_rnd = ((size_t)std::max(n_trials, 1) * 3 * sizeof(uint32_t) + 255) & ~size_t(255);
  const size_t b_out = ((size_t)std::max(cap, 1) * sizeof(frcnn_anchor_ref) + 255) & ~size_t(255);
  uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, b_rois + b_rnd + b_out + 256);
  frcnn::SampleNegativeParams p;
  p.w_lut = c->d_w_lut; p.h_lut = c->d_h_lut; p.n_scales = n_scales;
  for (int k = 0; k < 4; ++k) p.image_rect[k] = image_rect[k];
  p.rois = (const double*)mem; p.n_rois = n_rois; p.neg_threshold = neg_threshold; p.count = count;
  p.rnd = (const uint32_t*)(mem + b_rois); p.n_trials = n_trials; p.retry_in = retry_in < 0 ? 0 : retry_in;
  p.out = (frcnn_anchor_ref*)(mem + b_rois + b_rnd); p.cap = cap;
  p.result = (int*)(mem + b_rois + b_rnd + b_out);
  if (n_rois) FRCNN_CUDA_TRY(cudaMemcpyAsync(mem, rois_host, (size_t)n_rois * 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (n_trials) FRCNN_CUDA_TRY(cudaMemcpyAsync(mem + b_rois, rnd_host, (size_t)n_trials * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  frcnn::launch_sample_negative(p, c->stream);
  ++c->launches;
  int res[5];
  FRCNN_CUDA_TRY(cudaMemcpyAsync(res, p.result, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  FRCNN_REQUIRE(res[0] == 0 || out_host, FRCNN_E_INVALID, "null output");
  if (res[0]) FRCNN_CUDA_TRY(cudaMemcpy(out_host, p.out, (size_t)res[0] * sizeof(frcnn_anchor_ref), cudaMemcpyDeviceToHost));
  *n_out = res[0];
  *trials_consumed = res[1];
  *finished = res[2];
  if (retry_out) *retry_out = res[4];
  FRCNN_REQUIRE(count <= cap || res[0] < cap, FRCNN_E_OVERFLOW, "sample_negative: output capacity smaller than count");
  API_END(c)
}

int frcnn_rmsprop_step(frcnn_ctx* c, float* weights_dev, float* gradient_dev, float* state_m_dev, int64_t n, double grad_div, double lr,
                       double alpha, double epsilon, double weight_decay) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(weights_dev && gradient_dev && state_m_dev && n >= 0, FRCNN_E_INVALID, "null argument");
  FRCNN_REQUIRE(grad_div != 0.0, FRCNN_E_INVALID, "grad_div must be non-zero (pass 1 for no division)");
  if (n > 0) {
    frcnn::launch_rmsprop_step(weights_dev, gradient_dev, state_m_dev, (long)n, grad_div, lr, alpha, epsilon, weight_decay, c->sm_count,
                               c->stream);
    ++c->launches;
  }
  API_END(c)
}

int frcnn_set_schedule(frcnn_ctx* c, int schedule) {
  if (!c || (schedule != FRCNN_SCHED_LATENCY && schedule != FRCNN_SCHED_THROUGHPUT)) return FRCNN_E_INVALID;
  if (c->schedule != schedule) {
    c->schedule = schedule;
    for (auto& f : c->fcs)
      if (f.launch.p.m_limit) {
        f.launch.p.dyn_ctas = frcnn::cnet_ctas(c);
        const int total = f.launch.p.n_tiles_m * f.launch.p.n_tiles_n * f.launch.p.splits;
        f.launch.grid = std::min(std::min(total, c->sm_count), f.launch.p.dyn_ctas);
      }
    ++c->ws_gen;  // a captured detect graph holds the other schedule's launches: re-capture
  }
  return FRCNN_OK;
}

int frcnn_set_eval_precision(frcnn_ctx* c, int precision) {
  if (!c || (precision != FRCNN_PREC_BF16 && precision != FRCNN_PREC_FP16)) return FRCNN_E_INVALID;
  const int f16 = precision == FRCNN_PREC_FP16 ? 1 : 0;
  if (c->eval_f16 != f16) {
    c->eval_f16 = f16;
    ++c->ws_gen;  // a captured detect graph holds the other format's launches: re-capture
  }
  return FRCNN_OK;
}

int frcnn_set_graph_replay(frcnn_ctx* c, int enable) {
  if (!c) return FRCNN_E_INVALID;
  c->graph_enabled = enable != 0;
  return FRCNN_OK;
}

int frcnn_set_profiling(frcnn_ctx* c, int enable) {
  if (!c) return FRCNN_E_INVALID;
  c->profiling = enable != 0;
  return FRCNN_OK;
}

int frcnn_last_timings(const frcnn_ctx* c, float ms[6]) {
  if (!c || !ms) return FRCNN_E_INVALID;
  for (int i = 0; i < 6; ++i) ms[i] = c->timings[i];
  return FRCNN_OK;
}

int frcnn_last_conv_profile(const frcnn_ctx* c, float* ms, double* flops, int* launches) {
  if (!c) return FRCNN_E_INVALID;
  if (ms) *ms = c->conv_ms;
  if (flops) *flops = c->conv_flops;
  if (launches) *launches = c->conv_launches;
  return FRCNN_OK;
}

// ---- backward primitives of nn.SpatialConvolution (pnet:backward, objective.lua:189)
int frcnn_conv_dgrad_bf16(frcnn_ctx* c, const uint16_t* dy_dev, const float* w_dev, int n, int h, int w, int cin, int cout, int k,
                          int pad, uint16_t* dx_dev) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(dy_dev && w_dev && dx_dev, FRCNN_E_INVALID, "null argument");
  const int ho = h + 2 * pad - k + 1, wo = w + 2 * pad - k + 1;
  FRCNN_REQUIRE(ho > 0 && wo > 0 && k - 1 - pad >= 0, FRCNN_E_INVALID, "bad convolution geometry");
  const size_t wbytes = ((size_t)cout * cin * k * k * sizeof(frcnn::bf16) + 255) & ~size_t(255);
  frcnn::bf16* wp = (frcnn::bf16*)frcnn::ensure_scratch(c, wbytes + 256);
  frcnn::launch_pack_conv_weight_dgrad(w_dev, wp, cout, cin, k, k, c->stream);
  // transposed convolution = stride-1 convolution of dY with the flipped, channel-transposed filters, padding k-1-pad
  frcnn::ConvLaunch L;
  frcnn::conv_prepare(&L, (const frcnn::bf16*)dy_dev, wp, n, ho, wo, cout, cin, k, k, k - 1 - pad, k - 1 - pad, frcnn::EPI_STORE,
                      (frcnn::bf16*)dx_dev, c->sm_count, 0, 0, 0);
  frcnn::conv_launch(L, c->stream);
  c->launches += 2;
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int frcnn_conv_wgrad_bf16(frcnn_ctx* c, const uint16_t* x_dev, const uint16_t* dy_dev, int n, int h, int w, int cin, int cout, int k,
                          int pad, float* dw_dev) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(x_dev && dy_dev && dw_dev, FRCNN_E_INVALID, "null argument");
  const int ho = h + 2 * pad - k + 1, wo = w + 2 * pad - k + 1;
  FRCNN_REQUIRE(ho > 0 && wo > 0, FRCNN_E_INVALID, "bad convolution geometry");
  float* dw_taps = (float*)frcnn::ensure_scratch(c, (size_t)cout * k * k * cin * 4 + 256);
  FRCNN_CUDA_TRY(cudaMemsetAsync(dw_taps, 0, (size_t)cout * k * k * cin * 4, c->stream));
  frcnn::ConvLaunch L;
  frcnn::conv_wgrad_prepare(&L, (const frcnn::bf16*)dy_dev, (const frcnn::bf16*)x_dev, dw_taps, n, h, w, cin, cout, k, k, pad, pad,
                            c->sm_count);
  frcnn::conv_launch(L, c->stream);
  frcnn::launch_wgrad_finish(dw_taps, dw_dev, cout, cin, k, k, c->stream);
  c->launches += 2;
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  API_END(c)
}

int frcnn_conv_first_wgrad(frcnn_ctx* c, const uint16_t* dy_dev, const float* img_dev, int n, int h, int w, int pad,
                           float* dw_dev, int iters, float* elapsed_ms) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(dy_dev && img_dev && dw_dev, FRCNN_E_INVALID, "null argument");
  // the kernel indexes dy and the frame with the same (h, w): a 3 x 3 filter with padding 1, as both reference models have
  FRCNN_REQUIRE(n > 0 && h > 0 && w > 0 && pad == 1, FRCNN_E_INVALID, "first-layer wgrad: 3x3 filter, padding 1");
  if (iters < 1) iters = 1;
  cudaEvent_t e0, e1;
  FRCNN_CUDA_TRY(cudaEventCreate(&e0));
  FRCNN_CUDA_TRY(cudaEventCreate(&e1));
  FRCNN_CUDA_TRY(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) {
    frcnn::launch_first_wgrad((const frcnn::bf16*)dy_dev, img_dev, dw_dev, n, h, w, pad, c->sm_count, c->stream);
    ++c->launches;
  }
  FRCNN_CUDA_TRY(cudaEventRecord(e1, c->stream));
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  FRCNN_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  if (elapsed_ms) *elapsed_ms = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  API_END(c)
}

int frcnn_conv_first(frcnn_ctx* c, const float* img_dev, const float* w_dev, const float* bias_dev, const float* prelu_dev,
                     float scale, int n, int h, int w, int cout, int pad, int pool, uint16_t* out_dev, int iters, float* elapsed_ms) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(img_dev && w_dev && out_dev, FRCNN_E_INVALID, "null argument");
  uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, 2 * (size_t)cout * 32 * sizeof(frcnn::bf16) + 256);
  frcnn::bf16* wp = (frcnn::bf16*)mem;
  frcnn::launch_pack_first_conv_weight(w_dev, wp, cout, 3, 3, 3, c->stream);
  frcnn::ConvLaunch L;
  frcnn::conv_first_prepare(&L, wp, n, h, w, 3, cout, 3, 3, pad, pad, pool ? frcnn::EPI_POOL : frcnn::EPI_STORE, (frcnn::bf16*)out_dev,
                            c->sm_count);
  L.p.bias = bias_dev;
  L.p.prelu = prelu_dev;
  L.p.scale = scale;
  L.p.img = img_dev;
  if (iters < 1) iters = 1;
  cudaEvent_t e0, e1;
  FRCNN_CUDA_TRY(cudaEventCreate(&e0));
  FRCNN_CUDA_TRY(cudaEventCreate(&e1));
  FRCNN_CUDA_TRY(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) {
    frcnn::conv_launch(L, c->stream);
    ++c->launches;
  }
  FRCNN_CUDA_TRY(cudaEventRecord(e1, c->stream));
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (elapsed_ms) FRCNN_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  API_END(c)
}

int frcnn_conv_bf16(frcnn_ctx* c, const uint16_t* x_dev, const float* w_dev, const float* bias_dev, const float* prelu_dev, float scale,
                    int n, int h, int w, int cin, int cout, int k, int pad, int splits, int bn, int mt, int pool, uint16_t* out_dev,
                    int iters, float* elapsed_ms) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(x_dev && w_dev && out_dev, FRCNN_E_INVALID, "null argument");
  FRCNN_REQUIRE(bn == 0 || bn == 64 || bn == 128 || bn == 192 || bn == 256, FRCNN_E_INVALID, "bn must be 0, 64, 128, 192 or 256");
  FRCNN_REQUIRE(((mt >= 0 && mt <= 2) || (mt > 10 && mt < 50 && (mt % 10 == 1 || mt % 10 == 2)) || mt == 51 || mt == 61) && !(pool && splits > 1), FRCNN_E_INVALID,
                "bad mt / pool");
  const int ho = h + 2 * pad - k + 1, wo = w + 2 * pad - k + 1;
  FRCNN_REQUIRE(ho > 0 && wo > 0, FRCNN_E_INVALID, "input smaller than the kernel");
  const size_t wbytes = ((size_t)cout * cin * k * k * sizeof(frcnn::bf16) + 255) & ~size_t(255);
  const size_t abytes = (size_t)n * ho * wo * cout * sizeof(float);
  uint8_t* mem = (uint8_t*)frcnn::ensure_scratch(c, wbytes + abytes + 256);
  frcnn::bf16* wp = (frcnn::bf16*)mem;
  float* acc = (float*)(mem + wbytes);
  frcnn::launch_pack_conv_weight(w_dev, wp, cout, cin, k, k, c->stream);
  frcnn::ConvLaunch L;
  const bool split = splits > 1;
  frcnn::conv_prepare(&L, (const frcnn::bf16*)x_dev, wp, n, h, w, cin, cout, k, k, pad, pad,
                      split ? frcnn::EPI_F32_REDUCE : (pool ? frcnn::EPI_POOL : frcnn::EPI_STORE), (frcnn::bf16*)out_dev, c->sm_count,
                      split ? splits : 0, bn, mt);
  L.p.bias = split ? nullptr : bias_dev;
  L.p.prelu = split ? nullptr : prelu_dev;
  L.p.scale = scale;
  if (split) frcnn::conv_set_f32_output(&L, acc);
  if (iters < 1) iters = 1;
  cudaEvent_t e0, e1;
  FRCNN_CUDA_TRY(cudaEventCreate(&e0));
  FRCNN_CUDA_TRY(cudaEventCreate(&e1));
  if (split) FRCNN_CUDA_TRY(cudaMemsetAsync(acc, 0, abytes, c->stream));
  FRCNN_CUDA_TRY(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) {
    frcnn::conv_launch(L, c->stream);
    ++c->launches;
  }
  FRCNN_CUDA_TRY(cudaEventRecord(e1, c->stream));
  if (split) {
    if (iters > 1) {  // the timing loop accumulated `iters` times: redo once for the result
      FRCNN_CUDA_TRY(cudaMemsetAsync(acc, 0, abytes, c->stream));
      frcnn::conv_launch(L, c->stream);
    }
    long total = (long)n * ho * wo * cout;
    int blocks = (int)std::min<long>((total + 255) / 256, 148 * 16);
    frcnn::acc_tail_kernel<<<blocks, 256, 0, c->stream>>>(acc, bias_dev, prelu_dev, scale, (frcnn::bf16*)out_dev, total, cout);
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (elapsed_ms) FRCNN_CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (const char* path = getenv("FRCNN_CONV_TRACE")) {
    // measurement only: one more launch with per-CTA / per-unit globaltimer stamps (conv_halo_kernel), dumped to `path`
    const size_t words = (size_t)L.grid * 64;
    unsigned long long* d_trace = nullptr;
    FRCNN_CUDA_TRY(cudaMalloc(&d_trace, words * 8));
    FRCNN_CUDA_TRY(cudaMemset(d_trace, 0, words * 8));
    L.p.trace = d_trace;
    frcnn::conv_launch(L, c->stream);
    FRCNN_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<unsigned long long> t(words);
    FRCNN_CUDA_TRY(cudaMemcpy(t.data(), d_trace, words * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_trace);
    L.p.trace = nullptr;
    if (FILE* f = fopen(path, "wb")) {
      fwrite(t.data(), 8, t.size(), f);
      fclose(f);
    }
  }
  API_END(c)
}

// ---- data-parallel training ---------------------------------------------------------------------------------------
int frcnn_dp_unique_id(char id_out[128]) {
  try {
    FRCNN_REQUIRE(id_out != nullptr, FRCNN_E_INVALID, "null id buffer");
    frcnn::NcclUniqueId id;
    FRCNN_NCCL_TRY(frcnn::nccl().GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return FRCNN_OK;
  } catch (const frcnn::Error& e) {
    frcnn::set_global_error(e.msg);
    return e.code;
  }
}

int frcnn_dp_init_rank(frcnn_ctx* c, const char id[128], int rank, int nranks) {
  API_BEGIN(c)
  REQUIRE_DEVICE(c);
  FRCNN_REQUIRE(c->planned, FRCNN_E_STATE, "frcnn_model_plan must be called first");
  FRCNN_REQUIRE(id != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, FRCNN_E_INVALID, "bad rank / id");
  FRCNN_REQUIRE(c->dp_comm == nullptr, FRCNN_E_STATE, "the context already belongs to a communicator");
  frcnn::NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  FRCNN_NCCL_TRY(frcnn::nccl().CommInitRank(&c->dp_comm, nranks, uid, rank));
  c->dp_rank = rank;
  c->dp_nranks = nranks;
  frcnn::dp_setup_streams(c);
  API_END(c)
}

int frcnn_dp_init_all(frcnn_ctx* const* ctxs, int n) {
  if (!ctxs || n < 1 || n > 64) return FRCNN_E_INVALID;
  frcnn_ctx* c = ctxs[0];
  API_BEGIN(c)
  int devs[64];
  void* comms[64];
  for (int i = 0; i < n; ++i) {
    FRCNN_REQUIRE(ctxs[i] && ctxs[i]->device >= 0 && ctxs[i]->planned && !ctxs[i]->dp_comm, FRCNN_E_STATE,
                  "every context needs a device, a model plan and no communicator yet");
    devs[i] = ctxs[i]->device;
  }
  FRCNN_NCCL_TRY(frcnn::nccl().CommInitAll(comms, n, devs));
  for (int i = 0; i < n; ++i) {
    FRCNN_CUDA_TRY(cudaSetDevice(ctxs[i]->device));
    ctxs[i]->dp_comm = comms[i];
    ctxs[i]->dp_rank = i;
    ctxs[i]->dp_nranks = n;
    frcnn::dp_setup_streams(ctxs[i]);
  }
  FRCNN_CUDA_TRY(cudaSetDevice(c->device));
  API_END(c)
}

int frcnn_dp_set_overlap(frcnn_ctx* c, int enable) {
  if (!c) return FRCNN_E_INVALID;
  c->dp_overlap = enable != 0;
  return FRCNN_OK;
}

int frcnn_dp_allreduce(frcnn_ctx* const* ctxs, int n, float* const* counters_dev, int n_counters) {
  if (!ctxs || n < 1 || n > 64 || !ctxs[0]) return FRCNN_E_INVALID;
  frcnn_ctx* c = ctxs[0];
  API_BEGIN(c)
  for (int i = 0; i < n; ++i) {
    FRCNN_REQUIRE(ctxs[i] && ctxs[i]->dp_comm, FRCNN_E_STATE, "frcnn_dp_init_rank / frcnn_dp_init_all must be called first");
    for (auto g : ctxs[i]->grads) FRCNN_REQUIRE(g != nullptr, FRCNN_E_STATE, "frcnn_bind_grads must be called first");
  }
  FRCNN_REQUIRE(n_counters >= 0 && (n_counters == 0 || counters_dev != nullptr), FRCNN_E_INVALID, "bad counters");
  // one NCCL group over every context this thread drives (a single thread must not issue ungrouped collectives to
  // several devices): whatever pnet:backward has not sent yet, plus the counters
  FRCNN_NCCL_TRY(frcnn::nccl().GroupStart());
  try {
    for (int i = 0; i < n; ++i) {
      frcnn_ctx* x = ctxs[i];
      FRCNN_CUDA_TRY(cudaSetDevice(x->device));
      for (int b = 0; b < frcnn::dp_bucket_count(x); ++b)
        if (!x->dp_bucket_sent[b]) frcnn::dp_send_bucket(x, b);
      if (n_counters > 0) {
        FRCNN_CUDA_TRY(cudaEventRecord(x->dp_ready, x->stream));
        FRCNN_CUDA_TRY(cudaStreamWaitEvent(x->dp_stream, x->dp_ready, 0));
        FRCNN_NCCL_TRY(frcnn::nccl().AllReduce(counters_dev[i], counters_dev[i], (size_t)n_counters, frcnn::NCCL_FLOAT, frcnn::NCCL_SUM,
                                               x->dp_comm, x->dp_stream));
      }
    }
  } catch (...) {
    frcnn::nccl().GroupEnd();
    throw;
  }
  FRCNN_NCCL_TRY(frcnn::nccl().GroupEnd());
  for (int i = 0; i < n; ++i) {
    frcnn_ctx* x = ctxs[i];
    FRCNN_CUDA_TRY(cudaSetDevice(x->device));
    // the context's stream continues (optimiser step, next forward) only after the sums have arrived
    FRCNN_CUDA_TRY(cudaEventRecord(x->dp_done, x->dp_stream));
    FRCNN_CUDA_TRY(cudaStreamWaitEvent(x->stream, x->dp_done, 0));
    std::fill(x->dp_bucket_sent.begin(), x->dp_bucket_sent.end(), 0);
  }
  FRCNN_CUDA_TRY(cudaSetDevice(c->device));
  API_END(c)
}

int frcnn_dp_info(const frcnn_ctx* c, int* rank, int* nranks, int* nccl_version, int64_t* bytes_reduced) {
  if (!c) return FRCNN_E_INVALID;
  if (rank) *rank = c->dp_rank;
  if (nranks) *nranks = c->dp_comm ? c->dp_nranks : 0;
  if (bytes_reduced) *bytes_reduced = c->dp_bytes;
  if (nccl_version) {
    *nccl_version = 0;
    try {
      frcnn::nccl().GetVersion(nccl_version);
    } catch (...) {
    }
  }
  return FRCNN_OK;
}

}  // extern "C"
#pragma GCC visibility pop
