// Shared declarations of the B200-native Faster R-CNN hot-path library (internal; the public surface is
// include/frcnn_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/frcnn_b200.h"

namespace frcnn {

typedef __nv_bfloat16 bf16;

struct Error {
  int code;
  std::string msg;
};

void set_global_error(const std::string& msg);

#define FRCNN_CUDA_TRY(expr)                                                                            \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      throw ::frcnn::Error{FRCNN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                                             __FILE__ + ":" + std::to_string(__LINE__)};                \
    }                                                                                                   \
  } while (0)

#define FRCNN_REQUIRE(cond, code, text)                   \
  do {                                                    \
    if (!(cond)) throw ::frcnn::Error{(code), (text)};    \
  } while (0)

// cudaFuncSetAttribute is per device: a `static bool configured` would skip the second device of a multi-device
// process.  first_use_on_device(flags) is true once per (call site, device).
struct DeviceOnce {
  bool done[64] = {};
};
inline bool first_use_on_device(DeviceOnce& o) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (o.done[dev]) return false;
  o.done[dev] = true;
  return true;
}

// ------------------------------------------------------------------ conv / GEMM kernel interface
enum ConvEpilogue {
  EPI_STORE = 0,       // y = prelu(acc + bias) * scale -> bf16 NHWC, staged in shared memory, written by TMA store
  EPI_F32_REDUCE = 1,  // split-K partial sums added into a zeroed fp32 NHWC buffer [N][Hout][Wout][Cout] by TMA
                       // tensor reductions (cp.reduce.async.bulk.tensor .add); summation order not reproducible
  EPI_POOL = 2,        // as EPI_STORE followed by the 2x2 stride-2 ceil-mode max pool (model_utilities.lua:23):
                       // only the pooled map is written
  EPI_F32_SLICES = 3,  // deterministic split-K: raw fp32 partial sums, split s of image n -> slice [s * N + n] of an
                       // fp32 NHWC workspace [splits * N][Hout][Wout][Cout] (TMA store); summed by the consumer
  EPI_HEAD = 4,        // fused AnchorNetwork (model_utilities.lua:29-35): bias + PReLU + 1x1 conv to 18 channels in the
                       // epilogue, fp32; out = [N][18][Hout][Wout] fp32 (halo kernel, Cout = 256, no split-K)
};

struct ConvParams {
  int N, Hin, Win, Cin;     // input NHWC (Cin % 64 == 0)
  int Hout, Wout, Cout;     // output
  int KH, KW, padH, padW;   // stride 1
  int BW, BH, bw_shift;     // spatial shape of one 128-pixel M sub-tile (BW * BH == 128)
  int MT;                   // 128-row sub-tiles per CTA tile (1 or 2), stacked along H: the CTA tile is BW x (BH * MT)
  int tiles_w, tiles_h;     // ceil(Wout / BW), ceil(Hout / (BH * MT))
  int n_tiles_m, n_tiles_n; // N * tiles_h * tiles_w, ceil(Cout / BN)
  int cchunks;              // Cin / 64
  int k_iters;              // KH * KW * cchunks
  int splits, k_per_split;  // split-K
  int mode;                 // ConvEpilogue
  const float* bias;        // [Cout] or null
  const float* prelu;       // device pointer to the shared slope, or null (identity)
  float scale;              // post-activation scale (SpatialDropout eval factor), 1 if none
  const float* chan_scale;  // training: per-(image, channel) SpatialDropout mask [N][Cout] (0 / 1, no rescale), or null
  uint8_t* pool_arg;        // training, EPI_POOL: winner of every 2x2 window (0..3 = dy*2+dx) [N][Hp][Wp][Cout], or null
  void* out;                // EPI_STORE: bf16 NHWC [N][Hout][Wout][Cout]; EPI_POOL: its pooled map; fp32 modes: workspace
  const int* m_limit;       // optional device int: tiles whose first row >= *m_limit are skipped (GEMM rows)
  int dyn_ctas;             // with m_limit: > 0 = choose the split-K factor on the device so that the tiles of the
                            // *m_limit live rows fill dyn_ctas CTAs (host `splits` is then the upper bound)
  // weight-gradient GEMM (conv_wgrad_prepare): M = Cout rows, N = Cin columns of ONE filter tap, K = output pixels
  // in BW x BH = 64-pixel patches.  Both operands are read from the NHWC tensors as MN-major tiles ([64 pixels][64
  // channels] TMA boxes); the tap shift and the zero padding are TMA coordinates / OOB fill.  Reuses the fields:
  // n_tiles_m = taps * co_tiles, n_tiles_n = ci tiles, wchunks / tiles_h = patch columns / rows, k_iters = N * patches.
  int wgrad, co_tiles, wchunks;
  const float* img;         // first-layer kernel only: [N][Cimg][Hin][Win] fp32 input frames (Torch layout)
  int Cimg;                 // first-layer kernel only: image channels (3); K = Cimg * KH * KW <= 32
  // halo-tile kernel (conv_halo_kernel): the A operand of ALL filter taps of one 64-channel chunk is ONE TMA box
  // {64 ch, BW + KW - 1, BH * MT + KH - 1} = the CTA tile plus its halo; tap (kh, kw) is a row offset of the UMMA
  // descriptor into that box.  BW = 8 (one 8-row descriptor group per tile row), BH = 16.
  int halo;                 // 1: launched with conv_halo_kernel (with wgrad: conv_wgrad_halo_kernel, MT = taps per unit)
  int swap;                 // 1: conv_halo_kernel<128, 2, 3, 1, SWAP>: filters on the M side, 256 pixels on the N side (epilogue_swap)
  int pair;                 // 1: launched with conv_pair_kernel on CTA pairs (cta_group::2): a unit covers TWO CTA tiles stacked
                            // along H (cluster rank r computes rows [r, r + 1) * BH * MT of it); tiles_h counts pair tiles
  int occ;                  // CTAs resident per SM the launch is sized for (1, or 2: halved shared memory / TMEM per CTA; 3: conv_pair_bres_kernel)
  unsigned long long* trace; // measurement only (FRCNN_CONV_TRACE): per-CTA / per-unit globaltimer stamps of conv_halo_kernel; nullptr = off
  int a_slots;              // conv_pair_bres_kernel: depth of the activation ring (what the resident weights leave room for)
  int first_tma;            // first-layer kernel: eligible for conv_first_tma_kernel (16 x 8 tiles, Win % 4 == 0)
  int halo_desc;            // descriptor base-offset mode for the row-shifted start address (0: field left 0)
  const float* w2;          // EPI_HEAD: [18][256] weights of the 1x1 convolution (Torch layout)
  const float* b2;          // EPI_HEAD: [18] bias of the 1x1 convolution
  int f16;                  // 16-bit operand format of this launch: 0 = bf16 (training: gradients need the exponent range),
                            // 1 = fp16 (evaluate / detect: 11 significand bits instead of 8 at the same tensor-core rate; the
                            // precision contract of DESIGN.md).  Selects the weight copy (third coordinate of the weight
                            // tensor map), the MMA instruction descriptor's operand formats and the epilogue's conversion
  int slice_tile_major;     // EPI_F32_SLICES of a GEMM (BH == 1, N == 1) whose split count is chosen on the device: slice s of
                            // 128-row tile m is rows [(m * splits + s) * 128, +128) of the workspace, so that the workspace
                            // is bounded by max(M tiles, dyn_ctas / N tiles) tiles whatever the live row count
  int dbg;                  // FRCNN_CONV_DBG (measurement only): 1 skip the global stores, 2 skip the epilogue body, 4 skip the MMAs
};

struct TensorMapCache;

struct ConvLaunch {
  ConvParams p;
  int BN;
  bool first;               // fused first-layer kernel (in-kernel im2col of the fp32 image)
  CUtensorMap tmA, tmB, tmOut;
  const bf16* w_first;      // first-layer kernel: packed [2][Cout][32] weights (bf16 copy, fp16 copy)
  int w_copies;             // operand-format copies behind tmB: 1 = bf16 only, 2 = [bf16 | fp16] (ConvParams::f16 selects)
  int grid;
};

// Several independent convolutions executed by ONE launch of the conv kernel (the four anchor heads): work units
// (conv g, tile, split) are enumerated conv by conv and dealt round-robin to the persistent CTAs.
static constexpr int MAX_GROUP = 4;
struct ConvGroup {
  ConvParams p[MAX_GROUP];
  int unit_end[MAX_GROUP];  // exclusive prefix of the unit counts
  int n;
};
struct ConvMaps {
  CUtensorMap a[MAX_GROUP], b[MAX_GROUP], o[MAX_GROUP];
};

// Work schedule of conv_head_kernel (the four anchor networks in one launch): units = (head, image, 128-position tile,
// filter-row range), dealt to the persistent CTAs by the host.  units[i] = {head | slice << 8 | n_slices << 16, image, tile,
// kh0 | kh1 << 8}; CTA b processes units [cta_off[b], cta_off[b + 1]).
struct HeadSched {
  const int4* units;
  const int* cta_off;
  float* slices[MAX_GROUP];    // per head: [image][tile][slice][128][256] fp32 partial sums (split heads only)
  int* counters[MAX_GROUP];    // per head: [image][tile] arrivals; zero between launches
  int tiles[MAX_GROUP];        // tiles per image
  unsigned long long* trace;   // measurement only (FRCNN_HEAD_TRACE): [cta][8 units][8] globaltimer stamps; nullptr = off
};
// head_fixup_kernel: the anchor networks whose reduction conv_head_kernel split by filter rows (slices summed in ascending
// order + bias + PReLU + 1 x 1 conv).  Block b of head i handles 32 positions of tile b / 4.
struct HeadFixArgs {
  int n;                       // split heads
  int head[MAX_GROUP];         // index into the ConvGroup
  int nsl[MAX_GROUP];          // slices per tile
  int block_end[MAX_GROUP];    // running block count
};
// Builds the launch state of the fused anchor-network kernel for N frames of in_h x in_w (per head) and runs it.
struct HeadPlan {
  ConvMaps maps;
  ConvGroup grp;
  HeadSched sched;
  HeadFixArgs fix;
  int grid = 0;
  int n_units = 0;
  double flops = 0.0;
};

struct HeadDesc {
  const bf16* in;        // NHWC 16-bit input map [N][Hin][Win][Cin]
  const bf16* w_packed;  // [2][256][K][K][Cin] packed weights (bf16 | fp16 copies)
  int Hin, Win, Cin, K;
};
void conv_head_plan(HeadPlan* P, const HeadDesc* heads, int n_heads, int N, int num_sms, std::vector<int4>* units_out,
                    std::vector<int>* cta_off_out, size_t slice_floats[MAX_GROUP], int counter_ints[MAX_GROUP]);
void conv_launch_heads(const HeadPlan& P, cudaStream_t st);

// Host helpers (conv_igemm.cu)
void conv_choose_tile(int Hout, int Wout, int* BW, int* BH);
void make_tmap_act(CUtensorMap* m, const bf16* base, int N, int H, int W, int C, int BW, int BH);
void make_tmap_weight(CUtensorMap* m, const bf16* base, int Cout, int K, int BN, int copies);
void conv_prepare(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                  int KH, int KW, int padH, int padW, int mode, bf16* out, int num_sms, int force_splits, int force_bn,
                  int force_mt, int w_copies = 1);
// First layer (Cin = 3): reads the fp32 NCHW frames directly; w_packed32: [Cout][32] bf16 (K = Cin*KH*KW padded).
void conv_first_prepare(ConvLaunch* L, const bf16* w_packed32, int N, int Hin, int Win, int Cimg, int Cout, int KH,
                        int KW, int padH, int padW, int mode, bf16* out, int num_sms);
void conv_launch(const ConvLaunch& L, cudaStream_t st);
// all members: same BN, MT == 1, not the first-layer kernel; pass them heaviest (longest K per unit) first
void conv_launch_group(const ConvLaunch* const* Ls, int n, int num_sms, cudaStream_t st);
// Fused anchor head (EPI_HEAD): k x k valid conv (k <= 7) to 256 hidden channels + the 1x1 tail in the epilogue; set
// p.bias / p.prelu / p.w2 / p.b2 / p.out before launching.  conv_launch_head_group: up to 4 heads in ONE launch of the
// halo kernel, pass them heaviest first.
void conv_prepare_head(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int K, int num_sms);
void conv_launch_head_group(const ConvLaunch* const* Ls, int n, int num_sms, cudaStream_t st);
// dW[co][tap][ci] (fp32, zeroed by the caller) += sum over pixels dY[p][co] * X[p + tap][ci]; dy / x: NHWC bf16;
// always EPI_F32_REDUCE
void conv_wgrad_prepare(ConvLaunch* L, const bf16* dy, const bf16* x, float* dw_taps, int N, int Hin, int Win, int Cin,
                        int Cout, int KH, int KW, int padH, int padW, int num_sms);
// fp32 output map: EPI_F32_SLICES: ws is [splits * N][Hout][Wout][Cout]; EPI_F32_REDUCE: [N][Hout][Wout][Cout]
void conv_set_f32_output(ConvLaunch* L, float* ws);
int conv_smem_bytes(int BN);

// ------------------------------------------------------------------ element kernels (elementwise.cu)
// copies = 2: out is [2][...]: the bf16 copy followed by the fp16 copy (ConvParams::f16)
void launch_pack_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st, int copies = 1);
void launch_pack_first_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st);
// dgrad weights: out[ci][kh'][kw'][co] = w[co][ci][KH-1-kh'][KW-1-kw'] (bf16), the K-major B operand of the
// transposed convolution
void launch_pack_conv_weight_dgrad(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st);
// NHWC bf16 -> planar [N][C][H][pitch] bf16 (pitch >= W, padding columns zeroed)
void launch_nhwc_to_planar_bf16(const bf16* in, bf16* out, int N, int H, int W, int C, int pitch, cudaStream_t st);
// grad[co][ci][kh][kw] (Torch layout, fp32) += dw_taps[co][kh*KW+kw][ci]
void launch_wgrad_finish(const float* dw_taps, float* grad, int Cout, int Cin, int KH, int KW, cudaStream_t st);
void launch_pack_fc_weight(const float* w, bf16* out, int nout, int C, int bins, int permute, cudaStream_t st, int copies = 1);
void launch_maxpool2x2(const bf16* in, bf16* out, int N, int H, int W, int C, cudaStream_t st);
// AnchorNetwork tails of all heads in one launch (mid width 256, 18 outputs: model_utilities.lua:29-35)
struct HeadTail {
  const float* ws;      // [splits][npix][256] fp32 split-K slices of the k x k conv
  int splits;
  long slice_stride;    // npix * 256
  long npix;            // N * H * W
  int HW;
  const float *bias, *prelu, *w2, *b2;
  float* out;           // [N][18][H][W] fp32
  int block_end;        // filled by the launcher
};
struct HeadTailGroup {
  HeadTail h[4];
  int n;
};
void launch_head_tail_group(const HeadTailGroup& g, int num_sms, cudaStream_t st);
void launch_nhwc_bf16_to_chw_f32(const bf16* in, float* out, int N, int H, int W, int C, cudaStream_t st, int f16 = 0);
void launch_chw_f32_to_nhwc_bf16(const float* in, bf16* out, int N, int H, int W, int C, cudaStream_t st);
// split-K slices of a GEMM launched with ConvParams::slice_tile_major (k_iters == 0: acc is the finished sum)
struct FcSlices {
  int k_iters, host_splits, n_tiles_n, dyn_ctas;
};
void launch_fc_tail(const float* acc, const float* bias, const float* bn_w, const float* bn_b, const float* bn_mean,
                    const float* bn_var, const float* prelu, bf16* out_bf16, float* out_f32, int rows_max,
                    const int* rows_dev, int n, cudaStream_t st, const FcSlices* sl = nullptr, int grid_rows = 0, int f16 = 0);
void launch_cnet_out(const float* hidden, const float* w_reg, const float* b_reg, const float* w_cls,
                     const float* b_cls, float* reg_out, float* cls_out, int rows_max, const int* rows_dev, int nin,
                     int ncls, cudaStream_t st);

// ------------------------------------------------------------------ frame normalisation (preprocess_kernels.cu)
void launch_normalize_frame(float* img, int H, int W, int rgb2yuv, int centering, int scaling, const float* k1d_dev, int ksize,
                            float threshold, double* scratch, float* plane_tmp, cudaStream_t st);

// ------------------------------------------------------------------ optimiser (optim_kernels.cu)
void launch_scale_image(const float* src, int C, int sh, int sw, float* tmp, float* dst, int dh, int dw, cudaStream_t st);
void launch_rmsprop_step(float* w, float* g, float* m, long n, double grad_div, double lr, double alpha, double eps, double wd,
                         int num_sms, cudaStream_t st);

}  // namespace frcnn
