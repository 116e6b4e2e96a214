// HBM-bound helper kernels around the tensor-core conv: weight re-packing, 2x2 ceil-mode max pooling (standalone
// variant; the trunk uses the pool fused into the conv epilogue), anchor-head tail (bias + PReLU + 1x1 conv), cnet tails and layout converters.
// All are coalesced / 16-byte vectorised streaming kernels; none has data reuse worth staging in shared memory
// except the small weight matrices of the tails.
#include <algorithm>

#include "common.h"

namespace frcnn {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// Forward weights are packed TWICE: copy 0 = bf16 (training forward), copy 1 = fp16 (evaluate / detect,
// ConvParams::f16); `copy_stride` elements apart (0 = bf16 copy only).  fp16 conversion saturates to +-65504.
__device__ __forceinline__ void store_op16_copies(bf16* out, long i, long copy_stride, float v) {
  out[i] = __float2bfloat16_rn(v);
  if (copy_stride) reinterpret_cast<__half*>(out)[copy_stride + i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
}
__device__ __forceinline__ float load_op16(const bf16* p, int f16) {
  return f16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
}
__device__ __forceinline__ void store_op16(bf16* p, float v, int f16) {
  if (f16) *reinterpret_cast<__half*>(p) = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  else *p = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------ weight packing
// Torch conv weight [Cout][Cin][KH][KW] fp32 -> [Cout][KH][KW][Cin] bf16 (K-major GEMM B operand).
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int KH,
                                        int KW, int copies) {
  long total = (long)Cout * Cin * KH * KW;
  const long cs = copies > 1 ? total : 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = i % Cin;
    long r = i / Cin;
    int kw = r % KW;
    r /= KW;
    int kh = r % KH;
    int o = r / KH;
    store_op16_copies(out, i, cs, w[(((long)o * Cin + c) * KH + kh) * KW + kw]);
  }
}
void launch_pack_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st, int copies) {
  long total = (long)Cout * Cin * KH * KW;
  pack_conv_weight_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, Cout, Cin, KH, KW, copies);
}

// First layer (Cin = 3): the flat Torch weight row [Cin*KH*KW] is already the im2col K order; pad K to 32.
__global__ void pack_first_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 32) return;
  int k = i & 31, o = i >> 5;
  store_op16_copies(out, i, (long)Cout * 32, k < K ? w[o * K + k] : 0.f);   // always both copies: [2][Cout][32]
}
void launch_pack_first_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st) {
  pack_first_conv_weight_kernel<<<cdiv(Cout * 32, 256), 256, 0, st>>>(w, out, Cout, Cin * KH * KW);
}

__global__ void pack_conv_weight_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int KH,
                                              int KW) {
  long total = (long)Cout * Cin * KH * KW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int co = i % Cout;
    long r = i / Cout;
    int kw = r % KW;
    r /= KW;
    int kh = r % KH;
    int ci = r / KH;
    out[i] = __float2bfloat16_rn(w[(((long)co * Cin + ci) * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)]);
  }
}
void launch_pack_conv_weight_dgrad(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st) {
  long total = (long)Cout * Cin * KH * KW;
  pack_conv_weight_dgrad_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, Cout, Cin, KH, KW);
}

__global__ void wgrad_finish_kernel(const float* __restrict__ dw, float* __restrict__ grad, int Cout, int Cin, int taps) {
  // (a filter bank has far fewer than 2^31 weights: 32-bit index arithmetic)
  const int total = Cout * Cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % taps;
    const int r = i / taps;
    const int ci = r % Cin;
    const int co = r / Cin;
    grad[i] += dw[(co * taps + t) * Cin + ci];
  }
}
void launch_wgrad_finish(const float* dw_taps, float* grad, int Cout, int Cin, int KH, int KW, cudaStream_t st) {
  long total = (long)Cout * Cin * KH * KW;
  wgrad_finish_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(dw_taps, grad, Cout, Cin, KH * KW);
}

// NHWC bf16 -> planar [N][C][H][pitch] bf16 through a 32 x 32 shared-memory transpose of (pixel-in-row, channel)
__global__ void nhwc_to_planar_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H, int W, int C, int pitch) {
  __shared__ bf16 tile[32][34];
  const int n = blockIdx.z / H, h = blockIdx.z % H;
  const int w0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int w = w0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (w < W && c < C) ? in[(((long)n * H + h) * W + w) * C + c] : zero;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, w = w0 + threadIdx.x;
    if (w < pitch && c < C) out[(((long)n * C + c) * H + h) * pitch + w] = tile[threadIdx.x][j];
  }
}
void launch_nhwc_to_planar_bf16(const bf16* in, bf16* out, int N, int H, int W, int C, int pitch, cudaStream_t st) {
  dim3 grid(cdiv(pitch, 32), cdiv(C, 32), N * H), block(32, 8);
  nhwc_to_planar_kernel<<<grid, block, 0, st>>>(in, out, H, W, C, pitch);
}

// Linear weight [nout][K] fp32 -> bf16.  With permute: K index c*bins + b (reference ROI-pool flatten order,
// Detector.lua:97) -> b*C + c (the channel-contiguous order the ROI-pool kernel writes).
__global__ void pack_fc_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int nout, int C, int bins,
                                      int permute, int copies) {
  long K = (long)C * bins;
  long total = (long)nout * K;
  const long cs = copies > 1 ? total : 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long o = i / K;
    long k = i - o * K;
    long src = k;
    if (permute) {
      int b = k / C, c = k - (long)b * C;
      src = (long)c * bins + b;
    }
    store_op16_copies(out, i, cs, w[o * K + src]);
  }
}
void launch_pack_fc_weight(const float* w, bf16* out, int nout, int C, int bins, int permute, cudaStream_t st, int copies) {
  long total = (long)nout * C * bins;
  pack_fc_weight_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, nout, C, bins, permute, copies);
}

// ------------------------------------------------------------------------------------------ 2x2 ceil max pool
// nn.SpatialMaxPooling(2,2,2,2):ceil() (model_utilities.lua:23) on NHWC bf16; the last window is clipped.
__global__ void maxpool2x2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int N, int H, int W, int C, int Ho,
                                  int Wo) {
  int cv = C >> 3;
  long total = (long)N * Ho * Wo * cv;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c8 = i % cv;
    long r = i / cv;
    int xo = r % Wo;
    r /= Wo;
    int yo = r % Ho;
    int n = r / Ho;
    int y0 = yo * 2, x0 = xo * 2;
    const uint4* base = reinterpret_cast<const uint4*>(in);
    auto at = [&](int y, int x) { return base[(((long)n * H + y) * W + x) * cv + c8]; };
    uint4 m = at(y0, x0);
    auto mx = [&](uint4 a, uint4 b) {
      uint4 o;
      __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
      __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&b);
      __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) po[j] = __hmax2(pa[j], pb[j]);
      return o;
    };
    bool hx = x0 + 1 < W, hy = y0 + 1 < H;
    if (hx) m = mx(m, at(y0, x0 + 1));
    if (hy) m = mx(m, at(y0 + 1, x0));
    if (hx && hy) m = mx(m, at(y0 + 1, x0 + 1));
    reinterpret_cast<uint4*>(out)[i] = m;
  }
}
void launch_maxpool2x2(const bf16* in, bf16* out, int N, int H, int W, int C, cudaStream_t st) {
  int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  long total = (long)N * Ho * Wo * (C / 8);
  maxpool2x2_kernel<<<min(cdiv(total, 256), 148 * 16), 256, 0, st>>>(in, out, N, H, W, C, Ho, Wo);
}

// ------------------------------------------------------------------------------------------ anchor-head tail
// AnchorNetwork tail (model_utilities.lua:32-33) for all anchor heads in one launch: sum of the split-K slices of
// the k x k conv + bias -> PReLU -> 1x1 conv to 3*(2+4) = 18 channels, written in Torch layout [N][18][H][W] fp32.
// fp32 throughout, fixed summation order (deterministic).  A warp handles 4 pixels at a time: every lane owns 8 of
// the 256 mid channels, the 18 x 256 weights are read from shared memory once per 4 pixels, and the 72 partial
// sums are reduced across the warp with a transposing butterfly (31 shuffles per 32 values).
template <int NV>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
  // after the loop lane l holds the warp-wide total of v[l]
#pragma unroll
  for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const float send = upper ? v[i] : v[i + cnt];
      const float keep = upper ? v[i + cnt] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(256) head_tail_group_kernel(HeadTailGroup g) {
  constexpr int CO = 18, CM = 256;
  extern __shared__ float sw[];  // [18][256] weights, [256] bias
  float* sbias = sw + CO * CM;
  int hi = 0;
  while (hi + 1 < g.n && (int)blockIdx.x >= g.h[hi].block_end) ++hi;
  const HeadTail& H = g.h[hi];
  const int block0 = hi ? g.h[hi - 1].block_end : 0;
  // parameter pointers are views into Torch's flat buffer at arbitrary 4-byte offsets: vector loads only when the
  // view happens to be 16-byte aligned
  if ((reinterpret_cast<uintptr_t>(H.w2) & 15) == 0) {
    for (int i = threadIdx.x; i < CO * CM / 4; i += blockDim.x)
      reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(H.w2) + i);
  } else {
    for (int i = threadIdx.x; i < CO * CM; i += blockDim.x) sw[i] = __ldg(H.w2 + i);
  }
  for (int i = threadIdx.x; i < CM; i += blockDim.x) sbias[i] = __ldg(H.bias + i);
  __syncthreads();
  const float slope = H.prelu[0];
  const int lane = threadIdx.x & 31;
  const long warp_idx = ((long)(blockIdx.x - block0) * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)(H.block_end - block0) * blockDim.x) >> 5;
  for (long base = warp_idx * 4; base < H.npix; base += nwarps * 4) {
    float part[CO][4];
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int q = 0; q < 4; ++q) part[o][q] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = half * 128 + lane * 4;
      const float4 b = *reinterpret_cast<const float4*>(sbias + c);
      float4 hv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) hv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      // split-K slices summed in slice order (deterministic); the 4 pixels' loads of a slice are independent
#pragma unroll 2
      for (int s = 0; s < H.splits; ++s) {
        float4 t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long pix = base + q < H.npix ? base + q : H.npix - 1;
          t[q] = __ldg(reinterpret_cast<const float4*>(H.ws + (size_t)s * H.slice_stride + pix * CM + c));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          hv[q].x += t[q].x; hv[q].y += t[q].y; hv[q].z += t[q].z; hv[q].w += t[q].w;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 a = hv[q];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        a.x = a.x > 0.f ? a.x : a.x * slope;
        a.y = a.y > 0.f ? a.y : a.y * slope;
        a.z = a.z > 0.f ? a.z : a.z * slope;
        a.w = a.w > 0.f ? a.w : a.w * slope;
        hv[q] = a;
      }
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        const float4 w = *reinterpret_cast<const float4*>(sw + o * CM + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) part[o][q] += hv[q].x * w.x + hv[q].y * w.y + hv[q].z * w.z + hv[q].w * w.w;
      }
    }
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int idx = grp * 32 + i;
        v[i] = idx < CO * 4 ? part[idx >> 2 < CO ? idx >> 2 : 0][idx & 3] : 0.f;
      }
      const float tot = warp_transpose_reduce<32>(v, lane);
      const int idx = grp * 32 + lane;
      const int o = idx >> 2, q = idx & 3;
      const long pix = base + q;
      if (o < CO && pix < H.npix) {
        const long n = pix / H.HW, hw = pix - n * H.HW;
        H.out[(n * CO + o) * H.HW + hw] = tot + H.b2[o];
      }
    }
  }
}
void launch_head_tail_group(const HeadTailGroup& g_in, int num_sms, cudaStream_t st) {
  HeadTailGroup g = g_in;
  FRCNN_REQUIRE(g.n >= 1 && g.n <= 4, FRCNN_E_INVALID, "anchor head group: 1..4 heads");
  // blocks per head proportional to its pixels (8 warps x 4 pixels per block iteration), about two blocks per SM:
  // every block stages the 18 x 256 weights once, so fewer, longer-running blocks amortise that better
  long total_px = 0;
  for (int i = 0; i < g.n; ++i) total_px += g.h[i].npix;
  const long budget = (long)num_sms * 2;
  int end = 0;
  for (int i = 0; i < g.n; ++i) {
    long want = (g.h[i].npix + 31) / 32;
    long share = (budget * g.h[i].npix + total_px - 1) / total_px;
    long blocks = want < share ? want : share;
    if (blocks < 1) blocks = 1;
    end += (int)blocks;
    g.h[i].block_end = end;
  }
  const int smem = 19 * 256 * sizeof(float);
  head_tail_group_kernel<<<end, 256, smem, st>>>(g);
}

// ------------------------------------------------------------------------------------------ layout converters
// NHWC bf16 -> [N][C][H][W] fp32 (Torch layout) through a 32x32 shared-memory transpose of (pixel, channel).
__global__ void nhwc_to_chw_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C, int f16) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[j][threadIdx.x] = load_op16(in + ((long)n * HW + p) * C + c, f16);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) out[((long)n * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}
void launch_nhwc_bf16_to_chw_f32(const bf16* in, float* out, int N, int H, int W, int C, cudaStream_t st, int f16) {
  dim3 grid(cdiv((long)H * W, 32), cdiv(C, 32), N), block(32, 8);
  nhwc_to_chw_kernel<<<grid, block, 0, st>>>(in, out, H * W, C, f16);
}
__global__ void chw_to_nhwc_kernel(const float* __restrict__ in, bf16* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) tile[j][threadIdx.x] = in[((long)n * C + c) * HW + p];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[((long)n * HW + p) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}
void launch_chw_f32_to_nhwc_bf16(const float* in, bf16* out, int N, int H, int W, int C, cudaStream_t st) {
  dim3 grid(cdiv((long)H * W, 32), cdiv(C, 32), N), block(32, 8);
  chw_to_nhwc_kernel<<<grid, block, 0, st>>>(in, out, H * W, C);
}

// ------------------------------------------------------------------------------------------ cnet tails
// nn.Linear bias + nn.BatchNormalization (evaluate mode, eps 1e-5) + nn.PReLU (model_utilities.lua:82-86) on the
// fp32 split-K sums of a Linear layer.  Writes bf16 (operand of the next tensor-core GEMM) and/or fp32.
// acc: either the summed fp32 GEMM output [rows][n] (sl.k_iters == 0) or the tile-major split-K slices the conv kernel
// wrote with a device-chosen split count (ConvParams::slice_tile_major): the count is re-derived here from the live
// row count with the kernel's own formula (make_sched) and the slices are summed in ascending order -- a fixed
// summation order, so cnet:forward is reproducible run to run.
__global__ void fc_tail_kernel(const float* __restrict__ acc, const float* __restrict__ bias, const float* __restrict__ bn_w,
                               const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                               const float* __restrict__ bn_var, const float* __restrict__ prelu, bf16* __restrict__ out_bf16,
                               float* __restrict__ out_f32, int rows_max, const int* __restrict__ rows_dev, int n, FcSlices sl,
                               int f16) {
  int rows = rows_dev ? min(*rows_dev, rows_max) : rows_max;
  long total = (long)rows * n;
  const float slope = prelu[0];
  int splits = 0;
  if (sl.k_iters > 0) {
    const int m_tiles = max(1, (rows + 127) / 128);
    int want = sl.dyn_ctas / (m_tiles * sl.n_tiles_n);
    want = max(1, min(want, sl.host_splits));
    const int kps = (sl.k_iters + want - 1) / want;
    splits = (sl.k_iters + kps - 1) / kps;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = i % n;
    float x;
    if (splits > 0) {
      const long r = i / n;
      const float* src = acc + (((r >> 7) * splits) * 128 + (r & 127)) * n + c;
      x = src[0];
      for (int s = 1; s < splits; ++s) x += src[(long)s * 128 * n];
      x += bias[c];
    } else {
      x = acc[i] + bias[c];
    }
    if (bn_w) {
      float inv = 1.0f / sqrtf(bn_var[c] + 1e-5f);
      x = (x - bn_mean[c]) * inv * bn_w[c] + bn_b[c];
    }
    x = x > 0.f ? x : x * slope;
    if (out_bf16) store_op16(out_bf16 + i, x, f16);
    if (out_f32) out_f32[i] = x;
  }
}
void launch_fc_tail(const float* acc, const float* bias, const float* bn_w, const float* bn_b, const float* bn_mean,
                    const float* bn_var, const float* prelu, bf16* out_bf16, float* out_f32, int rows_max,
                    const int* rows_dev, int n, cudaStream_t st, const FcSlices* sl, int grid_rows, int f16) {
  FcSlices none = {0, 0, 0, 0};
  // grid_rows: rows the grid is sized for (the typical live count, not the capacity): the grid-stride loop covers more
  long total = (long)(grid_rows > 0 ? std::min(grid_rows, rows_max) : rows_max) * n;
  fc_tail_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(acc, bias, bn_w, bn_b, bn_mean, bn_var, prelu, out_bf16,
                                                                  out_f32, rows_max, rows_dev, n, sl ? *sl : none, f16);
}

// The two output branches of cnet (model_utilities.lua:96-105): Linear(nin -> 4) and Linear(nin -> ncls) +
// LogSoftMax, fp32.  One CTA per ROI row: the hidden vector sits in shared memory, one warp per output neuron
// (lanes stride the K dimension, coalesced weight reads), then a block-level log-softmax.
__global__ void cnet_out_kernel(const float* __restrict__ hidden, const float* __restrict__ w_reg, const float* __restrict__ b_reg,
                                const float* __restrict__ w_cls, const float* __restrict__ b_cls, float* __restrict__ reg_out,
                                float* __restrict__ cls_out, int rows_max, const int* __restrict__ rows_dev, int nin, int ncls) {
  extern __shared__ float sh[];  // [nin] hidden, [ncls + 4] logits
  int rows = rows_dev ? min(*rows_dev, rows_max) : rows_max;
  float* h = sh;
  float* logit = sh + nin;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
  __syncthreads();  // the previous row's logits / hidden vector have been consumed
  for (int i = threadIdx.x; i < nin; i += blockDim.x) h[i] = hidden[(long)r * nin + i];
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < ncls + 4; o += nw) {
    const float* wr = o < 4 ? w_reg + (long)o * nin : w_cls + (long)(o - 4) * nin;
    float s = 0.f;
    for (int k = lane; k < nin; k += 32) s += h[k] * wr[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) logit[o] = s + (o < 4 ? b_reg[o] : b_cls[o - 4]);
  }
  __syncthreads();
  if (threadIdx.x < 4) reg_out[(long)r * 4 + threadIdx.x] = logit[threadIdx.x];
  if (warp == 0) {
    float m = -INFINITY;
    for (int c = lane; c < ncls; c += 32) m = fmaxf(m, logit[4 + c]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float s = 0.f;
    for (int c = lane; c < ncls; c += 32) s += expf(logit[4 + c] - m);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    float lse = m + logf(s);
    for (int c = lane; c < ncls; c += 32) cls_out[(long)r * ncls + c] = logit[4 + c] - lse;
  }
  }
}
void launch_cnet_out(const float* hidden, const float* w_reg, const float* b_reg, const float* w_cls, const float* b_cls,
                     float* reg_out, float* cls_out, int rows_max, const int* rows_dev, int nin, int ncls, cudaStream_t st) {
  if (rows_max <= 0) return;
  int smem = (nin + ncls + 4) * sizeof(float);
  // the live row count is a device value: a grid of a few rows per SM strides over whatever it turns out to be
  cnet_out_kernel<<<std::min(rows_max, 148 * 4), 256, smem, st>>>(hidden, w_reg, b_reg, w_cls, b_cls, reg_out, cls_out, rows_max, rows_dev, nin,
                                                ncls);
}

}  // namespace frcnn
