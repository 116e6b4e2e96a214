// HBM-bound helper kernels around the tensor-core conv: weight re-packing, 2x2 ceil-mode max pooling (standalone
// variant; the trunk uses the pool fused into the conv epilogue), anchor-head tail (bias + PReLU + 1x1 conv), cnet tails and layout converters.
// All are coalesced / 16-byte vectorised streaming kernels; none has data reuse worth staging in shared memory
// except the small weight matrices of the tails.
#include "common.h"

namespace frcnn {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------ weight packing
// Torch conv weight [Cout][Cin][KH][KW] fp32 -> [Cout][KH][KW][Cin] bf16 (K-major GEMM B operand).
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int KH,
                                        int KW) {
  long total = (long)Cout * Cin * KH * KW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = i % Cin;
    long r = i / Cin;
    int kw = r % KW;
    r /= KW;
    int kh = r % KH;
    int o = r / KH;
    out[i] = __float2bfloat16_rn(w[(((long)o * Cin + c) * KH + kh) * KW + kw]);
  }
}
void launch_pack_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st) {
  long total = (long)Cout * Cin * KH * KW;
  pack_conv_weight_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, Cout, Cin, KH, KW);
}

// First layer (Cin = 3): the flat Torch weight row [Cin*KH*KW] is already the im2col K order; pad K to 32.
__global__ void pack_first_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int K) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 32) return;
  int k = i & 31, o = i >> 5;
  out[i] = __float2bfloat16_rn(k < K ? w[o * K + k] : 0.f);
}
void launch_pack_first_conv_weight(const float* w, bf16* out, int Cout, int Cin, int KH, int KW, cudaStream_t st) {
  pack_first_conv_weight_kernel<<<cdiv(Cout * 32, 256), 256, 0, st>>>(w, out, Cout, Cin * KH * KW);
}

// Linear weight [nout][K] fp32 -> bf16.  With permute: K index c*bins + b (reference ROI-pool flatten order,
// Detector.lua:97) -> b*C + c (the channel-contiguous order the ROI-pool kernel writes).
__global__ void pack_fc_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int nout, int C, int bins,
                                      int permute) {
  long K = (long)C * bins;
  long total = (long)nout * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long o = i / K;
    long k = i - o * K;
    long src = k;
    if (permute) {
      int b = k / C, c = k - (long)b * C;
      src = (long)c * bins + b;
    }
    out[i] = __float2bfloat16_rn(w[o * K + src]);
  }
}
void launch_pack_fc_weight(const float* w, bf16* out, int nout, int C, int bins, int permute, cudaStream_t st) {
  long total = (long)nout * C * bins;
  pack_fc_weight_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, nout, C, bins, permute);
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
// AnchorNetwork tail (model_utilities.lua:32-33): PReLU(acc + bias) followed by the 1x1 conv to 3*(2+4) = 18
// channels, written in Torch layout [N][18][H][W] fp32.  acc: fp32 [N*H*W][Cmid] split-K sums.  One warp per
// pixel; the 18 x Cmid matrix lives in shared memory; fp32 throughout.
template <int COUT2>
__global__ void head_tail_kernel(const float* __restrict__ acc, const float* __restrict__ bias, const float* __restrict__ prelu,
                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out,
                                 long npix, int HW, int Cmid) {
  // parameter pointers are views into Torch's flat weight buffer (utilities.lua:136-147) at arbitrary 4-byte
  // offsets: read them scalar-wise into shared memory, never with vector loads
  extern __shared__ float sw[];  // [COUT2][Cmid] weights, then [Cmid] bias
  float* sbias = sw + COUT2 * Cmid;
  for (int i = threadIdx.x; i < COUT2 * Cmid; i += blockDim.x) sw[i] = w2[i];
  for (int i = threadIdx.x; i < Cmid; i += blockDim.x) sbias[i] = bias[i];
  __syncthreads();
  const float slope = prelu[0];
  int lane = threadIdx.x & 31;
  long warp_global = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long pix = warp_global; pix < npix; pix += nwarps) {
    float part[COUT2];
#pragma unroll
    for (int o = 0; o < COUT2; ++o) part[o] = 0.f;
    for (int c = lane * 4; c < Cmid; c += 128) {
      float4 a = *reinterpret_cast<const float4*>(acc + pix * Cmid + c);
      float4 b = *reinterpret_cast<const float4*>(sbias + c);
      float h[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = h[e] > 0.f ? h[e] : h[e] * slope;
#pragma unroll
      for (int o = 0; o < COUT2; ++o) {
        const float* wr = sw + o * Cmid + c;
        part[o] += h[0] * wr[0] + h[1] * wr[1] + h[2] * wr[2] + h[3] * wr[3];
      }
    }
#pragma unroll
    for (int o = 0; o < COUT2; ++o) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) part[o] += __shfl_xor_sync(0xffffffffu, part[o], s);
    }
    long n = pix / HW;
    long hw = pix - n * HW;
    if (lane < COUT2) {
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < COUT2; ++o)
        if (lane == o) v = part[o];
      out[(n * COUT2 + lane) * HW + hw] = v + b2[lane];
    }
  }
}
void launch_head_tail(const float* acc, const float* bias, const float* prelu, const float* w2, const float* b2,
                      float* out_chw, int N, int H, int W, int Cmid, int Cout2, cudaStream_t st) {
  FRCNN_REQUIRE(Cout2 == 18, FRCNN_E_INVALID, "anchor head must have 18 outputs (model_utilities.lua:33)");
  FRCNN_REQUIRE(Cmid % 128 == 0, FRCNN_E_INVALID, "anchor head width must be a multiple of 128");
  long npix = (long)N * H * W;
  int smem = 19 * Cmid * sizeof(float);
  int blocks = min(cdiv(npix, 8), 148 * 4);
  head_tail_kernel<18><<<blocks, 256, smem, st>>>(acc, bias, prelu, w2, b2, out_chw, npix, H * W, Cmid);
}

// ------------------------------------------------------------------------------------------ layout converters
// NHWC bf16 -> [N][C][H][W] fp32 (Torch layout) through a 32x32 shared-memory transpose of (pixel, channel).
__global__ void nhwc_to_chw_kernel(const bf16* __restrict__ in, float* __restrict__ out, int HW, int C) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) tile[j][threadIdx.x] = __bfloat162float(in[((long)n * HW + p) * C + c]);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) out[((long)n * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}
void launch_nhwc_bf16_to_chw_f32(const bf16* in, float* out, int N, int H, int W, int C, cudaStream_t st) {
  dim3 grid(cdiv((long)H * W, 32), cdiv(C, 32), N), block(32, 8);
  nhwc_to_chw_kernel<<<grid, block, 0, st>>>(in, out, H * W, C);
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
__global__ void fc_tail_kernel(const float* __restrict__ acc, const float* __restrict__ bias, const float* __restrict__ bn_w,
                               const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                               const float* __restrict__ bn_var, const float* __restrict__ prelu, bf16* __restrict__ out_bf16,
                               float* __restrict__ out_f32, int rows_max, const int* __restrict__ rows_dev, int n) {
  int rows = rows_dev ? min(*rows_dev, rows_max) : rows_max;
  long total = (long)rows * n;
  const float slope = prelu[0];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int c = i % n;
    float x = acc[i] + bias[c];
    if (bn_w) {
      float inv = 1.0f / sqrtf(bn_var[c] + 1e-5f);
      x = (x - bn_mean[c]) * inv * bn_w[c] + bn_b[c];
    }
    x = x > 0.f ? x : x * slope;
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(x);
    if (out_f32) out_f32[i] = x;
  }
}
void launch_fc_tail(const float* acc, const float* bias, const float* bn_w, const float* bn_b, const float* bn_mean,
                    const float* bn_var, const float* prelu, bf16* out_bf16, float* out_f32, int rows_max,
                    const int* rows_dev, int n, cudaStream_t st) {
  long total = (long)rows_max * n;
  fc_tail_kernel<<<min(cdiv(total, 256), 148 * 8), 256, 0, st>>>(acc, bias, bn_w, bn_b, bn_mean, bn_var, prelu, out_bf16,
                                                                  out_f32, rows_max, rows_dev, n);
}

// The two output branches of cnet (model_utilities.lua:96-105): Linear(nin -> 4) and Linear(nin -> ncls) +
// LogSoftMax, fp32.  One CTA per ROI row: the hidden vector sits in shared memory, one warp per output neuron
// (lanes stride the K dimension, coalesced weight reads), then a block-level log-softmax.
__global__ void cnet_out_kernel(const float* __restrict__ hidden, const float* __restrict__ w_reg, const float* __restrict__ b_reg,
                                const float* __restrict__ w_cls, const float* __restrict__ b_cls, float* __restrict__ reg_out,
                                float* __restrict__ cls_out, int rows_max, const int* __restrict__ rows_dev, int nin, int ncls) {
  extern __shared__ float sh[];  // [nin] hidden, [ncls + 4] logits
  int rows = rows_dev ? min(*rows_dev, rows_max) : rows_max;
  int r = blockIdx.x;
  if (r >= rows) return;
  float* h = sh;
  float* logit = sh + nin;
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
void launch_cnet_out(const float* hidden, const float* w_reg, const float* b_reg, const float* w_cls, const float* b_cls,
                     float* reg_out, float* cls_out, int rows_max, const int* rows_dev, int nin, int ncls, cudaStream_t st) {
  if (rows_max <= 0) return;
  int smem = (nin + ncls + 4) * sizeof(float);
  cnet_out_kernel<<<rows_max, 256, smem, st>>>(hidden, w_reg, b_reg, w_cls, b_cls, reg_out, cls_out, rows_max, rows_dev, nin,
                                                ncls);
}

}  // namespace frcnn
