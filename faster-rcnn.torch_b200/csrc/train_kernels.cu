// Elementwise / reduction kernels of pnet:backward (objective.lua:189): everything between the tensor-core dgrad /
// wgrad GEMMs.  All are HBM-bound streaming kernels (16-byte vectors, coalesced); the per-channel / scalar parameter
// gradients are reduced in shared memory per CTA and added to the fp32 gradient buffers with one atomic per CTA and
// element (summation order therefore not reproducible run to run, like cunn's own accumulation).
//
// PReLU backward uses the sign of the STORED activation y = PReLU(x) * mask as the sign of x, which is exact for
// slopes > 0 (Torch initialises 0.25); a non-positive slope is reported by train_check_slopes.
#include "common.h"
#include "train.h"

namespace frcnn {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// ------------------------------------------------------------------------------------------ SpatialDropout masks
// nn.SpatialDropout(p) v1 at training time: one Bernoulli(1 - p) draw per (image, channel), no rescale.
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
__global__ void dropout_mask_kernel(float* __restrict__ mask, int n, float p, uint64_t seed, uint32_t layer) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = (mix32(seed * 0x9E3779B97F4A7C15ull + ((uint64_t)layer << 32) + (uint64_t)i) >> 8) * (1.0f / 16777216.0f);
  mask[i] = u >= p ? 1.0f : 0.0f;
}
void launch_dropout_mask(float* mask, int n, float p, uint64_t seed, uint32_t layer, cudaStream_t st) {
  dropout_mask_kernel<<<cdiv(n, 256), 256, 0, st>>>(mask, n, p, seed, layer);
}

// ------------------------------------------------------------------------------------------ pooled conv backward
// Backward of [PReLU -> SpatialDropout mask -> MaxPool 2x2 ceil] given the gradient wrt the POOLED output: only the
// winner of every window receives gradient; its pre-activation sign is the sign of the pooled value.
// g: fp32 [N][Hp][Wp][C]; arg: winners; yp: pooled activations; dpre: bf16 [N][H][W][C] (every element written).
// thread <-> (pooled pixel, 8 channels).
__global__ void __launch_bounds__(256) unpool_prelu_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ arg,
                                                               const bf16* __restrict__ yp, const float* __restrict__ slope_p,
                                                               const float* __restrict__ mask, bf16* __restrict__ dpre,
                                                               float* __restrict__ dbias, float* __restrict__ dslope, int N, int H,
                                                               int W, int C) {
  extern __shared__ float sb[];  // [C] per-CTA bias-gradient partials
  __shared__ float s_ds[8];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = 0.f;
  __syncthreads();
  const int Hp = (H + 1) >> 1, Wp = (W + 1) >> 1, cv = C >> 3;
  const float slope = slope_p[0];
  const float inv_slope = 1.0f / slope;
  const long total = (long)N * Hp * Wp * cv;
  float ds = 0.f;
  // the launcher makes the grid stride a multiple of C / 8, so a thread keeps its 8 channels over all iterations and
  // the bias-gradient partials live in registers (one shared-memory atomic per channel and thread at the end, not
  // one per element: the shared atomics paced this kernel)
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int c8_fixed = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) % cv);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = c8_fixed;
    long r = i / cv;
    const int pw = (int)(r % Wp);
    r /= Wp;
    const int ph = (int)(r % Hp);
    const int n = (int)(r / Hp);
    const long pbase = (((long)n * Hp + ph) * Wp + pw) * C + c8 * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(g + pbase), g1 = *reinterpret_cast<const float4*>(g + pbase + 4);
    const uint2 a8 = *reinterpret_cast<const uint2*>(arg + pbase);
    const uint4 y8 = *reinterpret_cast<const uint4*>(yp + pbase);
    const bf16* ye = reinterpret_cast<const bf16*>(&y8);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float val[8];
    uint32_t win[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float m = mask ? mask[(long)n * C + c8 * 8 + e] : 1.0f;
      const float y = __bfloat162float(ye[e]);
      const float gy = gv[e] * m;
      const bool neg = y < 0.f;
      val[e] = neg ? gy * slope : gy;
      if (neg) ds += gy * (y * inv_slope);  // d y / d slope = x * mask, x = y / (slope * mask)
      win[e] = ((e < 4 ? a8.x : a8.y) >> (8 * (e & 3))) & 3u;
      bacc[e] += val[e];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int h = 2 * ph + (q >> 1), w = 2 * pw + (q & 1);
      if (h < H && w < W) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = win[2 * j] == (uint32_t)q ? val[2 * j] : 0.f, hi = win[2 * j + 1] == (uint32_t)q ? val[2 * j + 1] : 0.f;
          __nv_bfloat162 t2 = __floats2bfloat162_rn(lo, hi);
          o[j] = *reinterpret_cast<uint32_t*>(&t2);
        }
        *reinterpret_cast<uint4*>(dpre + (((long)n * H + h) * W + w) * C + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (bacc[e] != 0.f) atomicAdd(&sb[c8_fixed * 8 + e], bacc[e]);
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) s_ds[threadIdx.x >> 5] = ds;
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    if (sb[i] != 0.f) atomicAdd(dbias + i, sb[i]);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_ds[w];
    if (t != 0.f) atomicAdd(dslope, t);
  }
}
// grid such that (grid * 256) % (C / 8) == 0: every thread then owns one 8-channel group for the whole launch
static int channel_stable_grid(long total, int cv, int num_sms) {
  int g = 1, a = cv, b = 256;
  while (b) { const int t = a % b; a = b; b = t; }   // a = gcd(cv, 256)
  g = cv / a;
  long blocks = std::min<long>(cdiv(total, 256), (long)num_sms * 8);
  blocks = std::max<long>(g, blocks / g * g);
  return (int)blocks;
}
void launch_unpool_prelu_bwd(const float* g, const uint8_t* arg, const bf16* yp, const float* slope, const float* mask, bf16* dpre,
                             float* dbias, float* dslope, int N, int H, int W, int C, int num_sms, cudaStream_t st) {
  const long total = (long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  const int blocks = channel_stable_grid(total, C / 8, num_sms);
  unpool_prelu_bwd_kernel<<<blocks, 256, C * sizeof(float), st>>>(g, arg, yp, slope, mask, dpre, dbias, dslope, N, H, W, C);
}

// ------------------------------------------------------------------------------------------ plain conv backward
// Backward of [PReLU -> mask] for a conv that is not followed by the pool: dpre = dy * mask * (y < 0 ? slope : 1), in
// place on the bf16 gradient tensor produced by the next conv's dgrad.  thread <-> (pixel, 8 channels).
__global__ void __launch_bounds__(256) prelu_bwd_kernel(bf16* __restrict__ d, const bf16* __restrict__ y, const float* __restrict__ slope_p,
                                                        const float* __restrict__ mask, float* __restrict__ dbias,
                                                        float* __restrict__ dslope, long npix_per_img, int N, int C) {
  extern __shared__ float sb[];
  __shared__ float s_ds[8];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = 0.f;
  __syncthreads();
  const int cv = C >> 3;
  const float slope = slope_p[0];
  const float inv_slope = 1.0f / slope;
  const long total = (long)N * npix_per_img * cv;
  float ds = 0.f;
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // see unpool_prelu_bwd_kernel
  const int c8_fixed = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) % cv);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c8 = c8_fixed;
    const long pix = i / cv;
    const int n = (int)(pix / npix_per_img);
    uint4 d8 = reinterpret_cast<uint4*>(d)[i];
    const uint4 y8 = reinterpret_cast<const uint4*>(y)[i];
    bf16* de = reinterpret_cast<bf16*>(&d8);
    const bf16* ye = reinterpret_cast<const bf16*>(&y8);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float m = mask ? mask[(long)n * C + c8 * 8 + e] : 1.0f;
      const float yv = __bfloat162float(ye[e]);
      const float gy = __bfloat162float(de[e]) * m;
      const bool neg = yv < 0.f;
      const float v = neg ? gy * slope : gy;
      if (neg) ds += gy * (yv * inv_slope);
      bacc[e] += v;
      de[e] = __float2bfloat16_rn(v);
    }
    reinterpret_cast<uint4*>(d)[i] = d8;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (bacc[e] != 0.f) atomicAdd(&sb[c8_fixed * 8 + e], bacc[e]);
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) s_ds[threadIdx.x >> 5] = ds;
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    if (sb[i] != 0.f) atomicAdd(dbias + i, sb[i]);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_ds[w];
    if (t != 0.f) atomicAdd(dslope, t);
  }
}
void launch_prelu_bwd(bf16* d, const bf16* y, const float* slope, const float* mask, float* dbias, float* dslope, int N, int H, int W,
                      int C, int num_sms, cudaStream_t st) {
  const long total = (long)N * H * W * (C / 8);
  const int blocks = channel_stable_grid(total, C / 8, num_sms);
  prelu_bwd_kernel<<<blocks, 256, C * sizeof(float), st>>>(d, y, slope, mask, dbias, dslope, (long)H * W, N, C);
}

// ------------------------------------------------------------------------------------------ anchor-head tail backward
// Backward of the AnchorNetwork tail (model_utilities.lua:32-33): d_out [N][18][HW] fp32 (the dense, mostly zero
// delta_outputs of objective.lua:78-84) -> gradient wrt the k x k conv's pre-activation output [N*HW][256] bf16 plus
// the parameter gradients of the 1x1 conv (w2, b2), the PReLU slope and the k x k conv's bias.  One warp per pixel;
// pixels whose 18 deltas are all zero (almost all of them) only write zeros.
__global__ void __launch_bounds__(256) head_tail_bwd_kernel(HeadTailBwd H) {
  constexpr int CO = 18, CM = 256;
  extern __shared__ float sw[];  // [18][256] w2, [256] bias
  float* sbias = sw + CO * CM;
  for (int i = threadIdx.x; i < CO * CM; i += blockDim.x) sw[i] = H.w2[i];
  for (int i = threadIdx.x; i < CM; i += blockDim.x) sbias[i] = H.bias[i];
  __syncthreads();
  const float slope = H.prelu[0];
  const int lane = threadIdx.x & 31;
  const long warp_idx = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long pix = warp_idx; pix < H.npix; pix += nwarps) {
    const long n = pix / H.HW, hw = pix - n * H.HW;
    float dv = lane < CO ? H.d_out[(n * CO + lane) * H.HW + hw] : 0.f;
    const unsigned nz = __ballot_sync(0xffffffffu, dv != 0.f);
    uint4* dst = reinterpret_cast<uint4*>(H.dpre + pix * CM) + lane;  // lane owns channels lane*8 .. +7
    if (nz == 0) {
      *dst = make_uint4(0, 0, 0, 0);
      continue;
    }
    float dout[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) dout[o] = __shfl_sync(0xffffffffu, dv, o);
    if (lane < CO) atomicAdd(H.db2 + lane, dv);
    float ds = 0.f;
    uint32_t packed[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = lane * 8 + e;
      float h = sbias[c];
      for (int s = 0; s < H.splits; ++s) h += H.ws[(size_t)s * H.slice_stride + pix * CM + c];
      // note: summation order bias-first differs from the forward (slices first) by fp32 rounding only
      const float a = h > 0.f ? h : h * slope;
      float dA = 0.f;
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        if (nz & (1u << o)) {
          dA += dout[o] * sw[o * CM + c];
          atomicAdd(H.dw2 + o * CM + c, dout[o] * a);
        }
      }
      const float dp = h > 0.f ? dA : dA * slope;
      if (!(h > 0.f)) ds += dA * h;
      atomicAdd(H.db1 + c, dp);
      const bf16 b = __float2bfloat16_rn(dp);
      const uint16_t bits = *reinterpret_cast<const uint16_t*>(&b);
      if (e & 1) packed[e >> 1] |= (uint32_t)bits << 16; else packed[e >> 1] = bits;
    }
    *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    ds = warp_sum(ds);
    if (lane == 0 && ds != 0.f) atomicAdd(H.dslope, ds);
  }
}
void launch_head_tail_bwd(const HeadTailBwd& H, int num_sms, cudaStream_t st) {
  const int smem = 19 * 256 * sizeof(float);
  const int blocks = (int)std::min<long>(cdiv(H.npix, 8), (long)num_sms * 4);
  head_tail_bwd_kernel<<<blocks, 256, smem, st>>>(H);
}

// ------------------------------------------------------------------------------------------ first-layer wgrad
// dW[co][c][kh][kw] += sum over pixels of dpre[n][h][w][co] * img[n][c][h + kh - pad][w + kw - pad] for the 3-channel
// first convolution (K = 27 is too narrow for the tensor-core wgrad).  Persistent CTAs over 8 x 32 pixel tiles: the
// gradient tile [256 px][64 co] and the (8+2) x (32+2) x 3 image patch are staged in shared memory, thread <-> (co,
// tap group) accumulates its 7 taps in registers across all its tiles, one atomic per accumulator at the end.
__global__ void __launch_bounds__(256) first_wgrad_kernel(const bf16* __restrict__ dpre, const float* __restrict__ img,
                                                          float* __restrict__ dw, int N, int H, int W, int pad) {
  // thread <-> (8 output channels, 7 of the 27 taps, one of 8 interleaved pixel subsets): per pixel ONE 16-byte
  // gradient read + 7 image reads feed 56 FMAs into a register tile (the one-channel-per-thread version issued
  // 8 shared-memory loads per 7 FMAs and ran at the LDS rate)
  constexpr int CO = 64, TAPS = 27, TG = 4, PER = 7, TH = 8, TW = 32, PS = 8;
  __shared__ __align__(16) bf16 s_d[TH * TW][CO];        // 32 KB
  __shared__ float s_img[3][TH + 2][TW + 2];
  const int cg = threadIdx.x & 7, tg = (threadIdx.x >> 3) & 3, ps = threadIdx.x >> 5;
  int t_off[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int tap = min(tg + j * TG, TAPS - 1);
    t_off[j] = ((tap / 9) * (TH + 2) + (tap % 9) / 3) * (TW + 2) + tap % 3;
  }
  float acc[8][PER];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < PER; ++j) acc[e][j] = 0.f;
  const float* simg = &s_img[0][0][0];
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
  const long total_tiles = (long)N * tiles_h * tiles_w;
  const long plane = (long)H * W;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const long r = tile / tiles_w;
    const int th = (int)(r % tiles_h), n = (int)(r / tiles_h);
    const int h0 = th * TH, w0 = tw * TW;
    __syncthreads();
    // gradient tile: 256 px x 128 B, 16-byte vectors
    for (int i = threadIdx.x; i < TH * TW * 8; i += 256) {
      const int px = i >> 3, c8 = i & 7;
      const int h = h0 + px / TW, w = w0 + px % TW;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (h < H && w < W) v = *reinterpret_cast<const uint4*>(dpre + (((long)n * H + h) * W + w) * CO + c8 * 8);
      *reinterpret_cast<uint4*>(&s_d[px][c8 * 8]) = v;
    }
    for (int i = threadIdx.x; i < 3 * (TH + 2) * (TW + 2); i += 256) {
      const int c = i / ((TH + 2) * (TW + 2)), rr = i % ((TH + 2) * (TW + 2));
      const int yy = h0 + rr / (TW + 2) - pad, xx = w0 + rr % (TW + 2) - pad;
      s_img[c][rr / (TW + 2)][rr % (TW + 2)] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + ((long)n * 3 + c) * plane + (long)yy * W + xx) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int px = ps; px < TH * TW; px += PS) {
      const uint4 raw = *reinterpret_cast<const uint4*>(&s_d[px][cg * 8]);
      const bf16* d8 = reinterpret_cast<const bf16*>(&raw);
      const int base = (px / TW) * (TW + 2) + (px % TW);
      float iv[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) iv[j] = simg[base + t_off[j]];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = __bfloat162float(d8[e]);
#pragma unroll
        for (int j = 0; j < PER; ++j) acc[e][j] = fmaf(d, iv[j], acc[e][j]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int tap = tg + j * TG;
      if (tap < TAPS && acc[e][j] != 0.f) atomicAdd(dw + (cg * 8 + e) * TAPS + tap, acc[e][j]);  // Torch layout [co][c][kh][kw]
    }
}
void launch_first_wgrad(const bf16* dpre, const float* img, float* dw, int N, int H, int W, int pad, int num_sms, cudaStream_t st) {
  const long tiles = (long)N * ((H + 7) / 8) * ((W + 31) / 32);
  const int grid = (int)std::min<long>(tiles, (long)num_sms * 4);
  first_wgrad_kernel<<<grid, 256, 0, st>>>(dpre, img, dw, N, H, W, pad);
}

// ------------------------------------------------------------------------------------------ misc
// dst (fp32 NHWC) += src (fp32 Torch layout [N][C][HW]): the ROI-pool gradient delta_outputs[5] (objective.lua:184)
__global__ void add_chw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int HW, int C) {
  __shared__ float tile[32][33];
  int n = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) tile[j][threadIdx.x] = src[((long)n * C + c) * HW + p];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) dst[((long)n * HW + p) * C + c] += tile[threadIdx.x][j];
  }
}
void launch_add_chw_to_nhwc(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st) {
  dim3 grid(cdiv((long)H * W, 32), cdiv(C, 32), N), block(32, 8);
  add_chw_to_nhwc_kernel<<<grid, block, 0, st>>>(src, dst, H * W, C);
}

// any PReLU slope <= 0 breaks the sign-of-y shortcut: flag it
__global__ void check_slopes_kernel(const float* const* slopes, int n, int* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(slopes[i][0] > 0.f)) atomicExch(flag, 1);
}
void launch_check_slopes(const float* const* slopes_dev, int n, int* flag, cudaStream_t st) {
  check_slopes_kernel<<<cdiv(n, 64), 64, 0, st>>>(slopes_dev, n, flag);
}

}  // namespace frcnn
