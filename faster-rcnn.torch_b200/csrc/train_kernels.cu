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
// the same draws for every frame of a batch in one launch (blockIdx.y = frame, the index restarts at 0 in each frame)
__global__ void dropout_mask_frames_kernel(float* __restrict__ mask, FrameList fl, int per_row, float p, FrameSeeds seeds, uint32_t layer) {
  const int fr = blockIdx.y;
  const long n = (long)fl.R[fr] * per_row;
  float* m = mask + (long)fl.off[fr] * per_row;
  const uint64_t seed = seeds.seed[fr];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float u = (mix32(seed * 0x9E3779B97F4A7C15ull + ((uint64_t)layer << 32) + (uint64_t)i) >> 8) * (1.0f / 16777216.0f);
    m[i] = u >= p ? 1.0f : 0.0f;
  }
}
void launch_dropout_mask_frames(float* mask, const FrameList& fl, int per_row, float p, const FrameSeeds& seeds, uint32_t layer,
                                cudaStream_t st) {
  const long n = (long)fl.max_R() * per_row;
  if (n <= 0) return;
  dropout_mask_frames_kernel<<<dim3((unsigned)std::min<long>(cdiv(n, 256), 4096), fl.nf), 256, 0, st>>>(mask, fl, per_row, p, seeds, layer);
}

// ------------------------------------------------------------------------------------------ pooled conv backward
// Backward of [PReLU -> SpatialDropout mask -> MaxPool 2x2 ceil] given the gradient wrt the POOLED output: only the
// winner of every window receives gradient; its pre-activation sign is the sign of the pooled value.
// g: fp32 [N][Hp][Wp][C]; arg: winners; yp: pooled activations; dpre: bf16 [N][H][W][C] (every element written).
// thread <-> (pooled pixel, 8 channels).  I = index type: 32-bit whenever the item count allows (a 64-bit division costs
// ~100 instructions; three of them per 16-byte item made these streaming kernels issue-bound).
template <typename I>
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
  const I total = (I)N * Hp * Wp * cv;
  float ds = 0.f;
  // the launcher makes the grid stride a multiple of C / 8, so a thread keeps its 8 channels over all iterations and
  // the bias-gradient partials live in registers (one shared-memory atomic per channel and thread at the end, not
  // one per element: the shared atomics paced this kernel)
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int c8_fixed = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) % cv);
  for (I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c8 = c8_fixed;
    I r = i / (I)cv;
    const int pw = (int)(r % (I)Wp);
    r /= (I)Wp;
    const int ph = (int)(r % (I)Hp);
    const int n = (int)(r / (I)Hp);
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
  if (total + (long)blocks * 256 < 0x7fffffffL)
    unpool_prelu_bwd_kernel<int><<<blocks, 256, C * sizeof(float), st>>>(g, arg, yp, slope, mask, dpre, dbias, dslope, N, H, W, C);
  else
    unpool_prelu_bwd_kernel<long><<<blocks, 256, C * sizeof(float), st>>>(g, arg, yp, slope, mask, dpre, dbias, dslope, N, H, W, C);
}

// ------------------------------------------------------------------------------------------ plain conv backward
// Backward of [PReLU -> mask] for a conv that is not followed by the pool: dpre = dy * mask * (y < 0 ? slope : 1), from
// the bf16 gradient tensor produced by the next conv's dgrad (d) into the tensor the previous conv's backward reads
// (out; may alias d).  thread <-> (pixel, 8 channels).
template <typename I>
__global__ void __launch_bounds__(256) prelu_bwd_kernel(const bf16* d, bf16* out, const bf16* __restrict__ y, const float* __restrict__ slope_p,
                                                        const float* __restrict__ mask, float* __restrict__ dbias,
                                                        float* __restrict__ dslope, long npix_per_img, int N, int C) {
  extern __shared__ float sb[];
  __shared__ float s_ds[8];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = 0.f;
  __syncthreads();
  const int cv = C >> 3;
  const float slope = slope_p[0];
  const float inv_slope = 1.0f / slope;
  const I total = (I)N * (I)npix_per_img * cv;
  float ds = 0.f;
  float bacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // see unpool_prelu_bwd_kernel
  const int c8_fixed = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) % cv);
  for (I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c8 = c8_fixed;
    const int n = mask ? (int)((i / (I)cv) / (I)npix_per_img) : 0;   // only the dropout mask is per frame
    uint4 d8 = reinterpret_cast<const uint4*>(d)[i];
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
    reinterpret_cast<uint4*>(out)[i] = d8;
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
void launch_prelu_bwd(const bf16* d, bf16* out, const bf16* y, const float* slope, const float* mask, float* dbias, float* dslope, int N, int H, int W,
                      int C, int num_sms, cudaStream_t st) {
  const long total = (long)N * H * W * (C / 8);
  const int blocks = channel_stable_grid(total, C / 8, num_sms);
  if (total + (long)blocks * 256 < 0x7fffffffL)
    prelu_bwd_kernel<int><<<blocks, 256, C * sizeof(float), st>>>(d, out, y, slope, mask, dbias, dslope, (long)H * W, N, C);
  else
    prelu_bwd_kernel<long><<<blocks, 256, C * sizeof(float), st>>>(d, out, y, slope, mask, dbias, dslope, (long)H * W, N, C);
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
// The same backward on a pixel list (sparse form).  Every listed pixel carries gradient, so the per-element atomics of
// the kernel above (18 x 256 addresses hit by every pixel) would serialise in L2: here thread <-> channel keeps its
// column of dw2, its bias and slope partials in registers over the CTA's pixels and issues one atomic per address at
// the end.  Per pixel the arithmetic (and its order) is the dense kernel's: adding the zero terms it skips changes nothing.
__global__ void __launch_bounds__(256) head_tail_bwd_list_kernel(HeadTailBwd H) {
  constexpr int CO = 18, CM = 256;
  __shared__ float s_dout[CO];
  __shared__ float s_red[CM / 32];
  const int c = threadIdx.x;
  float w2c[CO], acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) { w2c[o] = H.w2[o * CM + c]; acc[o] = 0.f; }
  const float bias = H.bias[c], slope = H.prelu[0];
  float db1 = 0.f, db2 = 0.f, ds = 0.f;
  for (int it = blockIdx.x; it < H.M; it += gridDim.x) {
    const long pix = H.list[it];
    const long n = pix / H.HW, hw = pix - n * H.HW;
    __syncthreads();   // the previous pixel's readers of s_dout are done
    if (c < CO) {
      const float dv = H.d_out[(n * CO + c) * H.HW + hw];
      s_dout[c] = dv;
      db2 += dv;
    }
    __syncthreads();
    float h = bias;
    if (H.ws_compact) h += H.ws[(long)it * CM + c];
    else
      for (int s = 0; s < H.splits; ++s) h += H.ws[(size_t)s * H.slice_stride + pix * CM + c];
    const float a = h > 0.f ? h : h * slope;
    float dA = 0.f;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      const float d = s_dout[o];
      dA += d * w2c[o];
      acc[o] += d * a;
    }
    const float dp = h > 0.f ? dA : dA * slope;
    if (!(h > 0.f)) ds += dA * h;
    db1 += dp;
    H.dpre[(long)it * CM + c] = __float2bfloat16_rn(dp);
  }
#pragma unroll
  for (int o = 0; o < CO; ++o)
    if (acc[o] != 0.f) atomicAdd(H.dw2 + o * CM + c, acc[o]);
  if (db1 != 0.f) atomicAdd(H.db1 + c, db1);
  if (c < CO && db2 != 0.f) atomicAdd(H.db2 + c, db2);
  ds = warp_sum(ds);
  if ((c & 31) == 0) s_red[c >> 5] = ds;
  __syncthreads();
  if (c == 0) {
    float t = 0.f;
    for (int i = 0; i < CM / 32; ++i) t += s_red[i];
    if (t != 0.f) atomicAdd(H.dslope, t);
  }
}
void launch_head_tail_bwd(const HeadTailBwd& H, int num_sms, cudaStream_t st) {
  if (H.list) {
    if (H.M > 0) head_tail_bwd_list_kernel<<<std::min(H.M, 4 * num_sms), 256, 0, st>>>(H);
    return;
  }
  const int smem = 19 * 256 * sizeof(float);
  if (H.npix <= 0) return;
  const int blocks = (int)std::min<long>(cdiv(H.npix, 8), (long)num_sms * 4);
  head_tail_bwd_kernel<<<blocks, 256, smem, st>>>(H);
}

// ------------------------------------------------------------------------------------------ sparse anchor-head passes
// CTA <-> listed pixel, thread <-> channel: 18 dot products over the 256 activations, reduced warp -> CTA
__global__ void __launch_bounds__(256) head_tail_list_kernel(HeadTailList H) {
  constexpr int CO = 18, CM = 256;
  __shared__ float s_part[CM / 32][CO];
  const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
  const float bias = H.bias[c], slope = H.prelu[0];
  for (int it = blockIdx.x; it < H.M; it += gridDim.x) {
    const long pix = H.list[it];
    const long n = pix / H.HW, hw = pix - n * H.HW;
    const float h = bias + H.hpre[(long)it * CM + c];
    const float a = h > 0.f ? h : h * slope;
    __syncthreads();   // the previous pixel's partial sums have been consumed
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      const float v = warp_sum(a * H.w2[o * CM + c]);
      if (lane == 0) s_part[warp][o] = v;
    }
    __syncthreads();
    if (c < CO) {
      float v = H.b2[c];
      for (int w = 0; w < CM / 32; ++w) v += s_part[w][c];
      H.out[(n * CO + c) * H.HW + hw] = v;
    }
  }
}
void launch_head_tail_list(const HeadTailList& H, cudaStream_t st) {
  if (H.M > 0) head_tail_list_kernel<<<H.M, 256, 0, st>>>(H);
}
__global__ void pack_head_weight_rows_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int taps) {
  const long total = (long)Cout * Cin * taps;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long r = i / Cout;
    const int ci = (int)(r % Cin), t = (int)(r / Cin);
    out[i] = __float2bfloat16_rn(w[((long)co * Cin + ci) * taps + t]);   // Torch layout [co][ci][kh][kw]
  }
}
void launch_pack_head_weight_rows(const float* w, bf16* out, int Cout, int Cin, int K, cudaStream_t st) {
  const long total = (long)Cout * Cin * K * K;
  pack_head_weight_rows_kernel<<<(int)std::min<long>(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, Cout, Cin, K * K);
}
// item <-> (listed pixel, filter tap, 8 channels): one 16-byte copy
template <typename I>
__global__ void head_gather_rows_kernel(const bf16* __restrict__ x, const int* __restrict__ list, int M, int hh, int hw, int Hin, int Win,
                                        int Cin, int K, bf16* __restrict__ rows) {
  const int cv = Cin >> 3, taps = K * K;
  const I total = (I)M * taps * cv;
  for (I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c8 = (int)(i % (I)cv);
    const I q = i / (I)cv;
    const int t = (int)(q % (I)taps), r = (int)(q / (I)taps);
    const int pix = list[r];
    const int n = pix / (hh * hw), rem = pix - n * hh * hw;
    const int y = rem / hw + t / K, xx = rem % hw + t % K;
    const uint4 v = *reinterpret_cast<const uint4*>(x + (((long)n * Hin + y) * Win + xx) * Cin + c8 * 8);
    *reinterpret_cast<uint4*>(rows + ((long)r * taps + t) * Cin + c8 * 8) = v;
  }
}
void launch_head_gather_rows(const bf16* x, const int* list, int M, int hh, int hw, int Hin, int Win, int Cin, int K, bf16* rows,
                             cudaStream_t st) {
  const long total = (long)M * K * K * (Cin / 8);
  if (total <= 0) return;
  const int blocks = (int)std::min<long>(cdiv(total, 256), 148 * 16);
  if (total + (long)blocks * 256 < 0x7fffffffL) head_gather_rows_kernel<int><<<blocks, 256, 0, st>>>(x, list, M, hh, hw, Hin, Win, Cin, K, rows);
  else head_gather_rows_kernel<long><<<blocks, 256, 0, st>>>(x, list, M, hh, hw, Hin, Win, Cin, K, rows);
}
// item <-> (listed pixel, filter tap, 4 channels): one 16-byte vector reduction (windows of listed pixels overlap)
template <typename I>
__global__ void head_scatter_rows_kernel(const float* __restrict__ g, const int* __restrict__ list, int M, int hh, int hw, int Hin,
                                         int Win, int Cin, int K, float* __restrict__ dx) {
  const int cv = Cin >> 2, taps = K * K;
  const I total = (I)M * taps * cv;
  for (I i = (I)blockIdx.x * (I)blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * (I)blockDim.x) {
    const int c4 = (int)(i % (I)cv);
    const I q = i / (I)cv;
    const int t = (int)(q % (I)taps), r = (int)(q / (I)taps);
    const int pix = list[r];
    const int n = pix / (hh * hw), rem = pix - n * hh * hw;
    const int y = rem / hw + t / K, xx = rem % hw + t % K;
    const float4 v = *reinterpret_cast<const float4*>(g + ((long)r * taps + t) * Cin + c4 * 4);
    atomicAdd(reinterpret_cast<float4*>(dx + (((long)n * Hin + y) * Win + xx) * Cin + c4 * 4), v);
  }
}
void launch_head_scatter_rows(const float* g, const int* list, int M, int hh, int hw, int Hin, int Win, int Cin, int K, float* dx,
                              cudaStream_t st) {
  const long total = (long)M * K * K * (Cin / 4);
  if (total <= 0) return;
  const int blocks = (int)std::min<long>(cdiv(total, 256), 148 * 16);
  if (total + (long)blocks * 256 < 0x7fffffffL) head_scatter_rows_kernel<int><<<blocks, 256, 0, st>>>(g, list, M, hh, hw, Hin, Win, Cin, K, dx);
  else head_scatter_rows_kernel<long><<<blocks, 256, 0, st>>>(g, list, M, hh, hw, Hin, Win, Cin, K, dx);
}

// ------------------------------------------------------------------------------------------ first-layer wgrad
// Ampere-style asynchronous copies (LDGSTS): src_bytes = 0 zero-fills the destination, which is how the out-of-frame
// pixels of a border tile become zeros without a branch around the copy.
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp-level tensor-core pieces of the first-layer weight gradient: the GEMM is [64 co] x [27 -> 32 taps] with K = pixels,
// far too narrow in M and N for a tcgen05 tile to pay for its TMEM round trip, and HBM-bound (one read of dY) once the
// products leave the fp32 pipe.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two fp32 values -> a bf16 pair (hi) and the bf16 pair of what the rounding dropped (lo): hi + lo carries 16 mantissa
// bits of the frame, so the products match the fp32 FMA version to ~2^-17 instead of 2^-9
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __uint_as_float(hi << 16), v1 - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int FW_CO = 64, FW_TH = 8, FW_TW = 32;
constexpr int FW_D_BYTES = FW_TH * FW_TW * FW_CO * 2;                      // 32 KB gradient tile
constexpr int FW_PITCH = 40, FW_CPITCH = (FW_TH + 2) * FW_PITCH;           // frame patch: [3][10][40] floats, 34 columns used
constexpr int FW_IMG_STRIDE = 1280;                                        // floats per stage (3 * 400 rounded up)
constexpr int FW_SMEM = 2 * FW_D_BYTES + 2 * FW_IMG_STRIDE * (int)sizeof(float);   // two stages: 74 KB, two CTAs per SM

// dW[co][c][kh][kw] += sum over pixels of dpre[n][h][w][co] * img[n][c][h + kh - pad][w + kw - pad] for the 3-channel
// first convolution.  Persistent CTAs over 8 x 32 pixel tiles, two shared-memory stages filled by cp.async (the next
// tile lands while this one is consumed).  Warp w owns tile row w = two K-steps of 16 pixels: A = dY^T fragments come
// out of the [px][co] tile with ldmatrix.trans (16-byte chunks XOR-swizzled by px & 7, so the eight row addresses of a
// matrix hit eight different bank groups), B = the frame at 32 tap offsets, split hi/lo in registers; 64 co x 32 taps
// accumulate in 64 registers per thread over all the CTA's tiles, then warps -> shared -> one global atomic per entry.
__global__ void __launch_bounds__(256, 2) first_wgrad_kernel(const bf16* __restrict__ dpre, const float* __restrict__ img,
                                                             float* __restrict__ dw, int N, int H, int W, int pad) {
  constexpr int CO = FW_CO, TAPS = 27, TH = FW_TH, TW = FW_TW;
  extern __shared__ __align__(128) unsigned char fw_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  int toff[4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int tap = min(nt * 8 + g, TAPS - 1);     // taps 27..31 are padding: computed on a valid address, never stored
    toff[nt] = (tap / 9) * FW_CPITCH + ((tap % 9) / 3) * FW_PITCH + tap % 3 + 2 * t;
  }
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][nt][e] = 0.f;
  const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH;
  const int total_tiles = N * tiles_h * tiles_w;
  const long plane = (long)H * W;

  auto stage = [&](int tile, int b) {
    const int tw = tile % tiles_w;
    const int r = tile / tiles_w;
    const int th = r % tiles_h, n = r / tiles_h;
    const int h0 = th * TH, w0 = tw * TW;
    unsigned char* sd = fw_smem + b * FW_D_BYTES;
    float* si = reinterpret_cast<float*>(fw_smem + 2 * FW_D_BYTES) + b * FW_IMG_STRIDE;
    // gradient tile: 256 px x 128 B, 16-byte vectors
    for (int i = threadIdx.x; i < TH * TW * 8; i += 256) {
      const int px = i >> 3, c8 = i & 7;
      const int h = h0 + px / TW, w = w0 + px % TW;
      const bool in = h < H && w < W;
      cp_async_16(sd + px * (CO * 2) + ((c8 ^ (px & 7)) << 4), in ? dpre + (((long)n * H + h) * W + w) * CO + c8 * 8 : dpre, in ? 16 : 0);
    }
    for (int i = threadIdx.x; i < 3 * (TH + 2) * (TW + 2); i += 256) {
      const int c = i / ((TH + 2) * (TW + 2)), rr = i % ((TH + 2) * (TW + 2));
      const int py = rr / (TW + 2), pxx = rr % (TW + 2);
      const int yy = h0 + py - pad, xx = w0 + pxx - pad;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      cp_async_4(si + c * FW_CPITCH + py * FW_PITCH + pxx, in ? img + ((long)n * 3 + c) * plane + (long)yy * W + xx : img, in ? 4 : 0);
    }
    cp_async_commit();
  };

  // ldmatrix row of this lane: matrix j = lane / 8 covers pixels (j / 2) * 8 .. + 7 and channel chunk 2 m + (j & 1)
  const int a_px = ((lane >> 4) << 3) + (lane & 7), a_ch = (lane >> 3) & 1, a_sw = lane & 7;
  int b = 0;
  if ((int)blockIdx.x < total_tiles) stage(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, b ^= 1) {
    const int nxt = tile + gridDim.x;
    if (nxt < total_tiles) { stage(nxt, b ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const uint32_t sd = (uint32_t)__cvta_generic_to_shared(fw_smem + b * FW_D_BYTES);
    const float* simg = reinterpret_cast<const float*>(fw_smem + 2 * FW_D_BYTES) + b * FW_IMG_STRIDE + warp * FW_PITCH;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float* s = simg + half * 16 + toff[nt];     // B[k][n]: k = 2t, 2t+1 (reg 0) and 2t+8, 2t+9 (reg 1), n = g
        split_bf16x2(s[0], s[1], bh[nt][0], bl[nt][0]);
        split_bf16x2(s[8], s[9], bh[nt][1], bl[nt][1]);
      }
      const uint32_t row = sd + (warp * 32 + half * 16 + a_px) * (CO * 2);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        ldmatrix_x4_trans(a, row + (((2 * m + a_ch) ^ a_sw) << 4));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_m16n8k16_bf16(acc[m][nt], a, bh[nt][0], bh[nt][1]);
          mma_m16n8k16_bf16(acc[m][nt], a, bl[nt][0], bl[nt][1]);
        }
      }
    }
    __syncthreads();   // stage b is refilled by the next iteration's copies
  }
  // D fragment: rows co = 16 m + g (+8), columns tap = 8 nt + 2t (+1)
  float* s_out = reinterpret_cast<float*>(fw_smem);
  for (int i = threadIdx.x; i < CO * 32; i += 256) s_out[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = m * 16 + g + (e >> 1) * 8, tap = nt * 8 + 2 * t + (e & 1);
        if (tap < TAPS && acc[m][nt][e] != 0.f) atomicAdd(&s_out[co * 32 + tap], acc[m][nt][e]);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * TAPS; i += 256) {   // Torch layout [co][c][kh][kw]
    const float v = s_out[(i / TAPS) * 32 + i % TAPS];
    if (v != 0.f) atomicAdd(dw + i, v);
  }
}
void launch_first_wgrad(const bf16* dpre, const float* img, float* dw, int N, int H, int W, int pad, int num_sms, cudaStream_t st) {
  static DeviceOnce configured;   // function attributes are per device: one process may drive several (frcnn_dp_init_all)
  if (first_use_on_device(configured)) cudaFuncSetAttribute(first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FW_SMEM);
  const long tiles = (long)N * ((H + 7) / 8) * ((W + 31) / 32);
  // the kernel indexes tiles in 32 bits: 2^31 tiles would be a 70 TB gradient tensor
  const int grid = (int)std::min<long>(tiles, (long)num_sms * 2);
  first_wgrad_kernel<<<grid, 256, FW_SMEM, st>>>(dpre, img, dw, N, H, W, pad);
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
