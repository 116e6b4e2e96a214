// Frame normalisation on the GPU: the tail of BatchIterator:processImage (BatchIterator.lua:146-161) and of
// load_image (utilities.lua:206-218) -- the step immediately before pnet:forward:
//   [image.rgb2yuv]                                   utilities.lua:211-212 (config color_space = 'yuv')
//   img[i] = img[i] - img[i]:mean()                   BatchIterator.lua:146-150 (normalization.centering)
//   s = img[i]:std(); if s > 1e-8: img[i] = img[i] / s   :152-159 (normalization.scaling; unbiased std)
//   img[1] = nn.SpatialContrastiveNormalization(1, image.gaussian1D(width)):forward(img[{{1}}])   :86,161
// `image` and `nn` are un-vendored Torch7 packages: their published algorithms are restated in
// oracle/preprocess.py, which is what this file is checked against (parity unpinned).  TH reduces means / variances
// in double; so do the two fixed-order reduction passes here.  The contrastive normalisation is two separable
// 7-tap passes with zero padding and the border coefficient map of nn.SpatialSubtractiveNormalization.
// HBM-bound: 4.3 MB frame, five passes (one fused read/write each).
#include "common.h"

namespace frcnn {

static constexpr int RED_BLOCKS = 148, RED_THREADS = 256;

// pass 1 / 2: per-channel sum of f(x) in double, fixed order: per-block partials, then one block folds them
template <int MODE>   // 0: sum(x)   1: sum((x - mean)^2) with x already centred in fp32 (the reference centres first)
__global__ void __launch_bounds__(RED_THREADS) channel_partial_kernel(const float* __restrict__ img, long plane, const double* __restrict__ mean,
                                                                      double* __restrict__ partial) {
  const int c = blockIdx.y;
  const float* p = img + (long)c * plane;
  const double mu = MODE == 1 ? mean[c] : 0.0;
  double s = 0.0;
  for (long i = blockIdx.x * (long)RED_THREADS + threadIdx.x; i < plane; i += (long)gridDim.x * RED_THREADS) {
    const double v = (double)p[i];
    s += MODE == 0 ? v : (v - mu) * (v - mu);
  }
  __shared__ double sh[RED_THREADS];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[c * gridDim.x + blockIdx.x] = sh[0];
}
// stats[c] = {mean of the channel (double), 1/divisor to apply} ; MODE 0 finalises the mean, MODE 1 the std
template <int MODE>
__global__ void channel_final_kernel(const double* __restrict__ partial, int nblocks, long plane, double* __restrict__ stat) {
  const int c = blockIdx.x;
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < nblocks; ++i) s += partial[c * nblocks + i];
  if (MODE == 0) stat[c] = s / (double)plane;                       // THTensor meanall: double accumulation
  else stat[c] = sqrt(s / (double)(plane - 1));                     // THTensor stdall (unbiased)
}

// rgb -> yuv (image.rgb2yuv) in place, then x - mean; or x / std
__global__ void rgb2yuv_kernel(float* __restrict__ img, long plane) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < plane; i += (long)gridDim.x * blockDim.x) {
    const float r = img[i], g = img[plane + i], b = img[2 * plane + i];
    // image.rgb2yuv: y = 0.299 r + 0.587 g + 0.114 b; u = -0.14713 r - 0.28886 g + 0.436 b; v = 0.615 r - 0.51499 g - 0.10001 b
    img[i] = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
    img[plane + i] = __fadd_rn(__fadd_rn(__fmul_rn(-0.14713f, r), __fmul_rn(-0.28886f, g)), __fmul_rn(0.436f, b));
    img[2 * plane + i] = __fadd_rn(__fadd_rn(__fmul_rn(0.615f, r), __fmul_rn(-0.51499f, g)), __fmul_rn(-0.10001f, b));
  }
}
template <int MODE>   // 0: x - (float)mean   1: x / (float)std when std > 1e-8
__global__ void channel_apply_kernel(float* __restrict__ img, long plane, const double* __restrict__ stat) {
  const int c = blockIdx.y;
  float* p = img + (long)c * plane;
  const double sd = stat[c];
  if (MODE == 1 && !(sd > 1e-8)) return;
  const float v = (float)sd;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < plane; i += (long)gridDim.x * blockDim.x)
    p[i] = MODE == 0 ? __fsub_rn(p[i], v) : __fdiv_rn(p[i], v);
}

// One stage of nn.SpatialContrastiveNormalization on a single plane, 32 x 8 output tile per CTA with a halo of R:
//   STAGE 0 (SpatialSubtractiveNormalization): out = x - (K * x) / coef
//   STAGE 1 (SpatialDivisiveNormalization):    out = x / max'(sqrt(K * x^2) / coef), max'(s) = s > thr ? s : thr
// K = separable kernel k (x) k, zero padding; coef = K * ones (the border attenuation map) = ch(x) * cv(y).
static constexpr int CN_TW = 32, CN_TH = 8, CN_MAXR = 8;
template <int STAGE>
__global__ void __launch_bounds__(CN_TW * CN_TH) contrastive_stage_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                                          const float* __restrict__ k1d, int ksize, float thr) {
  __shared__ float tile[CN_TH + 2 * CN_MAXR][CN_TW + 2 * CN_MAXR];
  __shared__ float hrow[CN_TH + 2 * CN_MAXR][CN_TW];
  __shared__ float kk[2 * CN_MAXR + 1];
  const int R = ksize / 2;
  const int tx = threadIdx.x % CN_TW, ty = threadIdx.x / CN_TW;
  const int x0 = blockIdx.x * CN_TW, y0 = blockIdx.y * CN_TH;
  if ((int)threadIdx.x < ksize) kk[threadIdx.x] = k1d[threadIdx.x];
  for (int i = threadIdx.x; i < (CN_TH + 2 * R) * (CN_TW + 2 * R); i += CN_TW * CN_TH) {
    const int yy = i / (CN_TW + 2 * R), xx = i % (CN_TW + 2 * R);
    const int gy = y0 + yy - R, gx = x0 + xx - R;
    float v = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? in[(long)gy * W + gx] : 0.f;
    tile[yy][xx] = STAGE == 1 ? v * v : v;
  }
  __syncthreads();
  // horizontal pass over every row of the tile (incl. the vertical halo), fixed tap order
  for (int i = threadIdx.x; i < (CN_TH + 2 * R) * CN_TW; i += CN_TW * CN_TH) {
    const int yy = i / CN_TW, xx = i % CN_TW;
    float s = 0.f;
    for (int j = 0; j < ksize; ++j) s = fmaf(kk[j], tile[yy][xx + j], s);
    hrow[yy][xx] = s;
  }
  __syncthreads();
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx < W && gy < H) {
    float s = 0.f, ch = 0.f, cv = 0.f;
    for (int j = 0; j < ksize; ++j) {
      s = fmaf(kk[j], hrow[ty + j][tx], s);
      const int xx = gx + j - R, yy = gy + j - R;
      if (xx >= 0 && xx < W) ch += kk[j];
      if (yy >= 0 && yy < H) cv += kk[j];
    }
    const float coef = ch * cv;
    const float x = in[(long)gy * W + gx];
    if (STAGE == 0) {
      out[(long)gy * W + gx] = x - s / coef;
    } else {
      const float sd = sqrtf(s) / coef;
      out[(long)gy * W + gx] = x / (sd > thr ? sd : thr);
    }
  }
}

void launch_normalize_frame(float* img, int H, int W, int rgb2yuv, int centering, int scaling, const float* k1d_dev, int ksize,
                            float threshold, double* scratch /* >= 3 * RED_BLOCKS + 8 doubles */, float* plane_tmp, cudaStream_t st) {
  const long plane = (long)H * W;
  double* partial = scratch;
  double* stat = scratch + 3 * RED_BLOCKS;
  const int eb = (int)std::min<long>((plane + 255) / 256, 148 * 8);
  if (rgb2yuv) rgb2yuv_kernel<<<eb, 256, 0, st>>>(img, plane);
  if (centering) {
    channel_partial_kernel<0><<<dim3(RED_BLOCKS, 3), RED_THREADS, 0, st>>>(img, plane, nullptr, partial);
    channel_final_kernel<0><<<3, 32, 0, st>>>(partial, RED_BLOCKS, plane, stat);
    channel_apply_kernel<0><<<dim3(eb, 3), 256, 0, st>>>(img, plane, stat);
  }
  if (scaling) {
    // std of the (already centred) channel about ITS mean, as img[i]:std() computes it
    channel_partial_kernel<0><<<dim3(RED_BLOCKS, 3), RED_THREADS, 0, st>>>(img, plane, nullptr, partial);
    channel_final_kernel<0><<<3, 32, 0, st>>>(partial, RED_BLOCKS, plane, stat);
    channel_partial_kernel<1><<<dim3(RED_BLOCKS, 3), RED_THREADS, 0, st>>>(img, plane, stat, partial);
    channel_final_kernel<1><<<3, 32, 0, st>>>(partial, RED_BLOCKS, plane, stat + 4);
    channel_apply_kernel<1><<<dim3(eb, 3), 256, 0, st>>>(img, plane, stat + 4);
  }
  if (ksize > 0) {
    FRCNN_REQUIRE(ksize % 2 == 1 && ksize <= 2 * CN_MAXR + 1, FRCNN_E_INVALID, "contrastive normalisation: odd kernel width <= 17");
    const dim3 grid((W + CN_TW - 1) / CN_TW, (H + CN_TH - 1) / CN_TH);
    contrastive_stage_kernel<0><<<grid, CN_TW * CN_TH, 0, st>>>(img, plane_tmp, H, W, k1d_dev, ksize, threshold);
    contrastive_stage_kernel<1><<<grid, CN_TW * CN_TH, 0, st>>>(plane_tmp, img, H, W, k1d_dev, ksize, threshold);
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------- resize
// image.scale(img, w, h) in its default 'bilinear' mode (BatchIterator.lua:49-52, utilities.lua:188-204): torch/image's
// Main_scaleLinear_rowcol applied along the rows, then along the columns.  Enlarging an axis = linear interpolation with
// scale (src - 1) / (dst - 1), the last sample copied; shrinking = the mean of the source interval [di * s, (di + 1) * s) with
// fractional end weights.  One thread per output sample walks its interval in the C loop's order with individually
// rounded fp32 operations (no FMA contraction), so the result equals the restated loop bit for bit.
__device__ __forceinline__ float scale_sample(const float* __restrict__ src, long stride, int src_len, int dst_len, int di) {
  if (dst_len > src_len) {
    if (src_len == 1 || di == dst_len - 1) return src[(long)(src_len - 1) * stride];
    const float scale = __fdiv_rn((float)(src_len - 1), (float)(dst_len - 1));
    float si_f = __fmul_rn((float)di, scale);
    const int si_i = (int)si_f;
    si_f = __fsub_rn(si_f, (float)si_i);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.f, si_f), src[(long)si_i * stride]), __fmul_rn(si_f, src[(long)(si_i + 1) * stride]));
  }
  if (dst_len < src_len) {
    const float scale = __fdiv_rn((float)src_len, (float)dst_len);
    float si0_f = __fmul_rn((float)di, scale);          // what the loop carries over from the previous sample
    const int si0_i = (int)si0_f;
    si0_f = __fsub_rn(si0_f, (float)si0_i);
    float si1_f = __fmul_rn((float)(di + 1), scale);
    const int si1_i = (int)si1_f;
    si1_f = __fsub_rn(si1_f, (float)si1_i);
    float acc = __fmul_rn(__fsub_rn(1.f, si0_f), src[(long)si0_i * stride]);
    float n = __fsub_rn(1.f, si0_f);
    for (int si = si0_i + 1; si < si1_i; ++si) {
      acc = __fadd_rn(acc, src[(long)si * stride]);
      n = __fadd_rn(n, 1.f);
    }
    if (si1_i < src_len) {
      acc = __fadd_rn(acc, __fmul_rn(si1_f, src[(long)si1_i * stride]));
      n = __fadd_rn(n, si1_f);
    }
    return __fdiv_rn(acc, n);
  }
  return src[(long)di * stride];
}
// pass 1: [C][sh][sw] -> [C][sh][dw]; pass 2: [C][sh][dw] -> [C][dh][dw]
__global__ void scale_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long planes_rows, int sw, int dw) {
  const long total = planes_rows * dw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / dw;
    const int x = (int)(i - row * dw);
    dst[i] = scale_sample(src + row * sw, 1, sw, dw, x);
  }
}
__global__ void scale_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int sh, int dh, int dw) {
  const long total = (long)C * dh * dw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % dw);
    const int y = (int)((i / dw) % dh);
    const int c = (int)(i / ((long)dw * dh));
    dst[i] = scale_sample(src + (long)c * sh * dw + x, dw, sh, dh, y);
  }
}
void launch_scale_image(const float* src, int C, int sh, int sw, float* tmp, float* dst, int dh, int dw, cudaStream_t st) {
  const long t1 = (long)C * sh * dw, t2 = (long)C * dh * dw;
  scale_rows_kernel<<<(int)std::min<long>((t1 + 255) / 256, 148 * 16), 256, 0, st>>>(src, tmp, (long)C * sh, sw, dw);
  scale_cols_kernel<<<(int)std::min<long>((t2 + 255) / 256, 148 * 16), 256, 0, st>>>(tmp, dst, C, sh, dh, dw);
}

}  // namespace frcnn
