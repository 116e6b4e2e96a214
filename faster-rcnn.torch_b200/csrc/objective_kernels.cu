// Kernels of the training objective (objective.lua:91-186) around the tensor-core GEMMs: RPN criteria on the listed
// anchors, training ROI pooling with winners and its scatter backward, the cnet training chains (Linear bias +
// BatchNorm batch statistics + PReLU + Dropout v2) forward / backward and the detection-stage criteria.
// Un-vendored torch7 `nn` criteria restated from their published definitions (oracle/objective.py header).
#include "train.h"
#include "roi_geom.cuh"

namespace frcnn {

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float smooth_l1(float d) { float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float smooth_l1_grad(float d) { return fminf(fmaxf(d, -1.f), 1.f); }

// ------------------------------------------------------------------------------------------ RPN criteria
// objective.lua:91-140, one thread per listed anchor: CrossEntropy on the (fg, bg) pair (target 1 = positive,
// 2 = negative), 10 * SmoothL1(sum) on the 4 regression outputs of the positives, deltas ADDED into delta_outputs
// (an anchor may be listed twice), detection-stage targets: class index / background and
// Anchors.inputToAnchor(Anchors.anchorToInput(anchor, reg_out), roi.rect) in double (Anchors.lua:237-252).
__global__ void rpn_loss_kernel(RpnLossParams p, FrameList fl) {
  // blockIdx.y = frame: its examples, its slice of the anchor-network outputs, its loss slots
  const int fr = blockIdx.y;
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  const int R = fl.R[fr];
  const int e = fl.off[fr] + local;                // row of the batch
  float l_cls = 0.f, l_reg = 0.f;
  if (local < R) {
    const ExampleDev& x = p.ex[e];
    const bool pos = local < fl.n_pos[fr];
    const int l = x.layer - 1, a = x.aspect - 1, yy = x.y - 1, xx = x.x - 1;
    bool ok = l >= 0 && l < MAX_HEADS && a >= 0 && a < 3;
    if (ok) ok = yy >= 0 && yy < p.hh[l] && xx >= 0 && xx < p.hw[l];
    const double* pool_rect = pos ? x.roi : x.anchor;
    for (int i = 0; i < 4; ++i) p.rects[e * 4 + i] = pool_rect[i];
    p.cctarget[e] = pos ? x.class_index - 1 : p.bg_class;
    for (int i = 0; i < 4; ++i) p.crtarget[e * 4 + i] = 0.f;
    if (!ok) {
      atomicExch(p.status, 1);
    } else {
      const long plane = (long)p.hh[l] * p.hw[l];
      const long base = ((long)fr * 18 + a * 6) * plane + (long)yy * p.hw[l] + xx;
      const float* v = p.out[l] + base;
      float* d = p.d_out[l] + base;
      const float v1 = v[0], v2 = v[plane];
      const float m = fmaxf(v1, v2);
      const float lse = m + logf(expf(v1 - m) + expf(v2 - m));
      const float p1 = expf(v1 - lse), p2 = expf(v2 - lse);
      l_cls = pos ? lse - v1 : lse - v2;                       // -log softmax[target]
      atomicAdd(d, p1 - (pos ? 1.f : 0.f));                    // softmax - onehot
      atomicAdd(d + plane, p2 - (pos ? 0.f : 1.f));
      if (pos) {
        float t[4];
        for (int i = 0; i < 4; ++i) {
          t[i] = v[(2 + i) * plane];
          const float df = t[i] - x.reg_target[i];
          l_reg += 10.f * smooth_l1(df);
          atomicAdd(d + (2 + i) * plane, 10.f * smooth_l1_grad(df));
        }
        // reg_proposal = Anchors.anchorToInput(anchor, reg_out); crtarget = Anchors.inputToAnchor(reg_proposal, roi)
        const double aw = __dsub_rn(x.anchor[2], x.anchor[0]), ah = __dsub_rn(x.anchor[3], x.anchor[1]);
        const double px = __dadd_rn(__dmul_rn((double)t[0], aw), x.anchor[0]);
        const double py = __dadd_rn(__dmul_rn((double)t[1], ah), x.anchor[1]);
        const double pw = __dmul_rn(exp((double)t[2]), aw), ph = __dmul_rn(exp((double)t[3]), ah);
        // Rect.fromXYWidthHeight(x, y, w, h) = (x, y, x + w, y + h); width() = maxX - minX
        const double pmaxx = __dadd_rn(px, pw), pmaxy = __dadd_rn(py, ph);
        const double w2 = __dsub_rn(pmaxx, px), h2 = __dsub_rn(pmaxy, py);
        p.crtarget[e * 4 + 0] = (float)((x.roi[0] - px) / w2);
        p.crtarget[e * 4 + 1] = (float)((x.roi[1] - py) / h2);
        p.crtarget[e * 4 + 2] = (float)log((x.roi[2] - x.roi[0]) / w2);
        p.crtarget[e * 4 + 3] = (float)log((x.roi[3] - x.roi[1]) / h2);
      }
    }
  }
  l_cls = wsum(l_cls);
  l_reg = wsum(l_reg);
  if ((threadIdx.x & 31) == 0) {
    if (l_cls != 0.f) atomicAdd(p.losses + 8 * fr + 0, l_cls);
    if (l_reg != 0.f) atomicAdd(p.losses + 8 * fr + 1, l_reg);
  }
}
void launch_rpn_loss(const RpnLossParams& p, const FrameList& fl, cudaStream_t st) {
  const int R = fl.max_R();
  if (R > 0) rpn_loss_kernel<<<dim3(cdiv(R, 128), fl.nf), 128, 0, st>>>(p, fl);
}

// ------------------------------------------------------------------------------------------ training ROI pooling
__global__ void __launch_bounds__(256) roi_pool_train_kernel(const bf16* __restrict__ fmap, int FH, int FW, int C, int kh, int kw,
                                                             LocalizerDev loc, const double* __restrict__ rects, FrameList fl,
                                                             bf16* __restrict__ out, int* __restrict__ argmax, int* status) {
  // CTA <-> ROI (the double-precision crop of extract_roi_pooling_input is evaluated once per ROI); thread <-> (bin, 8
  // channels): 16-byte feature reads, 16-byte row writes.  output [row][bin][C]
  __shared__ int s_rect[5];
  const int r = blockIdx.x;
  fmap += (long)frame_of_row(fl, r) * FH * FW * C;   // the feature map of the row's frame
  if (threadIdx.x == 0) {
    const double* q = rects + (long)r * 4;
    int y0, y1, x0, x1;
    const bool ok = roi_crop(loc, q[0], q[1], q[2], q[3], FH, FW, &y0, &y1, &x0, &x1);
    s_rect[0] = y0; s_rect[1] = y1; s_rect[2] = x0; s_rect[3] = x1; s_rect[4] = ok;
    if (!ok) atomicAdd(status, 1);
  }
  __syncthreads();
  const int y0 = s_rect[0], x0 = s_rect[2], ch = s_rect[1] - y0, cw = s_rect[3] - x0;
  const bool ok = s_rect[4] != 0;
  const int bins = kh * kw, cv = C >> 3;
  for (int item = threadIdx.x; item < bins * cv; item += blockDim.x) {
    const int bin = item / cv, c0 = (item - bin * cv) << 3;
    float m[8];
    int am[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { m[e] = 0.f; am[e] = 0; }
    if (ok) {
      const int by = bin / kw, bx = bin - by * kw;
      const int ys = (by * ch) / kh, ye = ((by + 1) * ch + kh - 1) / kh;
      const int xs = (bx * cw) / kw, xe = ((bx + 1) * cw + kw - 1) / kw;
      const int first = (y0 + ys) * FW + x0 + xs;
#pragma unroll
      for (int e = 0; e < 8; ++e) { m[e] = -INFINITY; am[e] = first; }
      for (int yy = ys; yy < ye; ++yy)
        for (int xx = xs; xx < xe; ++xx) {
          const int pos = (y0 + yy) * FW + x0 + xx;
          const uint4 raw = *reinterpret_cast<const uint4*>(fmap + (long)pos * C + c0);
          const bf16* v8 = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float v = __bfloat162float(v8[e]);
            if (v > m[e]) { m[e] = v; am[e] = pos; }   // first maximum in scan order, as nn.SpatialAdaptiveMaxPooling
          }
        }
    }
    uint4 o;
    bf16* o8 = reinterpret_cast<bf16*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = __float2bfloat16_rn(m[e]);
    const long base = (long)r * bins * C + (long)bin * C + c0;
    *reinterpret_cast<uint4*>(out + base) = o;
    *reinterpret_cast<int4*>(argmax + base) = make_int4(am[0], am[1], am[2], am[3]);
    *reinterpret_cast<int4*>(argmax + base + 4) = make_int4(am[4], am[5], am[6], am[7]);
  }
}
void launch_roi_pool_train(const bf16* fmap, int FH, int FW, int C, int kh, int kw, const LocalizerDev& loc, const double* rects_dev,
                           const FrameList& fl, bf16* out, int* argmax, int* status, cudaStream_t st) {
  FRCNN_REQUIRE(C % 8 == 0, FRCNN_E_INVALID, "training ROI pooling: channel count must be a multiple of 8");
  const int R = fl.rows();
  if (R > 0) roi_pool_train_kernel<<<R, 256, 0, st>>>(fmap, FH, FW, C, kh, kw, loc, rects_dev, fl, out, argmax, status);
}

__global__ void roi_pool_bwd_kernel(const float* __restrict__ d_rows, const int* __restrict__ argmax, FrameList fl, int feat, int C,
                                    float* __restrict__ dfeat, long fmap_elems) {
  const int fr = blockIdx.y;
  const long first = (long)fl.off[fr] * feat, total = (long)fl.R[fr] * feat;
  d_rows += first;
  argmax += first;
  dfeat += fr * fmap_elems;
  // (the launcher keeps a frame's element count below 2^31: 32-bit modulo instead of a 64-bit one per element)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float d = d_rows[i];
    if (d != 0.f) atomicAdd(dfeat + (long)argmax[i] * C + (int)((unsigned)i % (unsigned)C), d);  // ROIs overlap: atomics
  }
}
void launch_roi_pool_bwd(const float* d_rows, const int* argmax, const FrameList& fl, int bins, int C, float* dfeat, long fmap_elems,
                         cudaStream_t st) {
  const long total = (long)fl.max_R() * bins * C;   // (feat = bins * C is a multiple of C, so i % C is the channel in every frame)
  if (total <= 0) return;
  FRCNN_REQUIRE(total < 0x7fffffffL, FRCNN_E_INVALID, "ROI-pool backward: more than 2^31 pooled elements in one frame");
  const int bx = (int)std::min<long>(cdiv(total, 256), std::max(148 * 16 / fl.nf, 148));
  roi_pool_bwd_kernel<<<dim3(bx, fl.nf), 256, 0, st>>>(d_rows, argmax, fl, bins * C, C, dfeat, fmap_elems);
}

// ------------------------------------------------------------------------------------------ cnet chains
// One CTA = 32 feature columns x all R rows (32 x 8 threads: threadIdx.x = column, threadIdx.y strides the rows).
__global__ void __launch_bounds__(256) fc_train_fwd_kernel(FcTrainFwd p, FrameList fl) {
  // blockIdx.y = frame: BatchNormalization sees the ROI batch of ONE image (objective.lua:164 inside the per-image loop)
  __shared__ float s_a[8][33], s_b[8][33];
  const int fr = blockIdx.y, R = fl.R[fr];
  if (R <= 0) return;
  {
    const long o = (long)fl.off[fr] * p.n;
    p.acc += o; p.mask += o; p.pre += o;
    if (p.xhat) p.xhat += o;
    if (p.out_bf16) p.out_bf16 += o;
    if (p.out_f32) p.out_f32 += o;
    p.rstd += (long)fr * p.n;
  }
  const int col = blockIdx.x * 32 + threadIdx.x;
  const bool live = col < p.n;
  const float bias = live ? p.bias[col] : 0.f;
  float mean = 0.f, rstd = 1.f;
  if (p.bn_w) {
    // nn.BatchNormalization in training mode: batch mean, biased variance for the normalisation, running statistics
    // updated with momentum 0.1 (unbiased variance)
    float s = 0.f, ss = 0.f;
    if (live)
      for (int r = threadIdx.y; r < R; r += 8) {
        const float x = p.acc[(long)r * p.n + col] + bias;
        s += x;
        ss += x * x;
      }
    s_a[threadIdx.y][threadIdx.x] = s;
    s_b[threadIdx.y][threadIdx.x] = ss;
    __syncthreads();
    s = 0.f; ss = 0.f;
    for (int k = 0; k < 8; ++k) { s += s_a[k][threadIdx.x]; ss += s_b[k][threadIdx.x]; }
    mean = s / R;
    float var = ss / R - mean * mean;
    var = fmaxf(var, 0.f);
    rstd = rsqrtf(var + 1e-5f);
    if (live && threadIdx.y == 0) {
      p.rstd[col] = rstd;
      if (p.bn_mean) {
        const float unbiased = R > 1 ? var * R / (R - 1) : var;
        if (fl.nf == 1) {
          p.bn_mean[col] = 0.9f * p.bn_mean[col] + 0.1f * mean;
          p.bn_var[col] = 0.9f * p.bn_var[col] + 0.1f * unbiased;
        } else {   // the momentum recursion runs in frame order: bn_running_kernel, after this launch
          p.stat[((long)fr * 2 + 0) * p.n + col] = mean;
          p.stat[((long)fr * 2 + 1) * p.n + col] = unbiased;
        }
      }
    }
  }
  if (!live) return;
  const float slope = p.prelu[0];
  const float g = p.bn_w ? p.bn_w[col] : 1.f, b = p.bn_w ? p.bn_b[col] : 0.f;
  for (int r = threadIdx.y; r < R; r += 8) {
    const long i = (long)r * p.n + col;
    float x = p.acc[i] + bias;
    if (p.bn_w) {
      const float xh = (x - mean) * rstd;
      p.xhat[i] = xh;
      x = xh * g + b;
    }
    p.pre[i] = x;
    float a = x > 0.f ? x : x * slope;
    a *= p.mask[i] * p.keep_scale;   // nn.Dropout v2: scale by 1 / (1 - p) at training time
    if (p.out_bf16) p.out_bf16[i] = __float2bfloat16_rn(a);
    if (p.out_f32) p.out_f32[i] = a;
  }
}
// running statistics of nn.BatchNormalization, momentum 0.1, one update per frame in frame order (what the per-image
// loop of the reference does to them)
__global__ void bn_running_kernel(const float* __restrict__ stat, float* __restrict__ bn_mean, float* __restrict__ bn_var, int n, FrameList fl) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  float m = bn_mean[col], v = bn_var[col];
  for (int fr = 0; fr < fl.nf; ++fr) {
    if (fl.R[fr] <= 0) continue;
    m = 0.9f * m + 0.1f * stat[((long)fr * 2 + 0) * n + col];
    v = 0.9f * v + 0.1f * stat[((long)fr * 2 + 1) * n + col];
  }
  bn_mean[col] = m;
  bn_var[col] = v;
}
void launch_fc_train_fwd(const FcTrainFwd& p, const FrameList& fl, cudaStream_t st) {
  if (fl.max_R() <= 0) return;
  fc_train_fwd_kernel<<<dim3(cdiv(p.n, 32), fl.nf), dim3(32, 8), 0, st>>>(p, fl);
  if (p.bn_w && p.bn_mean && fl.nf > 1) bn_running_kernel<<<cdiv(p.n, 128), 128, 0, st>>>(p.stat, p.bn_mean, p.bn_var, p.n, fl);
}

__global__ void __launch_bounds__(256) fc_train_bwd_kernel(FcTrainBwd p, FrameList fl) {
  __shared__ float s_a[8][33], s_b[8][33], s_c[8][33];
  const int fr = blockIdx.y, R = fl.R[fr];
  if (R <= 0) return;
  {
    const long o = (long)fl.off[fr] * p.n;
    p.d_in += o; p.pre += o; p.mask += o; p.d_out_bf16 += o;
    if (p.xhat) p.xhat += o;
    p.rstd += (long)fr * p.n;
  }
  const int col = blockIdx.x * 32 + threadIdx.x;
  const bool live = col < p.n;
  const float slope = p.prelu[0];
  // pass 1: d wrt the PReLU input (= BN output), column sums for bias / BN parameter gradients, slope gradient
  float sum_d = 0.f, sum_dx = 0.f, ds = 0.f;
  if (live)
    for (int r = threadIdx.y; r < R; r += 8) {
      const long i = (long)r * p.n + col;
      const float x = p.pre[i];
      const float da = p.d_in[i] * p.mask[i] * p.keep_scale;
      const float d = x > 0.f ? da : da * slope;
      if (!(x > 0.f)) ds += da * x;
      sum_d += d;
      if (p.xhat) sum_dx += d * p.xhat[i];
    }
  s_a[threadIdx.y][threadIdx.x] = sum_d;
  s_b[threadIdx.y][threadIdx.x] = sum_dx;
  s_c[threadIdx.y][threadIdx.x] = ds;
  __syncthreads();
  sum_d = 0.f; sum_dx = 0.f; ds = 0.f;
  for (int k = 0; k < 8; ++k) { sum_d += s_a[k][threadIdx.x]; sum_dx += s_b[k][threadIdx.x]; ds += s_c[k][threadIdx.x]; }
  if (threadIdx.y == 0) {
    const float t = wsum(live ? ds : 0.f);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(p.g_prelu, t);
  }
  if (!live) return;
  const float g = p.xhat ? p.bn_w[col] : 1.f;
  const float rstd = p.xhat ? p.rstd[col] : 1.f;
  // Linear bias gradient = column sum of the gradient wrt the Linear output; with BatchNorm that sum is exactly zero
  // analytically (dx below sums to 0), as in the reference
  if (threadIdx.y == 0) {
    if (p.xhat) {
      atomicAdd(p.g_bn_w + col, sum_dx);
      atomicAdd(p.g_bn_b + col, sum_d);
    } else {
      atomicAdd(p.g_bias + col, sum_d);
    }
  }
  float col_dx = 0.f;
  for (int r = threadIdx.y; r < R; r += 8) {
    const long i = (long)r * p.n + col;
    const float x = p.pre[i];
    const float da = p.d_in[i] * p.mask[i] * p.keep_scale;
    float d = x > 0.f ? da : da * slope;
    if (p.xhat) d = g * rstd * (d - sum_d / R - p.xhat[i] * sum_dx / R);   // BatchNorm backward, batch statistics
    col_dx += d;
    p.d_out_bf16[i] = __float2bfloat16_rn(d);
  }
  if (p.xhat) {
    s_a[threadIdx.y][threadIdx.x] = col_dx;
    // (no barrier needed for correctness of g_bias: tiny residual of the analytic zero, added for completeness)
    atomicAdd(p.g_bias + col, col_dx);
  }
}
void launch_fc_train_bwd(const FcTrainBwd& p, const FrameList& fl, cudaStream_t st) {
  if (fl.max_R() <= 0) return;
  fc_train_bwd_kernel<<<dim3(cdiv(p.n, 32), fl.nf), dim3(32, 8), 0, st>>>(p, fl);
}

// objective.lua:166-177 + backward through Linear(nin -> 4) and Linear(nin -> ncls) + LogSoftMax.
// kernel 1: one CTA per row: outputs, losses, dz, d_hidden.  kernel 2: one CTA per output neuron: weight gradients.
__global__ void __launch_bounds__(256) cnet_loss_row_kernel(CnetLossParams p, FrameList fl) {
  extern __shared__ float sh[];  // [nin] hidden, [ncls + 4] outputs -> dz
  const int r = blockIdx.x;      // row of the batch; the criteria average over the rows of its frame
  const int fr = frame_of_row(fl, r), R = fl.R[fr];
  float* h = sh;
  float* z = sh + p.nin;
  for (int i = threadIdx.x; i < p.nin; i += blockDim.x) h[i] = p.hidden[(long)r * p.nin + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int no = p.ncls + 4;
  for (int o = warp; o < no; o += nw) {
    const float* wr = o < 4 ? p.w_reg + (long)o * p.nin : p.w_cls + (long)(o - 4) * p.nin;
    float s = 0.f;
    for (int k = lane; k < p.nin; k += 32) s += h[k] * wr[k];
    s = wsum(s);
    if (lane == 0) z[o] = s + (o < 4 ? p.b_reg[o] : p.b_cls[o - 4]);
  }
  __syncthreads();
  if (threadIdx.x == 0 && p.ext_dreg && p.ext_dcls) {
    // cnet:backward with the caller's output gradients: Linear(nin -> 4) takes d_reg as is; the class branch ends in
    // nn.LogSoftMax, y = z - lse(z): dz_c = dy_c - softmax(z)_c * sum_j dy_j
    for (int i = 0; i < 4; ++i) z[i] = p.ext_dreg[r * 4 + i];
    float m = -INFINITY;
    for (int c = 0; c < p.ncls; ++c) m = fmaxf(m, z[4 + c]);
    float s = 0.f;
    for (int c = 0; c < p.ncls; ++c) s += expf(z[4 + c] - m);
    const float lse = m + logf(s);
    float dsum = 0.f;
    for (int c = 0; c < p.ncls; ++c) dsum += p.ext_dcls[(long)r * p.ncls + c];
    for (int c = 0; c < p.ncls; ++c) z[4 + c] = p.ext_dcls[(long)r * p.ncls + c] - expf(z[4 + c] - lse) * dsum;
  } else if (threadIdx.x == 0) {
    const bool pos = r - fl.off[fr] < fl.n_pos[fr];
    float l_reg = 0.f;
    for (int i = 0; i < 4; ++i) {
      const float out = pos ? z[i] : 0.f;                       // crout of the negatives is zeroed (objective.lua:170)
      const float d = out - p.crtarget[r * 4 + i];
      l_reg += 10.f * smooth_l1(d);
      z[i] = pos ? 10.f * smooth_l1_grad(d) : 0.f;
    }
    float m = -INFINITY;
    for (int c = 0; c < p.ncls; ++c) m = fmaxf(m, z[4 + c]);
    float s = 0.f;
    for (int c = 0; c < p.ncls; ++c) s += expf(z[4 + c] - m);
    const float lse = m + logf(s);
    const int t = p.cctarget[r];
    const float l_cls = (lse - z[4 + t]) / R;                    // ClassNLLCriterion, sizeAverage
    for (int c = 0; c < p.ncls; ++c) z[4 + c] = (expf(z[4 + c] - lse) - (c == t ? 1.f : 0.f)) / R;
    float* losses = p.losses + (long)p.loss_stride * fr;
    if (l_reg != 0.f) atomicAdd(losses + 2, l_reg);
    atomicAdd(losses + 3, l_cls);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < no; o += blockDim.x) p.dz[(long)r * no + o] = z[o];
  for (int k = threadIdx.x; k < p.nin; k += blockDim.x) {
    float d = 0.f;
    for (int o = 0; o < 4; ++o) d += z[o] * p.w_reg[(long)o * p.nin + k];
    for (int c = 0; c < p.ncls; ++c) d += z[4 + c] * p.w_cls[(long)c * p.nin + k];
    p.d_hidden[(long)r * p.nin + k] = d;
  }
}
__global__ void __launch_bounds__(256) cnet_loss_wgrad_kernel(CnetLossParams p, int rows, int rows_per) {
  // CTA <-> (output o, 64 weights of its row, a slice of the batch's rows: dz already carries every frame's 1 / R, so
  // the sum runs over the rows of ALL frames); thread <-> (weight, one of 4 interleaved row groups): four independent
  // chains per weight instead of one (the serial chain paced this kernel), combined in fixed order, one atomic per
  // weight and slice
  const int o = blockIdx.x, kc = blockIdx.y, no = p.ncls + 4;
  const int r0 = blockIdx.z * rows_per, r1 = min(rows, r0 + rows_per);
  const int kl = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int k = kc * 64 + kl;
  __shared__ float part[4][64];
  __shared__ float bpart[256];
  float* gw = o < 4 ? p.g_w_reg + (long)o * p.nin : p.g_w_cls + (long)(o - 4) * p.nin;
  float s = 0.f;
  if (k < p.nin)
    for (int r = r0 + rg; r < r1; r += 4) s += p.dz[(long)r * no + o] * p.hidden[(long)r * p.nin + k];
  part[rg][kl] = s;
  if (kc == 0) {
    float b = 0.f;
    for (int r = r0 + threadIdx.x; r < r1; r += 256) b += p.dz[(long)r * no + o];
    bpart[threadIdx.x] = b;
  }
  __syncthreads();
  if (rg == 0 && k < p.nin) atomicAdd(gw + k, (part[0][kl] + part[1][kl]) + (part[2][kl] + part[3][kl]));
  if (kc == 0) {
    for (int off = 128; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) bpart[threadIdx.x] += bpart[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(o < 4 ? p.g_b_reg + o : p.g_b_cls + (o - 4), bpart[0]);
  }
}
void launch_cnet_loss_bwd(const CnetLossParams& p, const FrameList& fl, cudaStream_t st) {
  const int rows = fl.rows();
  if (rows <= 0) return;
  cnet_loss_row_kernel<<<rows, 256, (p.nin + p.ncls + 4) * sizeof(float), st>>>(p, fl);
  const int slices = std::min(16, cdiv(rows, 256));
  cnet_loss_wgrad_kernel<<<dim3(p.ncls + 4, (p.nin + 63) / 64, slices), 256, 0, st>>>(p, rows, cdiv(rows, slices));
}

// ------------------------------------------------------------------------------------------ fc weight layouts
__global__ void pack_fc_weight_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int nout, int C, int bins, int permute) {
  const long K = (long)C * bins, total = K * nout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int o = (int)(i % nout);
    const long k = i / nout;
    long src = k;
    if (permute) {
      const int b = (int)(k / C), c = (int)(k - (long)b * C);
      src = (long)c * bins + b;
    }
    out[i] = __float2bfloat16_rn(w[(long)o * K + src]);
  }
}
void launch_pack_fc_weight_dgrad(const float* w, bf16* out, int nout, int C, int bins, int permute, cudaStream_t st) {
  const long total = (long)nout * C * bins;
  pack_fc_weight_dgrad_kernel<<<(int)std::min<long>(cdiv(total, 256), 148 * 8), 256, 0, st>>>(w, out, nout, C, bins, permute);
}
// One CTA per output neuron: the row [bins][C] (ROI order) is staged in shared memory and added to the Torch-order
// row [C][bins] with coalesced accesses on both sides.
__global__ void __launch_bounds__(256) wgrad_finish_fc_kernel(const float* __restrict__ dw, float* __restrict__ grad, int nout, int C,
                                                              int bins, int permute) {
  extern __shared__ float srow[];
  const int K = C * bins;   // 32-bit index arithmetic throughout: the 64-bit divisions of the first version cost more than the traffic
  for (int o = blockIdx.x; o < nout; o += gridDim.x) {
    const float* src = dw + (long)o * K;
    float* dst = grad + (long)o * K;
    if (!permute) {
      for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] += src[k];
      continue;
    }
    __syncthreads();
    // one pad word per C-long bin row: consecutive threads read consecutive bins of one channel, C (a multiple of 32)
    // words apart -- the same bank without the pad
    for (int k = threadIdx.x; k < K; k += blockDim.x) srow[k + k / C] = src[k];
    __syncthreads();
    for (int d = threadIdx.x; d < K; d += blockDim.x) {
      const int c = d / bins, b = d - c * bins;
      dst[d] += srow[b * (C + 1) + c];
    }
  }
}
void launch_wgrad_finish_fc(const float* dw, float* grad, int nout, int C, int bins, int permute, cudaStream_t st) {
  static DeviceOnce configured;
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(wgrad_finish_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  }
  const size_t smem = permute ? (size_t)(C + 1) * bins * sizeof(float) : 0;
  FRCNN_REQUIRE(smem <= 96 * 1024, FRCNN_E_INVALID, "fc row too long for the transposing gradient accumulation");
  wgrad_finish_fc_kernel<<<std::min(nout, 148 * 4), 256, smem, st>>>(dw, grad, nout, C, bins, permute);
}

}  // namespace frcnn
