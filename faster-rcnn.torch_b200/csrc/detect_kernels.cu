// Detector-side kernels: everything Detector.lua:36-136 does in per-anchor / per-candidate Lua loops.
#include "detect.h"
#include "roi_geom.cuh"

namespace frcnn {

// ------------------------------------------------------------------------------------------------ RPN decode
// Detector.lua:36-66.  One thread per anchor in the reference's loop order (layer, y, x, aspect); the surviving
// anchors are written in that order (ordered stream compaction: ballot + a chained block prefix; block ids are
// handed out by an atomic ticket so that a block only ever waits for blocks that are already running).
__global__ void __launch_bounds__(256) rpn_decode_kernel(DecodeParams p) {
  __shared__ unsigned long long s_ticket;
  __shared__ int s_excl;
  __shared__ int warp_cnt[8];
  const int img = blockIdx.y;
  // tickets keep counting across launches (no per-launch parameter, so the launch can be replayed from a CUDA graph):
  // ticket t -> block id t % nblocks of launch number t / nblocks, which tags the scan-state words of that launch.
  // The counter is 64 bits wide: a 32-bit one wraps after 2^32 / nblocks launches and, unless nblocks divides 2^32,
  // the launch straddling the wrap would hand out duplicate block ids.  The tag is the low 32 bits of the launch
  // number + 1; launches 2^32 apart never overlap in time.
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.ticket[img], 1ull);
  __syncthreads();
  const unsigned long long ticket = s_ticket;
  const int bid = (int)(ticket % (unsigned long long)p.nblocks);
  const unsigned epoch = (unsigned)(ticket / (unsigned long long)p.nblocks + 1ull);
  const int idx = bid * 256 + threadIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  bool keep = false;
  double rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
  float logp = 0.f;
  int layer = 0, aspect = 0, y = 0, x = 0;
  if (idx < p.total) {
    while (layer + 1 < MAX_HEADS && idx >= p.offs[layer + 1]) ++layer;
    int rem = idx - p.offs[layer];
    const int hw = p.hw[layer], hh = p.hh[layer];
    aspect = rem % 3;
    rem /= 3;
    x = rem % hw;
    y = rem / hw;
    const long plane = (long)hh * hw;
    const float* base = p.head[layer] + ((long)img * 18 + aspect * 6) * plane + (long)y * hw + x;
    // nn.LogSoftMax over the (fg, bg) pair, element 1 (Detector.lua:47-52): evaluated in double, rounded once
    const double c1 = (double)__ldg(base), c2 = (double)__ldg(base + plane);
    const double m = fmax(c1, c2);
    const double lp = (c1 - m) - log(exp(c1 - m) + exp(c2 - m));
    logp = (float)lp;
    if (exp((double)logp) > p.threshold) {  // Detector.lua:54
      // Anchors:get (Anchors.lua:60-67): fp32 LUT entries read back as doubles
      const float* wl = p.w_lut + ((layer * 3 + aspect) * LUT_EXTENT + x) * 2;
      const float* hl = p.h_lut + ((layer * 3 + aspect) * LUT_EXTENT + y) * 2;
      const double aminx = (double)wl[0], amaxx = (double)wl[1], aminy = (double)hl[0], amaxy = (double)hl[1];
      const double aw = __dsub_rn(amaxx, aminx), ah = __dsub_rn(amaxy, aminy);
      const double t1 = (double)__ldg(base + 2 * plane), t2 = (double)__ldg(base + 3 * plane);
      const double t3 = (double)__ldg(base + 4 * plane), t4 = (double)__ldg(base + 5 * plane);
      // Anchors.anchorToInput (Anchors.lua:245-252), no FMA contraction
      rx0 = __dadd_rn(__dmul_rn(t1, aw), aminx);
      ry0 = __dadd_rn(__dmul_rn(t2, ah), aminy);
      rx1 = __dadd_rn(rx0, __dmul_rn(exp(t3), aw));
      ry1 = __dadd_rn(ry0, __dmul_rn(exp(t4), ah));
      // r:overlaps(input_rect), strict (Rect.lua:90-93) with input_rect = (0, 0, W, H)
      keep = rx0 < p.img_w && rx1 > 0.0 && ry0 < p.img_h && ry1 > 0.0;
    }
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int before = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    int c = warp_cnt[w];
    if (w < warp) before += c;
    block_total += c;
  }
  // Parallel look-back: every block publishes its own count (tagged with this launch's epoch) as soon as it is
  // known; block `bid` sums the counts of blocks 0..bid-1, each thread polling a strided subset.  Blocks with a
  // smaller ticket are already running (tickets are handed out in start order), so the waits terminate.
  volatile unsigned long long* status = p.status + (long)img * p.nblocks;
  if (threadIdx.x == 0) status[bid] = ((unsigned long long)epoch << 32) | (unsigned)block_total;
  unsigned part = 0;
  for (int b = threadIdx.x; b < bid; b += 256) {
    unsigned long long v;
    do {
      v = status[b];
    } while ((unsigned)(v >> 32) != epoch);
    part += (unsigned)v;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  __shared__ unsigned s_part[8];
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned excl = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) excl += s_part[w];
    const unsigned incl = excl + (unsigned)block_total;
    s_excl = (int)excl;
    if (bid == p.nblocks - 1) {
      p.cand_count[img] = min((int)incl, p.cap);
      if ((int)incl > p.cap) atomicExch(p.cand_overflow, 1);
    }
  }
  __syncthreads();
  if (keep) {
    const int slot = s_excl + before + __popc(bal & ((1u << lane) - 1u));
    if (slot < p.cap) {
      const long o = (long)img * p.cap + slot;
      double* r = p.cand_r + o * 4;
      r[0] = rx0; r[1] = ry0; r[2] = rx1; r[3] = ry1;
      p.cand_box[o] = make_float4((float)rx0, (float)ry0, (float)rx1, (float)ry1);  // Rect:totensor -> fp32
      p.cand_logp[o] = logp;
      p.cand_anchor[o] = make_int4(layer + 1, aspect + 1, y + 1, x + 1);
    }
  }
}
void launch_rpn_decode(const DecodeParams& p, int N, cudaStream_t st) {
  rpn_decode_kernel<<<dim3(p.nblocks, N), 256, 0, st>>>(p);
}

// One thread per NMS survivor: row index = (survivors of earlier images) + position in pick order; the crop rect is
// Localizer:inputToFeatureRect in double on the device.
__global__ void __launch_bounds__(256) roi_prepare_kernel(RoiParams p, int N) {
  const int img = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int base = 0, total = 0;
  for (int n = 0; n < N; ++n) {
    const int c = p.pick_count[n];
    if (n < img) base += c;
    total += c;
  }
  if (i == 0) {
    p.roi_base_out[img] = base;
    if (img == 0) *p.roi_total = min(total, p.total_cap);
  }
  if (i >= p.pick_count[img]) return;
  const int row = base + i;
  if (row >= p.total_cap) return;
  const int cand = p.pick[(long)img * p.cap + i];
  const double* r = p.cand_r + ((long)img * p.cap + cand) * 4;
  int y0, y1, x0, x1;
  const bool ok = roi_crop(p.loc, r[0], r[1], r[2], r[3], p.FH, p.FW, &y0, &y1, &x0, &x1);
  if (!ok) atomicAdd(p.status, 1);
  p.roi_rect[row] = ok ? make_int4(y0, y1, x0, x1) : make_int4(-1, -1, -1, -1);
  p.roi_img[row] = img;
  p.roi_cand[row] = cand;
}

// nn.SpatialAdaptiveMaxPooling(kw, kh) on the crop (Detector.lua:96-97), NHWC bf16 feature map, output
// [row][bin][C] bf16 (channel-contiguous; the first cnet weight matrix is permuted to match at pack time).
// Persistent CTAs over the ROI rows; a thread handles 8 channels (16 bytes) of one bin; max is exact in any format.
__global__ void __launch_bounds__(256) roi_pool_nhwc_kernel(RoiParams p) {
  const int total = *p.roi_total;
  const int cv = p.C >> 3;
  const int per_row = p.kh * p.kw * cv;
  const long n_items = (long)total * per_row;
  const uint32_t ni = p.f16 ? 0xFC00FC00u : 0xFF80FF80u;   // -inf pairs in the feature map's 16-bit format
  const uint4 ninf = make_uint4(ni, ni, ni, ni);
  // flat (row, bin, 8-channel group) index space over the whole grid: few large ROIs still fill the machine
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n_items; idx += (long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / per_row);
    const int item = (int)(idx - (long)row * per_row);
    const int4 rc = p.roi_rect[row];
    const int img = p.roi_img[row];
    const int c8 = item % cv, bin = item / cv;
    uint4 m = make_uint4(0, 0, 0, 0);
    if (rc.x >= 0) {
      const int y0 = rc.x, x0 = rc.z, ch = rc.y - rc.x, cw = rc.w - rc.z;
      const uint4* fm = reinterpret_cast<const uint4*>(p.fmap) + (long)img * p.FH * p.FW * cv + c8;
      const int by = bin / p.kw, bx = bin - by * p.kw;
      // adaptive pooling window: [floor(b*S/k), ceil((b+1)*S/k))
      const int ys = (by * ch) / p.kh, ye = ((by + 1) * ch + p.kh - 1) / p.kh;
      const int xs = (bx * cw) / p.kw, xe = ((bx + 1) * cw + p.kw - 1) / p.kw;
      m = ninf;
      for (int yy = ys; yy < ye; ++yy) {
        const uint4* rowp = fm + ((long)(y0 + yy) * p.FW + x0) * cv;
#pragma unroll 4
        for (int xx = xs; xx < xe; ++xx) {
          const uint4 v = __ldg(rowp + (long)xx * cv);
          if (p.f16) {
            __half2* pm = reinterpret_cast<__half2*>(&m);
            const __half2* pv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) pm[j] = __hmax2(pm[j], pv[j]);
          } else {
            __nv_bfloat162* pm = reinterpret_cast<__nv_bfloat162*>(&m);
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) pm[j] = __hmax2(pm[j], pv[j]);
          }
        }
      }
    }
    reinterpret_cast<uint4*>(p.out)[(long)row * per_row + item] = m;
  }
}
// ------------------------------------------------------------------------------------------ the `amp` module slot
// nn.SpatialAdaptiveMaxPooling(kw, kh):forward on a (possibly non-contiguous) [C][h][w] view, as objective.lua:118,138 and
// Detector.lua:97 call it on the crop extract_roi_pooling_input returns: bin (by, bx) covers rows
// [floor(by*h/kh), ceil((by+1)*h/kh)) and the same along x, the FIRST maximum in row-major order wins (strict >).  idx holds
// the winner's position inside the view (y*w + x) as a float -- the module's `indices` field, which objective.lua only
// clones and puts back (:119,139,183).
__global__ void adaptive_maxpool_fwd_kernel(const float* __restrict__ x, int C, int h, int w, long sc, long sh, long sw, int kh, int kw,
                                            float* __restrict__ out, float* __restrict__ idx) {
  const long total = (long)C * kh * kw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int bx = (int)(i % kw), by = (int)((i / kw) % kh), c = (int)(i / ((long)kw * kh));
    const int ys = (by * h) / kh, ye = ((by + 1) * h + kh - 1) / kh;
    const int xs = (bx * w) / kw, xe = ((bx + 1) * w + kw - 1) / kw;
    float best = -INFINITY;
    int arg = ys * w + xs;
    for (int yy = ys; yy < ye; ++yy)
      for (int xx = xs; xx < xe; ++xx) {
        const float v = __ldg(x + c * sc + yy * sh + xx * sw);
        if (v > best) { best = v; arg = yy * w + xx; }
      }
    out[i] = best;
    if (idx) idx[i] = (float)arg;
  }
}
// :backward -- gradInput [C][h][w] (contiguous, zeroed here) receives every output's gradient at its winner; bins overlap
// when the view is smaller than the grid, hence the atomic (cunn does the same)
__global__ void adaptive_maxpool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ idx, int C, int h, int w, int kh,
                                            int kw, float* __restrict__ dx) {
  const long total = (long)C * kh * kw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i / ((long)kw * kh));
    const int arg = (int)idx[i];
    if (arg >= 0 && arg < h * w) atomicAdd(dx + (long)c * h * w + arg, dout[i]);
  }
}
void launch_adaptive_maxpool_fwd(const float* x, int C, int h, int w, long sc, long sh, long sw, int kh, int kw, float* out,
                                 float* idx, cudaStream_t st) {
  const long total = (long)C * kh * kw;
  adaptive_maxpool_fwd_kernel<<<(int)std::min<long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(x, C, h, w, sc, sh, sw, kh, kw, out, idx);
}
void launch_adaptive_maxpool_bwd(const float* dout, const float* idx, int C, int h, int w, int kh, int kw, float* dx, cudaStream_t st) {
  const long total = (long)C * kh * kw;
  cudaMemsetAsync(dx, 0, (size_t)C * h * w * sizeof(float), st);
  adaptive_maxpool_bwd_kernel<<<(int)std::min<long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(dout, idx, C, h, w, kh, kw, dx);
}

void launch_roi_pool_nhwc(const RoiParams& p, int N, int num_sms, cudaStream_t st) {
  roi_prepare_kernel<<<dim3((p.cap + 255) / 256, N), 256, 0, st>>>(p, N);
  roi_pool_nhwc_kernel<<<num_sms * 4, 256, 0, st>>>(p);
}

// Same operation on a Torch-layout fp32 feature map [C][H][W] with the reference's output ordering
// (c*kh*kw + by*kw + bx) and its argmax indices -- the drop-in for the Lua `amp` slot.
__global__ void __launch_bounds__(256) roi_pool_chw_kernel(const float* __restrict__ fmap, int C, int H, int W, LocalizerDev loc,
                                                           const double* __restrict__ rects, int kh, int kw, float* __restrict__ out,
                                                           int32_t* __restrict__ argmax, int* status) {
  __shared__ int s_rect[5];
  const int r = blockIdx.x;
  if (threadIdx.x == 0) {
    const double* q = rects + (long)r * 4;
    int y0, y1, x0, x1;
    bool ok = roi_crop(loc, q[0], q[1], q[2], q[3], H, W, &y0, &y1, &x0, &x1);
    s_rect[0] = y0; s_rect[1] = y1; s_rect[2] = x0; s_rect[3] = x1; s_rect[4] = ok;
    if (!ok) atomicAdd(status, 1);
  }
  __syncthreads();
  const int y0 = s_rect[0], x0 = s_rect[2], ch = s_rect[1] - y0, cw = s_rect[3] - x0;
  const bool ok = s_rect[4] != 0;
  const int bins = kh * kw;
  for (int item = threadIdx.x; item < C * bins; item += blockDim.x) {
    const int c = item / bins, bin = item - c * bins;
    float m = 0.f;
    int am = 0;
    if (ok) {
      const int by = bin / kw, bx = bin - by * kw;
      const int ys = (by * ch) / kh, ye = ((by + 1) * ch + kh - 1) / kh;
      const int xs = (bx * cw) / kw, xe = ((bx + 1) * cw + kw - 1) / kw;
      m = -INFINITY;
      am = (y0 + ys) * W + x0 + xs;
      for (int yy = ys; yy < ye; ++yy)
        for (int xx = xs; xx < xe; ++xx) {
          float v = fmap[((long)c * H + y0 + yy) * W + x0 + xx];
          if (v > m) { m = v; am = (y0 + yy) * W + x0 + xx; }
        }
    }
    out[(long)r * C * bins + item] = m;
    if (argmax) argmax[(long)r * C * bins + item] = am;
  }
}
void launch_roi_pool_chw(const float* fmap, int C, int H, int W, const LocalizerDev& loc, const double* rects_dev, int R, int kh,
                         int kw, float* out, int32_t* argmax, int* status, cudaStream_t st) {
  if (R <= 0) return;
  roi_pool_chw_kernel<<<R, 256, 0, st>>>(fmap, C, H, W, loc, rects_dev, kh, kw, out, argmax, status);
}

// ------------------------------------------------------------------------------------------------ finalize
// Detector.lua:106-122 per candidate: refined rect r2 = anchorToInput(r, bbox_out[i]) in double, class = argmax of
// the log-softmax row (first maximum), accepted when class ~= bg and exp(confidence) > 0.2.
__global__ void finalize_kernel(FinalizeParams p) {
  const int R = *p.roi_total;
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < R; row += gridDim.x * blockDim.x) {
    const int img = p.roi_img[row], cand = p.roi_cand[row];
    const double* r = p.cand_r + ((long)img * p.cap + cand) * 4;
    const float* t = p.reg + (long)row * 4;
    const double aw = __dsub_rn(r[2], r[0]), ah = __dsub_rn(r[3], r[1]);
    const double x = __dadd_rn(__dmul_rn((double)t[0], aw), r[0]);
    const double y = __dadd_rn(__dmul_rn((double)t[1], ah), r[1]);
    const double w = __dmul_rn(exp((double)t[2]), aw), h = __dmul_rn(exp((double)t[3]), ah);
    double* o = p.fin_r2 + (long)row * 4;
    o[0] = x; o[1] = y; o[2] = __dadd_rn(x, w); o[3] = __dadd_rn(y, h);
    p.fin_box[row] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)o[3]);
    const float* c = p.cls + (long)row * p.ncls;
    int best = 0;
    float bv = c[0];
    for (int k = 1; k < p.ncls; ++k)
      if (c[k] > bv) { bv = c[k]; best = k; }
    const int cls1 = best + 1;
    const bool pass = cls1 != p.ncls && exp((double)bv) > p.class_prob;  // Detector.lua:115
    p.fin_cls[row] = pass ? cls1 : 0;
    p.fin_conf[row] = bv;
  }
}
void launch_finalize(const FinalizeParams& p, int R_cap, cudaStream_t st) {
  int blocks = (R_cap + 255) / 256;
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  finalize_kernel<<<blocks, 256, 0, st>>>(p);
}

// ------------------------------------------------------------------------------------------------ class grouping
// Detector.lua:115-121 builds yclass[class] lists in candidate order.  One CTA per image sorts the keys
// (class << 32 | candidate position) ascending in shared memory (rejected candidates get class 0xffff), writes the
// grouped boxes and the per-(image, class) segment table of the following NMS.
__global__ void __launch_bounds__(1024) group_by_class_kernel(GroupParams p, NmsState st) {
  extern __shared__ unsigned long long gkeys[];
  const int img = blockIdx.x;
  int n = min(p.pick_count[img], p.cap);
  if (n > NMS_CTA_MAX_SEG) {  // only reachable once the candidate capacity has grown past 8192 (api.cu)
    if (threadIdx.x == 0) atomicExch(p.overflow, 1);
    n = NMS_CTA_MAX_SEG;
  }
  const int base = p.roi_base[img];
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      int c = p.fin_cls[base + i];
      k = ((unsigned long long)(c > 0 ? c : 0xffff) << 32) | (unsigned)i;
    }
    gkeys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool asc = ((lo & size) == 0);
        unsigned long long a = gkeys[lo], b = gkeys[hi];
        if ((a > b) == asc) { gkeys[lo] = b; gkeys[hi] = a; }
      }
      __syncthreads();
    }
  }
  // segment s = img * n_classes + (class - 1)
  for (int c = threadIdx.x; c < p.n_classes; c += blockDim.x) {
    st.seg_beg[img * p.n_classes + c] = img * p.cap;
    st.seg_len[img * p.n_classes + c] = 0;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const unsigned long long key = gkeys[k];
    const int c = (int)(key >> 32);
    if (c == 0xffff) continue;
    const int i = (int)(key & 0xffffffffu);
    p.gbox[(long)img * p.cap + k] = p.fin_box[base + i];
    p.grow[(long)img * p.cap + k] = base + i;
    const int cprev = k > 0 ? (int)(gkeys[k - 1] >> 32) : -1;
    const int cnext = k + 1 < n ? (int)(gkeys[k + 1] >> 32) : 0xffff;
    if (cprev != c) st.seg_beg[img * p.n_classes + c - 1] = img * p.cap + k;
    if (cnext != c) atomicAdd(&st.seg_len[img * p.n_classes + c - 1], k + 1);   // end index ...
    if (cprev != c) atomicAdd(&st.seg_len[img * p.n_classes + c - 1], -k);      // ... minus begin index
  }
  if (threadIdx.x == 0) {
    int cnt = 0;
    // number of accepted candidates = first key with class 0xffff
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if ((int)(gkeys[mid] >> 32) == 0xffff) hi = mid; else lo = mid + 1;
    }
    cnt = lo;
    p.n_pass[img] = cnt;
  }
}
void launch_group_by_class(const GroupParams& p, NmsWorkspace* ws, int N, cudaStream_t st) {
  static DeviceOnce configured;
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(group_by_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
  }
  int n2 = 1;
  while (n2 < std::min(p.cap, NMS_CTA_MAX_SEG)) n2 <<= 1;
  group_by_class_kernel<<<N, 1024, n2 * 8, st>>>(p, ws->st);
}

// ------------------------------------------------------------------------------------------------ winners
// Detector.lua:125-136: winners = for each class, the picks of its NMS, in pick order.  Segments are ordered
// (image, class), so an exclusive scan of the pick counts gives every winner its output slot.
// One CTA per (image, class) segment: its output offset is the sum of the pick counts of all earlier segments.
__global__ void __launch_bounds__(256) assemble_kernel(AssembleParams p, NmsState st, int n_seg) {
  __shared__ int s_part[8];
  __shared__ int s_start;
  const int s = blockIdx.x;
  int local = 0;
  for (int i = threadIdx.x; i < s; i += blockDim.x) local += st.counts[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    s_start = t;
  }
  __syncthreads();
  const int start = s_start;
  const int cnt = st.counts[s];
  const int beg = st.seg_beg[s];
  if (s == n_seg - 1 && threadIdx.x == 0) *p.n_det = start + cnt;
  for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
    const int slot = start + k;
    if (slot >= p.det_cap) break;
    const int g = beg + st.pick[beg + k];
    const int row = p.grow[g];
    const int img = p.roi_img[row], cand = p.roi_cand[row];
    const long ci = (long)img * p.cap + cand;
    frcnn_detection d;
    for (int e = 0; e < 4; ++e) {
      d.r[e] = p.cand_r[ci * 4 + e];
      d.r2[e] = p.fin_r2[(long)row * 4 + e];
    }
    d.p = p.cand_logp[ci];
    d.confidence = p.fin_conf[row];
    d.cls = p.fin_cls[row];
    const int4 a = p.cand_anchor[ci];
    d.layer = a.x; d.aspect = a.y; d.y = a.z; d.x = a.w;
    d.image = img;
    p.det[slot] = d;
  }
}
void launch_assemble(const AssembleParams& p, NmsWorkspace* ws, int n_seg, cudaStream_t st) {
  assemble_kernel<<<n_seg, 256, 0, st>>>(p, ws->st, n_seg);
}

}  // namespace frcnn
