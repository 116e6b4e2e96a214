// Anchor labelling on the GPU: Anchors:findRangesXY / findPositive / sampleNegative (Anchors.lua:86-235), the nested
// Lua loops over anchors x ground-truth boxes that BatchIterator.lua:200-225 runs for every training image.
//
// Exactness contract (bit-exact index lists vs the oracle): anchors are the float32 LUT entries read as doubles
// (main.lua:51), Rect.IoU is evaluated in double exactly as Rect.lua:126-141 (intersection clamps to the empty rect
// (0,0,0,0); no +1), candidates are enumerated in the reference's loop order (scale, aspect, y, x) and every ordered
// list is built with scans, never with atomics.  None of this is HBM-bound: per ROI a few thousand IoU evaluations on
// LUT data that lives in shared memory.
#include "common.h"
#include "label.h"

namespace frcnn {

static constexpr int LB_THREADS = 256;

struct Range {
  int lx, ly, ux, uy;  // 1-based [lx, ux) x [ly, uy), Anchors.lua:118-138
};

// first 1-based index with t[i] >= value / > value over the 200 LUT cells (stride 2 floats: {min,max} pairs)
__device__ __forceinline__ int lower_bound_lut(const float* t, double value) {
  int low = 1, high = LUT_CELLS;
  while (low <= high) {
    const int mid = (low + high) >> 1;
    if ((double)t[(mid - 1) * 2] >= value) high = mid - 1;
    else low = mid + 1;
  }
  return low;
}
__device__ __forceinline__ int upper_bound_lut(const float* t, double value) {
  int low = 1, high = LUT_CELLS;
  while (low <= high) {
    const int mid = (low + high) >> 1;
    if ((double)t[(mid - 1) * 2] > value) high = mid - 1;
    else low = mid + 1;
  }
  return low;
}

// Anchors:findRangesXY for one (scale i, aspect j): false when the range is empty (Anchors.lua:134)
__device__ __forceinline__ bool find_range(const float* w_lut, const float* h_lut, int ij, const double rect[4], const double* clip, Range& r) {
  const float* w = w_lut + (size_t)ij * LUT_CELLS * 2;
  const float* h = h_lut + (size_t)ij * LUT_CELLS * 2;
  r.lx = upper_bound_lut(w + 1, rect[0]);   // a.maxX > r.minX
  r.ly = upper_bound_lut(h + 1, rect[1]);
  r.ux = lower_bound_lut(w, rect[2]);       // a.minX >= r.maxX
  r.uy = lower_bound_lut(h, rect[3]);
  if (clip) {
    r.lx = max(r.lx, lower_bound_lut(w, clip[0]));
    r.ly = max(r.ly, lower_bound_lut(h, clip[1]));
    r.ux = min(r.ux, upper_bound_lut(w + 1, clip[2]));
    r.uy = min(r.uy, upper_bound_lut(h + 1, clip[3]));
  }
  return r.ux > r.lx && r.uy > r.ly;
}

// Rect.IoU(a, b) (Rect.lua:126-141), doubles, individually rounded operations
__device__ __forceinline__ double rect_iou(const double a[4], double bx0, double by0, double bx1, double by1) {
  double minx = fmax(a[0], bx0), miny = fmax(a[1], by0), maxx = fmin(a[2], bx1), maxy = fmin(a[3], by1);
  double inter = 0.0;
  if (maxx >= minx && maxy >= miny) inter = __dmul_rn(__dsub_rn(maxx, minx), __dsub_rn(maxy, miny));
  const double area_a = __dmul_rn(__dsub_rn(a[2], a[0]), __dsub_rn(a[3], a[1]));
  const double area_b = __dmul_rn(__dsub_rn(bx1, bx0), __dsub_rn(by1, by0));
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_b), inter));
}

// block-wide exclusive scans over LB_THREADS values (sum of ints / max of doubles), result + total
__device__ __forceinline__ int block_excl_sum(int v, int* warp_buf, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_buf[warp] = inc;
  __syncthreads();
  int base = 0;
  total = 0;
#pragma unroll
  for (int k = 0; k < LB_THREADS / 32; ++k) {
    const int s = warp_buf[k];
    if (k < warp) base += s;
    total += s;
  }
  return base + inc - v;
}
__device__ __forceinline__ double block_excl_max(double v, double init, double* warp_buf, double& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = fmax(inc, t);
  }
  double excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = init;
  __syncthreads();
  if (lane == 31) warp_buf[warp] = inc;
  __syncthreads();
  double base = init;
  total = init;
#pragma unroll
  for (int k = 0; k < LB_THREADS / 32; ++k) {
    const double s = warp_buf[k];
    if (k < warp) base = fmax(base, s);
    total = fmax(total, s);
  }
  return fmax(base, excl);
}

// ------------------------------------------------------------------------------------------------ findPositive
// One CTA per ROI.  Output of ROI r goes to out[r * cap_per_roi ...] (n_out[r] entries): first the positives in
// enumeration order, or -- when there is none and include_best -- the best set (Anchors.lua:171-186).
__global__ void __launch_bounds__(LB_THREADS) find_positive_kernel(FindPositiveParams p) {
  __shared__ float sw[MAX_LABEL_IJ * LUT_CELLS * 2];
  __shared__ float sh[MAX_LABEL_IJ * LUT_CELLS * 2];
  __shared__ Range ranges[MAX_LABEL_IJ];
  __shared__ int r_ij[MAX_LABEL_IJ];
  __shared__ int r_off[MAX_LABEL_IJ + 1];
  __shared__ int n_ranges;
  __shared__ int warp_i[LB_THREADS / 32];
  __shared__ double warp_d[LB_THREADS / 32];
  const int roi = blockIdx.x;
  const int nij = p.n_scales * 3;
  for (int i = threadIdx.x; i < nij * LUT_CELLS * 2; i += LB_THREADS) {
    sw[i] = p.w_lut[i];
    sh[i] = p.h_lut[i];
  }
  __syncthreads();
  double rect[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) rect[k] = p.rois[(size_t)roi * 4 + k];
  if (threadIdx.x == 0) {
    int n = 0, off = 0;
    for (int ij = 0; ij < nij; ++ij) {
      Range r;
      if (find_range(sw, sh, ij, rect, p.has_clip ? p.clip : nullptr, r)) {
        ranges[n] = r;
        r_ij[n] = ij;
        r_off[n] = off;
        off += (r.ux - r.lx) * (r.uy - r.ly);
        ++n;
      }
    }
    r_off[n] = off;
    n_ranges = n;
  }
  __syncthreads();
  const int total = r_off[n_ranges];
  frcnn_anchor_ref* out = p.out + (size_t)roi * p.cap_per_roi;
  frcnn_anchor_ref* best = p.best_scratch + (size_t)roi * p.cap_per_roi;
  int n_pos = 0, n_best = 0;       // identical in every thread
  double run_max = -1.0;           // best_iou (Anchors.lua:153)
  for (int base = 0; base < total; base += LB_THREADS) {
    const int t = base + threadIdx.x;
    double v = -2.0;
    frcnn_anchor_ref a;
    a.layer = a.aspect = a.y = a.x = 0;
    if (t < total) {
      int j = 0;
      while (t >= r_off[j + 1]) ++j;
      const Range r = ranges[j];
      const int ij = r_ij[j];
      const int nx = r.ux - r.lx;
      const int local = t - r_off[j];
      const int y = r.ly + local / nx, x = r.lx + local % nx;   // 1-based LUT cells; x runs fastest (Anchors.lua:164-166)
      const float* w = sw + ((size_t)ij * LUT_CELLS + (x - 1)) * 2;
      const float* h = sh + ((size_t)ij * LUT_CELLS + (y - 1)) * 2;
      v = rect_iou(rect, (double)w[0], (double)h[0], (double)w[1], (double)h[1]);
      a.layer = ij / 3 + 1; a.aspect = ij % 3 + 1; a.y = y; a.x = x;
    }
    const bool is_pos = t < total && v > p.pos_threshold;
    const bool elig = t < total && !is_pos && v > p.neg_threshold;
    // positives: ordered compaction
    int chunk_pos;
    const int pos_at = block_excl_sum(is_pos ? 1 : 0, warp_i, chunk_pos);
    if (is_pos) {
      if (n_pos + pos_at < p.cap_per_roi) out[n_pos + pos_at] = a;
    }
    n_pos += chunk_pos;
    // best set: record when v >= running maximum of the earlier eligible candidates, reset when v - 0.025 exceeds it
    if (p.include_best) {
      double chunk_max;
      const double prev = block_excl_max(elig ? v : -1.0, run_max, warp_d, chunk_max);
      const bool rec = elig && v >= prev;
      const bool rst = rec && (__dsub_rn(v, 0.025) > prev);
      // last reset of the chunk: everything recorded before it is dropped
      int last_rst;
      {
        int cand = rst ? threadIdx.x : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) warp_i[threadIdx.x >> 5] = cand;
        __syncthreads();
        last_rst = -1;
#pragma unroll
        for (int k = 0; k < LB_THREADS / 32; ++k) last_rst = max(last_rst, warp_i[k]);
      }
      const bool keep = rec && (int)threadIdx.x >= last_rst;
      int chunk_keep;
      const int keep_at = block_excl_sum(keep ? 1 : 0, warp_i, chunk_keep);
      if (last_rst >= 0) n_best = 0;
      if (keep) {
        if (n_best + keep_at < p.cap_per_roi) best[n_best + keep_at] = a;
      }
      n_best += chunk_keep;
      run_max = chunk_max;
    }
  }
  __syncthreads();
  int n_out = n_pos;
  if (n_pos == 0 && p.include_best && n_best > 0 && run_max > 0.0) {   // Anchors.lua:183-187
    n_out = min(n_best, p.cap_per_roi);
    for (int i = threadIdx.x; i < n_out; i += LB_THREADS) out[i] = best[i];
  }
  if (threadIdx.x == 0) {
    p.n_out[roi] = min(n_out, p.cap_per_roi);
    if (n_pos > p.cap_per_roi || (n_pos == 0 && p.include_best && n_best > p.cap_per_roi)) atomicExch(p.status, 1);
  }
}

void launch_find_positive(const FindPositiveParams& p, int n_rois, cudaStream_t st) {
  if (n_rois <= 0) return;
  find_positive_kernel<<<n_rois, LB_THREADS, 0, st>>>(p);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ sampleNegative
// Anchors:sampleNegative (Anchors.lua:197-235) with the random stream supplied by the caller: trial t consumes the
// three consecutive torch.random() values rnd[3t .. 3t+2] (range, x, y).  All trials are evaluated in parallel; the
// sequential stopping rule (count accepted, or 500 consecutive rejections) is then resolved with scans.  One CTA.
__global__ void __launch_bounds__(LB_THREADS) sample_negative_kernel(SampleNegativeParams p) {
  __shared__ float sw[MAX_LABEL_IJ * LUT_CELLS * 2];
  __shared__ float sh[MAX_LABEL_IJ * LUT_CELLS * 2];
  __shared__ Range ranges[MAX_LABEL_IJ];
  __shared__ int r_ij[MAX_LABEL_IJ];
  __shared__ int n_ranges;
  __shared__ int warp_i[LB_THREADS / 32];
  const int nij = p.n_scales * 3;
  for (int i = threadIdx.x; i < nij * LUT_CELLS * 2; i += LB_THREADS) {
    sw[i] = p.w_lut[i];
    sh[i] = p.h_lut[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int ij = 0; ij < nij; ++ij) {
      Range r;
      if (find_range(sw, sh, ij, p.image_rect, p.image_rect, r)) {   // findRangesXY(image_rect, image_rect)
        ranges[n] = r;
        r_ij[n] = ij;
        ++n;
      }
    }
    n_ranges = n;
  }
  __syncthreads();
  int n_neg = 0, retry = p.retry_in, consumed = 0;   // identical in every thread
  bool done = n_ranges == 0 || p.count <= 0 || retry >= 500;
  for (int base = 0; base < p.n_trials && !done; base += LB_THREADS) {
    const int t = base + threadIdx.x;
    bool live = t < p.n_trials, reject = false;
    frcnn_anchor_ref a;
    a.layer = a.aspect = a.y = a.x = 0;
    if (live) {
      const Range r = ranges[p.rnd[3 * (size_t)t] % (uint32_t)n_ranges];
      const int ij = r_ij[p.rnd[3 * (size_t)t] % (uint32_t)n_ranges];
      const int x = r.lx + (int)(p.rnd[3 * (size_t)t + 1] % (uint32_t)(r.ux - r.lx));
      const int y = r.ly + (int)(p.rnd[3 * (size_t)t + 2] % (uint32_t)(r.uy - r.ly));
      const float* w = sw + ((size_t)ij * LUT_CELLS + (x - 1)) * 2;
      const float* h = sh + ((size_t)ij * LUT_CELLS + (y - 1)) * 2;
      for (int j = 0; j < p.n_rois && !reject; ++j) {
        double rect[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rect[k] = p.rois[(size_t)j * 4 + k];
        reject = rect_iou(rect, (double)w[0], (double)h[0], (double)w[1], (double)h[1]) > p.neg_threshold;
      }
      a.layer = ij / 3 + 1; a.aspect = ij % 3 + 1; a.y = y; a.x = x;
    }
    const bool acc = live && !reject;
    // sequential rule inside the chunk: the loop stops BEFORE trial t when #neg == count or retry == 500 at that point.
    // #neg before t = n_neg + (accepted before t in chunk); retry before t = run of rejections ending at t-1.
    int chunk_acc;
    const int acc_before = block_excl_sum(acc ? 1 : 0, warp_i, chunk_acc);
    // index (in chunk) of the last accepted trial before t, or -1: retry_before = (t_local - 1 - last_acc) [+ carried]
    int last_acc_incl = acc ? (int)threadIdx.x : -1;
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, last_acc_incl, o);
        if (lane >= o) last_acc_incl = max(last_acc_incl, v);
      }
      __syncthreads();
      if (lane == 31) warp_i[warp] = last_acc_incl;
      __syncthreads();
      int basev = -1;
#pragma unroll
      for (int k = 0; k < LB_THREADS / 32; ++k)
        if (k < warp) basev = max(basev, warp_i[k]);
      last_acc_incl = max(last_acc_incl, basev);
    }
    int last_acc_before;
    {
      // value of the previous thread across warp boundaries
      __shared__ int prev_buf[LB_THREADS];
      prev_buf[threadIdx.x] = last_acc_incl;
      __syncthreads();
      last_acc_before = threadIdx.x == 0 ? -1 : prev_buf[threadIdx.x - 1];
      __syncthreads();
    }
    const int retry_before = last_acc_before >= 0 ? ((int)threadIdx.x - 1 - last_acc_before) : (retry + (int)threadIdx.x);
    const bool stop_here = live && ((n_neg + acc_before >= p.count) || (retry_before >= 500));
    // first stopping trial of the chunk
    int first_stop = stop_here ? (int)threadIdx.x : LB_THREADS;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_stop = min(first_stop, __shfl_xor_sync(0xffffffffu, first_stop, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_i[threadIdx.x >> 5] = first_stop;
    __syncthreads();
    first_stop = LB_THREADS;
#pragma unroll
    for (int k = 0; k < LB_THREADS / 32; ++k) first_stop = min(first_stop, warp_i[k]);
    const int chunk_live = min(LB_THREADS, p.n_trials - base);
    const int executed = min(first_stop, chunk_live);   // trials of this chunk the reference loop actually runs
    if (acc && (int)threadIdx.x < executed && n_neg + acc_before < p.cap) p.out[n_neg + acc_before] = a;
    // carry the sequential state past the executed part of the chunk
    int acc_exec;
    {
      int tot;
      const int e = block_excl_sum((acc && (int)threadIdx.x < executed) ? 1 : 0, warp_i, tot);
      (void)e;
      acc_exec = tot;
    }
    // retry after the executed part: distance from the last accepted executed trial
    int last_exec_acc = (acc && (int)threadIdx.x < executed) ? (int)threadIdx.x : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last_exec_acc = max(last_exec_acc, __shfl_xor_sync(0xffffffffu, last_exec_acc, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_i[threadIdx.x >> 5] = last_exec_acc;
    __syncthreads();
    last_exec_acc = -1;
#pragma unroll
    for (int k = 0; k < LB_THREADS / 32; ++k) last_exec_acc = max(last_exec_acc, warp_i[k]);
    retry = last_exec_acc >= 0 ? (executed - 1 - last_exec_acc) : (retry + executed);
    n_neg += acc_exec;
    consumed += executed;
    if (first_stop < chunk_live) done = true;
    __syncthreads();
  }
  if (!done && (n_neg >= p.count || retry >= 500)) done = true;   // the rule fired exactly at the end of the stream
  if (threadIdx.x == 0) {
    p.result[0] = min(n_neg, p.cap);
    p.result[1] = consumed;
    p.result[2] = done ? 1 : 0;     // 0: the random stream ran out before the loop's stopping rule fired
    p.result[3] = n_ranges;
    p.result[4] = retry;            // consecutive rejections at the end (continue a run with retry_in = this)
  }
}

void launch_sample_negative(const SampleNegativeParams& p, cudaStream_t st) {
  sample_negative_kernel<<<1, LB_THREADS, 0, st>>>(p);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------- findNearby
// BatchIterator.lua:206-217 (cfg.nearby_aversion): for every positive anchor p, Anchors:findNearby(p:center()) -- the
// anchors whose cell centre shares p's 16-pixel bin on both axes (Anchors.lua:22-30,69-84), in the order the Lua tables
// yield them: (scale, aspect, y cell, x cell) ascending -- kept when Rect.IoU(p, a) < negative_threshold.  One CTA, a thread
// per positive, ordered output through a block scan (positives in list order, at most 4 x 3 x 2 x 2 entries each).
static constexpr int NEARBY_BIN = 16;   // Anchors.lua:5
static constexpr int NEARBY_MAX = 64;   // per positive and (scale, aspect): cells per bin along an axis are <= 16 / stride

__device__ __forceinline__ void anchor_rect(const float* w_lut, const float* h_lut, int ij, int y, int x, double r[4]) {
  const float* w = w_lut + ((size_t)ij * LUT_CELLS + (x - 1)) * 2;
  const float* h = h_lut + ((size_t)ij * LUT_CELLS + (y - 1)) * 2;
  r[0] = (double)w[0]; r[1] = (double)h[0]; r[2] = (double)w[1]; r[3] = (double)h[1];
}

template <bool WRITE>
__device__ __forceinline__ int nearby_of(const FindNearbyParams& p, int pi, frcnn_anchor_ref* out, int* out_pos, int at) {
  const frcnn_anchor_ref a = p.pos[pi];
  double pr[4];
  anchor_rect(p.w_lut, p.h_lut, (a.layer - 1) * 3 + (a.aspect - 1), a.y, a.x, pr);
  const double cx = (pr[0] + pr[2]) / 2, cy = (pr[1] + pr[3]) / 2;     // Rect:center (Rect.lua:65-67)
  const double bx = floor(cx / NEARBY_BIN), by = floor(cy / NEARBY_BIN);
  int n = 0;
  for (int i = 0; i < p.n_scales; ++i) {
    const double* ceny = p.cen_y + (size_t)i * LUT_CELLS;
    const double* cenx = p.cen_x + (size_t)i * LUT_CELLS;
    for (int j = 0; j < 3; ++j) {
      for (int vy = 1; vy <= LUT_CELLS; ++vy) {
        if (floor(ceny[vy - 1] / NEARBY_BIN) != by) continue;
        for (int vx = 1; vx <= LUT_CELLS; ++vx) {
          if (floor(cenx[vx - 1] / NEARBY_BIN) != bx) continue;
          double r[4];
          anchor_rect(p.w_lut, p.h_lut, i * 3 + j, vy, vx, r);
          if (rect_iou(pr, r[0], r[1], r[2], r[3]) < p.neg_threshold) {
            if (WRITE && at + n < p.cap) {
              out[at + n] = frcnn_anchor_ref{i + 1, j + 1, vy, vx};
              out_pos[at + n] = pi;
            }
            ++n;
          }
        }
      }
    }
  }
  return n;
}

__global__ void __launch_bounds__(LB_THREADS) find_nearby_kernel(FindNearbyParams p) {
  __shared__ int warp_buf[LB_THREADS / 32];
  int base = 0;
  for (int p0 = 0; p0 < p.n_pos; p0 += LB_THREADS) {
    const int pi = p0 + (int)threadIdx.x;
    const int mine = pi < p.n_pos ? nearby_of<false>(p, pi, nullptr, nullptr, 0) : 0;
    int total;
    const int at = block_excl_sum(mine, warp_buf, total);
    if (pi < p.n_pos && mine) nearby_of<true>(p, pi, p.out, p.out_pos, base + at);
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.result[0] = base;
}
void launch_find_nearby(const FindNearbyParams& p, cudaStream_t st) { find_nearby_kernel<<<1, LB_THREADS, 0, st>>>(p); }

}  // namespace frcnn
