// Exact greedy non-maximum suppression on the GPU, batched over class segments.
//
// Replaces the global nms(boxes, overlap, scores) of the reference (nms.lua:23-102) and the per-class loop around
// it (Detector.lua:125-136).  The result is bit-identical to the sequential algorithm: same fp32 operation order
// for area / intersection / IoU (nms.lua:35,85-94, no FMA contraction, IEEE division), same `IoU <= overlap`
// keep rule (nms.lua:96), same processing order (ascending sort, pop from the end; ties broken by (key, index),
// see oracle/nms.py), same pick order.
//
// Algorithm (per segment, all segments concurrently):
//   sort      positions in priority order (key descending, index descending)
//   rounds    until no candidate is undecided:
//     select    the first B (=1024) still-alive candidates in priority order
//     mask      B x B lower-triangular suppression bit matrix of the selection (all SMs)
//     resolve   the greedy recursion "kept(j) <=> no kept i<j suppresses j" evaluated as a fixed point over the
//               bit matrix in shared memory (each sweep decides every candidate whose predecessors are decided),
//               appends the keepers to the pick list in priority order
//     filter    every later candidate is tested against this round's keepers (all SMs); suppressed ones die
//   Only candidates that were alive against ALL earlier keepers are ever selected, so the result is exactly the
//   sequential greedy one while the expensive N x kept tests run fully parallel.
// For segments of at most 8192 boxes (the detector's case) sorting happens inside one CTA and the number of rounds
// is bounded on the host, so the whole NMS is a fixed launch sequence without any host synchronisation.
#include "common.h"
#include "nms.h"

namespace frcnn {

static constexpr int NMS_B = 1024;          // selection size per round
static constexpr int NMS_ROW_WORDS = NMS_B / 32;
static constexpr int SORT_CTA_MAX = 8192;   // largest segment sorted inside one CTA

__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// nms.lua:35: area = (x2 - x1 + 1) * (y2 - y1 + 1), fp32 op by op
__device__ __forceinline__ float box_area(float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// true iff the reference would drop box j after picking box i: NOT (IoU <= overlap)   (nms.lua:72-96)
__device__ __forceinline__ bool suppresses(float4 bi, float ai, float4 bj, float aj, float thr) {
  float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);
  float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
  float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  float inter = __fmul_rn(w, h);
  float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aj, ai), inter));
  return !(iou <= thr);
}

__device__ __forceinline__ float4 load_box(const float* boxes, long row, int row_stride) {
  const float* p = boxes + row * row_stride;
  return make_float4(p[0], p[1], p[2], p[3]);
}

__device__ __forceinline__ float order_key(const float* boxes, long row, int row_stride, int order_mode, int order_col) {
  if (order_mode == FRCNN_NMS_ORDER_AREA) return box_area(load_box(boxes, row, row_stride));
  if (order_mode == FRCNN_NMS_ORDER_COLUMN) return boxes[row * row_stride + order_col];
  return boxes[row * row_stride + 3];  // nms.lua:41-42: y2
}

// ------------------------------------------------------------------------------------------------ init
// Resets the per-segment state; with n_total_dev the segment table is a single segment [0, *n_total_dev).
__global__ void nms_init_kernel(NmsState st, int n_seg) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg) {
    st.cursor[s] = 0;
    st.counts[s] = 0;
    st.newk_cnt[s] = 0;
    st.sel_cnt[s] = 0;
  }
  if (s == 0) *st.remaining = 0;
}

// ------------------------------------------------------------------------------------------------ sort (one CTA)
// Bitonic sort of the 64-bit keys (orderable(key) << 32 | index), descending, in shared memory.
__global__ void __launch_bounds__(1024) nms_sort_cta_kernel(NmsState st, const float* __restrict__ boxes, int row_stride,
                                                            int order_mode, int order_col, int cap_len) {
  extern __shared__ unsigned long long skeys[];
  const int s = blockIdx.x;
  const int beg = st.seg_beg[s];
  const int len = min(st.seg_len[s], cap_len);  // cap_len: what the shared-memory allocation holds
  if (len <= 0) return;
  int n2 = 1;
  while (n2 < len) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < len) k = ((unsigned long long)orderable(order_key(boxes, beg + i, row_stride, order_mode, order_col)) << 32) | (unsigned)i;
    skeys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` cleared
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = skeys[lo], b = skeys[hi];
        if ((a < b) == desc) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    st.order[beg + i] = (int)(unsigned)(skeys[i] & 0xffffffffull);
    st.alive[beg + i] = 1;
  }
}

// ------------------------------------------------------------------------------------------------ sort (global)
// LSD radix sort, 8 bits per pass, stable.  Keys: pass 0..3 = bytes of ~orderable(key) (ascending => key
// descending), pass 4 = segment id.  The initial arrangement lists every segment in DESCENDING index order, so that
// stability yields the (key desc, index desc) priority order.
__global__ void radix_prepare_kernel(NmsState st, const float* __restrict__ boxes, int row_stride, int order_mode,
                                     int order_col, int n_seg, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                     uint8_t* __restrict__ segid) {
  int s = blockIdx.y;
  int beg = st.seg_beg[s];
  int len = st.seg_len[s];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < len; p += gridDim.x * blockDim.x) {
    int idx = len - 1 - p;
    keys[beg + p] = ~orderable(order_key(boxes, beg + idx, row_stride, order_mode, order_col));
    vals[beg + p] = (uint32_t)idx;
    segid[beg + p] = (uint8_t)s;
    st.alive[beg + p] = 1;
  }
}

static constexpr int RADIX_TILE = 2048;  // items per block
static constexpr int RADIX_THREADS = 256;

// pass < 4: digit = byte `pass` of keys[i]; pass == 4: digit = segment id carried alongside
__global__ void __launch_bounds__(RADIX_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys,
                                                                    const uint8_t* __restrict__ segid, int n, int pass,
                                                                    uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * RADIX_TILE;
  for (int i = base + threadIdx.x; i < min(n, base + RADIX_TILE); i += RADIX_THREADS) {
    uint32_t d = pass < 4 ? ((keys[i] >> (8 * pass)) & 255u) : (uint32_t)segid[i];
    atomicAdd(&h[d], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist[256 * nblocks] in place (single CTA)
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ hist, int total) {
  __shared__ uint32_t partial[1024];
  int per = (total + 1023) / 1024;
  int b = threadIdx.x * per, e = min(total, b + per);
  uint32_t s = 0;
  for (int i = b; i < e; ++i) s += hist[i];
  partial[threadIdx.x] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = threadIdx.x >= off ? partial[threadIdx.x - off] : 0;
    __syncthreads();
    partial[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = partial[threadIdx.x] - s;
  for (int i = b; i < e; ++i) {
    uint32_t v = hist[i];
    hist[i] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(RADIX_THREADS) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                       const uint32_t* __restrict__ vals_in,
                                                                       const uint8_t* __restrict__ seg_in,
                                                                       uint32_t* __restrict__ keys_out,
                                                                       uint32_t* __restrict__ vals_out,
                                                                       uint8_t* __restrict__ seg_out, int n, int pass,
                                                                       const uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t digit_base[256];            // running global offset of each digit for this block
  __shared__ uint32_t warp_hist[RADIX_THREADS / 32][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  digit_base[threadIdx.x] = hist[threadIdx.x * nblocks + blockIdx.x];
  int base = blockIdx.x * RADIX_TILE;
  int end = min(n, base + RADIX_TILE);
  for (int sub = base; sub < end; sub += RADIX_THREADS) {
    for (int w = 0; w < RADIX_THREADS / 32; ++w) warp_hist[w][threadIdx.x] = 0;
    __syncthreads();
    int i = sub + threadIdx.x;
    bool valid = i < end;
    uint32_t k = 0, v = 0, d = 0;
    uint8_t sg = 0;
    if (valid) {
      k = keys_in[i];
      v = vals_in[i];
      sg = seg_in[i];
      d = pass < 4 ? ((k >> (8 * pass)) & 255u) : (uint32_t)sg;
    }
    // rank inside the warp among equal digits (stable: lower lanes first)
    uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
    uint32_t lower = peers & ((1u << lane) - 1u);
    uint32_t rank_in_warp = __popc(lower);
    if (valid && lower == 0) warp_hist[warp][d] = __popc(peers);
    __syncthreads();
    // per digit: exclusive prefix over warps, and advance the running base
    {
      uint32_t dsum = 0;
      for (int w = 0; w < RADIX_THREADS / 32; ++w) {
        uint32_t c = warp_hist[w][threadIdx.x];
        warp_hist[w][threadIdx.x] = dsum;
        dsum += c;
      }
      // digit_base is advanced after the scatter below (needs the old value) -> stash the sum in a register
      __syncthreads();
      if (valid) {
        uint32_t pos = digit_base[d] + warp_hist[warp][d] + rank_in_warp;
        keys_out[pos] = k;
        vals_out[pos] = v;
        seg_out[pos] = sg;
      }
      __syncthreads();
      digit_base[threadIdx.x] += dsum;
    }
    __syncthreads();
  }
}

__global__ void radix_finish_kernel(NmsState st, const uint32_t* __restrict__ vals, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st.order[i] = (int)vals[i];
}

// ------------------------------------------------------------------------------------------------ select
// Gathers the first NMS_B alive candidates (priority order) of every segment, starting at cursor[s].
__global__ void __launch_bounds__(1024) nms_select_kernel(NmsState st, const float* __restrict__ boxes, int row_stride) {
  __shared__ int warp_cnt[32];
  __shared__ int s_taken, s_next;
  const int s = blockIdx.x;
  const int beg = st.seg_beg[s];
  const int len = st.seg_len[s];
  const int cur = st.cursor[s];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    s_taken = 0;
    s_next = len;
  }
  __syncthreads();
  if (cur >= len) {
    if (threadIdx.x == 0) st.sel_cnt[s] = 0;
    return;
  }
  int taken = 0;
  for (int base = cur; base < len && taken < NMS_B; base += 1024) {
    int pos = base + threadIdx.x;
    bool a = pos < len && st.alive[beg + pos] != 0;
    unsigned bal = __ballot_sync(0xffffffffu, a);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, chunk_total = 0;
    for (int w = 0; w < 32; ++w) {
      int c = warp_cnt[w];
      if (w < warp) before += c;
      chunk_total += c;
    }
    int slot = taken + before + __popc(bal & ((1u << lane) - 1u));
    if (a && slot < NMS_B) {
      int local = st.order[beg + pos];
      float4 b = load_box(boxes, beg + local, row_stride);
      st.sel_pos[(long)s * NMS_B + slot] = pos;
      st.sel_box[(long)s * NMS_B + slot] = b;
      st.sel_area[(long)s * NMS_B + slot] = box_area(b);
      if (slot == NMS_B - 1) s_next = pos + 1;  // selection full: everything before pos+1 is decided or selected
    }
    taken += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    st.sel_cnt[s] = min(taken, NMS_B);
    st.cursor[s] = s_next;
  }
}

// ------------------------------------------------------------------------------------------------ mask
// mask[s][j][wi] bit b  <=>  selected candidate i = 32*wi + b (i < j) would suppress selected candidate j.
// Thread <-> (j, wi) with consecutive threads on consecutive j: box i is a warp-wide broadcast.
__global__ void __launch_bounds__(256) nms_mask_kernel(NmsState st, float thr) {
  const int s = blockIdx.y;
  const int m = st.sel_cnt[s];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = idx % NMS_B;
  const int wi = idx / NMS_B;
  const int j_warp_max = (j | 31);
  if (wi * 32 > j_warp_max || (j & ~31) >= m) return;  // warp-uniform: upper triangle or beyond the selection
  const float4* sb = st.sel_box + (long)s * NMS_B;
  const float* sa = st.sel_area + (long)s * NMS_B;
  uint32_t word = 0;
  if (j < m) {
    float4 bj = sb[j];
    float aj = sa[j];
    int i_end = min(32, j - wi * 32);  // only i < j
    for (int b = 0; b < i_end; ++b) {
      int i = wi * 32 + b;
      if (suppresses(sb[i], sa[i], bj, aj, thr)) word |= 1u << b;
    }
  }
  if (j < m) st.mask[((long)s * NMS_B + j) * NMS_ROW_WORDS + wi] = word;
}

// ------------------------------------------------------------------------------------------------ resolve
__global__ void __launch_bounds__(1024) nms_resolve_kernel(NmsState st) {
  extern __shared__ uint32_t smask[];  // [NMS_B][33] padded rows (bank-conflict-free column walks)
  __shared__ uint32_t kept[NMS_ROW_WORDS], undec[NMS_ROW_WORDS];
  __shared__ int warp_cnt[32];
  const int s = blockIdx.x;
  const int m = st.sel_cnt[s];
  const int j = threadIdx.x;
  const int warp = j >> 5, lane = j & 31;
  if (m == 0) {
    if (j == 0) st.newk_cnt[s] = 0;
    return;
  }
  const int beg = st.seg_beg[s];
  // rows < m, words <= row/32 are valid in global memory; everything else is treated as zero
  for (int idx = j; idx < NMS_B * NMS_ROW_WORDS; idx += 1024) {
    int r = idx >> 5, w = idx & 31;
    uint32_t v = 0;
    if (r < m && w * 32 <= r) v = st.mask[((long)s * NMS_B + r) * NMS_ROW_WORDS + w];
    smask[r * 33 + w] = v;
  }
  if (j < NMS_ROW_WORDS) {
    kept[j] = 0;
    int lo = j * 32;
    undec[j] = m >= lo + 32 ? 0xffffffffu : (m > lo ? ((1u << (m - lo)) - 1u) : 0u);
  }
  __syncthreads();
  const int nwords = (j >> 5) + 1;  // predecessors of j live in words 0..j/32
  bool undecided = j < m;
  bool is_kept = false;
  for (;;) {
    int decision = 0;  // 1 keep, 2 drop
    if (undecided) {
      bool hit_kept = false, hit_undec = false;
      for (int w = 0; w < nwords; ++w) {
        uint32_t row = smask[j * 33 + w];
        hit_kept |= (row & kept[w]) != 0;
        hit_undec |= (row & undec[w]) != 0;
      }
      if (hit_kept) decision = 2;
      else if (!hit_undec) decision = 1;
    }
    int progress = __syncthreads_or(decision != 0);  // also separates the read phase from the write phase
    if (decision != 0) {
      atomicAnd(&undec[warp], ~(1u << lane));
      if (decision == 1) {
        atomicOr(&kept[warp], 1u << lane);
        is_kept = true;
      }
      undecided = false;
    }
    int any_left = __syncthreads_or(undecided);
    if (!any_left) break;
    if (!progress) break;  // cannot happen (the first undecided candidate is always decidable); guards against hangs
  }
  // append the keepers in priority order
  unsigned bal = __ballot_sync(0xffffffffu, is_kept);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    int c = warp_cnt[w];
    if (w < warp) before += c;
    total += c;
  }
  const int base_count = st.counts[s];
  if (j < m) {
    int pos = st.sel_pos[(long)s * NMS_B + j];
    st.alive[beg + pos] = 0;  // every selected candidate is decided now
    if (is_kept) {
      int slot = before + __popc(bal & ((1u << lane) - 1u));
      st.pick[beg + base_count + slot] = st.order[beg + pos];
      st.newk_box[(long)s * NMS_B + slot] = st.sel_box[(long)s * NMS_B + j];
      st.newk_area[(long)s * NMS_B + slot] = st.sel_area[(long)s * NMS_B + j];
    }
  }
  __syncthreads();
  if (j == 0) {
    st.counts[s] = base_count + total;
    st.newk_cnt[s] = total;
  }
}

// ------------------------------------------------------------------------------------------------ filter
// Tests every still-alive candidate behind the cursor against this round's keepers.
__global__ void __launch_bounds__(256) nms_filter_kernel(NmsState st, const float* __restrict__ boxes, int row_stride, float thr) {
  __shared__ float4 kb[256];
  __shared__ float ka[256];
  const int s = blockIdx.y;
  const int beg = st.seg_beg[s];
  const int len = st.seg_len[s];
  const int cur = st.cursor[s];
  const int nk = st.newk_cnt[s];
  if (cur >= len) return;
  const int per_block = (len - cur + gridDim.x - 1) / gridDim.x;
  const int p_beg = cur + blockIdx.x * per_block;
  const int p_end = min(len, p_beg + per_block);
  if (p_beg >= p_end) return;
  bool any_alive = false;
  for (int base = p_beg; base < p_end; base += 256) {
    int pos = base + threadIdx.x;
    bool a = pos < p_end && st.alive[beg + pos] != 0;
    float4 bj = make_float4(0, 0, 0, 0);
    float aj = 0;
    if (a) {
      bj = load_box(boxes, beg + st.order[beg + pos], row_stride);
      aj = box_area(bj);
    }
    for (int k0 = 0; k0 < nk; k0 += 256) {
      __syncthreads();
      if (k0 + threadIdx.x < nk) {
        kb[threadIdx.x] = st.newk_box[(long)s * NMS_B + k0 + threadIdx.x];
        ka[threadIdx.x] = st.newk_area[(long)s * NMS_B + k0 + threadIdx.x];
      }
      __syncthreads();
      int kn = min(256, nk - k0);
      if (a) {
        for (int k = 0; k < kn; ++k) {
          if (suppresses(kb[k], ka[k], bj, aj, thr)) {
            a = false;
            break;
          }
        }
      }
    }
    if (pos < p_end && !a && st.alive[beg + pos] != 0) st.alive[beg + pos] = 0;
    any_alive |= a;
  }
  if (__syncthreads_or(any_alive) && threadIdx.x == 0) atomicOr(st.remaining, 1);
}

// ------------------------------------------------------------------------------------------------ CTA-level path
// Segments of at most SORT_CTA_MAX boxes (the detector's case): one CTA per segment runs whole algorithm steps
// back to back instead of one launch per step.  Three launches per NMS -- sort + first selection (one CTA per
// segment), the 1024 x 1024 suppression matrix of the first selection on all SMs, resolve + filter + every further
// round inside the CTA -- or ONE launch (fused) when the segments are known to be small (per-class NMS).
__device__ void cta_sort(unsigned long long* skeys, const NmsState& st, int s, const float* __restrict__ boxes, int row_stride,
                         int order_mode, int order_col, int cap_len) {
  const int beg = st.seg_beg[s];
  const int len = min(st.seg_len[s], cap_len);
  if (len <= 0) return;  // uniform
  int n2 = 1;
  while (n2 < len) n2 <<= 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < len) k = ((unsigned long long)orderable(order_key(boxes, beg + i, row_stride, order_mode, order_col)) << 32) | (unsigned)i;
    skeys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        unsigned long long a = skeys[lo], b = skeys[hi];
        if ((a < b) == desc) {
          skeys[lo] = b;
          skeys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    st.order[beg + i] = (int)(unsigned)(skeys[i] & 0xffffffffull);
    st.alive[beg + i] = 1;
  }
  __syncthreads();
}

// first NMS_B alive candidates behind the cursor -> sel_*; returns the selection size (uniform)
__device__ int cta_select(const NmsState& st, int s, const float* __restrict__ boxes, int row_stride) {
  __shared__ int warp_cnt[32];
  __shared__ int s_next;
  const int beg = st.seg_beg[s];
  const int len = st.seg_len[s];
  const int cur = st.cursor[s];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (threadIdx.x == 0) s_next = len;
  __syncthreads();
  int taken = 0;
  for (int base = cur; base < len && taken < NMS_B; base += 1024) {
    int pos = base + threadIdx.x;
    bool a = pos < len && st.alive[beg + pos] != 0;
    unsigned bal = __ballot_sync(0xffffffffu, a);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, chunk_total = 0;
    for (int w = 0; w < 32; ++w) {
      int c = warp_cnt[w];
      if (w < warp) before += c;
      chunk_total += c;
    }
    int slot = taken + before + __popc(bal & ((1u << lane) - 1u));
    if (a && slot < NMS_B) {
      int local = st.order[beg + pos];
      float4 b = load_box(boxes, beg + local, row_stride);
      st.sel_pos[(long)s * NMS_B + slot] = pos;
      st.sel_box[(long)s * NMS_B + slot] = b;
      st.sel_area[(long)s * NMS_B + slot] = box_area(b);
      if (slot == NMS_B - 1) s_next = pos + 1;
    }
    taken += chunk_total;
    __syncthreads();
  }
  const int m = min(taken, NMS_B);
  if (threadIdx.x == 0) {
    st.sel_cnt[s] = m;
    st.cursor[s] = s_next;
  }
  __syncthreads();
  return m;
}

// suppression matrix of the selection computed by this CTA, straight into shared memory
__device__ void cta_mask(uint32_t* smask, const NmsState& st, int s, int m, float thr) {
  const float4* sb = st.sel_box + (long)s * NMS_B;
  const float* sa = st.sel_area + (long)s * NMS_B;
  const int j = threadIdx.x;
  if (j < m) {
    const float4 bj = sb[j];
    const float aj = sa[j];
    for (int wi = 0; wi < NMS_ROW_WORDS; ++wi) {
      uint32_t word = 0;
      const int i_end = min(32, j - wi * 32);
      for (int b = 0; b < i_end; ++b) {
        const int i = wi * 32 + b;
        if (suppresses(sb[i], sa[i], bj, aj, thr)) word |= 1u << b;
      }
      smask[j * 33 + wi] = word;
    }
  } else {
    for (int wi = 0; wi < NMS_ROW_WORDS; ++wi) smask[j * 33 + wi] = 0;
  }
  __syncthreads();
}

__device__ void cta_load_mask(uint32_t* smask, const NmsState& st, int s, int m) {
  for (int idx = threadIdx.x; idx < NMS_B * NMS_ROW_WORDS; idx += 1024) {
    int r = idx >> 5, w = idx & 31;
    uint32_t v = 0;
    if (r < m && w * 32 <= r) v = st.mask[((long)s * NMS_B + r) * NMS_ROW_WORDS + w];
    smask[r * 33 + w] = v;
  }
  __syncthreads();
}

// greedy recursion over the selection as a fixed point; appends the keepers; returns their number (uniform)
__device__ int cta_resolve(const uint32_t* smask, const NmsState& st, int s, int m) {
  __shared__ uint32_t kept[NMS_ROW_WORDS], undec[NMS_ROW_WORDS];
  __shared__ int warp_cnt[32];
  const int j = threadIdx.x;
  const int warp = j >> 5, lane = j & 31;
  const int beg = st.seg_beg[s];
  __syncthreads();
  if (j < NMS_ROW_WORDS) {
    kept[j] = 0;
    int lo = j * 32;
    undec[j] = m >= lo + 32 ? 0xffffffffu : (m > lo ? ((1u << (m - lo)) - 1u) : 0u);
  }
  __syncthreads();
  const int nwords = (j >> 5) + 1;
  bool undecided = j < m;
  bool is_kept = false;
  for (;;) {
    int decision = 0;
    if (undecided) {
      bool hit_kept = false, hit_undec = false;
      for (int w = 0; w < nwords; ++w) {
        uint32_t row = smask[j * 33 + w];
        hit_kept |= (row & kept[w]) != 0;
        hit_undec |= (row & undec[w]) != 0;
      }
      if (hit_kept) decision = 2;
      else if (!hit_undec) decision = 1;
    }
    int progress = __syncthreads_or(decision != 0);
    if (decision != 0) {
      atomicAnd(&undec[warp], ~(1u << lane));
      if (decision == 1) {
        atomicOr(&kept[warp], 1u << lane);
        is_kept = true;
      }
      undecided = false;
    }
    int any_left = __syncthreads_or(undecided);
    if (!any_left) break;
    if (!progress) break;
  }
  unsigned bal = __ballot_sync(0xffffffffu, is_kept);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    int c = warp_cnt[w];
    if (w < warp) before += c;
    total += c;
  }
  const int base_count = st.counts[s];
  if (j < m) {
    int pos = st.sel_pos[(long)s * NMS_B + j];
    st.alive[beg + pos] = 0;
    if (is_kept) {
      int slot = before + __popc(bal & ((1u << lane) - 1u));
      st.pick[beg + base_count + slot] = st.order[beg + pos];
      st.newk_box[(long)s * NMS_B + slot] = st.sel_box[(long)s * NMS_B + j];
      st.newk_area[(long)s * NMS_B + slot] = st.sel_area[(long)s * NMS_B + j];
    }
  }
  __syncthreads();
  if (j == 0) {
    st.counts[s] = base_count + total;
    st.newk_cnt[s] = total;
  }
  __syncthreads();
  return total;
}

// every alive candidate behind the cursor is tested against this round's nk keepers
__device__ void cta_filter(const NmsState& st, int s, const float* __restrict__ boxes, int row_stride, float thr, int nk) {
  __shared__ float4 kb[256];
  __shared__ float ka[256];
  const int beg = st.seg_beg[s];
  const int len = st.seg_len[s];
  const int cur = st.cursor[s];
  for (int base = cur; base < len; base += 1024) {
    int pos = base + threadIdx.x;
    bool a = pos < len && st.alive[beg + pos] != 0;
    float4 bj = make_float4(0, 0, 0, 0);
    float aj = 0;
    if (a) {
      bj = load_box(boxes, beg + st.order[beg + pos], row_stride);
      aj = box_area(bj);
    }
    bool dead = false;
    for (int k0 = 0; k0 < nk; k0 += 256) {
      __syncthreads();
      if (threadIdx.x < 256 && k0 + threadIdx.x < nk) {
        kb[threadIdx.x] = st.newk_box[(long)s * NMS_B + k0 + threadIdx.x];
        ka[threadIdx.x] = st.newk_area[(long)s * NMS_B + k0 + threadIdx.x];
      }
      __syncthreads();
      int kn = min(256, nk - k0);
      if (a && !dead) {
        for (int k = 0; k < kn; ++k) {
          if (suppresses(kb[k], ka[k], bj, aj, thr)) {
            dead = true;
            break;
          }
        }
      }
    }
    if (a && dead) st.alive[beg + pos] = 0;
  }
  __syncthreads();
}

enum { NMS_PH_SORT_SELECT = 0, NMS_PH_RESOLVE_LOOP = 1, NMS_PH_FUSED = 2 };

struct NmsCtaArgs {
  NmsState st;
  const float* boxes;
  int row_stride, order_mode, order_col;
  float thr;
  int cap_len;            // upper bound of the segment lengths (shared-memory sort capacity)
  const int* seg_counts;  // optional: segment s = rows [s * seg_stride, s * seg_stride + min(seg_counts[s], seg_stride))
  int seg_stride;
};

template <int PHASE>
__global__ void __launch_bounds__(1024) nms_cta_kernel(NmsCtaArgs a) {
  extern __shared__ unsigned long long nms_smem[];
  uint32_t* smask = reinterpret_cast<uint32_t*>(nms_smem);
  const NmsState& st = a.st;
  const int s = blockIdx.x;
  if (PHASE != NMS_PH_RESOLVE_LOOP) {
    if (threadIdx.x == 0) {
      if (a.seg_counts) {
        st.seg_beg[s] = s * a.seg_stride;
        st.seg_len[s] = min(a.seg_counts[s], a.seg_stride);
      }
      st.cursor[s] = 0;
      st.counts[s] = 0;
      st.newk_cnt[s] = 0;
      st.sel_cnt[s] = 0;
    }
    __syncthreads();
    cta_sort(nms_smem, st, s, a.boxes, a.row_stride, a.order_mode, a.order_col, a.cap_len);
    const int m = cta_select(st, s, a.boxes, a.row_stride);
    if (PHASE == NMS_PH_SORT_SELECT) return;
    if (m == 0) return;
    cta_mask(smask, st, s, m, a.thr);
    const int nk = cta_resolve(smask, st, s, m);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
  } else {
    const int m = st.sel_cnt[s];
    if (m == 0) return;
    cta_load_mask(smask, st, s, m);
    const int nk = cta_resolve(smask, st, s, m);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
  }
  // further rounds (rare: more than NMS_B candidates survive the first round's keepers) stay inside the CTA
  for (;;) {
    const int m = cta_select(st, s, a.boxes, a.row_stride);
    if (m == 0) break;
    cta_mask(smask, st, s, m, a.thr);
    const int nk = cta_resolve(smask, st, s, m);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
  }
}

// ------------------------------------------------------------------------------------------------ host driver
size_t nms_workspace_bytes(int cap_total, int cap_seg) {
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 255) & ~size_t(255); };
  add(sizeof(int) * (cap_seg + 1) * 2);                   // seg_beg, seg_len
  add(sizeof(int) * cap_total);                           // order
  add(cap_total);                                         // alive
  add(sizeof(int) * cap_seg * 4);                         // cursor, counts, sel_cnt, newk_cnt
  add(sizeof(int) * (size_t)cap_seg * NMS_B);             // sel_pos
  add(sizeof(float4) * (size_t)cap_seg * NMS_B * 2);      // sel_box, newk_box
  add(sizeof(float) * (size_t)cap_seg * NMS_B * 2);       // sel_area, newk_area
  add(sizeof(uint32_t) * (size_t)cap_seg * NMS_B * NMS_ROW_WORDS);  // mask
  add(sizeof(int) * cap_total);                           // pick
  add(256);                                               // remaining
  // radix sort double buffers
  add(sizeof(uint32_t) * (size_t)cap_total * 4);
  add((size_t)cap_total * 2);
  add(sizeof(uint32_t) * 256 * (size_t)((cap_total + RADIX_TILE - 1) / RADIX_TILE + 1));
  return b + 4096;
}

void nms_workspace_init(NmsWorkspace* ws, void* mem, size_t bytes, int cap_total, int cap_seg) {
  uint8_t* p = static_cast<uint8_t*>(mem);
  auto take = [&](size_t x) {
    uint8_t* r = p;
    p += (x + 255) & ~size_t(255);
    return r;
  };
  NmsState& st = ws->st;
  st.seg_beg = (int*)take(sizeof(int) * (cap_seg + 1) * 2);
  st.seg_len = st.seg_beg + cap_seg + 1;
  st.order = (int*)take(sizeof(int) * cap_total);
  st.alive = (uint8_t*)take(cap_total);
  int* four = (int*)take(sizeof(int) * cap_seg * 4);
  st.cursor = four;
  st.counts = four + cap_seg;
  st.sel_cnt = four + 2 * cap_seg;
  st.newk_cnt = four + 3 * cap_seg;
  st.sel_pos = (int*)take(sizeof(int) * (size_t)cap_seg * NMS_B);
  float4* boxes2 = (float4*)take(sizeof(float4) * (size_t)cap_seg * NMS_B * 2);
  st.sel_box = boxes2;
  st.newk_box = boxes2 + (size_t)cap_seg * NMS_B;
  float* areas2 = (float*)take(sizeof(float) * (size_t)cap_seg * NMS_B * 2);
  st.sel_area = areas2;
  st.newk_area = areas2 + (size_t)cap_seg * NMS_B;
  st.mask = (uint32_t*)take(sizeof(uint32_t) * (size_t)cap_seg * NMS_B * NMS_ROW_WORDS);
  st.pick = (int*)take(sizeof(int) * cap_total);
  st.remaining = (int*)take(256);
  uint32_t* r4 = (uint32_t*)take(sizeof(uint32_t) * (size_t)cap_total * 4);
  ws->rkeys[0] = r4;
  ws->rkeys[1] = r4 + cap_total;
  ws->rvals[0] = r4 + 2 * (size_t)cap_total;
  ws->rvals[1] = r4 + 3 * (size_t)cap_total;
  uint8_t* s2 = take((size_t)cap_total * 2);
  ws->rseg[0] = s2;
  ws->rseg[1] = s2 + cap_total;
  ws->rhist = (uint32_t*)take(sizeof(uint32_t) * 256 * (size_t)((cap_total + RADIX_TILE - 1) / RADIX_TILE + 1));
  ws->cap_total = cap_total;
  ws->cap_seg = cap_seg;
  FRCNN_REQUIRE((size_t)(p - static_cast<uint8_t*>(mem)) <= bytes, FRCNN_E_NOMEM, "nms workspace too small");
}

// seg_beg / seg_len must already be on the device in ws->st.  max_seg_len: host upper bound of the longest segment.
// Returns the number of kernels launched.
int nms_run(NmsWorkspace* ws, const float* boxes_dev, int row_stride, int n_seg, int n_total_cap, int max_seg_len,
            float thr, int order_mode, int order_col, cudaStream_t st_, int* h_remaining_pinned) {
  NmsState& st = ws->st;
  int launches = 0;
  if (n_seg <= 0 || n_total_cap <= 0) return 0;
  FRCNN_REQUIRE(n_seg <= ws->cap_seg && n_total_cap <= ws->cap_total, FRCNN_E_INVALID, "nms: workspace capacity exceeded");
  static bool configured = false;
  if (!configured) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_B * 33 * 4));
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_sort_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_CTA_MAX * 8));
    configured = true;
  }
  const bool small = max_seg_len <= SORT_CTA_MAX;
  if (small) {
    // CTA-level path: 3 launches (or 1 when fused)
    static bool cta_configured = false;
    const int mask_bytes = NMS_B * 33 * 4;
    if (!cta_configured) {
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_SORT_SELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_CTA_MAX * 8));
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_RESOLVE_LOOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, mask_bytes));
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, mask_bytes));
      cta_configured = true;
    }
    int n2 = 1;
    while (n2 < max_seg_len) n2 <<= 1;
    NmsCtaArgs a;
    a.st = st; a.boxes = boxes_dev; a.row_stride = row_stride; a.order_mode = order_mode; a.order_col = order_col; a.thr = thr;
    a.cap_len = n2; a.seg_counts = ws->seg_counts; a.seg_stride = ws->seg_stride;
    ws->seg_counts = nullptr;
    if (ws->fused) {
      nms_cta_kernel<NMS_PH_FUSED><<<n_seg, 1024, std::max(mask_bytes, n2 * 8), st_>>>(a);
      FRCNN_CUDA_TRY(cudaGetLastError());
      return 1;
    }
    nms_cta_kernel<NMS_PH_SORT_SELECT><<<n_seg, 1024, n2 * 8, st_>>>(a);
    nms_mask_kernel<<<dim3(NMS_B * NMS_ROW_WORDS / 256, n_seg), 256, 0, st_>>>(st, thr);
    nms_cta_kernel<NMS_PH_RESOLVE_LOOP><<<n_seg, 1024, mask_bytes, st_>>>(a);
    FRCNN_CUDA_TRY(cudaGetLastError());
    return 3;
  } else {
    FRCNN_REQUIRE(n_seg <= 256, FRCNN_E_INVALID, "nms: at most 256 segments in the large-N path");
    FRCNN_REQUIRE(ws->seg_counts == nullptr, FRCNN_E_INVALID, "nms: device-side segment counts need the CTA-level path");
    nms_init_kernel<<<(n_seg + 255) / 256, 256, 0, st_>>>(st, n_seg);
    ++launches;
    int n = n_total_cap;
    dim3 g((max_seg_len + 255) / 256 < 1024 ? (max_seg_len + 255) / 256 : 1024, n_seg);
    radix_prepare_kernel<<<g, 256, 0, st_>>>(st, boxes_dev, row_stride, order_mode, order_col, n_seg, ws->rkeys[0],
                                             ws->rvals[0], ws->rseg[0]);
    ++launches;
    int nblocks = (n + RADIX_TILE - 1) / RADIX_TILE;
    int cur = 0;
    int npass = n_seg > 1 ? 5 : 4;
    for (int pass = 0; pass < npass; ++pass) {
      radix_hist_kernel<<<nblocks, RADIX_THREADS, 0, st_>>>(ws->rkeys[cur], ws->rseg[cur], n, pass, ws->rhist, nblocks);
      radix_scan_kernel<<<1, 1024, 0, st_>>>(ws->rhist, 256 * nblocks);
      radix_scatter_kernel<<<nblocks, RADIX_THREADS, 0, st_>>>(ws->rkeys[cur], ws->rvals[cur], ws->rseg[cur],
                                                               ws->rkeys[cur ^ 1], ws->rvals[cur ^ 1], ws->rseg[cur ^ 1], n,
                                                               pass, ws->rhist, nblocks);
      launches += 3;
      cur ^= 1;
    }
    radix_finish_kernel<<<(n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184, 256, 0, st_>>>(st, ws->rvals[cur], n);
    ++launches;
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  const int max_rounds = (max_seg_len + NMS_B - 1) / NMS_B;
  const bool need_filter = max_seg_len > NMS_B;
  int filter_blocks = (max_seg_len + 2047) / 2048;
  if (filter_blocks > 592) filter_blocks = 592;
  for (int round = 0; round < max_rounds; ++round) {
    nms_select_kernel<<<n_seg, 1024, 0, st_>>>(st, boxes_dev, row_stride);
    nms_mask_kernel<<<dim3(NMS_B * NMS_ROW_WORDS / 256, n_seg), 256, 0, st_>>>(st, thr);
    nms_resolve_kernel<<<n_seg, 1024, NMS_B * 33 * 4, st_>>>(st);
    launches += 3;
    if (need_filter) {
      if (h_remaining_pinned) FRCNN_CUDA_TRY(cudaMemsetAsync(st.remaining, 0, sizeof(int), st_));
      nms_filter_kernel<<<dim3(filter_blocks, n_seg), 256, 0, st_>>>(st, boxes_dev, row_stride, thr);
      ++launches;
      if (h_remaining_pinned) {
        // large-N path: stop as soon as no candidate is left alive anywhere
        FRCNN_CUDA_TRY(cudaMemcpyAsync(h_remaining_pinned, st.remaining, sizeof(int), cudaMemcpyDeviceToHost, st_));
        FRCNN_CUDA_TRY(cudaStreamSynchronize(st_));
        if (*h_remaining_pinned == 0) break;
      }
    }
  }
  FRCNN_CUDA_TRY(cudaGetLastError());
  return launches;
}

// int32 segment-local picks -> the int64 layout of the public API
__global__ void nms_export_kernel(NmsState st, int n_seg, long long* __restrict__ pick64, long long* __restrict__ counts64) {
  int s = blockIdx.y;
  int beg = st.seg_beg[s];
  int cnt = st.counts[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) pick64[beg + i] = st.pick[beg + i];
  if (blockIdx.x == 0 && threadIdx.x == 0) counts64[s] = cnt;
}
void nms_export(NmsWorkspace* ws, int n_seg, int max_seg_len, int64_t* pick64, int64_t* counts64, cudaStream_t st_) {
  int gx = (max_seg_len + 255) / 256;
  if (gx < 1) gx = 1;
  if (gx > 256) gx = 256;
  nms_export_kernel<<<dim3(gx, n_seg), 256, 0, st_>>>(ws->st, n_seg, (long long*)pick64, (long long*)counts64);
}

// Device-resident segment counts (detector pipeline): segment s = rows [s * stride, s * stride + min(counts[s],
// stride)); the table is filled by the first kernel of the next nms_run (CTA-level path).
void nms_set_segments_from_counts(NmsWorkspace* ws, const int* counts_dev, int n_seg, int stride, cudaStream_t st_) {
  (void)n_seg;
  (void)st_;
  ws->seg_counts = counts_dev;
  ws->seg_stride = stride;
}

}  // namespace frcnn
