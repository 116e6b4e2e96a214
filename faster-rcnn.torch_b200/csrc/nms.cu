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
static constexpr int SORT_CTA_MAX = NMS_CTA_MAX_SEG;   // largest segment sorted inside one CTA

__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// nms.lua:35: area = (x2 - x1 + 1) * (y2 - y1 + 1), fp32 op by op
__device__ __forceinline__ float box_area(float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// true iff the reference would drop box j after picking box i: NOT (IoU <= overlap)   (nms.lua:72-96)
// A zero intersection (the common case) never reaches the divider: +0 / d is NaN for d == 0 or NaN and +-0 otherwise,
// so the IEEE result of the comparison is known; those lanes divide 1 / 1 instead, which keeps the whole warp off the
// slow path the division takes for zero / denormal numerators.  Bit-identical to dividing.
__device__ __forceinline__ bool suppresses(float4 bi, float ai, float4 bj, float aj, float thr) {
  float xx1 = fmaxf(bj.x, bi.x), yy1 = fmaxf(bj.y, bi.y);
  float xx2 = fminf(bj.z, bi.z), yy2 = fminf(bj.w, bi.w);
  float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  float inter = __fmul_rn(w, h);
  float denom = __fsub_rn(__fadd_rn(aj, ai), inter);
  const bool zero = inter == 0.0f;
  const float iou = __fdiv_rn(zero ? 1.0f : inter, zero ? 1.0f : denom);
  const bool sup_zero = (denom != denom) || (denom == 0.0f) || !(0.0f <= thr);
  return zero ? sup_zero : !(iou <= thr);
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
  __shared__ float4 ib[32];
  __shared__ float ia[32];
  const int s = blockIdx.y;
  const int m = st.sel_cnt[s];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = idx % NMS_B;           // a block covers 256 consecutive j of one word index
  const int wi = idx / NMS_B;
  const int j_block_max = (j - (int)threadIdx.x) + 255;
  if (wi * 32 > j_block_max || (j - (int)threadIdx.x) >= m) return;  // block-uniform: upper triangle or beyond the selection
  const float4* sb = st.sel_box + (long)s * NMS_B;
  const float* sa = st.sel_area + (long)s * NMS_B;
  if (threadIdx.x < 32) {
    const int i = wi * 32 + threadIdx.x;
    ib[threadIdx.x] = i < m ? sb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    ia[threadIdx.x] = i < m ? sa[i] : 0.f;
  }
  __syncthreads();
  if (j >= m || wi * 32 > j) return;
  const float4 bj = sb[j];
  const float aj = sa[j];
  uint32_t word = 0;
#pragma unroll 8
  for (int b = 0; b < 32; ++b) word |= suppresses(ib[b], ia[b], bj, aj, thr) ? (1u << b) : 0u;
  word &= (wi == (j >> 5)) ? ((1u << (j & 31)) - 1u) : 0xffffffffu;
  st.mask[((long)s * NMS_B + j) * NMS_ROW_WORDS + wi] = word;
}

// ------------------------------------------------------------------------------------------------ resolve
__device__ void cta_load_mask(uint32_t* smask, const NmsState& st, int s, int m);
__device__ int cta_resolve(const uint32_t* smask, const NmsState& st, int s, int m);

__global__ void __launch_bounds__(1024) nms_resolve_kernel(NmsState st) {
  extern __shared__ uint32_t smask[];  // [NMS_B][33] padded rows (bank-conflict-free column walks)
  const int s = blockIdx.x;
  const int m = st.sel_cnt[s];
  if (m == 0) {
    if (threadIdx.x == 0) st.newk_cnt[s] = 0;
    return;
  }
  cta_load_mask(smask, st, s, m);
  cta_resolve(smask, st, s, m);
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
// Bitonic sort (descending) of <= SORT_CTA_MAX 64-bit keys.  Thread t keeps the adjacent pairs (c * 2048 + 2t, + 1) in
// registers: the compare-exchange distances 1..32 run on shuffles without a block barrier, only distances >= 64 go
// through shared memory (15 block-level stages instead of 66 for 2048 keys).
__device__ __forceinline__ void cx_reg(unsigned long long& mine, unsigned long long other, bool take_max) {
  const bool other_bigger = other > mine;
  if (other_bigger == take_max) mine = other;
}
__device__ void cta_sort(unsigned long long* skeys, const NmsState& st, int s, const float* __restrict__ boxes, int row_stride,
                         int order_mode, int order_col, int cap_len) {
  constexpr int MAXC = SORT_CTA_MAX / 2048;
  const int beg = st.seg_beg[s];
  const int len = min(st.seg_len[s], cap_len);
  if (len <= 0) return;  // uniform
  int n2 = 2;
  while (n2 < len) n2 <<= 1;
  const int t = threadIdx.x, lane = t & 31;
  const int nchunks = (n2 + 2047) >> 11;
  unsigned long long k0[MAXC], k1[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int e = c * 2048 + 2 * t;
    k0[c] = 0ull;
    k1[c] = 0ull;
    if (c < nchunks) {
      if (e < len) k0[c] = ((unsigned long long)orderable(order_key(boxes, beg + e, row_stride, order_mode, order_col)) << 32) | (unsigned)e;
      if (e + 1 < len) k1[c] = ((unsigned long long)orderable(order_key(boxes, beg + e + 1, row_stride, order_mode, order_col)) << 32) | (unsigned)(e + 1);
    }
  }
  for (int size = 2; size <= n2; size <<= 1) {
    int stride = size >> 1;
    if (stride >= 64) {
      // block-level distances through shared memory
      __syncthreads();
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < nchunks && c * 2048 + 2 * t < n2) {
          skeys[c * 2048 + 2 * t] = k0[c];
          skeys[c * 2048 + 2 * t + 1] = k1[c];
        }
      __syncthreads();
      for (; stride >= 64; stride >>= 1) {
        for (int u = t; u < (n2 >> 1); u += 1024) {
          const int lo = 2 * u - (u & (stride - 1));
          const int hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long a = skeys[lo], b = skeys[hi];
          if ((a < b) == desc) {
            skeys[lo] = b;
            skeys[hi] = a;
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < nchunks && c * 2048 + 2 * t < n2) {
          k0[c] = skeys[c * 2048 + 2 * t];
          k1[c] = skeys[c * 2048 + 2 * t + 1];
        }
    }
    // distances 32..2 between lanes, distance 1 inside the thread
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c >= nchunks) continue;  // uniform
      const int e = c * 2048 + 2 * t;
      const bool desc = ((e & size) == 0);
      for (int sd = min(stride, 32); sd >= 2; sd >>= 1) {
        const int x = sd >> 1;
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0[c], x);
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1[c], x);
        const bool is_lo = (lane & x) == 0;
        cx_reg(k0[c], o0, desc == is_lo);
        cx_reg(k1[c], o1, desc == is_lo);
      }
      if ((k0[c] < k1[c]) == desc) {
        const unsigned long long tmp = k0[c];
        k0[c] = k1[c];
        k1[c] = tmp;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int e = c * 2048 + 2 * t;
    if (c < nchunks) {
      if (e < len) {
        st.order[beg + e] = (int)(unsigned)(k0[c] & 0xffffffffull);
        st.alive[beg + e] = 1;
      }
      if (e + 1 < len) {
        st.order[beg + e + 1] = (int)(unsigned)(k1[c] & 0xffffffffull);
        st.alive[beg + e + 1] = 1;
      }
    }
  }
  __syncthreads();
}

// first NMS_B alive candidates behind the cursor -> sel_*; returns the selection size (uniform)
struct CtaSel {  // this round's selection, kept on chip for the in-CTA suppression matrix
  float4 box[NMS_B];
  float area[NMS_B];
};
__device__ int cta_select(const NmsState& st, int s, const float* __restrict__ boxes, int row_stride, CtaSel* sel) {
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
    const bool in = pos < len;
    const int local = in ? st.order[beg + pos] : 0;  // independent of the alive flag: one memory round trip
    bool a = in && st.alive[beg + pos] != 0;
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
      float4 b = load_box(boxes, beg + local, row_stride);
      const float ar = box_area(b);
      st.sel_pos[(long)s * NMS_B + slot] = pos;
      st.sel_box[(long)s * NMS_B + slot] = b;
      st.sel_area[(long)s * NMS_B + slot] = ar;
      sel->box[slot] = b;
      sel->area[slot] = ar;
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

// suppression matrix of the selection computed by this CTA from the on-chip selection, straight into shared memory
// (rows < m, words <= row / 32: all the resolve reads).  Thread <-> (word, row) with consecutive threads on
// consecutive rows: box i is a warp-wide shared-memory broadcast.
__device__ void cta_mask(uint32_t* smask, const CtaSel* sel, int m, float thr) {
  const int nw = (m + 31) >> 5;
  for (int idx = threadIdx.x; idx < m * nw; idx += 1024) {
    const int wi = idx / m, j = idx - wi * m;
    if (wi > (j >> 5)) continue;
    const float4 bj = sel->box[j];
    const float aj = sel->area[j];
    // all 32 tests of the word, independent of each other (entries i >= j are masked off afterwards; slots beyond m
    // hold stale boxes, which is harmless for the same reason)
    uint32_t word = 0;
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
      const int i = wi * 32 + b;
      const bool in = i < j;  // later slots may hold stale boxes: test the box against itself instead
      word |= suppresses(in ? sel->box[i] : bj, in ? sel->area[i] : aj, bj, aj, thr) ? (1u << b) : 0u;
    }
    word &= (wi == (j >> 5)) ? ((1u << (j & 31)) - 1u) : 0xffffffffu;
    smask[j * 33 + wi] = word;
  }
  __syncthreads();
}

__device__ void cta_load_mask(uint32_t* smask, const NmsState& st, int s, int m) {
  const uint32_t* g = st.mask + (long)s * NMS_B * NMS_ROW_WORDS;
  const int total = m * NMS_ROW_WORDS;
  for (int base = threadIdx.x; base < total; base += 8 * 1024) {
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 1024;
      const int r = idx >> 5, w = idx & 31;
      v[u] = (idx < total && w * 32 <= r) ? g[idx] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 1024;
      const int r = idx >> 5, w = idx & 31;
      if (idx < total && w * 32 <= r) smask[r * 33 + w] = v[u];
    }
  }
  __syncthreads();
}

// Greedy recursion over the selection; appends the keepers; returns their number (uniform).  Warp w owns candidates
// 32w..32w+31: it folds in the keeper words of the earlier warps as they are published (spin on a shared 64-bit
// {flag, word} slot -- all 32 warps of the CTA are resident and warp w only ever waits for warps < w), then settles its
// own 32 candidates with a warp-local fixed point on ballots (as many iterations as the longest suppression chain
// inside the block, usually 1-3).  The critical path is 32 x (one word hand-over + that fixed point), not one
// block-wide barrier pair per level of the whole dependency graph.
__device__ int cta_resolve(const uint32_t* smask, const NmsState& st, int s, int m) {
  __shared__ volatile unsigned long long kept_slot[NMS_ROW_WORDS];  // bit 32 = published, low word = keepers
  __shared__ int warp_cnt[32];
  const int j = threadIdx.x;
  const int warp = j >> 5, lane = j & 31;
  const int beg = st.seg_beg[s];
  const int base_count = st.counts[s];
  const int my_pos = j < m ? st.sel_pos[(long)s * NMS_B + j] : 0;
  __syncthreads();
  if (j < NMS_ROW_WORDS) kept_slot[j] = 0ull;
  __syncthreads();
  const int nblocks = (m + 31) >> 5;
  bool is_kept = false;
  if (warp < nblocks) {
    bool dead = j >= m;
    const uint32_t* row = smask + j * 33;
    const uint32_t d = dead ? 0u : (row[warp] & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) {
      const uint32_t r = dead ? 0u : row[w];
      unsigned long long slot;
      do {
        slot = kept_slot[w];
      } while ((slot >> 32) == 0ull);
      dead |= (r & (uint32_t)slot) != 0;
    }
    uint32_t undec = ~__ballot_sync(0xffffffffu, dead);
    uint32_t kw = 0;
    while (undec != 0u) {  // uniform
      const bool me = (undec >> lane) & 1u;
      const bool hit_kept = (d & kw) != 0u;
      const bool hit_undec = (d & undec) != 0u;
      const uint32_t k = __ballot_sync(0xffffffffu, me && !hit_kept && !hit_undec);
      const uint32_t dr = __ballot_sync(0xffffffffu, me && hit_kept);
      kw |= k;
      undec &= ~(k | dr);
    }
    is_kept = (kw >> lane) & 1u;
    if (lane == 0) kept_slot[warp] = (1ull << 32) | kw;
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
  if (j < m) {
    const int pos = my_pos;
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

enum { NMS_PH_SORT_SELECT = 0, NMS_PH_RESOLVE_LOOP = 1, NMS_PH_FUSED = 2, NMS_PH_RESOLVE_SELECT = 3 };

// development aid (FRCNN_NMS_PROF=1): phase time stamps of CTA 0, printed by the host after the launch
__device__ long long g_nms_prof[64];
__device__ __forceinline__ void prof_mark(int enabled, int& slot) {
  if (enabled && threadIdx.x == 0 && slot < 64) g_nms_prof[slot] = clock64();
  ++slot;
}

struct NmsCtaArgs {
  NmsState st;
  const float* boxes;
  int row_stride, order_mode, order_col;
  float thr;
  int cap_len;            // upper bound of the segment lengths (shared-memory sort capacity)
  const int* seg_counts;  // optional: segment s = rows [s * seg_stride, s * seg_stride + min(seg_counts[s], seg_stride))
  int seg_stride;
  int prof;
};

template <int PHASE>
__global__ void __launch_bounds__(1024) nms_cta_kernel(NmsCtaArgs a) {
  extern __shared__ unsigned long long nms_smem[];
  __shared__ CtaSel sel;
  uint32_t* smask = reinterpret_cast<uint32_t*>(nms_smem);
  const NmsState& st = a.st;
  const int s = blockIdx.x;
  int ps = 0;
  int prof_on = 0;
  if (a.prof) {
    const int len0 = a.seg_counts ? min(a.seg_counts[s], a.seg_stride) : st.seg_len[s];
    prof_on = a.prof > 1000 ? (len0 > 4) : ((int)blockIdx.x == a.prof - 1);
  }
  prof_mark(prof_on, ps);
  if (PHASE != NMS_PH_RESOLVE_LOOP && PHASE != NMS_PH_RESOLVE_SELECT) {
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
    prof_mark(prof_on, ps);
    const int m = cta_select(st, s, a.boxes, a.row_stride, &sel);
    prof_mark(prof_on, ps);
    if (PHASE == NMS_PH_SORT_SELECT) return;
    if (m == 0) return;
    cta_mask(smask, &sel, m, a.thr);
    prof_mark(prof_on, ps);
    const int nk = cta_resolve(smask, st, s, m);
    prof_mark(prof_on, ps);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
    prof_mark(prof_on, ps);
  } else {
    const int m = st.sel_cnt[s];
    if (m == 0) return;
    cta_load_mask(smask, st, s, m);
    prof_mark(prof_on, ps);
    const int nk = cta_resolve(smask, st, s, m);
    prof_mark(prof_on, ps);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
    prof_mark(prof_on, ps);
    if (PHASE == NMS_PH_RESOLVE_SELECT) {
      // segments that can hold more than NMS_B boxes: the SECOND round's selection is made here and its 1024 x 1024
      // suppression matrix goes to all SMs as well (next launch); the in-CTA matrix of the loop below took 136 us at
      // 2 377 matches (vgg_large) on this one CTA
      cta_select(st, s, a.boxes, a.row_stride, &sel);
      return;
    }
  }
  // further rounds (rare: more than NMS_B candidates survive the first round's keepers) stay inside the CTA
  for (;;) {
    const int m = cta_select(st, s, a.boxes, a.row_stride, &sel);
    prof_mark(prof_on, ps);
    if (m == 0) break;
    cta_mask(smask, &sel, m, a.thr);
    prof_mark(prof_on, ps);
    const int nk = cta_resolve(smask, st, s, m);
    prof_mark(prof_on, ps);
    cta_filter(st, s, a.boxes, a.row_stride, a.thr, nk);
    prof_mark(prof_on, ps);
  }
  if (prof_on && threadIdx.x == 0) g_nms_prof[63] = ps;
}

static void nms_prof_dump(const char* what, cudaStream_t st_) {
  long long h[64];
  cudaStreamSynchronize(st_);
  cudaMemcpyFromSymbol(h, g_nms_prof, sizeof(h));
  const int n = (int)std::min<long long>(h[63], 62);
  fprintf(stderr, "[nms prof] %s:", what);
  for (int i = 1; i < n; ++i) fprintf(stderr, " %.1f", (double)(h[i] - h[i - 1]) / 1965.0);
  fprintf(stderr, " us\n");
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
  static DeviceOnce configured;
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_B * 33 * 4));
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_sort_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_CTA_MAX * 8));
  }
  const bool small = max_seg_len <= SORT_CTA_MAX;
  if (small) {
    // CTA-level path: 3 launches (or 1 when fused)
    static DeviceOnce cta_configured;
    const int mask_bytes = NMS_B * 33 * 4;
    if (first_use_on_device(cta_configured)) {
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_SORT_SELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_CTA_MAX * 8));
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_RESOLVE_LOOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, mask_bytes));
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, mask_bytes));
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(nms_cta_kernel<NMS_PH_RESOLVE_SELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mask_bytes));
    }
    int n2 = 1;
    while (n2 < max_seg_len) n2 <<= 1;
    NmsCtaArgs a;
    a.st = st; a.boxes = boxes_dev; a.row_stride = row_stride; a.order_mode = order_mode; a.order_col = order_col; a.thr = thr;
    a.cap_len = n2; a.seg_counts = ws->seg_counts; a.seg_stride = ws->seg_stride;
    static const int prof = getenv("FRCNN_NMS_PROF") ? atoi(getenv("FRCNN_NMS_PROF")) + 1 : 0;  // value = CTA to time
    a.prof = prof;  // > 1000: every CTA that gets past the first selection (racy, last writer wins)
    ws->seg_counts = nullptr;
    if (ws->fused) {
      nms_cta_kernel<NMS_PH_FUSED><<<n_seg, 1024, std::max(mask_bytes, n2 * 8), st_>>>(a);
      FRCNN_CUDA_TRY(cudaGetLastError());
      if (prof) nms_prof_dump("fused sort|select|mask|resolve|filter|select...", st_);
      return 1;
    }
    nms_cta_kernel<NMS_PH_SORT_SELECT><<<n_seg, 1024, n2 * 8, st_>>>(a);
    if (prof) nms_prof_dump("phase0 sort|select", st_);
    nms_mask_kernel<<<dim3(NMS_B * NMS_ROW_WORDS / 256, n_seg), 256, 0, st_>>>(st, thr);
    int launched = 3;
    static const int two_rounds = getenv("FRCNN_NMS_TWO_ROUNDS") ? atoi(getenv("FRCNN_NMS_TWO_ROUNDS")) : 1;
    if (max_seg_len > NMS_B && two_rounds) {
      // a second round is possible: its matrix on all SMs too (two more launches; a no-op when the round is empty)
      nms_cta_kernel<NMS_PH_RESOLVE_SELECT><<<n_seg, 1024, mask_bytes, st_>>>(a);
      if (prof) nms_prof_dump("phase1a load_mask|resolve|filter|select", st_);
      nms_mask_kernel<<<dim3(NMS_B * NMS_ROW_WORDS / 256, n_seg), 256, 0, st_>>>(st, thr);
      launched += 2;
    }
    nms_cta_kernel<NMS_PH_RESOLVE_LOOP><<<n_seg, 1024, mask_bytes, st_>>>(a);
    FRCNN_CUDA_TRY(cudaGetLastError());
    if (prof) nms_prof_dump("phase1 load_mask|resolve|filter|select|mask|resolve|filter|select...", st_);
    return launched;
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
