// Implicit-GEMM convolution / GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue -> TMA store.
//
// Replaces nn.SpatialConvolution forward as pnet uses it (models/model_utilities.lua:8,31; stride 1, square or
// rectangular kernels, symmetric zero padding) fused with the nn.PReLU, the evaluate-mode nn.SpatialDropout scale
// and -- for the last conv of a block -- the nn.SpatialMaxPooling(2,2,2,2):ceil() that follow it
// (model_utilities.lua:9-12,23), and nn.Linear forward as cnet uses it (model_utilities.lua:82).
//
// Data layout in HBM
//   activations  NHWC bf16, C % 64 == 0          (a GEMM operand [rows][K] is the case H = 1, W = rows)
//   weights      [Cout][KH][KW][Cin] bf16        (K-major rows of K = KH*KW*Cin)
// GEMM view      M = N*Hout*Wout output pixels, N = Cout, K = KH*KW*Cin.
//
// One CTA tile is a (BH*MT) x BW rectangle of 128*MT output pixels of one image, so that the A operand of filter tap
// (kh, kw) and channel chunk c is ONE 4-D TMA box {64 ch, BW, BH*MT, 1} at (c, w0 + kw - padW, h0 + kh - padH, n):
// zero padding and image borders are the TMA unit's out-of-bounds zero fill, no im2col buffer exists.  The box
// lands in shared memory as rows of 128 bytes with the 128-byte swizzle -- the canonical K-major UMMA operand
// layout -- and is consumed by four tcgen05.mma (M=128, N=BN, K=16) per 64-channel chunk and 128-row sub-tile.
// MT = 2 halves the weight traffic per MAC for the narrow (Cout <= 128) layers, which are L2->SM bandwidth bound.
//
// Kernel structure (persistent, warp-specialised, 384 threads, 1 or 2 CTAs / SM):
//   warp 0    TMA producer   (one elected lane)        smem ring: full[s] / empty[s] mbarriers
//   warp 1    MMA issuer     (the whole warp converged, one lane elected inside the asm: ptx::mma_bf16_ss_w -- a loop wrapped
//             in `if (lane == 0)` costs ~130 clocks of operand waterfall per tcgen05.mma, profiles/r2_mma_issue.md)
//                                                      TMEM accumulators double-buffered: tmem_full / tmem_empty
//   warp 2    TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld 32x32b -> bias (smem) + PReLU + scale -> bf16 -> swizzled smem tile
//             [-> 2x2 max pool in smem] -> coalesced 16-byte global stores (borders clipped);
//             or fp32 slices / TMA reduce-add for split-K (cnet, anchor heads, weight and data gradients).
//
// Kernels of this file (all share the warp roles and, except the first-layer ones, epilogue_loop):
//   conv_igemm_kernel<BN, MT>        one TMA box per filter tap (any k x k, GEMMs, split-K, weight-gradient mode)
//   conv_halo_kernel<BN, MT, KMAX, OCC, SWAP>   tile + halo in ONE box per 64-channel chunk, taps = row-shifted UMMA
//                                    descriptors (the trunk's 3x3 layers; OCC = 2: two CTAs per SM; KMAX = 7: the fused
//                                    anchor heads of the training forward with epilogue_head; SWAP: filters as the M side
//                                    -- a measured, unselected variant); 8 x 16 tiles use the warp-local epilogue_tile_warp
//   conv_pair_kernel<BN, MT, KMAX, OCC>   the halo kernel on CTA pairs (cta_group::2, M = 256, half weight boxes per CTA)
//   conv_pair_bres_kernel<BN>        CTA pairs with the layer's weights resident in shared memory (measured, unselected)
//   conv_head_kernel + head_fixup_kernel   the four anchor networks of evaluate mode in one launch: linear 128-position
//                                    tiles on CTA pairs, reduction split by filter rows, tail in the epilogue / fix-up
//   conv_wgrad_halo_kernel<BN, T>    weight gradient, T filter taps per unit sharing one dY box and one X halo box
//   conv_first_kernel                the 3-channel first layer: the A operand (K = 27 padded to 32) is built in shared
//                                    memory by producer warps straight from the fp32 NCHW frame (any frame)
//   conv_first_tma_kernel            the same with the input patch brought in by TMA, the bias through the tensor core
//                                    and four accumulator stages (16-byte aligned frames with W % 4 == 0)
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace frcnn {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;                           // bf16 elements = 128 bytes = one swizzle row
static constexpr int A_SUB_BYTES = BLOCK_M * BLOCK_K * 2;    // 16 KB per 128-row sub-tile
static constexpr int STAGE_TILE_BYTES = 128 * 128;           // epilogue staging: 128 px x 64 ch bf16
static constexpr int MAX_BIAS = 512;
static constexpr int SMEM_LIMIT = 227 * 1024;
static constexpr int SMEM_FIXED = STAGE_TILE_BYTES + MAX_BIAS * 4 + 512 + 1024;   // staging tile | bias | <= 63 mbarriers + TMEM slot | alignment slack

__host__ __device__ constexpr int stage_bytes(int BN, int MT) { return MT * A_SUB_BYTES + BN * BLOCK_K * 2; }
__host__ __device__ constexpr int conv_stages(int BN, int MT) {
  return (SMEM_LIMIT - SMEM_FIXED) / stage_bytes(BN, MT) > 8 ? 8 : (SMEM_LIMIT - SMEM_FIXED) / stage_bytes(BN, MT);
}
__host__ __device__ constexpr int tmem_cols(int BN, int MT) {
  return 2 * BN * MT <= 64 ? 64 : (2 * BN * MT <= 128 ? 128 : (2 * BN * MT <= 256 ? 256 : 512));
}
static int conv_smem_bytes_mt(int BN, int MT) { return conv_stages(BN, MT) * stage_bytes(BN, MT) + SMEM_FIXED; }
int conv_smem_bytes(int BN) { return conv_smem_bytes_mt(BN, 1); }

struct TileCoord {
  int n_img, h0, w0, n0, k_begin, k_end, m0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile, int BN, int splits, int k_per_split) {
  TileCoord t;
  int ks = tile % splits;
  int r = tile / splits;
  int nt = r % p.n_tiles_n;
  int mt = r / p.n_tiles_n;
  int tw = mt % p.tiles_w;
  int r2 = mt / p.tiles_w;
  int th = r2 % p.tiles_h;
  t.n_img = r2 / p.tiles_h;
  t.h0 = th * p.BH * p.MT;
  if (p.pair) t.h0 = (2 * th + (int)ptx::cluster_ctarank()) * p.BH * p.MT;  // CTA pair: this CTA's half of the pair tile
  t.w0 = tw * p.BW;
  t.n0 = nt * BN;
  t.k_begin = ks * k_per_split;
  t.k_end = min(p.k_iters, t.k_begin + k_per_split);
  t.m0 = t.w0;  // first GEMM row of the tile; only used with m_limit (GEMM use: BH == 1, N == 1, tiles_h == 1)
  if (p.wgrad) {
    // M index = (tap [group], 128-row Cout tile): the epilogue's reduce-add coordinates are (ci, tap, co, 0)
    t.w0 = (mt / p.co_tiles) * (p.halo ? p.MT : 1);   // (first) filter tap
    t.h0 = (mt % p.co_tiles) * BLOCK_M;               // first output channel
    t.n_img = 0;
    t.m0 = 0;
  }
  return t;
}

// Work units are dealt round-robin to the persistent CTAs -- or, on the CTA-pair kernel, to the pairs (both CTAs of a
// pair walk the same unit sequence).
__device__ __forceinline__ int unit_first(const ConvGroup& grp) { return grp.p[0].pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }
__device__ __forceinline__ int unit_stride(const ConvGroup& grp) { return grp.p[0].pair ? (int)(gridDim.x >> 1) : (int)gridDim.x; }

// Per-conv schedule values every warp role derives identically at kernel start.
struct GroupSched {
  int m_limit[MAX_GROUP];
  int splits[MAX_GROUP];
  int kps[MAX_GROUP];
};
__device__ __forceinline__ void make_sched(const ConvGroup& grp, int BN, GroupSched& sc) {
#pragma unroll
  for (int g = 0; g < MAX_GROUP; ++g) {
    const ConvParams& p = grp.p[g];
    const bool live = g < grp.n;
    sc.m_limit[g] = (live && p.m_limit) ? *p.m_limit : 0x7fffffff;
    sc.splits[g] = p.splits;
    sc.kps[g] = p.k_per_split;
    if (live && p.m_limit && p.dyn_ctas > 0) {
      const int m_tiles = max(1, (sc.m_limit[g] + BLOCK_M - 1) / BLOCK_M);
      int want = p.dyn_ctas / (m_tiles * p.n_tiles_n);
      want = max(1, min(want, p.splits));
      const int kps = (p.k_iters + want - 1) / want;
      sc.kps[g] = kps;
      sc.splits[g] = (p.k_iters + kps - 1) / kps;
    }
  }
}
// unit -> (conv index, tile coordinates); units of conv g are enumerated with its HOST split count (upper bound):
// unit indices beyond the device-chosen count decode to tiles at or beyond m_limit and are skipped by every role
__device__ __forceinline__ bool next_unit(const ConvGroup& grp, const GroupSched& sc, int unit, int BN, int& gi, TileCoord& t) {
  gi = 0;
  while (unit >= grp.unit_end[gi]) ++gi;
  const int local = unit - (gi ? grp.unit_end[gi - 1] : 0);
  const ConvParams& p = grp.p[gi];
  if (local >= p.n_tiles_m * p.n_tiles_n * sc.splits[gi]) return false;
  t = decode_tile(p, local, BN, sc.splits[gi], sc.kps[gi]);
  return t.m0 < sc.m_limit[gi];
}

// ------------------------------------------------------------------------------------------------- epilogue
// Runs on EIGHT epilogue warps (256 threads): warp e handles TMEM lane quarter (e & 3) -- thread = pixel row of the
// 128-row sub-tile -- and the column half (e >> 2) of every 64-channel group.  With one epilogue warp per scheduler
// the fixed-latency dependency stalls of the bias / PReLU / convert math paced every layer; two warps per
// scheduler interleave.  tile_buf: 1024-byte aligned 16 KB staging tile of 128 rows x 128 bytes (64 bf16 or 32 fp32
// channels of 128 pixels), 16-byte chunks XOR-swizzled by (row & 7) so that both the row-wise writes (thread =
// pixel) and the chunk-wise reads (8 lanes = one 128-byte pixel segment) are bank-conflict free.  Output leaves
// the SM as fully coalesced 128-byte segments written with plain 16-byte st.global by all 256 threads.
static constexpr int EPI_THREADS = 256;
__device__ __forceinline__ unsigned long long conv_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace layout: [cta][8 units][8 stamps]; unit slot 7 stamp 7 = CTA start, stamp 6 = CTA end
#define CONV_TRACE(P, local, k) \
  do { if ((P).trace && (local) < 7) (P).trace[((size_t)blockIdx.x * 8 + (local)) * 8 + (k)] = conv_now(); } while (0)

// Warp-local bf16 epilogue of one 128-pixel sub-tile of the halo kernels (tile 8 wide x 16 tall: a warp's 32 TMEM lanes are
// four complete tile rows, so every 2 x 2 pooling window and every output row lies inside ONE warp).  The CTA-wide version
// below needs two 256-thread barriers per 64 channels; its per-tile latency chain (TMEM load -> activation -> staging ->
// barrier -> read-back -> store, ~2 us per 128 x 128 tile) paced the narrow layers even with two accumulator stages.  Here
// every warp drains its lane quarter (q) and channel half (hf) on its own, 32 channels per pass, through a private 2 KB
// staging area (32 rows x 64 B, 16-byte chunks XOR-swizzled by (row >> 1) & 3), with __syncwarp only.
template <int BN>
__device__ __forceinline__ void epilogue_tile_warp(const ConvParams& p, const TileCoord& t, int hbase, uint32_t taddr_mt, uint8_t* tile_buf,
                                                   uint32_t bias_addr, int ewarp, int lane, float k_neg, float k_pos) {
  const int q = ewarp & 3, hf = ewarp >> 2;
  const int row = q * 32 + lane;
  const int dy = row >> 3, dx = row & 7;
  const bool valid = (hbase + dy < p.Hout) && (t.w0 + dx < p.Wout);
  const uint32_t st = ptx::smem_u32(tile_buf) + ewarp * 2048;
  const uint32_t my_row = st + lane * 64;
  const int sw = (lane >> 1) & 3;
  const int Hp = (p.Hout + 1) >> 1, Wp = (p.Wout + 1) >> 1;
  constexpr int PASSES = BN / 64;            // 32-channel passes per warp: channels [hf * BN / 2, (hf + 1) * BN / 2)
#pragma unroll 1
  for (int ps = 0; ps < PASSES; ++ps) {
    const int col = hf * (BN / 2) + ps * 32;  // first accumulator column = channel offset inside the N tile
    const int cbase = t.n0 + col;
    uint32_t o[16];
    {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(taddr_mt + col, v);
      const uint32_t baddr = bias_addr + (uint32_t)cbase * 4u;
      uint4 bq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bq[j] = ptx::ld_shared_v4(baddr + j * 16);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x0 = __uint_as_float(v[4 * j]) + __uint_as_float(bq[j].x);
        const float x1 = __uint_as_float(v[4 * j + 1]) + __uint_as_float(bq[j].y);
        const float x2 = __uint_as_float(v[4 * j + 2]) + __uint_as_float(bq[j].z);
        const float x3 = __uint_as_float(v[4 * j + 3]) + __uint_as_float(bq[j].w);
        float y0 = fmaf(fminf(x0, 0.f), k_neg, x0 * k_pos);
        float y1 = fmaf(fminf(x1, 0.f), k_neg, x1 * k_pos);
        float y2 = fmaf(fminf(x2, 0.f), k_neg, x2 * k_pos);
        float y3 = fmaf(fminf(x3, 0.f), k_neg, x3 * k_pos);
        if (p.chan_scale) {  // training: SpatialDropout mask of this image's channels
          const float4 m = __ldg(reinterpret_cast<const float4*>(p.chan_scale + (size_t)t.n_img * p.Cout + cbase) + j);
          y0 *= m.x; y1 *= m.y; y2 *= m.z; y3 *= m.w;
        }
        o[2 * j] = ptx::pack_op16x2(y0, y1, p.f16);
        o[2 * j + 1] = ptx::pack_op16x2(y2, y3, p.f16);
      }
    }
    if (p.mode == EPI_POOL && !valid) {
      const uint32_t ninf = p.f16 ? 0xFC00FC00u : 0xFF80FF80u;  // -inf: outside the map, never wins a ceil-mode window
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = ninf;
    }
    __syncwarp();   // the previous pass's staging reads are done
#pragma unroll
    for (int c = 0; c < 4; ++c) ptx::st_shared_v4(my_row + ((c ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    __syncwarp();
    if (cbase >= p.Cout || (p.dbg & 1)) continue;
    if (p.mode == EPI_POOL) {
      // this warp's 4 tile rows x 8 pixels = 2 x 4 pooled pixels x 4 chunks of 8 channels: one item per lane
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * Hp * Wp * p.Cout;
      const int pp = lane >> 2, c = lane & 3;
      const int ppy = pp >> 2, ppx = pp & 3;
      const int ph = ((hbase + q * 4) >> 1) + ppy, pw = (t.w0 >> 1) + ppx;
      if (ph < Hp && pw < Wp) {
        const int r00 = (2 * ppy) * 8 + 2 * ppx, r01 = r00 + 1, r10 = r00 + 8, r11 = r10 + 1;
        uint4 a = ptx::ld_shared_v4(st + r00 * 64 + ((c ^ ((r00 >> 1) & 3)) << 4));
        const uint4 b = ptx::ld_shared_v4(st + r01 * 64 + ((c ^ ((r01 >> 1) & 3)) << 4));
        const uint4 cc = ptx::ld_shared_v4(st + r10 * 64 + ((c ^ ((r10 >> 1) & 3)) << 4));
        const uint4 d = ptx::ld_shared_v4(st + r11 * 64 + ((c ^ ((r11 >> 1) & 3)) << 4));
        if (p.pool_arg) {
          // training (bf16): remember the winner (first maximum in window scan order, as nn.SpatialMaxPooling)
          const bf16* ea = reinterpret_cast<const bf16*>(&a);
          const bf16* eb = reinterpret_cast<const bf16*>(&b);
          const bf16* ec = reinterpret_cast<const bf16*>(&cc);
          const bf16* ed = reinterpret_cast<const bf16*>(&d);
          uint32_t lo = 0, hi = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float best = __bfloat162float(ea[e]);
            uint32_t arg = 0;
            const float vb = __bfloat162float(eb[e]), vc = __bfloat162float(ec[e]), vd = __bfloat162float(ed[e]);
            if (vb > best) { best = vb; arg = 1; }
            if (vc > best) { best = vc; arg = 2; }
            if (vd > best) { best = vd; arg = 3; }
            if (e < 4) lo |= arg << (8 * e); else hi |= arg << (8 * (e - 4));
          }
          *reinterpret_cast<uint2*>(p.pool_arg + (((size_t)t.n_img * Hp + ph) * Wp + pw) * p.Cout + cbase + c * 8) = make_uint2(lo, hi);
        }
        a.x = ptx::max4_op16x2(a.x, b.x, cc.x, d.x, p.f16);
        a.y = ptx::max4_op16x2(a.y, b.y, cc.y, d.y, p.f16);
        a.z = ptx::max4_op16x2(a.z, b.z, cc.z, d.z, p.f16);
        a.w = ptx::max4_op16x2(a.w, b.w, cc.w, d.w, p.f16);
        *reinterpret_cast<uint4*>(out_img + ((size_t)ph * Wp + pw) * p.Cout + cbase + c * 8) = a;
      }
    } else {
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * p.Hout * p.Wout * p.Cout;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane;
        const int lr = idx >> 2, c = idx & 3;
        const int r = q * 32 + lr;
        const int hh = hbase + (r >> 3), ww = t.w0 + (r & 7);
        if (hh < p.Hout && ww < p.Wout) {
          const uint4 val = ptx::ld_shared_v4(st + lr * 64 + ((c ^ ((lr >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(out_img + ((size_t)hh * p.Wout + ww) * p.Cout + cbase + c * 8) = val;
        }
      }
    }
  }
}

template <int BN, int MT, int ACC = 2>
__device__ __forceinline__ void epilogue_loop(const ConvGroup& grp, const CUtensorMap* tmOuts, uint8_t* tile_buf,
                                              const float* sbias, uint32_t tmem_base, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, const GroupSched& sc, int ewarp, int lane,
                                              uint32_t empty_leader_addr = 0) {
  // empty_leader_addr (CTA-pair kernel): shared::cluster address of the LEADER's tmem_empty[0] -- the MMA issuer lives in the
  // leader CTA and must see both CTAs' accumulator stages drained
  const int q = ewarp & 3;
  const int half = ewarp >> 2;
  const int row = q * 32 + lane;    // TMEM lane == pixel row of the sub-tile
  const int tid = ewarp * 32 + lane;
  const bool store_thread = (tid == 0);
  const uint32_t tile_addr = ptx::smem_u32(tile_buf);
  const uint32_t bias_addr = ptx::smem_u32(sbias);
  const uint32_t my_row_addr = tile_addr + row * 128;
  const int sw = row & 7;
  const int total_units = grp.unit_end[grp.n - 1];
  int seq = 0;  // index of the unit in this CTA's sequence of live units: accumulator stage = seq & 1 (ACC = 2)
  for (int unit = unit_first(grp); unit < total_units; unit += unit_stride(grp)) {
    int gi;
    TileCoord t;
    if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
    const int acc = ACC == 2 ? (seq & 1) : 0;
    const uint32_t acc_phase = (uint32_t)(ACC == 2 ? (seq >> 1) : seq) & 1u;
    ++seq;
    const ConvParams& p = grp.p[gi];
    const CUtensorMap* tmOut = tmOuts + gi;
    const int dy = row >> p.bw_shift;
    const int dx = row & (p.BW - 1);
    // PReLU(x) * scale = scale * x + scale * (slope - 1) * min(x, 0); no PReLU: slope = 1
    const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
    const float k_neg = (slope - 1.0f) * p.scale;
    const float k_pos = p.scale;
    ptx::mbar_wait(&tmem_full[acc], acc_phase);
    ptx::tc_fence_after();
    if (tid == 0) CONV_TRACE(p, seq - 1, 3);
    // weight-gradient tap groups (conv_wgrad_halo_kernel): sub-tile mt is filter tap t.w0 + mt of the same Cout rows
    const int mt_count = (p.wgrad && p.halo) ? min(MT, p.KH * p.KW - t.w0) : MT;
#pragma unroll 1
    for (int mt = 0; mt < ((p.dbg & 2) ? 0 : mt_count); ++mt) {
      const int hbase = (p.wgrad && p.halo) ? t.h0 : t.h0 + mt * p.BH;
      const int wcoord = (p.wgrad && p.halo) ? t.w0 + mt : t.w0;
      const int h = hbase + dy, w = t.w0 + dx;
      const bool valid = (h < p.Hout) && (w < p.Wout);
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BN * MT + mt * BN) + ((uint32_t)(q * 32) << 16);
      if (p.mode == EPI_F32_SLICES || p.mode == EPI_F32_REDUCE) {
        // raw fp32 partial sums: 32 columns = one 128-byte staging row (16 per column half).  SLICES: plain stores
        // into slice split * N + image (deterministic); REDUCE: TMA tensor reduction (add) into the image's map
        const int slice = p.mode == EPI_F32_SLICES ? (t.k_begin / sc.kps[gi]) * p.N + t.n_img : t.n_img;
        float* out_img = reinterpret_cast<float*>(p.out) + (size_t)slice * p.Hout * p.Wout * p.Cout;
        if (p.slice_tile_major)  // GEMM rows: tile-major slices, addressed below with the tile-local row (ww - t.w0)
          out_img = reinterpret_cast<float*>(p.out) +
                    (((long long)(t.w0 / BLOCK_M) * sc.splits[gi] + slice) * BLOCK_M - (long long)t.w0) * p.Cout;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(taddr + c0 + half * 16, v);
          ptx::tmem_ld_wait();
          if (p.mode == EPI_F32_REDUCE && store_thread) ptx::tma_store_wait_read();
          ptx::named_bar_sync(1, EPI_THREADS);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ptx::st_shared_v4(my_row_addr + (((half * 4 + c) ^ sw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          if (p.mode == EPI_F32_REDUCE) {
            ptx::fence_proxy_async();
            ptx::named_bar_sync(1, EPI_THREADS);
            if (store_thread) {
              ptx::tma_reduce_add_4d(tmOut, tile_buf, t.n0 + c0, wcoord, hbase, slice);
              ptx::tma_store_commit();
            }
          } else {
            ptx::named_bar_sync(1, EPI_THREADS);
            if (t.n0 + c0 < p.Cout) {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int idx = it * EPI_THREADS + tid;
                const int r = idx >> 3, c = idx & 7;
                const int hh = hbase + (r >> p.bw_shift), ww = t.w0 + (r & (p.BW - 1));
                if (hh < p.Hout && ww < p.Wout) {
                  const uint4 val = ptx::ld_shared_v4(tile_addr + r * 128 + ((c ^ (r & 7)) << 4));
                  *reinterpret_cast<uint4*>(out_img + ((size_t)hh * p.Wout + ww) * p.Cout + t.n0 + c0 + c * 4) = val;
                }
              }
            }
          }
        }
      } else if (p.halo && !p.wgrad && p.bw_shift == 3 && BN % 64 == 0 && !(p.dbg & 64)) {
        // 8 x 16-pixel halo tiles: every warp on its own (no CTA-wide barriers)
        epilogue_tile_warp<BN>(p, t, hbase, taddr, tile_buf, bias_addr, ewarp, lane, k_neg, k_pos);
      } else {
        const int Hp = (p.Hout + 1) >> 1, Wp = (p.Wout + 1) >> 1;
        bf16* out_img = reinterpret_cast<bf16*>(p.out) +
                        (size_t)t.n_img * (p.mode == EPI_POOL ? (size_t)Hp * Wp : (size_t)p.Hout * p.Wout) * p.Cout;
#pragma unroll 1
        for (int g = 0; g < BN / 64; ++g) {
          const int cbase = t.n0 + g * 64;
          uint32_t o[16];
          {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(taddr + g * 64 + half * 32, v);
            const uint32_t baddr = bias_addr + (uint32_t)(cbase + half * 32) * 4u;
            uint4 bq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) bq[j] = ptx::ld_shared_v4(baddr + j * 16);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x0 = __uint_as_float(v[4 * j]) + __uint_as_float(bq[j].x);
              const float x1 = __uint_as_float(v[4 * j + 1]) + __uint_as_float(bq[j].y);
              const float x2 = __uint_as_float(v[4 * j + 2]) + __uint_as_float(bq[j].z);
              const float x3 = __uint_as_float(v[4 * j + 3]) + __uint_as_float(bq[j].w);
              float y0 = fmaf(fminf(x0, 0.f), k_neg, x0 * k_pos);
              float y1 = fmaf(fminf(x1, 0.f), k_neg, x1 * k_pos);
              float y2 = fmaf(fminf(x2, 0.f), k_neg, x2 * k_pos);
              float y3 = fmaf(fminf(x3, 0.f), k_neg, x3 * k_pos);
              if (p.chan_scale) {  // training: SpatialDropout mask of this image's channels
                const float4 m = __ldg(reinterpret_cast<const float4*>(p.chan_scale + (size_t)t.n_img * p.Cout + cbase + half * 32) + j);
                y0 *= m.x; y1 *= m.y; y2 *= m.z; y3 *= m.w;
              }
              o[2 * j] = ptx::pack_op16x2(y0, y1, p.f16);
              o[2 * j + 1] = ptx::pack_op16x2(y2, y3, p.f16);
            }
          }
          if (p.mode == EPI_POOL && !valid) {
            const uint32_t ninf = p.f16 ? 0xFC00FC00u : 0xFF80FF80u;  // -inf: outside the map, never wins a ceil-mode window
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = ninf;
          }
          ptx::named_bar_sync(1, EPI_THREADS);  // every thread has finished reading the previous staging pass
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ptx::st_shared_v4(my_row_addr + (((half * 4 + c) ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          ptx::named_bar_sync(1, EPI_THREADS);
          if (cbase < p.Cout && !(p.dbg & 1)) {
            if (p.mode == EPI_POOL) {
              // 2x2 stride-2 ceil-mode max pool: 32 pooled pixels x 8 chunks, one per thread, straight to HBM
              const int pw_shift = p.bw_shift - 1;  // pooled tile width BW / 2
              const int pp = tid >> 3, c = tid & 7;
              const int ppy = pp >> pw_shift, ppx = pp & ((p.BW >> 1) - 1);
              const int ph = (hbase >> 1) + ppy, pw = (t.w0 >> 1) + ppx;
              if (ph < Hp && pw < Wp) {
                const int r00 = (2 * ppy) * p.BW + 2 * ppx;
                const int r01 = r00 + 1, r10 = r00 + p.BW, r11 = r10 + 1;
                uint4 a = ptx::ld_shared_v4(tile_addr + r00 * 128 + ((c ^ (r00 & 7)) << 4));
                const uint4 b = ptx::ld_shared_v4(tile_addr + r01 * 128 + ((c ^ (r01 & 7)) << 4));
                const uint4 cc = ptx::ld_shared_v4(tile_addr + r10 * 128 + ((c ^ (r10 & 7)) << 4));
                const uint4 d = ptx::ld_shared_v4(tile_addr + r11 * 128 + ((c ^ (r11 & 7)) << 4));
                if (p.pool_arg) {
                  // training (bf16): remember the winner (first maximum in window scan order, as nn.SpatialMaxPooling)
                  const bf16* ea = reinterpret_cast<const bf16*>(&a);
                  const bf16* eb = reinterpret_cast<const bf16*>(&b);
                  const bf16* ec = reinterpret_cast<const bf16*>(&cc);
                  const bf16* ed = reinterpret_cast<const bf16*>(&d);
                  uint32_t lo = 0, hi = 0;
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    float best = __bfloat162float(ea[e]);
                    uint32_t arg = 0;
                    const float vb = __bfloat162float(eb[e]), vc = __bfloat162float(ec[e]), vd = __bfloat162float(ed[e]);
                    if (vb > best) { best = vb; arg = 1; }
                    if (vc > best) { best = vc; arg = 2; }
                    if (vd > best) { best = vd; arg = 3; }
                    if (e < 4) lo |= arg << (8 * e); else hi |= arg << (8 * (e - 4));
                  }
                  *reinterpret_cast<uint2*>(p.pool_arg + (((size_t)t.n_img * Hp + ph) * Wp + pw) * p.Cout + cbase + c * 8) = make_uint2(lo, hi);
                }
                a.x = ptx::max4_op16x2(a.x, b.x, cc.x, d.x, p.f16);
                a.y = ptx::max4_op16x2(a.y, b.y, cc.y, d.y, p.f16);
                a.z = ptx::max4_op16x2(a.z, b.z, cc.z, d.z, p.f16);
                a.w = ptx::max4_op16x2(a.w, b.w, cc.w, d.w, p.f16);
                *reinterpret_cast<uint4*>(out_img + ((size_t)ph * Wp + pw) * p.Cout + cbase + c * 8) = a;
              }
            } else {
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int idx = it * EPI_THREADS + tid;
                const int r = idx >> 3, c = idx & 7;
                const int hh = hbase + (r >> p.bw_shift), ww = t.w0 + (r & (p.BW - 1));
                if (hh < p.Hout && ww < p.Wout) {
                  const uint4 val = ptx::ld_shared_v4(tile_addr + r * 128 + ((c ^ (r & 7)) << 4));
                  *reinterpret_cast<uint4*>(out_img + ((size_t)hh * p.Wout + ww) * p.Cout + cbase + c * 8) = val;
                }
              }
            }
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (empty_leader_addr) ptx::mbar_arrive_cluster(empty_leader_addr + (uint32_t)acc * 8u);
      else ptx::mbar_arrive(&tmem_empty[acc]);
    }
    if (tid == 0) CONV_TRACE(p, seq - 1, 4);
  }
  if (store_thread) ptx::tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------- main kernel
static constexpr int CONV_THREADS = 128 + EPI_THREADS;
template <int BN, int MT>
__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_igemm_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp) {
  constexpr int STAGES = conv_stages(BN, MT);
  constexpr int A_STAGE_BYTES = MT * A_SUB_BYTES;
  constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  constexpr uint32_t TX_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, BN);
  constexpr uint32_t IDESC_MN = ptx::make_idesc_bf16(BLOCK_M, BN, 1, 1);  // both operands MN-major (wgrad)
  constexpr int TMEM_COLS = tmem_cols(BN, MT);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* tile_buf = smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
  float* sbias = reinterpret_cast<float*>(tile_buf + STAGE_TILE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = grp.unit_end[grp.n - 1];
  GroupSched sc;
  make_sched(grp, BN, sc);
  const ConvParams& p0 = grp.p[0];

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n; ++g) {
      ptx::tma_prefetch_desc(&maps.a[g]);
      ptx::tma_prefetch_desc(&maps.b[g]);
      if (grp.p[g].mode == EPI_F32_REDUCE) ptx::tma_prefetch_desc(&maps.o[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], EPI_THREADS / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (p0.mode == EPI_STORE || p0.mode == EPI_POOL) {
    // (single-conv launches only) parameter pointers are views into Torch's flat weight buffer at arbitrary
    // 4-byte offsets: scalar loads
    for (int i = threadIdx.x; i < MAX_BIAS; i += blockDim.x) sbias[i] = (p0.bias && i < p0.Cout) ? p0.bias[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const CUtensorMap& tmA = maps.a[gi];
        const CUtensorMap& tmB = maps.b[gi];
        for (int k = t.k_begin; k < t.k_end; ++k) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
          if (p.wgrad) {
            // K = output pixels in patches of BW x BH = 64: k -> (image, patch row, patch column).  A = dY patch
            // [64 px][128 co], B = X patch shifted by the tap [64 px][BN ci]: MN-major operands, one 8 KB TMA box per
            // 64 channels; the tap shift / zero padding are coordinates of the (unconstrained) W, H dimensions
            const int per_img = p.tiles_h * p.wchunks;
            const int n = k / per_img;
            const int r = k - n * per_img;
            const int ph = r / p.wchunks;
            const int pw = r - ph * p.wchunks;
            const int kh = t.w0 / p.KW, kw = t.w0 - kh * p.KW;
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i)
              ptx::tma_load_4d(smem_a + stage * A_STAGE_BYTES + i * 8192, &tmA, &full_bar[stage], t.h0 + 64 * i, pw * p.BW, ph * p.BH, n);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_4d(smem_b + stage * B_STAGE_BYTES + j * 8192, &tmB, &full_bar[stage], t.n0 + 64 * j, pw * p.BW + kw - p.padW,
                               ph * p.BH + kh - p.padH, n);
          } else {
            int tap = k / p.cchunks;
            int cc = k - tap * p.cchunks;
            int kh = tap / p.KW;
            int kw = tap - kh * p.KW;
            ptx::tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cc * BLOCK_K, t.w0 + kw - p.padW,
                             t.h0 + kh - p.padH, t.n_img);
            ptx::tma_load_3d(smem_b + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], k * BLOCK_K, t.n0, p.f16);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (true) {   // the WHOLE warp, converged: one elected lane issues (ptx::mma_bf16_ss_w)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN * MT;
        for (int k = t.k_begin; k < t.k_end; ++k) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (grp.p[gi].wgrad) {
            // MN-major operands: [64 K rows (pixels)][64 channels = 128 B] atoms, 8-row groups 1024 B apart (SBO), the
            // next 64 channels 8 KB further (LBO); one K = 16 step = 16 rows = 2048 B
            const uint64_t da = ptx::make_desc_mn_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES), 8192);
            const uint64_t db = ptx::make_desc_mn_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES), 8192);
#pragma unroll
            for (int j = 0; j < BLOCK_K / 16; ++j)
              ptx::mma_bf16_ss_w(d_tmem, da + (uint64_t)(j * 2048 >> 4), db + (uint64_t)(j * 2048 >> 4), IDESC_MN, (k > t.k_begin || j > 0) ? 1u : 0u);
          } else {
            const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES));
            const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES));
            const uint32_t idesc = grp.p[gi].f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
#pragma unroll
            for (int j = 0; j < BLOCK_K / 16; ++j) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                // +32 bytes along K inside the 128-byte swizzle row = +2 in the (addr >> 4) field; the second
                // 128-row sub-tile starts 16 KB further
                if (!(grp.p[gi].dbg & 4))
                  ptx::mma_bf16_ss_w(d_tmem + mt * BN, da + 2 * j + mt * (A_SUB_BYTES >> 4), db + 2 * j, idesc,
                                   (k > t.k_begin || j > 0) ? 1u : 0u);
              }
            }
          }
          ptx::mma_commit_w(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (k == t.k_end - 1) ptx::mma_commit_w(&tmem_full[acc]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN, MT>(grp, maps.o, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- halo kernel
// The k x k convolutions of the trunk re-read every activation k*k times when each filter tap is its own TMA box
// (conv_igemm_kernel): at 800x450 the block-2 layers then move 9 x 23 MB = 207 MB of A operand L2 -> SM per launch
// and run at the L2 bandwidth, not at the tensor pipe.  Here the CTA tile (8 wide x 16*MT tall) is loaded ONCE per
// 64-channel chunk together with its halo -- one TMA box {64 ch, 8 + KW - 1, 16*MT + KH - 1}, 1.33x the tile instead
// of 9x -- and filter tap (kh, kw) of sub-tile mt is just a UMMA descriptor whose start address is moved by
// ((mt*16 + kh) * pitch + kw) rows of 128 bytes into that box: with a tile width of 8 pixels every 8-row group of
// the M = 128 operand is one tile row, so the group stride (SBO) is the halo pitch and the canonical K-major
// 128-byte-swizzle layout holds for every tap (the swizzle is a function of the absolute shared-memory address, the
// box starts on a 1024-byte boundary).  Weights stream through their own ring, one {64 x BN} box per (chunk, tap).
// Same warp roles and the same epilogue as conv_igemm_kernel; BN x MT = 512 columns run with ONE accumulator stage.
static constexpr int HALO_BW = 8, HALO_BH = 16, HALO_MAXK = 3, HEAD_MAXK = 7;
// extra shared memory of the fused anchor-head epilogue (EPI_HEAD): 1x1 weights [256][20] fp32 + the second column
// half's partial sums [128][19] fp32 (the 16 KB staging tile of the bf16 epilogues is not used in that mode)
static constexpr int HEAD_CO = 18, HEAD_CM = 256, HEAD_W2_PITCH = 20, HEAD_PART_PITCH = 19;
static constexpr int HEAD_SMEM = HEAD_CM * HEAD_W2_PITCH * 4 + 128 * HEAD_PART_PITCH * 4 + 128;
__host__ __device__ constexpr int halo_a_slot(int MT, int KMAX) {
  return (((HALO_BH * MT + KMAX - 1) * (HALO_BW + KMAX - 1) * 128) + 1023) & ~1023;
}
__host__ __device__ constexpr int halo_a_slots(int MT, int KMAX) { return (MT == 2 || KMAX > HALO_MAXK) ? 2 : 3; }
__host__ __device__ constexpr int halo_extra(int KMAX) { return KMAX > HALO_MAXK ? HEAD_SMEM - STAGE_TILE_BYTES : 0; }
__host__ __device__ constexpr int halo_b_slots(int BN, int MT, int KMAX) {
  return (SMEM_LIMIT - SMEM_FIXED - halo_extra(KMAX) - halo_a_slots(MT, KMAX) * halo_a_slot(MT, KMAX)) / (BN * 128) > 8
             ? 8
             : (SMEM_LIMIT - SMEM_FIXED - halo_extra(KMAX) - halo_a_slots(MT, KMAX) * halo_a_slot(MT, KMAX)) / (BN * 128);
}
__host__ __device__ constexpr int halo_acc_stages(int BN, int MT) { return 2 * BN * MT <= 512 ? 2 : 1; }
__host__ __device__ constexpr int halo_tmem_cols(int BN, int MT) {
  return halo_acc_stages(BN, MT) * BN * MT <= 128 ? 128 : (halo_acc_stages(BN, MT) * BN * MT <= 256 ? 256 : 512);
}
static int halo_smem_bytes(int BN, int MT, int KMAX) {
  return halo_a_slots(MT, KMAX) * halo_a_slot(MT, KMAX) + halo_b_slots(BN, MT, KMAX) * BN * 128 + SMEM_FIXED + halo_extra(KMAX);
}
// OCC = 2: TWO CTAs resident per SM (half the shared memory, half the TMEM columns, <= 85 registers per thread).  A CTA
// spends a third to a half of its life outside its MMA phase -- prologue, the latency of its first loads, the epilogue of
// its last unit -- and with one CTA per SM the tensor pipe idles meanwhile (ncu, batch 1: pipe active 50-63 % of a CTA's
// cycles).  A second resident CTA (another unit of the same layer, or a layer of another frame in flight) fills those
// phases; the rings are shallower, which the co-resident CTA covers as well.
__host__ __device__ constexpr int occ_smem_limit(int OCC) { return OCC == 2 ? 115712 : SMEM_LIMIT; }   // (228 KB - 2 x 1 KB reserved) / 2
__host__ __device__ constexpr int occ_fixed(int OCC) { return OCC == 2 ? SMEM_FIXED - 1024 : SMEM_FIXED; }  // OCC 2: no alignment slack
__host__ __device__ constexpr int occ_a_slots(int MT, int KMAX, int OCC) { return OCC == 2 ? 2 : halo_a_slots(MT, KMAX); }
__host__ __device__ constexpr int occ_b_slots(int b_slot_bytes, int MT, int KMAX, int OCC, int cap) {
  const int n = (occ_smem_limit(OCC) - occ_fixed(OCC) - halo_extra(KMAX) - occ_a_slots(MT, KMAX, OCC) * halo_a_slot(MT, KMAX)) / b_slot_bytes;
  return n > cap ? cap : n;
}
__host__ __device__ constexpr int occ_acc_stages(int BN, int MT, int OCC) { return 2 * BN * MT <= 512 / OCC ? 2 : 1; }
__host__ __device__ constexpr int occ_tmem_cols(int BN, int MT, int OCC) {
  const int c = occ_acc_stages(BN, MT, OCC) * BN * MT;
  return c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512));
}
static int occ_smem_bytes(int b_slot_bytes, int MT, int KMAX, int OCC, int cap) {
  return occ_a_slots(MT, KMAX, OCC) * halo_a_slot(MT, KMAX) + occ_b_slots(b_slot_bytes, MT, KMAX, OCC, cap) * b_slot_bytes + occ_fixed(OCC) +
         halo_extra(KMAX);
}
__device__ __forceinline__ uint8_t* occ_smem_base(uint8_t* raw, int OCC) {
  if (OCC == 2) {  // no slack to round up into: the dynamic shared memory window itself must be 1024-byte aligned
    if ((ptx::smem_u32(raw) & 1023u) != 0u) __trap();
    return raw;
  }
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}

// Fused AnchorNetwork epilogue (model_utilities.lua:29-35): the k x k conv's 256 fp32 sums per pixel never leave the
// SM -- bias + PReLU and the 1x1 convolution to 18 channels run on the CUDA cores straight out of TMEM, in fp32 with
// a fixed summation order (channels ascending inside a column half, then half 0 + half 1), and only the 18-channel
// map [N][18][H][W] (Torch layout, what Detector.lua:47-49 indexes) is written.  No split-K, no slice workspace.
// Thread = pixel row of the tile (TMEM lane); warp e covers lane quarter (e & 3) and column half (e >> 2).
template <int ACC>
__device__ __forceinline__ void epilogue_head(const ConvGroup& grp, float* w2s, float* parts, float* sb, uint32_t tmem_base,
                                              uint64_t* tmem_full, uint64_t* tmem_empty, const GroupSched& sc, int ewarp, int lane) {
  constexpr int BN = HEAD_CM;
  const int q = ewarp & 3, half = ewarp >> 2;
  const int row = q * 32 + lane;
  const int tid = ewarp * 32 + lane;
  const int total_units = grp.unit_end[grp.n - 1];
  int seq = 0, cur = -1;
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    int gi;
    TileCoord t;
    if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
    const ConvParams& p = grp.p[gi];
    const int acc = ACC == 2 ? (seq & 1) : 0;
    const uint32_t acc_phase = (uint32_t)(ACC == 2 ? (seq >> 1) : seq) & 1u;
    ++seq;
    if (gi != cur) {  // another head: its 1x1 weights [18][256] -> [256][20], hidden bias [256], output bias [18]
      ptx::named_bar_sync(1, EPI_THREADS);
      for (int i = tid; i < HEAD_CO * HEAD_CM; i += EPI_THREADS) {
        const int o = i / HEAD_CM, c = i - o * HEAD_CM;
        w2s[c * HEAD_W2_PITCH + o] = __ldg(p.w2 + i);
      }
      for (int i = tid; i < HEAD_CM; i += EPI_THREADS) sb[i] = __ldg(p.bias + i);
      if (tid < HEAD_CO) sb[HEAD_CM + tid] = __ldg(p.b2 + tid);
      ptx::named_bar_sync(1, EPI_THREADS);
      cur = gi;
    }
    const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
    float o18[HEAD_CO];
#pragma unroll
    for (int o = 0; o < HEAD_CO; ++o) o18[o] = 0.f;
    ptx::mbar_wait(&tmem_full[acc], acc_phase);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + half * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(taddr + c0, v);
      ptx::tmem_ld_wait();
      const float* wrow = w2s + (half * 128 + c0) * HEAD_W2_PITCH;
      const float* brow = sb + half * 128 + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float h = __uint_as_float(v[j]) + brow[j];
        h = h > 0.f ? h : h * slope;
        const float4* w4 = reinterpret_cast<const float4*>(wrow + j * HEAD_W2_PITCH);
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          const float4 w = w4[g];
          o18[4 * g] = fmaf(h, w.x, o18[4 * g]);
          o18[4 * g + 1] = fmaf(h, w.y, o18[4 * g + 1]);
          if (g < 4) {
            o18[4 * g + 2] = fmaf(h, w.z, o18[4 * g + 2]);
            o18[4 * g + 3] = fmaf(h, w.w, o18[4 * g + 3]);
          }
        }
      }
    }
    // all TMEM reads of this thread are done: the accumulator stage can be refilled while the halves are combined
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    if (half == 1) {
#pragma unroll
      for (int o = 0; o < HEAD_CO; ++o) parts[row * HEAD_PART_PITCH + o] = o18[o];
    }
    ptx::named_bar_sync(1, EPI_THREADS);
    if (half == 0) {
      const int h = t.h0 + (row >> p.bw_shift), w = t.w0 + (row & (p.BW - 1));
      if (h < p.Hout && w < p.Wout) {
        const size_t HW = (size_t)p.Hout * p.Wout;
        float* o_px = reinterpret_cast<float*>(p.out) + (size_t)t.n_img * HEAD_CO * HW + (size_t)h * p.Wout + w;
#pragma unroll
        for (int o = 0; o < HEAD_CO; ++o) o_px[o * HW] = (o18[o] + parts[row * HEAD_PART_PITCH + o]) + sb[HEAD_CM + o];
      }
    }
    ptx::named_bar_sync(1, EPI_THREADS);  // `parts` may be overwritten by the next unit
  }
}

// ------------------------------------------------------------------------------------------------- swapped operands
// Epilogue of conv_halo_kernel<128, 2, 3, 1, SWAP = true>.  Measured on B200 (tools/micro/mma_issue_bench.cu): one
// tcgen05.mma with M = 128 and shared-memory operands takes ~130 clocks per K16 step WHATEVER N is (64, 128 or 256) -- the
// 4 KB A tile is read at 32 B/clk -- so an N = 128 tile (Cout = 128: conv2_x; 3 x 128 = 384: conv4_x) runs the tensor
// pipe at half rate.  With the operands swapped the 128 filters are the M side (A = the weight box) and 256 pixels (an
// 8 x 32 halo tile, tap = descriptor offset exactly as before) the N side: the same instruction does twice the work.
// The accumulator then holds D[filter][pixel]: TMEM lane = output channel, column = pixel y * 8 + x of the tile.
//   thread <-> channel: bias / PReLU slope / dropout mask are per-thread scalars; the 2 x 2 max-pool windows (columns
//   j, j + 1, j + 8, j + 9) lie in the thread's own registers; values go through the 16 KB staging tile transposed
//   ([pixel][128 channels], 2-byte stores: a warp writes 64 contiguous bytes) and leave as 16-byte coalesced stores.
// Warp e covers lane quarter (e & 3) = channels, column half (e >> 2) = tile rows [16 half, 16 half + 16).
__device__ __forceinline__ void epilogue_swap(const ConvGroup& grp, uint8_t* tile_buf, const float* sbias, uint32_t tmem_base,
                                              uint64_t* tmem_full, uint64_t* tmem_empty, const GroupSched& sc, int ewarp, int lane) {
  constexpr int BN = 128;                 // filters per unit (the M side)
  const int q = ewarp & 3, half = ewarp >> 2;
  const int ch = q * 32 + lane;           // TMEM lane == output channel within the unit's filter tile
  const int tid = ewarp * 32 + lane;
  const uint32_t tile_addr = ptx::smem_u32(tile_buf);
  const int total_units = grp.unit_end[grp.n - 1];
  const ConvParams& p = grp.p[0];
  const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
  const float k_neg = (slope - 1.0f) * p.scale, k_pos = p.scale;
  const int Hp = (p.Hout + 1) >> 1, Wp = (p.Wout + 1) >> 1;
  int seq = 0;
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    int gi;
    TileCoord t;
    if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
    const int acc = seq & 1;
    const uint32_t acc_phase = (uint32_t)(seq >> 1) & 1u;
    ++seq;
    const int cg = t.n0 + ch;             // global output channel
    const float bias = sbias[cg];
    const float mask = p.chan_scale ? __ldg(p.chan_scale + (size_t)t.n_img * p.Cout + cg) : 1.0f;
    ptx::mbar_wait(&tmem_full[acc], acc_phase);
    ptx::tc_fence_after();
    if (tid == 0) CONV_TRACE(p, seq - 1, 3);
    const uint32_t taddr = tmem_base + (uint32_t)(acc * 256 + half * 128) + ((uint32_t)(q * 32) << 16);
    if (p.mode == EPI_POOL) {
      // ---- 2 x 2 ceil-mode max pool in registers: chunk k = tile rows 4k .. 4k + 3 of this half -> 2 x 4 pooled pixels
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * Hp * Wp * p.Cout;
      const uint16_t ninf = p.f16 ? (uint16_t)0xFC00u : (uint16_t)0xFF80u;
      ptx::named_bar_sync(1, EPI_THREADS);   // the previous unit's staging reads are done
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + k * 32, v);
        ptx::tmem_ld_wait();
        uint16_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int y = half * 16 + k * 4 + (i >> 3), x = i & 7;
          const float a = __uint_as_float(v[i]) + bias;
          const float o = fmaf(fminf(a, 0.f), k_neg, a * k_pos) * mask;
          const bool valid = (t.h0 + y < p.Hout) && (t.w0 + x < p.Wout);
          r[i] = valid ? ptx::float_to_op16(o, p.f16) : ninf;   // outside the map: never wins a window
        }
#pragma unroll
        for (int py = 0; py < 2; ++py) {
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const int i00 = (2 * py) * 8 + 2 * px;
            // winner = first maximum in window scan order, decided on the rounded values (as nn.SpatialMaxPooling would
            // on the stored activations)
            float best = ptx::op16_to_float(r[i00], p.f16);
            uint16_t bv = r[i00];
            uint32_t arg = 0;
            const float fb = ptx::op16_to_float(r[i00 + 1], p.f16), fc = ptx::op16_to_float(r[i00 + 8], p.f16), fd = ptx::op16_to_float(r[i00 + 9], p.f16);
            if (fb > best) { best = fb; bv = r[i00 + 1]; arg = 1; }
            if (fc > best) { best = fc; bv = r[i00 + 8]; arg = 2; }
            if (fd > best) { best = fd; bv = r[i00 + 9]; arg = 3; }
            const int ppy = half * 8 + k * 2 + py;                  // pooled row / column inside the tile (16 x 4)
            const int prow = ppy * 4 + px;                          // staging row
            *reinterpret_cast<uint16_t*>(tile_buf + prow * 256 + ch * 2) = bv;
            if (p.pool_arg) {
              const int ph = (t.h0 >> 1) + ppy, pw = (t.w0 >> 1) + px;
              if (ph < Hp && pw < Wp) p.pool_arg[(((size_t)t.n_img * Hp + ph) * Wp + pw) * p.Cout + cg] = (uint8_t)arg;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);   // all TMEM reads done: the stage goes back before the stores
      ptx::named_bar_sync(1, EPI_THREADS);
      // 64 pooled pixels x 256 bytes: 16-byte coalesced stores
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * EPI_THREADS + tid;
        const int prow = idx >> 4, c16 = idx & 15;
        const int ph = (t.h0 >> 1) + (prow >> 2), pw = (t.w0 >> 1) + (prow & 3);
        if (ph < Hp && pw < Wp && !(p.dbg & 1)) {
          const uint4 val = ptx::ld_shared_v4(tile_addr + prow * 256 + c16 * 16);
          *reinterpret_cast<uint4*>(out_img + ((size_t)ph * Wp + pw) * p.Cout + t.n0 + c16 * 8) = val;
        }
      }
    } else {
      // ---- EPI_STORE: four staging passes of 64 pixels (chunk k of both column halves)
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * p.Hout * p.Wout * p.Cout;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + k * 32, v);
        ptx::tmem_ld_wait();
        if (k == 3) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
        }
        ptx::named_bar_sync(1, EPI_THREADS);   // the previous pass's staging reads are done
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float a = __uint_as_float(v[i]) + bias;
          const float o = fmaf(fminf(a, 0.f), k_neg, a * k_pos) * mask;
          *reinterpret_cast<uint16_t*>(tile_buf + (half * 32 + i) * 256 + ch * 2) = ptx::float_to_op16(o, p.f16);
        }
        ptx::named_bar_sync(1, EPI_THREADS);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int idx = it * EPI_THREADS + tid;
          const int srow = idx >> 4, c16 = idx & 15;
          const int hf = srow >> 5, i = srow & 31;
          const int hh = t.h0 + hf * 16 + k * 4 + (i >> 3), ww = t.w0 + (i & 7);
          if (hh < p.Hout && ww < p.Wout && !(p.dbg & 1)) {
            const uint4 val = ptx::ld_shared_v4(tile_addr + srow * 256 + c16 * 16);
            *reinterpret_cast<uint4*>(out_img + ((size_t)hh * p.Wout + ww) * p.Cout + t.n0 + c16 * 8) = val;
          }
        }
      }
    }
    if (tid == 0) CONV_TRACE(p, seq - 1, 4);
  }
}

template <int BN, int MT, int KMAX, int OCC = 1, bool SWAP = false>
__global__ void __launch_bounds__(CONV_THREADS, OCC)
    conv_halo_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp) {
  constexpr int A_SLOT = halo_a_slot(MT, KMAX), A_SLOTS = occ_a_slots(MT, KMAX, OCC);
  constexpr int B_SLOT = BN * 128, B_SLOTS = occ_b_slots(B_SLOT, MT, KMAX, OCC, 8);
  constexpr int ACC = occ_acc_stages(BN, MT, OCC);
  constexpr int TMEM_COLS = occ_tmem_cols(BN, MT, OCC);
  static_assert(TMEM_COLS <= 512 / OCC, "accumulators of all resident CTAs must fit the 512 TMEM columns");
  // SWAP: the weight box is the A operand (M = BN = 128 filters), the 8 x 32-pixel halo tile the B operand (N = 256)
  static_assert(!SWAP || (BN == 128 && MT == 2 && OCC == 1 && KMAX == HALO_MAXK), "swapped operands: 128 filters x 256 pixels");
  constexpr uint32_t IDESC = SWAP ? ptx::make_idesc_bf16(BLOCK_M, 256) : ptx::make_idesc_bf16(BLOCK_M, BN);
  constexpr bool HEAD = KMAX > HALO_MAXK;  // the fused anchor-head configuration: EPI_HEAD units of up to 4 convs
  static_assert(B_SLOTS >= 2, "weight ring too small");
  static_assert(!HEAD || (BN == HEAD_CM && MT == 1), "anchor heads: 256 hidden channels, one sub-tile");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = occ_smem_base(smem_raw, OCC);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * A_SLOT;
  uint8_t* tile_buf = smem_b + B_SLOTS * B_SLOT;   // bf16 epilogues: 16 KB staging tile; EPI_HEAD: w2 | partial sums
  uint8_t* after_tile = tile_buf + (HEAD ? HEAD_SMEM : STAGE_TILE_BYTES);
  float* sbias = reinterpret_cast<float*>(after_tile);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_a = full_a + A_SLOTS;
  uint64_t* full_b = empty_a + A_SLOTS;
  uint64_t* empty_b = full_b + B_SLOTS;
  uint64_t* tmem_full = empty_b + B_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (threadIdx.x == 0 && grp.p[0].trace) grp.p[0].trace[((size_t)blockIdx.x * 8 + 7) * 8 + 7] = conv_now();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = grp.unit_end[grp.n - 1];
  GroupSched sc;
  make_sched(grp, BN, sc);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n; ++g) {
      ptx::tma_prefetch_desc(&maps.a[g]);
      ptx::tma_prefetch_desc(&maps.b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < B_SLOTS; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], EPI_THREADS / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (!HEAD) {
    const ConvParams& p0 = grp.p[0];
    for (int i = threadIdx.x; i < MAX_BIAS; i += blockDim.x) sbias[i] = (p0.bias && i < p0.Cout) ? p0.bias[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // Items are issued in consumption order, except that the A box of the NEXT chunk goes out after the first
    // `kpre` weight boxes of the current one: its slot frees when the previous chunk retires, i.e. about when the
    // MMA warp starts on the weight boxes already in flight, so neither ring starves the other (and the order
    // cannot deadlock: everything the MMA warp needs to retire the previous chunk was issued before).
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int ua = blockIdx.x, ca = 0;  // cursor of the A ring: one chunk ahead of the weight ring
      auto issue_a = [&]() {
        int ga;
        TileCoord ta;
        while (ua < total_units && !next_unit(grp, sc, ua, BN, ga, ta)) ua += gridDim.x;
        if (ua >= total_units) return;
        const ConvParams& pa = grp.p[ga];
        const uint32_t a_tx = (uint32_t)((HALO_BW + pa.KW - 1) * (HALO_BH * MT + pa.KH - 1)) * 128u;
        ptx::mbar_wait(&empty_a[as], aph ^ 1);
        ptx::mbar_arrive_expect_tx(&full_a[as], a_tx);
        ptx::tma_load_4d(smem_a + as * A_SLOT, &maps.a[ga], &full_a[as], ca * BLOCK_K, ta.w0 - pa.padW, ta.h0 - pa.padH, ta.n_img);
        if (++as == A_SLOTS) {
          as = 0;
          aph ^= 1;
        }
        if (++ca == pa.cchunks) {
          ca = 0;
          ua += gridDim.x;
        }
      };
      issue_a();
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int taps = p.KH * p.KW;
        const int kpre = min(B_SLOTS - 1, taps - 1);
        for (int c = 0; c < p.cchunks; ++c) {
          for (int tap = 0; tap < taps; ++tap) {
            if (tap == kpre) issue_a();
            ptx::mbar_wait(&empty_b[bs], bph ^ 1);
            ptx::mbar_arrive_expect_tx(&full_b[bs], B_SLOT);
            ptx::tma_load_3d(smem_b + bs * B_SLOT, &maps.b[gi], &full_b[bs], (tap * p.cchunks + c) * BLOCK_K, t.n0, p.f16);
            if (++bs == B_SLOTS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (true) {   // the WHOLE warp, converged: one elected lane issues (ptx::mma_bf16_ss_w)
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int seq = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int PW = HALO_BW + p.KW - 1;  // halo pitch, pixels
        const uint32_t sbo = (uint32_t)PW * 128u;
        const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
        const int acc = ACC == 2 ? (seq & 1) : 0;
        const uint32_t acc_phase = (uint32_t)(ACC == 2 ? (seq >> 1) : seq) & 1u;
        ++seq;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) CONV_TRACE(p, seq - 1, 0);
        const uint32_t d_tmem = tmem_base + acc * BN * MT;
        for (int c = 0; c < p.cchunks; ++c) {
          ptx::mbar_wait(&full_a[as], aph);
          const uint32_t a_addr = ptx::smem_u32(smem_a + as * A_SLOT);
          for (int kh = 0; kh < p.KH; ++kh) {
            for (int kw = 0; kw < p.KW; ++kw) {
              ptx::mbar_wait(&full_b[bs], bph);
              ptx::tc_fence_after();
              if (c == 0 && kh == 0 && kw == 0) CONV_TRACE(p, seq - 1, 1);
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + bs * B_SLOT));
              const uint32_t first = (c == 0 && kh == 0 && kw == 0) ? 0u : 1u;
              // consecutive tcgen05.mma into the SAME accumulator serialise on its read-modify-write (measured: ~190 clocks
              // per dependent K16 step whatever N is): the sub-tiles' accumulators alternate instruction by instruction
              if constexpr (SWAP) {
                // A = the weight box (canonical K-major tile), B = the pixel tile read from the halo box at the tap's row
                // offset: 32 groups of 8 rows (one tile row each), SBO = halo pitch
                const int row0 = kh * PW + kw;
                const uint64_t dpx = ptx::make_desc_k_sw128_sbo(a_addr + row0 * 128, sbo, p.halo_desc ? (uint32_t)row0 : 0u);
#pragma unroll
                for (int j = 0; j < BLOCK_K / 16; ++j)
                  if (!(p.dbg & 4)) ptx::mma_bf16_ss_w(d_tmem, db + 2 * j, dpx + 2 * j, idesc, j > 0 ? 1u : first);
              } else {
              uint64_t da[MT];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const int row0 = (mt * HALO_BH + kh) * PW + kw;  // first 128-byte row of this tap's operand
                da[mt] = ptx::make_desc_k_sw128_sbo(a_addr + row0 * 128, sbo, p.halo_desc ? (uint32_t)row0 : 0u);
              }
#pragma unroll
              for (int j = 0; j < BLOCK_K / 16; ++j) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                  if (!(p.dbg & 4)) ptx::mma_bf16_ss_w(d_tmem + mt * BN, da[mt] + 2 * j, db + 2 * j, idesc, j > 0 ? 1u : first);
              }
              }
              ptx::mma_commit_w(&empty_b[bs]);
              if (++bs == B_SLOTS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          ptx::mma_commit_w(&empty_a[as]);
          if (++as == A_SLOTS) {
            as = 0;
            aph ^= 1;
          }
        }
        ptx::mma_commit_w(&tmem_full[acc]);
        if (lane == 0) CONV_TRACE(p, seq - 1, 2);
      }
    }
  } else if (warp >= 4) {
    if constexpr (HEAD) {
      float* w2s = reinterpret_cast<float*>(tile_buf);
      float* parts = w2s + HEAD_CM * HEAD_W2_PITCH;
      epilogue_head<ACC>(grp, w2s, parts, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane);
    } else if constexpr (SWAP) {
      epilogue_swap(grp, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane);
    } else {
      epilogue_loop<BN, MT, ACC>(grp, maps.o, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0 && grp.p[0].trace) grp.p[0].trace[((size_t)blockIdx.x * 8 + 7) * 8 + 6] = conv_now();
}

// ------------------------------------------------------------------------------------------------- halo kernel on CTA pairs
// conv_halo_kernel with cta_group::2.  What bounds the halo kernel (profiles/r1b_conv_sweep.md, r1b_ncu_full_*): every
// CTA streams the WHOLE weight tensor of its N tile through its SM's L2 port (one CTA tile per unit at batch 1: 576 KB of
// weights against 87 KB of activations for conv3_2) and reads A + B operands from shared memory at the 128 B/clk limit.
// A CTA pair (two SMs of a TPC) computes a 2 x (128 * MT)-pixel tile with M = 256 MMAs: each CTA loads its own pixel
// tile (with halo) but only HALF of every weight box, the tensor core reads the other half from the peer's shared
// memory -- weight ingress and B operand reads per SM are halved.
//   * cluster {2,1,1}; both CTAs run producer + epilogue warps, only the leader (rank 0) runs the MMA issuer;
//   * full_a / full_b live in the leader: its producer arms them with the byte count of BOTH CTAs, the peer's TMA loads
//     complete on them (cp.async.bulk.tensor .cta_group::2 with the leader's barrier address);
//   * tcgen05.commit multicasts to empty_a / empty_b / tmem_full of both CTAs (same shared-memory offsets);
//   * the peer's epilogue warps release the accumulator stage with a remote arrive on the leader's tmem_empty;
//   * TMEM is allocated with cta_group::2 by the same warp of both CTAs; cluster barriers bracket the kernel.
template <int BN, int MT, int KMAX, int OCC = 1>
__global__ void __launch_bounds__(CONV_THREADS, OCC)
    conv_pair_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp) {
  constexpr int A_SLOT = halo_a_slot(MT, KMAX), A_SLOTS = occ_a_slots(MT, KMAX, OCC);
  constexpr int B_SLOT = BN * 64, B_SLOTS = occ_b_slots(B_SLOT, MT, KMAX, OCC, 12);   // half a weight box: BN / 2 rows of 128 bytes
  constexpr int ACC = occ_acc_stages(BN, MT, OCC);
  constexpr int TMEM_COLS = occ_tmem_cols(BN, MT, OCC);
  static_assert(TMEM_COLS <= 512 / OCC, "accumulators of all resident CTAs must fit the 512 TMEM columns");
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(2 * BLOCK_M, BN);       // M = 256 over the pair
  static_assert(B_SLOTS >= 2 && BN % 16 == 0, "weight ring too small / N must be a multiple of 16 for cta_group::2");
  static_assert((2 * A_SLOTS + 2 * B_SLOTS + 4) * 8 + 4 <= 512, "mbarrier area of SMEM_FIXED");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = occ_smem_base(smem_raw, OCC);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * A_SLOT;
  uint8_t* tile_buf = smem_b + B_SLOTS * B_SLOT;
  float* sbias = reinterpret_cast<float*>(tile_buf + STAGE_TILE_BYTES);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_a = full_a + A_SLOTS;
  uint64_t* full_b = empty_a + A_SLOTS;
  uint64_t* empty_b = full_b + B_SLOTS;
  uint64_t* tmem_full = empty_b + B_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = grp.unit_end[grp.n - 1];
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  GroupSched sc;
  make_sched(grp, BN, sc);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n; ++g) {
      ptx::tma_prefetch_desc(&maps.a[g]);
      ptx::tma_prefetch_desc(&maps.b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < B_SLOTS; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * (EPI_THREADS / 32));   // the epilogue warps of BOTH CTAs (used in the leader only)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish_2cta();
  }
  {
    const ConvParams& p0 = grp.p[0];
    for (int i = threadIdx.x; i < MAX_BIAS; i += blockDim.x) sbias[i] = (p0.bias && i < p0.Cout) ? p0.bias[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int u0 = unit_first(grp), ustride = unit_stride(grp);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs): own A box, half of every B box
    if (lane == 0) {
      const uint32_t full_a_leader = ptx::mapa_shared(ptx::smem_u32(full_a), 0);
      const uint32_t full_b_leader = ptx::mapa_shared(ptx::smem_u32(full_b), 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int ua = u0, ca = 0;  // cursor of the A ring: one chunk ahead of the weight ring
      auto issue_a = [&]() {
        int ga;
        TileCoord ta;
        while (ua < total_units && !next_unit(grp, sc, ua, BN, ga, ta)) ua += ustride;
        if (ua >= total_units) return;
        const ConvParams& pa = grp.p[ga];
        const uint32_t a_tx = (uint32_t)((HALO_BW + pa.KW - 1) * (HALO_BH * MT + pa.KH - 1)) * 128u;
        ptx::mbar_wait(&empty_a[as], aph ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&full_a[as], 2 * a_tx);
        ptx::tma_load_4d_2sm(smem_a + as * A_SLOT, &maps.a[ga], full_a_leader + (uint32_t)as * 8u, ca * BLOCK_K, ta.w0 - pa.padW,
                             ta.h0 - pa.padH, ta.n_img);
        if (++as == A_SLOTS) {
          as = 0;
          aph ^= 1;
        }
        if (++ca == pa.cchunks) {
          ca = 0;
          ua += ustride;
        }
      };
      issue_a();
      for (int unit = u0; unit < total_units; unit += ustride) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int taps = p.KH * p.KW;
        const int kpre = min(B_SLOTS - 1, taps - 1);
        for (int c = 0; c < p.cchunks; ++c) {
          for (int tap = 0; tap < taps; ++tap) {
            if (tap == kpre) issue_a();
            ptx::mbar_wait(&empty_b[bs], bph ^ 1);
            if (leader) ptx::mbar_arrive_expect_tx(&full_b[bs], 2 * B_SLOT);
            ptx::tma_load_3d_2sm(smem_b + bs * B_SLOT, &maps.b[gi], full_b_leader + (uint32_t)bs * 8u, (tap * p.cchunks + c) * BLOCK_K,
                                 t.n0 + (int)rank * (BN / 2), p.f16);
            if (++bs == B_SLOTS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {   // the WHOLE warp, converged: one elected lane issues
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int seq = 0;
      for (int unit = u0; unit < total_units; unit += ustride) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int PW = HALO_BW + p.KW - 1;  // halo pitch, pixels
        const uint32_t sbo = (uint32_t)PW * 128u;
        const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
        const int acc = ACC == 2 ? (seq & 1) : 0;
        const uint32_t acc_phase = (uint32_t)(ACC == 2 ? (seq >> 1) : seq) & 1u;
        ++seq;
        ptx::mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN * MT;
        for (int c = 0; c < p.cchunks; ++c) {
          ptx::mbar_wait(&full_a[as], aph);
          const uint32_t a_addr = ptx::smem_u32(smem_a + as * A_SLOT);
          for (int kh = 0; kh < p.KH; ++kh) {
            for (int kw = 0; kw < p.KW; ++kw) {
              ptx::mbar_wait(&full_b[bs], bph);
              ptx::tc_fence_after();
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + bs * B_SLOT));
              const uint32_t first = (c == 0 && kh == 0 && kw == 0) ? 0u : 1u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const int row0 = (mt * HALO_BH + kh) * PW + kw;  // first 128-byte row of this tap's operand
                const uint64_t da = ptx::make_desc_k_sw128_sbo(a_addr + row0 * 128, sbo, 0u);
#pragma unroll
                for (int j = 0; j < BLOCK_K / 16; ++j)
                  if (!(p.dbg & 4)) ptx::mma_bf16_ss_2cta_w(d_tmem + mt * BN, da + 2 * j, db + 2 * j, idesc, j > 0 ? 1u : first);
              }
              ptx::mma_commit_2cta_w(&empty_b[bs], 3);
              if (++bs == B_SLOTS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
          ptx::mma_commit_2cta_w(&empty_a[as], 3);
          if (++as == A_SLOTS) {
            as = 0;
            aph ^= 1;
          }
        }
        ptx::mma_commit_2cta_w(&tmem_full[acc], 3);
      }
    }
  } else if (warp >= 4) {
    const uint32_t empty_leader = ptx::mapa_shared(ptx::smem_u32(tmem_empty), 0);
    epilogue_loop<BN, MT, ACC>(grp, maps.o, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane, empty_leader);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the peer may still read its shared memory / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- CTA pairs, weights resident
// conv_pair_kernel for the narrow layers (Cout = 128: conv2_x) with the WHOLE weight tensor of the layer resident in the
// pair's shared memory.  What bounds those layers (profiles/r2_ncu_full_b1.md: tensor pipe active 30-43 % of a CTA's cycles):
// one M128 x N128 x K16 MMA reads 8 KB of operands in 64 clocks = the entire 128 B/clk shared-memory bandwidth, while the
// weight boxes streaming through the ring (16 KB per 256 MMA clocks) and the A boxes write into the same memory.  Here
//   * each CTA of a pair keeps its 64-filter half of ALL taps x channel chunks (9 x Cin / 64 boxes of 8 KB: 72 KB for
//     conv2_1, 144 KB for conv2_2), loaded once while the first unit runs -- no weight traffic afterwards;
//   * the M = 256 pair MMA reads 4 KB (A) + 2 KB (B half) per SM and K16 step: 96 B/clk;
//   * only the pixel halo boxes (23 KB per 64-channel chunk and 36 MMAs) keep arriving.
// One CTA per SM (214 KB), persistent over ~10 units, two accumulator stages.  Requires 9 * Cin / 64 <= 18 boxes.
static constexpr int BRES_MAX_BOXES = 18, BRES_MAX_A_SLOTS = 6;
static int bres_a_slots(int BN, int boxes) {
  const int n = (SMEM_LIMIT - SMEM_FIXED - boxes * BN * 64) / halo_a_slot(1, HALO_MAXK);
  return n > BRES_MAX_A_SLOTS ? BRES_MAX_A_SLOTS : n;
}
static int bres_smem_bytes(int BN, int boxes) { return bres_a_slots(BN, boxes) * halo_a_slot(1, HALO_MAXK) + boxes * BN * 64 + SMEM_FIXED; }
template <int BN>
__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_pair_bres_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp) {
  constexpr int MT = 1, KMAX = HALO_MAXK, OCC = 1;
  constexpr int A_SLOT = halo_a_slot(MT, KMAX), A_SLOTS_MAX = BRES_MAX_A_SLOTS;
  const int A_SLOTS = grp.p[0].a_slots;   // what the resident weights leave room for (2 .. BRES_MAX_A_SLOTS)
  constexpr int B_SLOT = BN * 64, B_SLOTS = BRES_MAX_BOXES;   // half a weight box: BN / 2 rows of 128 bytes; ALL boxes resident
  constexpr int ACC = occ_acc_stages(BN, MT, OCC);
  constexpr int TMEM_COLS = occ_tmem_cols(BN, MT, OCC);
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(2 * BLOCK_M, BN);       // M = 256 over the pair
  static_assert(B_SLOTS >= 2 && BN % 16 == 0, "weight ring too small / N must be a multiple of 16 for cta_group::2");
  static_assert((2 * A_SLOTS_MAX + 2 * B_SLOTS + 4) * 8 + 4 <= 512, "mbarrier area of SMEM_FIXED");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = occ_smem_base(smem_raw, OCC);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_SLOTS * A_SLOT;
  uint8_t* tile_buf = smem_b + grp.p[0].KH * grp.p[0].KW * grp.p[0].cchunks * B_SLOT;   // only the boxes the layer has
  float* sbias = reinterpret_cast<float*>(tile_buf + STAGE_TILE_BYTES);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_a = full_a + A_SLOTS_MAX;
  uint64_t* full_b = empty_a + A_SLOTS_MAX;
  uint64_t* empty_b = full_b + B_SLOTS;
  uint64_t* tmem_full = empty_b + B_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (threadIdx.x == 0 && grp.p[0].trace) grp.p[0].trace[((size_t)blockIdx.x * 8 + 7) * 8 + 7] = conv_now();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = grp.unit_end[grp.n - 1];
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  GroupSched sc;
  make_sched(grp, BN, sc);

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n; ++g) {
      ptx::tma_prefetch_desc(&maps.a[g]);
      ptx::tma_prefetch_desc(&maps.b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_SLOTS; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < B_SLOTS; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * (EPI_THREADS / 32));   // the epilogue warps of BOTH CTAs (used in the leader only)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish_2cta();
  }
  {
    const ConvParams& p0 = grp.p[0];
    for (int i = threadIdx.x; i < MAX_BIAS; i += blockDim.x) sbias[i] = (p0.bias && i < p0.Cout) ? p0.bias[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int u0 = unit_first(grp), ustride = unit_stride(grp);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs): own A box, half of every B box
    if (lane == 0) {
      const uint32_t full_a_leader = ptx::mapa_shared(ptx::smem_u32(full_a), 0);
      const uint32_t full_b_leader = ptx::mapa_shared(ptx::smem_u32(full_b), 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int ua = u0, ca = 0;  // cursor of the A ring: one chunk ahead of the weight ring
      auto issue_a = [&]() {
        int ga;
        TileCoord ta;
        while (ua < total_units && !next_unit(grp, sc, ua, BN, ga, ta)) ua += ustride;
        if (ua >= total_units) return;
        const ConvParams& pa = grp.p[ga];
        const uint32_t a_tx = (uint32_t)((HALO_BW + pa.KW - 1) * (HALO_BH * MT + pa.KH - 1)) * 128u;
        if (pa.dbg & 8) { if (++ca == pa.cchunks) { ca = 0; ua += ustride; } return; }   // measurement: no pixel loads at all
        ptx::mbar_wait(&empty_a[as], aph ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&full_a[as], 2 * a_tx);
        ptx::tma_load_4d_2sm(smem_a + as * A_SLOT, &maps.a[ga], full_a_leader + (uint32_t)as * 8u, ca * BLOCK_K, ta.w0 - pa.padW,
                             ta.h0 - pa.padH, ta.n_img);
        if (++as == A_SLOTS) {
          as = 0;
          aph ^= 1;
        }
        if (++ca == pa.cchunks) {
          ca = 0;
          ua += ustride;
        }
      };
      issue_a();
      bool weights_loaded = false;
      for (int unit = u0; unit < total_units; unit += ustride) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int taps = p.KH * p.KW;
        for (int c = 0; c < p.cchunks; ++c) {
          if (!weights_loaded) {
            // the CTA's first unit: every weight box of its filter half goes into its own slot, in consumption order,
            // and STAYS there (one N tile: the boxes are the same for every later unit)
            for (int tap = 0; tap < taps; ++tap) {
              const int slot = c * taps + tap;
              if (leader) ptx::mbar_arrive_expect_tx(&full_b[slot], 2 * B_SLOT);
              ptx::tma_load_3d_2sm(smem_b + slot * B_SLOT, &maps.b[gi], full_b_leader + (uint32_t)slot * 8u, (tap * p.cchunks + c) * BLOCK_K,
                                   t.n0 + (int)rank * (BN / 2), p.f16);
              if (tap == 0) issue_a();
            }
          } else {
            issue_a();
          }
        }
        weights_loaded = true;
      }
      (void)bs; (void)bph;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {   // the WHOLE warp, converged: one elected lane issues
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int seq = 0;
      bool weights_ready = false;
      for (int unit = u0; unit < total_units; unit += ustride) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const ConvParams& p = grp.p[gi];
        const int taps = p.KH * p.KW;
        const int PW = HALO_BW + p.KW - 1;  // halo pitch, pixels
        const uint32_t sbo = (uint32_t)PW * 128u;
        const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
        const int acc = ACC == 2 ? (seq & 1) : 0;
        const uint32_t acc_phase = (uint32_t)(ACC == 2 ? (seq >> 1) : seq) & 1u;
        ++seq;
        ptx::mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) CONV_TRACE(p, seq - 1, 0);
        const uint32_t d_tmem = tmem_base + acc * BN * MT;
        for (int c = 0; c < p.cchunks; ++c) {
          if (!(p.dbg & 8)) ptx::mbar_wait(&full_a[as], aph);
          const uint32_t a_addr = ptx::smem_u32(smem_a + as * A_SLOT);
          for (int kh = 0; kh < p.KH; ++kh) {
            for (int kw = 0; kw < p.KW; ++kw) {
              const int slot = c * taps + kh * p.KW + kw;
              if (!weights_ready) ptx::mbar_wait(&full_b[slot], 0u);
              ptx::tc_fence_after();
              if (slot == 0) CONV_TRACE(p, seq - 1, 1);
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + ((p.dbg & 16) ? 0 : slot) * B_SLOT));
              const uint32_t first = (c == 0 && kh == 0 && kw == 0) ? 0u : 1u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const int row0 = (mt * HALO_BH + kh) * PW + kw;  // first 128-byte row of this tap's operand
                const uint64_t da = (p.dbg & 32) ? ptx::make_desc_k_sw128(a_addr) : ptx::make_desc_k_sw128_sbo(a_addr + row0 * 128, sbo, 0u);
#pragma unroll
                for (int j = 0; j < BLOCK_K / 16; ++j)
                  if (!(p.dbg & 4)) ptx::mma_bf16_ss_2cta_w(d_tmem + mt * BN, da + 2 * j, db + 2 * j, idesc, j > 0 ? 1u : first);
              }
            }
          }
          if (!(p.dbg & 8)) ptx::mma_commit_2cta_w(&empty_a[as], 3);
          if (++as == A_SLOTS) {
            as = 0;
            aph ^= 1;
          }
        }
        ptx::mma_commit_2cta_w(&tmem_full[acc], 3);
        if (lane == 0) CONV_TRACE(p, seq - 1, 2);
        weights_ready = true;
      }
      (void)bs; (void)bph;
    }
  } else if (warp >= 4) {
    const uint32_t empty_leader = ptx::mapa_shared(ptx::smem_u32(tmem_empty), 0);
    epilogue_loop<BN, MT, ACC>(grp, maps.o, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane, empty_leader);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the peer may still read its shared memory / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0 && grp.p[0].trace) grp.p[0].trace[((size_t)blockIdx.x * 8 + 7) * 8 + 6] = conv_now();
}

// ------------------------------------------------------------------------------------------------- anchor networks
// The four AnchorNetworks (model_utilities.lua:29-35: k x k valid conv to 256 channels -> PReLU -> 1 x 1 conv to 18) of a
// frame in ONE launch, fused down to the 18-channel maps Detector.lua:47-49 reads.  What the halo-kernel version lost
// (profiles/r1b_ncu_full_b1.md: 136 us, 0.13 of the tensor peak, 88 CTAs): 8 x 16 pixel tiles on 23..27-row maps waste
// 16-34 % of every MMA; one CTA per tile makes the 7 x 7 head's 18 816-deep reduction a 78-135 us critical path; and
// every CTA streams the whole 256-filter weight tensor of its head through its SM's L2 port (32 KB per 512 MMA clocks =
// 64 B/clk against the ~42 B/clk an SM gets when all 148 pull: the MMA phase ran at half speed).
//   * LINEAR tiles: the input map is the 2-D matrix [N*Hin*Win pixels][Cin] and a tile is 128 CONSECUTIVE pixel positions
//     p = y*Win + x.  For filter row kh the A operand of ALL kw taps is one slab of 128 + k - 1 consecutive rows starting at
//     p0 + kh*Win (one TMA box {64 ch, 136 rows}); tap kw is the same slab read from row kw on (UMMA descriptor start
//     + kw*128 B, canonical 8-row groups).  Positions whose window wraps around the row end (x > Win - k) produce
//     garbage that is never stored: 4-12 % waste instead of 16-34 %.
//   * CTA PAIRS (cta_group::2, cluster {2,1,1}): a pair computes two neighbouring tiles with M = 256 MMAs; each CTA loads
//     its own slab but only HALF of every weight box (128 of the 256 filters, 16 KB), the tensor core reads the other half
//     from the peer's shared memory: weight ingress per SM halved.
//   * K SPLIT BY FILTER ROWS with an in-kernel fix-up: a unit is (head, image, tile pair, kh range).  Split units write their
//     fp32 partial sums to a slice in L2 ([column group of 4][128 rows][4]: a warp stores / loads 512 contiguous bytes); the
//     CTA that arrives LAST at its tile's counter sums the slices in ascending order (fixed order: the result does not
//     depend on who is last), applies bias + PReLU + the 1 x 1 conv and stores.  Units are dealt to the persistent pairs
//     by a host-side longest-processing-time schedule.
// Warp roles as conv_pair_kernel: TMA producer (both CTAs), single-thread MMA issuer (leader CTA, M256 x N256 x K16), two
// accumulator stages in TMEM, eight epilogue warps per CTA.
static constexpr int HEADK_SLAB_ROWS = 136;                                  // 128 + (7 - 1), padded to a multiple of 8
static constexpr int HEADK_A_SLOT = HEADK_SLAB_ROWS * 128;                   // 17 408 B = 17 KB
static constexpr int HEADK_A_SLOTS = 3, HEADK_B_SLOTS = 8;
static constexpr int HEADK_B_SLOT = (HEAD_CM / 2) * 128;                     // half a weight box: 128 filters x 128 B = 16 KB
static constexpr int HEADK_SMEM = HEADK_A_SLOTS * HEADK_A_SLOT + HEADK_B_SLOTS * HEADK_B_SLOT + HEAD_SMEM + MAX_BIAS * 4 + 512 + 1024;
static constexpr int HEADK_MAX_SLICES = HEAD_MAXK;                           // one slice per filter row at most
static_assert(HEADK_SMEM <= 227 * 1024, "anchor-network kernel: shared memory");
static_assert((2 * HEADK_A_SLOTS + 2 * HEADK_B_SLOTS + 4) * 8 + 8 <= 512, "anchor-network kernel: mbarrier area");

// u.x bit 24: the unit stores its fp32 sums as a slice even when it is the only one (the tail then runs in head_fixup_kernel)
__device__ __forceinline__ void headk_decode(const int4 u, int& head, int& s, int& nsl, int& img, int& tile, int& kh0, int& kh1) {
  head = u.x & 0xff; s = (u.x >> 8) & 0xff; nsl = (u.x >> 16) & 0xff;
  img = u.y; tile = u.z; kh0 = u.w & 0xff; kh1 = (u.w >> 8) & 0xff;
}

__device__ __forceinline__ unsigned long long headk_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define HEADK_TRACE(ui, k, v) \
  do { if (hs.trace && (ui) - u_begin < 8) hs.trace[((size_t)blockIdx.x * 8 + ((ui) - u_begin)) * 8 + (k)] = (v); } while (0)

__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_head_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp, const HeadSched hs) {
  constexpr int BN = HEAD_CM;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(2 * BLOCK_M, BN);          // M = 256 over the pair
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + HEADK_A_SLOTS * HEADK_A_SLOT;
  float* w2s = reinterpret_cast<float*>(smem_b + HEADK_B_SLOTS * HEADK_B_SLOT);
  float* parts = w2s + HEAD_CM * HEAD_W2_PITCH;
  float* sb = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(w2s) + HEAD_SMEM);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sb + MAX_BIAS);
  uint64_t* empty_a = full_a + HEADK_A_SLOTS;
  uint64_t* full_b = empty_a + HEADK_A_SLOTS;
  uint64_t* empty_b = full_b + HEADK_B_SLOTS;
  uint64_t* tmem_full = empty_b + HEADK_B_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_base_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int u_begin = hs.cta_off[pair], u_end = hs.cta_off[pair + 1];
  if (threadIdx.x == 0 && hs.trace) hs.trace[((size_t)blockIdx.x * 8 + 7) * 8 + 7] = headk_now();

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < grp.n; ++g) {
      ptx::tma_prefetch_desc(&maps.a[g]);
      ptx::tma_prefetch_desc(&maps.b[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < HEADK_A_SLOTS; ++i) {
      ptx::mbar_init(&full_a[i], 1);
      ptx::mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < HEADK_B_SLOTS; ++i) {
      ptx::mbar_init(&full_b[i], 1);
      ptx::mbar_init(&empty_b[i], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * (EPI_THREADS / 32));   // the epilogue warps of BOTH CTAs (used in the leader only)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_base_slot, 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs): per (kh, chunk) step the CTA's
    // own slab + its half of k weight boxes.  The slab of step i + 1 goes out BEFORE the weight boxes of step i (its own
    // cursor, one step ahead): otherwise it would queue behind weight boxes that wait for ring slots.
    if (lane == 0) {
      const uint32_t full_a_leader = ptx::mapa_shared(ptx::smem_u32(full_a), 0);
      const uint32_t full_b_leader = ptx::mapa_shared(ptx::smem_u32(full_b), 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int a_ui = u_begin, a_kh = -1, a_c = 0;   // cursor of the slab ring
      auto issue_a = [&]() {
        while (a_ui < u_end) {
          int g, s, nsl, img, tile, kh0, kh1;
          headk_decode(hs.units[a_ui], g, s, nsl, img, tile, kh0, kh1);
          const ConvParams& p = grp.p[g];
          if (a_kh < 0) { a_kh = kh0; a_c = 0; }
          if (a_kh >= kh1) { ++a_ui; a_kh = -1; continue; }
          ptx::mbar_wait(&empty_a[as], aph ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&full_a[as], 2 * HEADK_A_SLOT);
          // rows past the end of the map (the odd tile of a pair, the last slab rows) are zero-filled by the TMA unit
          ptx::tma_load_2d_2sm(smem_a + as * HEADK_A_SLOT, &maps.a[g], full_a_leader + (uint32_t)as * 8u, a_c * BLOCK_K,
                               img * p.Hin * p.Win + (2 * tile + (int)rank) * BLOCK_M + a_kh * p.Win);
          if (++as == HEADK_A_SLOTS) {
            as = 0;
            aph ^= 1;
          }
          if (++a_c == p.cchunks) { a_c = 0; ++a_kh; }
          return;
        }
      };
      issue_a();
      for (int ui = u_begin; ui < u_end; ++ui) {
        int g, s, nsl, img, tile, kh0, kh1;
        headk_decode(hs.units[ui], g, s, nsl, img, tile, kh0, kh1);
        const ConvParams& p = grp.p[g];
        HEADK_TRACE(ui, 6, headk_now());
        for (int kh = kh0; kh < kh1; ++kh) {
          for (int c = 0; c < p.cchunks; ++c) {
            issue_a();   // the NEXT step's slab
            for (int kw = 0; kw < p.KW; ++kw) {
              ptx::mbar_wait(&empty_b[bs], bph ^ 1);
              if (leader) ptx::mbar_arrive_expect_tx(&full_b[bs], 2 * HEADK_B_SLOT);
              ptx::tma_load_3d_2sm(smem_b + bs * HEADK_B_SLOT, &maps.b[g], full_b_leader + (uint32_t)bs * 8u,
                                   ((kh * p.KW + kw) * p.cchunks + c) * BLOCK_K, (int)rank * (BN / 2), p.f16);
              if (++bs == HEADK_B_SLOTS) {
                bs = 0;
                bph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {   // the WHOLE warp, converged: one elected lane issues
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int seq = 0;
      for (int ui = u_begin; ui < u_end; ++ui, ++seq) {
        int g, s, nsl, img, tile, kh0, kh1;
        headk_decode(hs.units[ui], g, s, nsl, img, tile, kh0, kh1);
        const ConvParams& p = grp.p[g];
        const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
        const int acc = seq & 1;
        ptx::mbar_wait_cluster(&tmem_empty[acc], ((uint32_t)(seq >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        if (lane == 0) HEADK_TRACE(ui, 0, headk_now());
        if (lane == 0) HEADK_TRACE(ui, 5, (unsigned long long)(g | (s << 8) | (nsl << 16) | (tile << 24)));
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t first = 0u;
        for (int kh = kh0; kh < kh1; ++kh) {
          for (int c = 0; c < p.cchunks; ++c) {
            ptx::mbar_wait(&full_a[as], aph);
            const uint32_t a_addr = ptx::smem_u32(smem_a + as * HEADK_A_SLOT);
            for (int kw = 0; kw < p.KW; ++kw) {
              ptx::mbar_wait(&full_b[bs], bph);
              ptx::tc_fence_after();
              // tap kw = the slab read from row kw on: start address + kw * 128 B, canonical 8-row groups (SBO 1024)
              const uint64_t da = ptx::make_desc_k_sw128(a_addr + kw * 128);
              const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + bs * HEADK_B_SLOT));
              if (lane == 0 && hs.trace && ui == u_begin && first == 0u) hs.trace[((size_t)blockIdx.x * 8 + 7) * 8 + 6] = headk_now();   // first operands landed
#pragma unroll
              for (int j = 0; j < BLOCK_K / 16; ++j) ptx::mma_bf16_ss_2cta_w(d_tmem, da + 2 * j, db + 2 * j, idesc, j > 0 ? 1u : first);
              first = 1u;
              ptx::mma_commit_2cta_w(&empty_b[bs], 3);
              if (++bs == HEADK_B_SLOTS) {
                bs = 0;
                bph ^= 1;
              }
            }
            ptx::mma_commit_2cta_w(&empty_a[as], 3);
            if (++as == HEADK_A_SLOTS) {
              as = 0;
              aph ^= 1;
            }
          }
        }
        ptx::mma_commit_2cta_w(&tmem_full[acc], 3);
        if (lane == 0) HEADK_TRACE(ui, 1, headk_now());
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, each on its own tile): thread = pixel
    // row (TMEM lane), column half = ewarp >> 2
    const uint32_t empty_leader = ptx::mapa_shared(ptx::smem_u32(tmem_empty), 0);
    const int ewarp = warp - 4;
    const int q = ewarp & 3, half = ewarp >> 2;
    const int row = q * 32 + lane;
    const int tid = ewarp * 32 + lane;
    int cur = -1, seq = 0;
    for (int ui = u_begin; ui < u_end; ++ui, ++seq) {
      int g, s, nsl, img, ptile, kh0, kh1;
      headk_decode(hs.units[ui], g, s, nsl, img, ptile, kh0, kh1);
      const ConvParams& p = grp.p[g];
      const int acc = seq & 1;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * BN + half * 128) + ((uint32_t)(q * 32) << 16);
      const int tiles_img = hs.tiles[g];
      const int tile = 2 * ptile + (int)rank;
      const long tile_lin = (long)img * tiles_img + tile;
      ptx::mbar_wait(&tmem_full[acc], (uint32_t)(seq >> 1) & 1u);
      ptx::tc_fence_after();
      if (tid == 0) HEADK_TRACE(ui, 2, headk_now());
      if (tile >= tiles_img) {
        // the odd tile of the last pair: nothing to keep, only the accumulator stage to hand back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(empty_leader + (uint32_t)acc * 8u);
        continue;
      }
      // partial sums / fix-up layout of a slice: [64 column groups][128 rows][4 floats]
      float* slice0 = hs.slices[g] + (tile_lin * nsl) * (long)(BLOCK_M * BN) + (long)(half * 32) * (BLOCK_M * 4) + row * 4;
      if (nsl > 1 || ((hs.units[ui].x >> 24) & 1)) {
        // ---- partial sums of this filter-row range -> slice s of the tile (a warp stores 512 contiguous bytes)
        float* sl = slice0 + (long)s * (BLOCK_M * BN);
#pragma unroll 2
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(taddr + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            __stcg(reinterpret_cast<uint4*>(sl + (c0 / 4 + j) * (BLOCK_M * 4)), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(empty_leader + (uint32_t)acc * 8u);   // the stage is free for the next unit's MMAs
        if (tid == 0) HEADK_TRACE(ui, 3, headk_now());
        continue;   // head_fixup_kernel (next launch) sums the slices and applies the tail
      }
      if (g != cur) {  // another head: its 1x1 weights [18][256] -> [256][20], hidden bias [256], output bias [18]
        ptx::named_bar_sync(1, EPI_THREADS);
        for (int i = tid; i < HEAD_CO * HEAD_CM; i += EPI_THREADS) {
          const int o = i / HEAD_CM, c = i - o * HEAD_CM;
          w2s[c * HEAD_W2_PITCH + o] = __ldg(p.w2 + i);
        }
        for (int i = tid; i < HEAD_CM; i += EPI_THREADS) sb[i] = __ldg(p.bias + i);
        if (tid < HEAD_CO) sb[HEAD_CM + tid] = __ldg(p.b2 + tid);
        ptx::named_bar_sync(1, EPI_THREADS);
        cur = g;
      }
      const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
      // 18 outputs as 9 packed pairs: one FFMA2 per pair and hidden channel (each half rounds like a scalar fmaf)
      float2 o9[HEAD_CO / 2];
#pragma unroll
      for (int o = 0; o < HEAD_CO / 2; ++o) o9[o] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(taddr + c0, v);
        ptx::tmem_ld_wait();
        const float* wrow = w2s + (half * 128 + c0) * HEAD_W2_PITCH;
        const float* brow = sb + half * 128 + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float h = __uint_as_float(v[j]) + brow[j];
          h = h > 0.f ? h : h * slope;
          const float2 hh = make_float2(h, h);
          const float4* w4 = reinterpret_cast<const float4*>(wrow + j * HEAD_W2_PITCH);
#pragma unroll
          for (int gq = 0; gq < 5; ++gq) {
            const float4 w = w4[gq];
            o9[2 * gq] = __ffma2_rn(hh, make_float2(w.x, w.y), o9[2 * gq]);
            if (gq < 4) o9[2 * gq + 1] = __ffma2_rn(hh, make_float2(w.z, w.w), o9[2 * gq + 1]);
          }
        }
      }
      float o18[HEAD_CO];
#pragma unroll
      for (int o = 0; o < HEAD_CO / 2; ++o) { o18[2 * o] = o9[o].x; o18[2 * o + 1] = o9[o].y; }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(empty_leader + (uint32_t)acc * 8u);
      if (half == 1) {
#pragma unroll
        for (int o = 0; o < HEAD_CO; ++o) parts[row * HEAD_PART_PITCH + o] = o18[o];
      }
      ptx::named_bar_sync(1, EPI_THREADS);
      if (half == 0) {
        const int pp = tile * BLOCK_M + row;         // linear position on the INPUT grid
        const int y = pp / p.Win, x = pp - y * p.Win;
        if (y < p.Hout && x < p.Wout) {
          const size_t HW = (size_t)p.Hout * p.Wout;
          float* o_px = reinterpret_cast<float*>(p.out) + (size_t)img * HEAD_CO * HW + (size_t)y * p.Wout + x;
#pragma unroll
          for (int o = 0; o < HEAD_CO; ++o) o_px[o * HW] = (o18[o] + parts[row * HEAD_PART_PITCH + o]) + sb[HEAD_CM + o];
        }
      }
      ptx::named_bar_sync(1, EPI_THREADS);  // `parts` may be overwritten by the next unit
      if (tid == 0) HEADK_TRACE(ui, 4, headk_now());
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the peer may still read its shared memory / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
  if (threadIdx.x == 0) HEADK_TRACE(u_begin, 7, headk_now());
}

// The tail of the anchor networks conv_head_kernel split by filter rows: sum of the slices in ascending order (fixed
// summation order) + bias -> PReLU -> 1 x 1 conv to 18 channels (model_utilities.lua:32-33), fp32.  A block = 32 positions
// of one 128-position tile: lane <-> position (the slice layout [column group][row][4] makes a warp's load 512 contiguous
// bytes), warp <-> 32 of the 256 hidden channels; the eight partial 18-vectors of a position are added in warp order.
__global__ void __launch_bounds__(256) head_fixup_kernel(const __grid_constant__ ConvGroup grp, const HeadSched hs, const HeadFixArgs fx) {
  __shared__ __align__(16) float w2s[HEAD_CM * HEAD_W2_PITCH];
  __shared__ float sb[HEAD_CM + HEAD_CO];
  __shared__ float red[8][32][HEAD_PART_PITCH];
  int hi = 0;
  while (hi + 1 < fx.n && (int)blockIdx.x >= fx.block_end[hi]) ++hi;
  const int g = fx.head[hi], nsl = fx.nsl[hi];
  const int b = (int)blockIdx.x - (hi ? fx.block_end[hi - 1] : 0);
  const ConvParams& p = grp.p[g];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < HEAD_CO * HEAD_CM; i += 256) {
    const int o = i / HEAD_CM, c = i - o * HEAD_CM;
    w2s[c * HEAD_W2_PITCH + o] = __ldg(p.w2 + i);
  }
  for (int i = tid; i < HEAD_CM; i += 256) sb[i] = __ldg(p.bias + i);
  if (tid < HEAD_CO) sb[HEAD_CM + tid] = __ldg(p.b2 + tid);
  __syncthreads();
  const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
  const long tile_lin = b >> 2;
  const int row = (b & 3) * 32 + lane;
  const float* slice0 = hs.slices[g] + (tile_lin * nsl) * (long)(BLOCK_M * HEAD_CM) + (long)(warp * 8) * (BLOCK_M * 4) + row * 4;
  float2 o9[HEAD_CO / 2];
#pragma unroll
  for (int o = 0; o < HEAD_CO / 2; ++o) o9[o] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int cg = 0; cg < 8; cg += 2) {
    float4 t[HEADK_MAX_SLICES][2];
#pragma unroll
    for (int s2 = 0; s2 < HEADK_MAX_SLICES; ++s2) {
      if (s2 < nsl) {
        const float* slp = slice0 + (long)s2 * (BLOCK_M * HEAD_CM) + cg * (BLOCK_M * 4);
        t[s2][0] = __ldcg(reinterpret_cast<const float4*>(slp));
        t[s2][1] = __ldcg(reinterpret_cast<const float4*>(slp + BLOCK_M * 4));
      }
    }
    float x[8] = {t[0][0].x, t[0][0].y, t[0][0].z, t[0][0].w, t[0][1].x, t[0][1].y, t[0][1].z, t[0][1].w};
#pragma unroll
    for (int s2 = 1; s2 < HEADK_MAX_SLICES; ++s2) {
      if (s2 < nsl) {
        x[0] += t[s2][0].x; x[1] += t[s2][0].y; x[2] += t[s2][0].z; x[3] += t[s2][0].w;
        x[4] += t[s2][1].x; x[5] += t[s2][1].y; x[6] += t[s2][1].z; x[7] += t[s2][1].w;
      }
    }
    const int c0 = (warp * 8 + cg) * 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float h = x[j] + sb[c0 + j];
      h = h > 0.f ? h : h * slope;
      const float2 hh = make_float2(h, h);
      const float4* w4 = reinterpret_cast<const float4*>(w2s + (c0 + j) * HEAD_W2_PITCH);
#pragma unroll
      for (int gq = 0; gq < 5; ++gq) {
        const float4 w = w4[gq];
        o9[2 * gq] = __ffma2_rn(hh, make_float2(w.x, w.y), o9[2 * gq]);
        if (gq < 4) o9[2 * gq + 1] = __ffma2_rn(hh, make_float2(w.z, w.w), o9[2 * gq + 1]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < HEAD_CO / 2; ++o) {
    red[warp][lane][2 * o] = o9[o].x;
    red[warp][lane][2 * o + 1] = o9[o].y;
  }
  __syncthreads();
  const int img = (int)(tile_lin / hs.tiles[g]), tile = (int)(tile_lin - (long)img * hs.tiles[g]);
  const size_t HW = (size_t)p.Hout * p.Wout;
  for (int i = tid; i < 32 * HEAD_CO; i += 256) {
    const int o = i >> 5, r = i & 31;
    float v = red[0][r][o];
#pragma unroll
    for (int w = 1; w < 8; ++w) v += red[w][r][o];
    const int pp = tile * BLOCK_M + (b & 3) * 32 + r;   // linear position on the INPUT grid
    const int y = pp / p.Win, x = pp - y * p.Win;
    if (y < p.Hout && x < p.Wout)
      reinterpret_cast<float*>(p.out)[((size_t)img * HEAD_CO + o) * HW + (size_t)y * p.Wout + x] = v + sb[HEAD_CM + o];
  }
}

// ------------------------------------------------------------------------------------------------- weight gradient, tap groups
// conv_igemm_kernel's weight-gradient mode makes one unit per filter tap: dY and X are re-read nine times and every
// K step moves 16 KB + BN/64 x 8 KB through the SM's L2 port for 4 MMAs -- 94 to 188 B/clk against the ~50 B/clk the port
// sustains (profiles/r1b_conv_sweep.md).  Here a unit covers a GROUP of T taps of one (Cout tile, Cin tile): per K step
// (an 8 x 8 pixel patch) ONE dY box and ONE X box with the patch's halo are loaded, and tap (kh, kw) is an MN-major UMMA
// descriptor whose start is moved (kh * pitch + kw) pixel rows into the X box (8-pixel patch rows = 8-row descriptor
// groups, SBO = halo pitch: the scheme of conv_halo_kernel applied to the K dimension).  T accumulators of BN columns
// live in TMEM (T x BN <= 512, one stage); the epilogue reduce-adds them into dW[co][tap][ci] tap by tap.
__host__ __device__ constexpr int wgh_b_atom(int KMAX) { return (((8 + KMAX - 1) * (8 + KMAX - 1) * 128) + 1023) & ~1023; }
__host__ __device__ constexpr int wgh_stage_bytes(int BN) { return A_SUB_BYTES + (BN / 64) * wgh_b_atom(HALO_MAXK); }
__host__ __device__ constexpr int wgh_stages(int BN) {
  return (SMEM_LIMIT - SMEM_FIXED) / wgh_stage_bytes(BN) > 6 ? 6 : (SMEM_LIMIT - SMEM_FIXED) / wgh_stage_bytes(BN);
}
static int wgh_smem_bytes(int BN) { return wgh_stages(BN) * wgh_stage_bytes(BN) + SMEM_FIXED; }

template <int BN, int T>
__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_wgrad_halo_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvGroup grp) {
  constexpr int STAGES = wgh_stages(BN);
  constexpr int B_ATOM = wgh_b_atom(HALO_MAXK);
  constexpr int STAGE_BYTES = wgh_stage_bytes(BN);
  constexpr uint32_t IDESC_MN = ptx::make_idesc_bf16(BLOCK_M, BN, 1, 1);
  static_assert(T * BN <= 512 && STAGES >= 2, "accumulators must fit TMEM; at least two stages");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tile_buf = smem + STAGES * STAGE_BYTES;
  float* sbias = reinterpret_cast<float*>(tile_buf + STAGE_TILE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_units = grp.unit_end[0];
  GroupSched sc;
  make_sched(grp, BN, sc);
  const ConvParams& p = grp.p[0];
  const int PW = 8 + p.KW - 1, PH = 8 + p.KH - 1;
  const uint32_t tx_bytes = (uint32_t)(A_SUB_BYTES + (BN / 64) * PW * PH * 128);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&maps.a[0]);
    ptx::tma_prefetch_desc(&maps.b[0]);
    ptx::tma_prefetch_desc(&maps.o[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], EPI_THREADS / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        for (int k = t.k_begin; k < t.k_end; ++k) {
          // k -> (image, patch row, patch column) of the 8 x 8 output-pixel patch
          const int per_img = p.tiles_h * p.wchunks;
          const int n = k / per_img;
          const int r = k - n * per_img;
          const int ph = r / p.wchunks;
          const int pw = r - ph * p.wchunks;
          uint8_t* st = smem + stage * STAGE_BYTES;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
          for (int i = 0; i < BLOCK_M / 64; ++i)
            ptx::tma_load_4d(st + i * 8192, &maps.a[0], &full_bar[stage], t.h0 + 64 * i, pw * 8, ph * 8, n);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            ptx::tma_load_4d(st + A_SUB_BYTES + j * B_ATOM, &maps.b[0], &full_bar[stage], t.n0 + 64 * j, pw * 8 - p.padW, ph * 8 - p.padH, n);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (true) {   // the WHOLE warp, converged: one elected lane issues (ptx::mma_bf16_ss_w)
      int stage = 0;
      uint32_t phase = 0;
      int seq = 0;
      const int taps = p.KH * p.KW;
      const uint32_t sbo = (uint32_t)PW * 128u;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int gi;
        TileCoord t;
        if (!next_unit(grp, sc, unit, BN, gi, t)) continue;
        const uint32_t acc_phase = (uint32_t)seq & 1u;
        ++seq;
        ptx::mbar_wait(&tmem_empty[0], acc_phase ^ 1);
        ptx::tc_fence_after();
        const int ntap = min(T, taps - t.w0);
        for (int k = t.k_begin; k < t.k_end; ++k) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_SUB_BYTES;
          const uint64_t da = ptx::make_desc_mn_sw128(a_addr, 8192);
          for (int ti = 0; ti < ntap; ++ti) {
            const int tap = t.w0 + ti;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
#pragma unroll
            for (int j = 0; j < BLOCK_K / 16; ++j) {
              // K step j = patch rows 2j, 2j + 1: A rows [16j, 16j + 16) contiguous; X rows (kh + 2j) * PW + kw of the halo box
              const uint64_t db = ptx::make_desc_mn_sw128_sbo(b_addr + (uint32_t)(((kh + 2 * j) * PW + kw) * 128), B_ATOM, sbo);
              ptx::mma_bf16_ss_w(tmem_base + ti * BN, da + (uint64_t)(j * 2048 >> 4), db, IDESC_MN, (k > t.k_begin || j > 0) ? 1u : 0u);
            }
          }
          ptx::mma_commit_w(&empty_bar[stage]);
          if (k == t.k_end - 1) ptx::mma_commit_w(&tmem_full[0]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN, T, 1>(grp, maps.o, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, sc, warp - 4, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// Warp-local epilogue of the first layer (64 output channels, tile width BW <= 16 so that every 2x2 pooling window lies
// inside one warp's 32 pixel rows).  Warp e drains TMEM lane quarter (e & 3) of accumulator stage (e >> 2) -- the two
// stages are drained concurrently -- through its own 4 KB staging tile with __syncwarp only: the per-tile latency
// chain of the CTA-wide epilogue (two 256-thread barriers per tile) paced this K = 27 layer.
template <bool FOLDED, int NACC = 2>
__device__ __forceinline__ void epilogue_first_warp(const ConvParams& p, uint8_t* stage_base, const float* sbias, uint32_t tmem_base,
                                                    uint64_t* tmem_full, uint64_t* tmem_empty, int total_tiles, int ewarp, int lane) {
  constexpr int BN = 64;
  const int q = ewarp & 3, stg = ewarp >> 2;
  const int row = q * 32 + lane;
  const uint32_t st_addr = ptx::smem_u32(stage_base + ewarp * 4096);
  const uint32_t bias_addr = ptx::smem_u32(sbias);
  const int sw = lane & 7;
  const int dy = row >> p.bw_shift, dx = row & (p.BW - 1);
  const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
  const float k_neg = (slope - 1.0f) * p.scale, k_pos = p.scale;
  const int Hp = (p.Hout + 1) >> 1, Wp = (p.Wout + 1) >> 1;
  const int rows_per_warp = 32 >> p.bw_shift;  // tile rows covered by this warp (>= 2, even)
  int seq = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++seq) {
    if ((seq % NACC) != stg) continue;
    const uint32_t acc_phase = (uint32_t)(seq / NACC) & 1u;
    const TileCoord t = decode_tile(p, tile, BN, 1, 1);
    ptx::mbar_wait(&tmem_full[stg], acc_phase);
    ptx::tc_fence_after();
    const int h = t.h0 + dy, w = t.w0 + dx;
    const bool valid = (h < p.Hout) && (w < p.Wout);
    const uint32_t taddr = tmem_base + (uint32_t)(stg * BN) + ((uint32_t)(q * 32) << 16);
    __syncwarp();  // the previous tile's staging reads are done
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(taddr + half * 32, v);
      uint4 bq[8];
      if (!FOLDED) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = ptx::ld_shared_v4(bias_addr + (uint32_t)(half * 32) * 4u + j * 16);
      }
      ptx::tmem_ld_wait();
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x0 = __uint_as_float(v[4 * j]), x1 = __uint_as_float(v[4 * j + 1]);
        float x2 = __uint_as_float(v[4 * j + 2]), x3 = __uint_as_float(v[4 * j + 3]);
        float y0, y1, y2, y3;
        if (FOLDED) {  // bias arrived through the tensor core (two spare K slots); scale == 1: x + min(x, 0) * (slope - 1)
          y0 = fmaf(fminf(x0, 0.f), k_neg, x0);
          y1 = fmaf(fminf(x1, 0.f), k_neg, x1);
          y2 = fmaf(fminf(x2, 0.f), k_neg, x2);
          y3 = fmaf(fminf(x3, 0.f), k_neg, x3);
        } else {
          x0 += __uint_as_float(bq[j].x);
          x1 += __uint_as_float(bq[j].y);
          x2 += __uint_as_float(bq[j].z);
          x3 += __uint_as_float(bq[j].w);
          y0 = fmaf(fminf(x0, 0.f), k_neg, x0 * k_pos);
          y1 = fmaf(fminf(x1, 0.f), k_neg, x1 * k_pos);
          y2 = fmaf(fminf(x2, 0.f), k_neg, x2 * k_pos);
          y3 = fmaf(fminf(x3, 0.f), k_neg, x3 * k_pos);
        }
        if (p.chan_scale) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(p.chan_scale + (size_t)t.n_img * p.Cout + half * 32) + j);
          y0 *= m.x; y1 *= m.y; y2 *= m.z; y3 *= m.w;
        }
        o[2 * j] = ptx::pack_op16x2(y0, y1, p.f16);
        o[2 * j + 1] = ptx::pack_op16x2(y2, y3, p.f16);
      }
      if (p.mode == EPI_POOL && !valid) {
        const uint32_t ninf = p.f16 ? 0xFC00FC00u : 0xFF80FF80u;
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = ninf;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        ptx::st_shared_v4(st_addr + lane * 128 + (((half * 4 + c) ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    }
    // all TMEM reads of this warp's quarter are done: release the accumulator stage early
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&tmem_empty[stg]);
    if (p.mode == EPI_POOL) {
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * Hp * Wp * p.Cout;
      const int pw_shift = p.bw_shift - 1;
      const int pooled = 8;  // 32 pixels = 8 windows per warp
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int item = it * 32 + lane;           // (window, chunk)
        const int pp = item >> 3, c = item & 7;
        const int lppy = pp >> pw_shift, ppx = pp & ((p.BW >> 1) - 1);
        const int ph = ((t.h0 + q * rows_per_warp) >> 1) + lppy, pw = (t.w0 >> 1) + ppx;
        if (pp < pooled && ph < Hp && pw < Wp) {
          const int r00 = (2 * lppy) * p.BW + 2 * ppx;   // local row inside the warp's 32 rows
          const int r01 = r00 + 1, r10 = r00 + p.BW, r11 = r10 + 1;
          uint4 a = ptx::ld_shared_v4(st_addr + r00 * 128 + ((c ^ (r00 & 7)) << 4));
          const uint4 b = ptx::ld_shared_v4(st_addr + r01 * 128 + ((c ^ (r01 & 7)) << 4));
          const uint4 cc = ptx::ld_shared_v4(st_addr + r10 * 128 + ((c ^ (r10 & 7)) << 4));
          const uint4 d = ptx::ld_shared_v4(st_addr + r11 * 128 + ((c ^ (r11 & 7)) << 4));
          if (p.pool_arg) {
            const bf16* ea = reinterpret_cast<const bf16*>(&a);
            const bf16* eb = reinterpret_cast<const bf16*>(&b);
            const bf16* ec = reinterpret_cast<const bf16*>(&cc);
            const bf16* ed = reinterpret_cast<const bf16*>(&d);
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float best = __bfloat162float(ea[e]);
              uint32_t arg = 0;
              const float vb = __bfloat162float(eb[e]), vc = __bfloat162float(ec[e]), vd = __bfloat162float(ed[e]);
              if (vb > best) { best = vb; arg = 1; }
              if (vc > best) { best = vc; arg = 2; }
              if (vd > best) { best = vd; arg = 3; }
              if (e < 4) lo |= arg << (8 * e); else hi |= arg << (8 * (e - 4));
            }
            *reinterpret_cast<uint2*>(p.pool_arg + (((size_t)t.n_img * Hp + ph) * Wp + pw) * p.Cout + c * 8) = make_uint2(lo, hi);
          }
          a.x = ptx::max4_op16x2(a.x, b.x, cc.x, d.x, p.f16);
          a.y = ptx::max4_op16x2(a.y, b.y, cc.y, d.y, p.f16);
          a.z = ptx::max4_op16x2(a.z, b.z, cc.z, d.z, p.f16);
          a.w = ptx::max4_op16x2(a.w, b.w, cc.w, d.w, p.f16);
          *reinterpret_cast<uint4*>(out_img + ((size_t)ph * Wp + pw) * p.Cout + c * 8) = a;
        }
      }
    } else {
      bf16* out_img = reinterpret_cast<bf16*>(p.out) + (size_t)t.n_img * p.Hout * p.Wout * p.Cout;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 32 + lane;
        const int lr = idx >> 3, c = idx & 7;
        const int r = q * 32 + lr;
        const int hh = t.h0 + (r >> p.bw_shift), ww = t.w0 + (r & (p.BW - 1));
        if (hh < p.Hout && ww < p.Wout) {
          const uint4 val = ptx::ld_shared_v4(st_addr + lr * 128 + ((c ^ (lr & 7)) << 4));
          *reinterpret_cast<uint4*>(out_img + ((size_t)hh * p.Wout + ww) * p.Cout + c * 8) = val;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- first layer
// 3 x (3x3) first convolution on the fp32 NCHW frame (Detector.lua:32-33 uploads exactly this tensor).
// 640 threads: warps 0-7 build the im2col rows (K = 27 -> 32, 64 bytes of each 128-byte swizzled row) in two groups
// that alternate tiles, each thread prefetching its next tile's 27 taps into registers before it publishes the
// current one; warp 8 issues two tcgen05.mma (M128 x N64 x K16) per tile, warp 9 owns TMEM; warps 12-19 run the
// shared epilogue.  K is tiny, so this layer is paced by instruction latency of producers and epilogue: two warps
// of each role per scheduler.
static constexpr int FIRST_STAGES = 6;
static constexpr int FIRST_BN = 64;
static constexpr int FIRST_THREADS = 640;
static constexpr int FIRST_SMEM = FIRST_STAGES * A_SUB_BYTES + FIRST_BN * 128 + SMEM_FIXED + STAGE_TILE_BYTES;  // 8 x 4 KB warp staging

__device__ __forceinline__ void first_load_taps(const ConvParams& p, int tile, int dy, int dx, float (&v)[27]) {
  const TileCoord t = decode_tile(p, tile, FIRST_BN, 1, 1);
  const int h = t.h0 + dy - p.padH, w = t.w0 + dx - p.padW;  // top-left tap in input coordinates
  const size_t plane = (size_t)p.Hin * p.Win;
  const float* base = p.img + (size_t)t.n_img * 3 * plane;
  if (h >= 0 && h + 2 < p.Hin && w >= 0 && w + 2 < p.Win) {  // interior: no predicates
    const float* q = base + (size_t)h * p.Win + w;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) v[c * 9 + kh * 3 + kw] = __ldg(q + c * plane + kh * p.Win + kw);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int yy = h + kh, xx = w + kw;
          const bool in = yy >= 0 && yy < p.Hin && xx >= 0 && xx < p.Win;
          v[c * 9 + kh * 3 + kw] = in ? __ldg(base + c * plane + (size_t)yy * p.Win + xx) : 0.f;
        }
  }
}

__global__ void __launch_bounds__(FIRST_THREADS, 1)
    conv_first_kernel(const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvGroup grp,
                      const bf16* __restrict__ w32) {
  constexpr int BN = FIRST_BN;
  const ConvParams& p = grp.p[0];
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, BN);
  constexpr int TMEM_COLS = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + FIRST_STAGES * A_SUB_BYTES;
  uint8_t* tile_buf = smem_b + BN * 128;                      // 8 warps x 4 KB
  float* sbias = reinterpret_cast<float*>(tile_buf + 2 * STAGE_TILE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_bar = full_bar + FIRST_STAGES;
  uint64_t* tmem_full = empty_bar + FIRST_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles_m;

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < FIRST_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 128);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);  // four warps (one per TMEM lane quarter) drain each accumulator stage
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < MAX_BIAS; i += blockDim.x) sbias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  // weights [64][32] bf16 -> K-major 128-byte-swizzled rows (chunks 0..3 of every row)
  for (int i = threadIdx.x; i < BN * 4; i += blockDim.x) {
    const int n = i >> 2, c = i & 3;
    const uint4 v = reinterpret_cast<const uint4*>(w32 + (p.f16 ? BN * 32 : 0))[i];   // [bf16 copy | fp16 copy]
    *reinterpret_cast<uint4*>(smem_b + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp < 8) {
    // ------------------------------------------------------------ im2col producers: thread <-> pixel row
    const int group = warp >> 2;          // tiles with an even / odd local index
    const int row = threadIdx.x & 127;
    const int dy = row >> p.bw_shift, dx = row & (p.BW - 1);
    const int sw = row & 7;
    const int stride = 2 * gridDim.x;
    int tile = blockIdx.x + group * gridDim.x;
    int local = group;                    // position of `tile` in this CTA's tile sequence
    float v[27];
    if (tile < total_tiles) first_load_taps(p, tile, dy, dx, v);
    for (; tile < total_tiles; tile += stride, local += 2) {
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 13; ++j) o[j] = ptx::pack_op16x2(v[2 * j], v[2 * j + 1], p.f16);
      o[13] = ptx::pack_op16x2(v[26], 0.f, p.f16);
      o[14] = 0u;
      o[15] = 0u;
      if (tile + stride < total_tiles) first_load_taps(p, tile + stride, dy, dx, v);  // in flight across the wait
      const int stage = local % FIRST_STAGES;
      const uint32_t phase = (uint32_t)(local / FIRST_STAGES) & 1u;
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t dst = ptx::smem_u32(smem_a + stage * A_SUB_BYTES) + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::st_shared_v4(dst + ((c ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&full_bar[stage]);
    }
  } else if (warp == 8) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b));
      const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_SUB_BYTES));
        ptx::mma_bf16_ss(tmem_base + acc * BN, da, db, idesc, 0u);
        ptx::mma_bf16_ss(tmem_base + acc * BN, da + 2, db + 2, idesc, 1u);
        ptx::mma_commit(&empty_bar[stage]);
        ptx::mma_commit(&tmem_full[acc]);
        if (++stage == FIRST_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 12) {
    epilogue_first_warp<false>(p, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, total_tiles, warp - 12, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- first layer, TMA patches
// Same layer as conv_first_kernel, restructured after its ncu capture showed it instruction-bound (about 1000 warp
// instructions per warp and tile: 27 predicated scalar global loads with 64-bit address arithmetic per pixel, tile
// decoding by integer division, and an epilogue of four fp32 operations per output):
//  * the 3 x 10 x 18 input patch of a 16 x 8 tile arrives by ONE TMA load from the fp32 NCHW frame (tensor map over
//    the caller's frame, box {24, 10, 3}; zero padding and the borders are the TMA unit's out-of-bounds fill), so a
//    producer thread only gathers its 27 taps from shared memory;
//  * the bias rides through the tensor core: K slots 27 and 28 of every A row hold 1.0 and the matching weight columns
//    the bf16 pair (hi, lo) of the fp32 bias (hi + lo reproduces it to 2^-17 relative), which removes the bias add;
//  * with scale == 1 (block 1 has no dropout) PReLU is x + min(x, 0) * (slope - 1): two operations.
// Requires a 16-byte aligned frame pointer and Win % 4 == 0 (TMA global strides); other frames take conv_first_kernel.
// Measured on B200: cp.async.bulk.tensor raises "illegal instruction" when the byte offset of the box start along the
// innermost dimension is not a multiple of 16 -- the box therefore starts 4 pixels (16 bytes) left of the tile, not 1.
static constexpr int FT_PATCH_W = 24, FT_PATCH_H = 10, FT_PATCH_X0 = 4;
static constexpr int FT_PATCH_BYTES = FT_PATCH_W * FT_PATCH_H * 3 * 4;   // 2880
static constexpr int FT_PATCH_PITCH = 2944;  // 24 floats x 10 rows x 3 planes = 2880, 128-byte aligned
// The layer is paced by the latency of one tile's trip through an epilogue warp (TMEM load, activation, staging,
// pooling, store: ~3600 clocks), not by instruction issue: FOUR accumulator stages, each drained by its own four warps.
static constexpr int FT_ACC = 4;
// The input patches are tiny (2.9 KB each, 30 row segments of 96 B): with the patch ring as deep as the A ring (6) an SM had
// 17 KB in flight against a DRAM round trip of more than a microsecond -- the layer ran at 0.8 TB/s (profiles/
// r1b_ncu_full_b1.md).  The rings are decoupled: FT_PATCH_STAGES patches in flight, FT_A_STAGES im2col tiles.
static constexpr int FT_PATCH_STAGES = 16, FT_A_STAGES = 6;
static constexpr int FT_THREADS = 384 + FT_ACC * 128;   // 8 gather warps, MMA / TMEM / TMA / spare, 4 epilogue warps per stage
static constexpr int FT_SMEM = FT_A_STAGES * A_SUB_BYTES + FT_PATCH_STAGES * FT_PATCH_PITCH + FIRST_BN * 128 + FT_ACC * 4 * 4096 +
                               MAX_BIAS * 4 + 512 + 1024;

__global__ void __launch_bounds__(FT_THREADS, 1)
    conv_first_tma_kernel(const __grid_constant__ CUtensorMap tmImg, const __grid_constant__ ConvGroup grp, const bf16* __restrict__ w32) {
  constexpr int BN = FIRST_BN;
  const ConvParams& p = grp.p[0];
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, BN);
  constexpr int TMEM_COLS = FT_ACC * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + FT_A_STAGES * A_SUB_BYTES;
  uint8_t* patches = smem_b + BN * 128;
  uint8_t* tile_buf = patches + FT_PATCH_STAGES * FT_PATCH_PITCH;  // 16 warps x 4 KB
  float* sbias = reinterpret_cast<float*>(tile_buf + FT_ACC * 4 * 4096);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + MAX_BIAS);
  uint64_t* empty_bar = full_bar + FT_A_STAGES;
  uint64_t* patch_full = empty_bar + FT_A_STAGES;
  uint64_t* patch_empty = patch_full + FT_PATCH_STAGES;
  uint64_t* tmem_full = patch_empty + FT_PATCH_STAGES;
  uint64_t* tmem_empty = tmem_full + FT_ACC;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + FT_ACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles_m;

  if (warp == 8 && lane == 0) {
    ptx::tma_prefetch_desc(&tmImg);
    for (int s = 0; s < FT_A_STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 4);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < FT_PATCH_STAGES; ++s) {
      ptx::mbar_init(&patch_full[s], 1);
      ptx::mbar_init(&patch_empty[s], 4);   // the four warps of the gathering group
    }
    for (int a = 0; a < FT_ACC; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // weights [64][32] bf16 -> K-major 128-byte-swizzled rows (chunks 0..3 of every row); K slots 27 / 28 <- bias (hi, lo)
  for (int i = threadIdx.x; i < BN * 4; i += blockDim.x) {
    const int n = i >> 2, c = i & 3;
    uint4 v = reinterpret_cast<const uint4*>(w32 + (p.f16 ? BN * 32 : 0))[i];   // [bf16 copy | fp16 copy]
    if (c == 3) {  // elements 24..31: 27 = high half of v.y, 28 = low half of v.z
      const float b = (p.bias && n < p.Cout) ? p.bias[n] : 0.f;
      const uint16_t hi = ptx::float_to_op16(b, p.f16);
      const uint16_t lo = ptx::float_to_op16(b - ptx::op16_to_float(hi, p.f16), p.f16);
      v.y = (v.y & 0x0000FFFFu) | ((uint32_t)hi << 16);
      v.z = (v.z & 0xFFFF0000u) | (uint32_t)lo;
    }
    *reinterpret_cast<uint4*>(smem_b + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp < 8) {
    // ------------------------------------------------------------ gatherers: thread <-> pixel row, taps from the patch
    const int group = warp >> 2;          // tiles with an even / odd local index
    const int row = threadIdx.x & 127;
    const int dy = row >> p.bw_shift, dx = row & (p.BW - 1);
    const int sw = row & 7;
    int local = group;
    for (int tile = blockIdx.x + group * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, local += 2) {
      const int stage = local % FT_A_STAGES;
      const uint32_t phase = (uint32_t)(local / FT_A_STAGES) & 1u;
      const int pstage = local % FT_PATCH_STAGES;
      const uint32_t pphase = (uint32_t)(local / FT_PATCH_STAGES) & 1u;
      ptx::mbar_wait(&patch_full[pstage], pphase);
      constexpr int pw = FT_PATCH_W;
      const float* pt = reinterpret_cast<const float*>(patches + pstage * FT_PATCH_PITCH) + dy * pw + dx + (FT_PATCH_X0 - 1);
      float v[27];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) v[c * 9 + kh * 3 + kw] = pt[(c * FT_PATCH_H + kh) * pw + kw];
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&patch_empty[pstage]);
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 13; ++j) o[j] = ptx::pack_op16x2(v[2 * j], v[2 * j + 1], p.f16);
      o[13] = ptx::pack_op16x2(v[26], 1.0f, p.f16);   // K slot 27: 1.0 x bias_hi
      o[14] = ptx::pack_op16x2(1.0f, 0.f, p.f16);     // K slot 28: 1.0 x bias_lo
      o[15] = 0u;
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      const uint32_t dst = ptx::smem_u32(smem_a + stage * A_SUB_BYTES) + row * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::st_shared_v4(dst + ((c ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
      ptx::fence_proxy_async();   // every writer orders its own stores before the async proxy (the MMA) reads them
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&full_bar[stage]);   // one arrival per warp instead of 32
    }
  } else if (warp == 8) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b));
      const uint32_t idesc = p.f16 ? ptx::idesc_to_f16(IDESC) : IDESC;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_SUB_BYTES));
        ptx::mma_bf16_ss(tmem_base + acc * BN, da, db, idesc, 0u);
        ptx::mma_bf16_ss(tmem_base + acc * BN, da + 2, db + 2, idesc, 1u);
        ptx::mma_commit(&empty_bar[stage]);
        ptx::mma_commit(&tmem_full[acc]);
        if (++stage == FT_A_STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++acc == FT_ACC) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------ patch loader: one TMA box per tile
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile, BN, 1, 1);
        ptx::mbar_wait(&patch_empty[stage], phase ^ 1);
        ptx::mbar_arrive_expect_tx(&patch_full[stage], FT_PATCH_BYTES);
        ptx::tma_load_4d(patches + stage * FT_PATCH_PITCH, &tmImg, &patch_full[stage], t.w0 - FT_PATCH_X0, t.h0 - p.padH, 0, t.n_img);
        if (++stage == FT_PATCH_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 12) {
    epilogue_first_warp<true, FT_ACC>(p, tile_buf, sbias, tmem_base, tmem_full, tmem_empty, total_tiles, warp - 12, lane);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FRCNN_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    FRCNN_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, FRCNN_E_CUDA,
                  "cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// NHWC bf16 tensor [N][H][W][C], box {64 ch, BW, BH, 1}, 128-byte swizzle: used for loads (A operand) and stores
void make_tmap_act(CUtensorMap* m, const bf16* base, int N, int H, int W, int C, int BW, int BH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)BW, (cuuint32_t)BH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA,
                "cuTensorMapEncodeTiled(activation) failed, CUresult " + std::to_string((int)r) + " dims " +
                    std::to_string(C) + "x" + std::to_string(W) + "x" + std::to_string(H) + "x" + std::to_string(N) + " box " +
                    std::to_string(BW) + "x" + std::to_string(BH));
}

// K-major weight rows [copies][Cout][K] of 16-bit operands: box {64, BN, 1}; the third coordinate selects the copy
// (0 = bf16, 1 = fp16 where the packer wrote both; ConvParams::f16).  The element type only sets the element size.
void make_tmap_weight(CUtensorMap* m, const bf16* base, int Cout, int K, int BN, int copies) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, (cuuint64_t)copies};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Cout * K * 2};
  cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA,
                "cuTensorMapEncodeTiled(weight) failed, CUresult " + std::to_string((int)r));
}

// Pick the BW x BH (= 128) rectangle that wastes the fewest padded pixels.  mt: sub-tiles stacked along H;
// even: both sides even (needed by the fused 2x2 pool).
static void choose_tile(int Hout, int Wout, int mt, bool even, int* BW, int* BH, int max_bw = 128) {
  long best = -1;
  for (int bw = std::min(even ? 64 : 128, max_bw); bw >= (even ? 2 : 1); bw >>= 1) {
    int bh = 128 / bw;
    long th = (long)bh * mt;
    long padded = (long)((Wout + bw - 1) / bw) * bw * (long)((Hout + th - 1) / th) * th;
    if (best < 0 || padded < best) {
      best = padded;
      *BW = bw;
      *BH = bh;
    }
  }
}
void conv_choose_tile(int Hout, int Wout, int* BW, int* BH) { choose_tile(Hout, Wout, 1, false, BW, BH); }

static int choose_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 192 == 0) return 192;
  if (Cout % 128 == 0) return 128;
  if (Cout <= 64) return 64;
  return 128;
}
static void fill_geometry(ConvParams& p, int N, int Hin, int Win, int Cin, int Cout, int KH, int KW, int padH, int padW,
                          int mode, int MT, int max_bw = 128) {
  p = ConvParams();
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin;
  p.Hout = Hin + 2 * padH - KH + 1;
  p.Wout = Win + 2 * padW - KW + 1;
  FRCNN_REQUIRE(p.Hout > 0 && p.Wout > 0, FRCNN_E_INVALID, "conv: input smaller than the kernel");
  p.Cout = Cout; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  p.MT = MT;
  choose_tile(p.Hout, p.Wout, MT, mode == EPI_POOL || max_bw < 128, &p.BW, &p.BH, max_bw);
  p.bw_shift = 0;
  while ((1 << p.bw_shift) < p.BW) ++p.bw_shift;
  p.tiles_w = (p.Wout + p.BW - 1) / p.BW;
  p.tiles_h = (p.Hout + p.BH * MT - 1) / (p.BH * MT);
  p.n_tiles_m = N * p.tiles_h * p.tiles_w;
  p.mode = mode;
  p.scale = 1.0f;
}

static void make_out_map(ConvLaunch* L, bf16* out) {
  L->p.out = out;          // bf16 NHWC map (EPI_STORE) or its 2x2-pooled map (EPI_POOL); fp32 modes: conv_set_f32_output
  L->tmOut = L->tmB;       // only EPI_F32_REDUCE uses an output tensor map
}

void conv_set_f32_output(ConvLaunch* L, float* ws) {
  const ConvParams& p = L->p;
  FRCNN_REQUIRE((p.mode == EPI_F32_SLICES || p.mode == EPI_F32_REDUCE) && ws != nullptr, FRCNN_E_INVALID,
                "conv: not an fp32 split-K launch");
  L->p.out = ws;
  if (p.mode == EPI_F32_SLICES) return;  // plain stores, no tensor map
  const int S = p.N;
  cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wout, (cuuint64_t)p.Hout, (cuuint64_t)S};
  cuuint64_t strides[3] = {(cuuint64_t)p.Cout * 4, (cuuint64_t)p.Wout * p.Cout * 4, (cuuint64_t)p.Hout * p.Wout * p.Cout * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)p.BH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&L->tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA, "cuTensorMapEncodeTiled(fp32 slices) failed, CUresult " + std::to_string((int)r));
}

// Environment switches for A/B measurements: FRCNN_CONV_HALO=0 keeps every layer on the tap-per-box kernel;
// FRCNN_HALO_DESC=1 fills the descriptor's base-offset field with the row phase of the shifted start address.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && v[0] ? atoi(v) : dflt;
}

static bool halo_cfg_ok(int Cout, int BN, int MT) {
  return Cout % BN == 0 && BN * MT <= 512 && (BN == 64 || BN == 128 || BN == 192 || BN == 256) && (MT == 1 || MT == 2);
}

static void conv_prepare_halo(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                              int KH, int KW, int padH, int padW, int mode, bf16* out, int num_sms, int BN, int MT, int w_copies,
                              int occ = 1) {
  FRCNN_REQUIRE(occ == 1 || (occ == 2 && BN * MT <= 256), FRCNN_E_INVALID, "conv: two CTAs per SM need BN x MT <= 256 TMEM columns");
  L->w_copies = w_copies;
  FRCNN_REQUIRE(halo_cfg_ok(Cout, BN, MT), FRCNN_E_INVALID, "conv (halo kernel): unsupported (BN, MT) for this Cout");
  L->BN = BN;
  L->first = false;
  L->w_first = nullptr;
  ConvParams& p = L->p;
  p = ConvParams();
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin;
  p.Hout = Hin + 2 * padH - KH + 1;
  p.Wout = Win + 2 * padW - KW + 1;
  FRCNN_REQUIRE(p.Hout > 0 && p.Wout > 0, FRCNN_E_INVALID, "conv: input smaller than the kernel");
  p.Cout = Cout; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  p.MT = MT;
  p.BW = HALO_BW; p.BH = HALO_BH; p.bw_shift = 3;
  p.tiles_w = (p.Wout + p.BW - 1) / p.BW;
  p.tiles_h = (p.Hout + p.BH * MT - 1) / (p.BH * MT);
  p.n_tiles_m = N * p.tiles_h * p.tiles_w;
  p.n_tiles_n = Cout / BN;
  p.cchunks = Cin / 64;
  p.k_iters = KH * KW * p.cchunks;
  p.splits = 1;
  p.k_per_split = p.k_iters;
  p.mode = mode;
  p.scale = 1.0f;
  p.halo = 1;
  p.halo_desc = env_int("FRCNN_HALO_DESC", 0);
  p.dbg = env_int("FRCNN_CONV_DBG", 0);
  make_tmap_act(&L->tmA, in, N, Hin, Win, Cin, HALO_BW + KW - 1, HALO_BH * MT + KH - 1);
  make_tmap_weight(&L->tmB, w_packed, Cout, KH * KW * Cin, BN, w_copies);
  make_out_map(L, out);
  p.occ = occ;
  const int total = p.n_tiles_m * p.n_tiles_n;
  L->grid = total < occ * num_sms ? total : occ * num_sms;
}

// conv_pair_kernel (cta_group::2): same tiling as the halo kernel, a unit = two CTA tiles stacked along H, the weight
// box is half an N tile (each CTA of the pair loads BN / 2 rows)
static void conv_prepare_pair(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                              int KH, int KW, int padH, int padW, int mode, bf16* out, int num_sms, int BN, int MT, int w_copies,
                              int occ = 1) {
  FRCNN_REQUIRE(halo_cfg_ok(Cout, BN, MT) && BN % 16 == 0, FRCNN_E_INVALID, "conv (pair kernel): unsupported (BN, MT) for this Cout");
  conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, BN, MT, w_copies, occ);
  ConvParams& p = L->p;
  p.pair = 1;
  p.tiles_h = (p.Hout + 2 * p.BH * MT - 1) / (2 * p.BH * MT);
  p.n_tiles_m = N * p.tiles_h * p.tiles_w;
  make_tmap_weight(&L->tmB, w_packed, Cout, KH * KW * Cin, BN / 2, w_copies);
  L->tmOut = L->tmB;
  const int total = p.n_tiles_m * p.n_tiles_n;
  const int pairs = std::max(1, occ * num_sms / 2);
  L->grid = 2 * (total < pairs ? total : pairs);
}

// conv_pair_bres_kernel: one N tile (BN == Cout), all 9 * Cin / 64 half boxes resident
static bool bres_ok(int Cin, int Cout, int BN, int KH, int KW) {
  return (BN == 64 || BN == 128) && Cout % BN == 0 && KH == 3 && KW == 3 && KH * KW * (Cin / 64) <= BRES_MAX_BOXES &&
         bres_a_slots(BN, KH * KW * (Cin / 64)) >= 2;
}
static void conv_prepare_pair_bres(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                                   int KH, int KW, int padH, int padW, int mode, bf16* out, int num_sms, int BN, int w_copies) {
  FRCNN_REQUIRE(bres_ok(Cin, Cout, BN, KH, KW), FRCNN_E_INVALID, "conv (pair kernel, resident weights): 3x3, N tile 64 / 128, at most 18 weight boxes");
  conv_prepare_pair(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, BN, 1, w_copies, 1);
  // a pair keeps the weights of ONE N tile: units are enumerated N-tile-fastest and dealt round-robin, so the pair count
  // must be a multiple of the N tile count for every pair to stay on its tile
  const int ntn = L->p.n_tiles_n;
  int pairs = L->grid / 2;
  if (pairs > ntn) pairs -= pairs % ntn;
  L->grid = 2 * pairs;
  FRCNN_REQUIRE(pairs % ntn == 0 || pairs < ntn, FRCNN_E_INVALID, "conv (pair kernel, resident weights): pair count / N tile mismatch");
  L->p.occ = 3;
  L->p.a_slots = bres_a_slots(BN, KH * KW * (Cin / 64));
}

void conv_prepare_head(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int K, int num_sms) {
  FRCNN_REQUIRE(Cin % 64 == 0 && K >= 1 && K <= HEAD_MAXK, FRCNN_E_INVALID, "fused anchor head: Cin % 64 == 0, kernel size <= 7");
  FRCNN_REQUIRE(Hin >= K && Win >= K, FRCNN_E_INVALID, "fused anchor head: input smaller than the kernel");
  L->BN = HEAD_CM;
  L->first = false;
  L->w_first = nullptr;
  ConvParams& p = L->p;
  p = ConvParams();
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin;
  p.Hout = Hin - K + 1;
  p.Wout = Win - K + 1;
  p.Cout = HEAD_CM; p.KH = K; p.KW = K; p.padH = 0; p.padW = 0;
  p.MT = 1;
  p.BW = HALO_BW; p.BH = HALO_BH; p.bw_shift = 3;
  p.tiles_w = (p.Wout + p.BW - 1) / p.BW;
  p.tiles_h = (p.Hout + p.BH - 1) / p.BH;
  p.n_tiles_m = N * p.tiles_h * p.tiles_w;
  p.n_tiles_n = 1;
  p.cchunks = Cin / 64;
  p.k_iters = K * K * p.cchunks;
  p.splits = 1;
  p.k_per_split = p.k_iters;
  p.mode = EPI_HEAD;
  p.scale = 1.0f;
  p.halo = 1;
  p.halo_desc = env_int("FRCNN_HALO_DESC", 0);
  make_tmap_act(&L->tmA, in, N, Hin, Win, Cin, HALO_BW + K - 1, HALO_BH + K - 1);
  make_tmap_weight(&L->tmB, w_packed, HEAD_CM, K * K * Cin, HEAD_CM, 2);   // forward weights: [bf16 | fp16] copies
  L->w_copies = 2;
  L->tmOut = L->tmB;
  L->grid = p.n_tiles_m < num_sms ? p.n_tiles_m : num_sms;
}

void conv_prepare(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                  int KH, int KW, int padH, int padW, int mode, bf16* out, int num_sms, int force_splits, int force_bn,
                  int force_mt, int w_copies) {
  L->w_copies = w_copies;
  FRCNN_REQUIRE(Cin % 64 == 0, FRCNN_E_INVALID, "conv: Cin must be a multiple of 64");
  const bool f32 = mode == EPI_F32_REDUCE || mode == EPI_F32_SLICES;
  FRCNN_REQUIRE(f32 ? Cout % 32 == 0 : (Cout % 64 == 0 && Cout <= MAX_BIAS), FRCNN_E_INVALID,
                "conv: Cout must be a multiple of 64 (<= 512) for the bf16 epilogues, of 32 for split-K");
  FRCNN_REQUIRE(f32 || out != nullptr, FRCNN_E_INVALID, "conv: null output");
  // halo-tile kernel: k x k (k = 2, 3) filters with a bf16 epilogue.  force_mt: 0 = automatic, 1 / 2 = tap-per-box
  // kernel with that MT, 11 / 12 = halo kernel with MT 1 / 2
  {
    // (also the unsplit fp32 reduce-add mode: the data gradients that accumulate into a block gradient, where the
    // tap-per-box kernel's narrow N tiles are the most ingress-bound)
    const bool f32_single = mode == EPI_F32_REDUCE && force_splits == 1;
    const bool eligible = (!f32 || f32_single) && KH <= HALO_MAXK && KW <= HALO_MAXK && KH * KW > 1 && Cout % 64 == 0;
    const bool forced = force_mt >= 10;
    FRCNN_REQUIRE(!forced || eligible, FRCNN_E_INVALID, "conv: the halo kernel needs 2x2..3x3 filters and a bf16 epilogue");
    // CTA-pair kernel (cta_group::2): force_mt 21 / 22, or automatically (FRCNN_CONV_PAIR, see the selection rule below)
    const bool pair_ok = eligible && !f32 && Cout % 64 == 0;
    if (pair_ok && force_mt == 61) {   // swapped operands: 128 filters (M) x 256 pixels (N)
      FRCNN_REQUIRE(Cout % 128 == 0 && KH == 3 && KW == 3, FRCNN_E_INVALID, "conv (swapped operands): 3x3, Cout a multiple of 128");
      conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, 128, 2, w_copies, 1);
      L->p.swap = 1;
      return;
    }
    if (pair_ok && force_mt == 51) {   // CTA pairs with the layer's weights resident in shared memory
      conv_prepare_pair_bres(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms,
                             force_bn > 0 ? force_bn : (Cout % 128 == 0 ? 128 : 64), w_copies);
      return;
    }
    if (pair_ok && force_mt > 20) {
      // 21 / 22: CTA pairs; 31 / 32: halo kernel with two CTAs per SM; 41 / 42: CTA pairs with two CTAs per SM
      const int mt = force_mt % 10, kind = force_mt / 10;
      FRCNN_REQUIRE((mt == 1 || mt == 2) && kind >= 2 && kind <= 4, FRCNN_E_INVALID, "conv: bad kernel selector");
      int bn = force_bn > 0 ? force_bn : (Cout % 256 == 0 ? 256 : (Cout % 192 == 0 ? 192 : (Cout % 128 == 0 ? 128 : 64)));
      if (kind >= 3 && bn * mt > 256 && force_bn == 0) bn = Cout % 128 == 0 ? 128 : 64;
      FRCNN_REQUIRE(halo_cfg_ok(Cout, bn, mt), FRCNN_E_INVALID, "conv: no pair-kernel tile for this (Cout, bn, mt)");
      if (kind == 3) conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, bn, mt, w_copies, 2);
      else conv_prepare_pair(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, bn, mt, w_copies, kind == 4 ? 2 : 1);
      return;
    }
    FRCNN_REQUIRE(force_mt < 20, FRCNN_E_INVALID, "conv: the pair kernel needs 2x2..3x3 filters and a bf16 epilogue");
    // Automatic choice (FRCNN_CONV_DUO=0 restores the one-CTA-per-SM rule below).  Measured on B200
    // (tools/bench_conv_layers.py, profiles/r2_conv_sweep.md): every trunk layer is at least as fast with TWO resident
    // CTAs per SM (the tensor pipe of a lone CTA idles through its prologue, first-load latency and last epilogue), and
    // kernels sized that way can share an SM with the kernels of the other frames in flight:
    //   * wide layers with enough work for two waves of CTA pairs: conv_pair_kernel, BN = 256 / 192, M = 256 MMAs;
    //   * otherwise conv_halo_kernel with BN = 128 (Cout = 128 layers; the wide layers at batch 1), or BN = 64.
    //   * force_mt = -1 (the throughput schedule: several frames in flight, the SMs a launch leaves idle are filled by the
    //     other frames' kernels, so a launch costs its SM-time = sum of CTA residencies, not its duration): wide layers
    //     WITHOUT two waves of pairs run on conv_halo_kernel with the 256-pixel tile and ONE CTA per SM -- few, long-lived
    //     CTAs amortise prologue / first-load latency / last epilogue (profiles/r2_conv_smtime.md: conv4_2 at batch 1
    //     25.4 -> 13.1 SM-us although the launch alone takes 42 instead of 31 us).
    const bool sm_time = force_mt == -1;
    if (sm_time) force_mt = 0;
    if (pair_ok && force_mt == 0 && force_bn == 0 && env_int("FRCNN_CONV_DUO", 1)) {
      const int Ho = Hin + 2 * padH - KH + 1, Wo = Win + 2 * padW - KW + 1;
      const int wide = Cout % 256 == 0 ? 256 : (Cout % 192 == 0 ? 192 : 0);
      // Re-measured after the issue-rate fix (converged-warp tcgen05.mma issue; profiles/r2_conv_smtime2_*.md):
      //   * swapped operands (filters as M, 256 pixels as N) win where the reduction is short or the N tile would be 192:
      //     the first conv of block 2 (Cin = 64: 19.5 instead of 21.9 SM-us at batch 1, 123 instead of 135 us at batch 8)
      //     and the Cout = 384 layers once there are two waves of units (conv4_x at batch 8: 65 / 92 instead of 72 / 96 us);
      //   * Cout = 256 layers with two waves of pairs: CTA pairs with ONE CTA per SM (87 / 146 us at batch 8 against 98 / 156
      //     with two per SM: the deeper rings matter more than the overlap now that the MMA phase is shorter);
      //   * Cout = 64 (vgg_large conv1_2): CTA pairs with 256-pixel tiles, two per SM (65 instead of 73 SM-us).
      const bool swap_ok = KH == 3 && KW == 3 && Cout % 128 == 0 && env_int("FRCNN_CONV_SWAP", 1);
      auto prepare_swap = [&]() {
        conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, 128, 2, w_copies, 1);
        L->p.swap = 1;
      };
      if (swap_ok && Cin == 64) {
        prepare_swap();
        return;
      }
      if (wide) {
        const long pair_ctas = 2L * N * ((Ho + 2 * HALO_BH - 1) / (2 * HALO_BH)) * ((Wo + HALO_BW - 1) / HALO_BW) * (Cout / wide);
        if (pair_ctas >= 4L * num_sms || Cout % 128 != 0) {
          if (wide == 192 && swap_ok) prepare_swap();
          else conv_prepare_pair(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, wide, 1, w_copies, wide == 256 ? 1 : 2);
          return;
        }
        if (sm_time && halo_cfg_ok(Cout, wide, 2) && env_int("FRCNN_CONV_SMTIME", 1)) {
          conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, wide, 2, w_copies, 1);
          return;
        }
        if (wide == 192 && swap_ok) {   // latency schedule, few units: 22 / 31 us alone against 25 / 31 with two CTAs per SM
          prepare_swap();
          return;
        }
      }
      if (Cout == 64 && halo_cfg_ok(Cout, 64, 2)) {
        conv_prepare_pair(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, 64, 2, w_copies, 2);
        return;
      }
      conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, Cout % 128 == 0 ? 128 : 64, 1,
                        w_copies, 2);
      return;
    }
    if (eligible && (forced || (force_mt == 0 && env_int("FRCNN_CONV_HALO", 1)))) {
      // Measured on B200 (tools/bench_conv_layers.py sweep, profiles/r1_conv_sweep.md): the halo kernel wins where the
      // N tile is wide (BN >= 192: the M128 x N128 MMA is bound by shared-memory operand bandwidth whichever way its A
      // operand arrives, and the tap-per-box kernel is as fast there); with BN = 256 two accumulator stages (MT = 1)
      // beat the larger tile as soon as there are two waves of units, with BN = 192 the 256-pixel tile (one stage of
      // 384 columns) wins from two waves on.
      int bn = 0, mt = 0;
      const int Ho = Hin + 2 * padH - KH + 1, Wo = Win + 2 * padW - KW + 1;
      auto units = [&](int cbn, int cmt) {
        return (long)N * ((Ho + HALO_BH * cmt - 1) / (HALO_BH * cmt)) * ((Wo + HALO_BW - 1) / HALO_BW) * (Cout / cbn);
      };
      if (forced) {
        mt = force_mt - 10;
        bn = force_bn > 0 ? force_bn : (Cout % 256 == 0 ? 256 : (Cout % 192 == 0 ? 192 : (Cout % 128 == 0 ? 128 : 64)));
        if (!halo_cfg_ok(Cout, bn, mt)) bn = 0;
      } else if (f32_single && force_bn == 0 && Cout % 192 != 0 && Cout % 256 != 0) {
        bn = Cout % 128 == 0 ? 128 : 64;
        mt = 2;
      } else if (force_bn == 0 || force_bn >= 192) {
        if ((force_bn == 0 || force_bn == 256) && Cout % 256 == 0) {
          bn = 256;
          mt = units(256, 1) >= 2L * num_sms ? 1 : 2;
        } else if ((force_bn == 0 || force_bn == 192) && Cout % 192 == 0) {
          bn = 192;
          mt = units(192, 2) >= 2L * num_sms ? 2 : 1;
        }
      }
      if (bn) {
        conv_prepare_halo(L, in, w_packed, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, out, num_sms, bn, mt, w_copies);
        return;
      }
      FRCNN_REQUIRE(!forced, FRCNN_E_INVALID, "conv: no halo-kernel tile for this (Cout, bn, mt)");
    }
  }
  // the widest tile dividing Cout wins even when it leaves SMs idle (measured at batch 1, conv4_x: 100 CTAs of BN=192
  // take 20/26 us, 135 CTAs of BN=128 take 34/46 us -- operand traffic per MAC, not occupancy, bounds these layers)
  const int BN = force_bn > 0 ? force_bn : choose_bn(Cout);
  L->BN = BN;
  L->first = false;
  L->w_first = nullptr;
  // two 128-row sub-tiles per CTA for the narrow layers (halves the weight traffic per MAC) when there are enough
  // tiles left to fill the machine at least twice
  int MT = 1;
  if (force_mt > 0) MT = force_mt;
  else if (!f32 && BN <= 128) {
    long px = (long)N * (Hin + 2 * padH - KH + 1) * (Win + 2 * padW - KW + 1);
    if (px / 256 * ((Cout + BN - 1) / BN) >= 2L * num_sms) MT = 2;
  }
  FRCNN_REQUIRE(MT == 1 || (MT == 2 && BN <= 128), FRCNN_E_INVALID, "conv: MT = 2 needs BN <= 128 (TMEM columns)");
  fill_geometry(L->p, N, Hin, Win, Cin, Cout, KH, KW, padH, padW, mode, MT);
  ConvParams& p = L->p;
  p.n_tiles_n = (Cout + BN - 1) / BN;
  p.cchunks = Cin / 64;
  p.k_iters = KH * KW * p.cchunks;
  int splits = 1;
  if (mode == EPI_F32_REDUCE || mode == EPI_F32_SLICES) {
    if (force_splits > 0) {
      splits = force_splits;
    } else {
      // enough CTAs to fill the machine, but keep >= 8 K iterations per split
      int base = p.n_tiles_m * p.n_tiles_n;
      splits = (num_sms + base - 1) / base;
      int max_splits = p.k_iters / 8 > 0 ? p.k_iters / 8 : 1;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > p.k_iters) splits = p.k_iters;
  p.k_per_split = (p.k_iters + splits - 1) / splits;
  p.splits = (p.k_iters + p.k_per_split - 1) / p.k_per_split;
  p.dbg = env_int("FRCNN_CONV_DBG", 0);
  make_tmap_act(&L->tmA, in, N, Hin, Win, Cin, p.BW, p.BH * MT);
  make_tmap_weight(&L->tmB, w_packed, Cout, KH * KW * Cin, BN, w_copies);
  make_out_map(L, out);
  int total = p.n_tiles_m * p.n_tiles_n * p.splits;
  L->grid = total < num_sms ? total : num_sms;
}

void conv_wgrad_prepare(ConvLaunch* L, const bf16* dy, const bf16* x, float* dw_taps, int N, int Hin, int Win, int Cin, int Cout,
                        int KH, int KW, int padH, int padW, int num_sms) {
  FRCNN_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, FRCNN_E_INVALID, "wgrad: channel counts must be multiples of 64");
  ConvParams& p = L->p;
  p = ConvParams();
  L->first = false;
  L->w_first = nullptr;
  L->w_copies = 1;
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  p.Hout = Hin + 2 * padH - KH + 1;
  p.Wout = Win + 2 * padW - KW + 1;
  FRCNN_REQUIRE(p.Hout > 0 && p.Wout > 0, FRCNN_E_INVALID, "wgrad: input smaller than the kernel");
  const int BN = Cin % 256 == 0 ? 256 : (Cin % 192 == 0 ? 192 : (Cin % 128 == 0 ? 128 : 64));
  L->BN = BN;
  p.MT = 1;
  p.wgrad = 1;
  if (KH <= HALO_MAXK && KW <= HALO_MAXK && KH * KW > 1 && env_int("FRCNN_WGRAD_HALO", 1)) {
    // conv_wgrad_halo_kernel: units of T filter taps sharing one dY box and one X halo box per 8 x 8 pixel patch
    const int T = BN == 64 ? 5 : (BN == 128 ? 3 : 2);
    const int groups = (KH * KW + T - 1) / T;
    p.halo = 1;
    p.MT = T;
    p.BW = 8; p.BH = 8; p.bw_shift = 3;
    p.co_tiles = (Cout + BLOCK_M - 1) / BLOCK_M;
    p.wchunks = (p.Wout + 7) / 8;
    p.tiles_h = (p.Hout + 7) / 8;
    p.tiles_w = groups * p.co_tiles;
    p.n_tiles_m = groups * p.co_tiles;
    p.n_tiles_n = (Cin + BN - 1) / BN;
    p.cchunks = 1;
    p.k_iters = N * p.tiles_h * p.wchunks;
    p.mode = EPI_F32_REDUCE;
    p.scale = 1.f;
    const int base = p.n_tiles_m * p.n_tiles_n;
    int splits = (2 * num_sms + base - 1) / base;
    splits = std::max(1, std::min(splits, std::max(1, p.k_iters / 8)));
    p.k_per_split = (p.k_iters + splits - 1) / splits;
    p.splits = (p.k_iters + p.k_per_split - 1) / p.k_per_split;
    make_tmap_act(&L->tmA, dy, N, p.Hout, p.Wout, Cout, 8, 8);
    make_tmap_act(&L->tmB, x, N, Hin, Win, Cin, 8 + KW - 1, 8 + KH - 1);
    p.out = dw_taps;
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)(KH * KW), (cuuint64_t)Cout, 1};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)KH * KW * Cin * 4, (cuuint64_t)Cout * KH * KW * Cin * 4};
    cuuint32_t box[4] = {32, 1, BLOCK_M, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode()(&L->tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dw_taps, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA, "cuTensorMapEncodeTiled(wgrad out) failed, CUresult " + std::to_string((int)r));
    const int total = p.n_tiles_m * p.n_tiles_n * p.splits;
    L->grid = total < num_sms ? total : num_sms;
    return;
  }
  // K chunks = BW x BH = 64 output pixels: the rectangle that wastes the fewest padded pixels
  long best = -1;
  for (int bw = 64; bw >= 1; bw >>= 1) {
    const int bh = 64 / bw;
    const long padded = (long)((p.Wout + bw - 1) / bw) * bw * (long)((p.Hout + bh - 1) / bh) * bh;
    if (best < 0 || padded < best) {
      best = padded;
      p.BW = bw;
      p.BH = bh;
    }
  }
  p.bw_shift = 0;
  while ((1 << p.bw_shift) < p.BW) ++p.bw_shift;
  p.co_tiles = (Cout + BLOCK_M - 1) / BLOCK_M;
  p.wchunks = (p.Wout + p.BW - 1) / p.BW;           // patch columns
  p.tiles_h = (p.Hout + p.BH - 1) / p.BH;           // patch rows
  p.tiles_w = KH * KW * p.co_tiles;
  p.n_tiles_m = KH * KW * p.co_tiles;
  p.n_tiles_n = (Cin + BN - 1) / BN;
  p.cchunks = 1;
  p.k_iters = N * p.tiles_h * p.wchunks;
  p.mode = EPI_F32_REDUCE;
  p.scale = 1.f;
  // split the pixel dimension so that the units fill the machine about twice, at least 8 K iterations each
  const int base = p.n_tiles_m * p.n_tiles_n;
  int splits = (2 * num_sms + base - 1) / base;
  splits = std::max(1, std::min(splits, std::max(1, p.k_iters / 8)));
  p.k_per_split = (p.k_iters + splits - 1) / splits;
  p.splits = (p.k_iters + p.k_per_split - 1) / p.k_per_split;
  make_tmap_act(&L->tmA, dy, N, p.Hout, p.Wout, Cout, p.BW, p.BH);
  make_tmap_act(&L->tmB, x, N, Hin, Win, Cin, p.BW, p.BH);
  // output dW[co][tap][ci] fp32: dims {Cin, taps, Cout, 1}, box {32 ci, 1 tap, 128 co, 1} = the staging tile
  p.out = dw_taps;
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)(KH * KW), (cuuint64_t)Cout, 1};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)KH * KW * Cin * 4, (cuuint64_t)Cout * KH * KW * Cin * 4};
  cuuint32_t box[4] = {32, 1, BLOCK_M, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&L->tmOut, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dw_taps, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA, "cuTensorMapEncodeTiled(wgrad out) failed, CUresult " + std::to_string((int)r));
  const int total = p.n_tiles_m * p.n_tiles_n * p.splits;
  L->grid = total < num_sms ? total : num_sms;
}

void conv_first_prepare(ConvLaunch* L, const bf16* w_packed32, int N, int Hin, int Win, int Cimg, int Cout, int KH,
                        int KW, int padH, int padW, int mode, bf16* out, int num_sms) {
  FRCNN_REQUIRE(Cimg == 3 && KH == 3 && KW == 3 && Cout == FIRST_BN, FRCNN_E_INVALID,
                "first-layer kernel: 3 input planes, 3x3 filters, 64 outputs (models/vgg_*.lua)");
  FRCNN_REQUIRE(mode == EPI_STORE || mode == EPI_POOL, FRCNN_E_INVALID, "first-layer kernel: bf16 epilogues only");
  L->BN = FIRST_BN;
  L->first = true;
  L->w_first = w_packed32;   // [2][Cout][32]: bf16 copy, fp16 copy
  L->w_copies = 2;
  fill_geometry(L->p, N, Hin, Win, 64, Cout, KH, KW, padH, padW, mode, 1, 16);  // BW <= 16: warp-local pooling windows
  ConvParams& p = L->p;
  if (Win % 4 == 0 && padH == 1 && padW == 1 && env_int("FRCNN_FIRST_TMA", 1)) {
    // conv_first_tma_kernel: fixed 16 x 8 tiles (its patch box is {20, 10, 3}); chosen at launch when the frame
    // pointer is 16-byte aligned and scale == 1
    p.BW = 16; p.BH = 8; p.bw_shift = 4;
    p.tiles_w = (p.Wout + p.BW - 1) / p.BW;
    p.tiles_h = (p.Hout + p.BH - 1) / p.BH;
    p.n_tiles_m = N * p.tiles_h * p.tiles_w;
    p.first_tma = 1;
  }
  p.Cimg = Cimg;
  p.n_tiles_n = 1;
  p.cchunks = 1;
  p.k_iters = 1;
  p.splits = 1;
  p.k_per_split = 1;
  make_out_map(L, out);
  L->tmA = L->tmOut;
  L->tmB = L->tmOut;
  L->grid = p.n_tiles_m < num_sms ? p.n_tiles_m : num_sms;
}

template <int BN, int MT>
static void launch_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  int smem = conv_smem_bytes_mt(BN, MT);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  conv_igemm_kernel<BN, MT><<<grid, CONV_THREADS, smem, st>>>(maps, grp);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

template <int BN, int MT, int KMAX, int OCC = 1>
static void launch_halo_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  const int smem = OCC == 2 ? occ_smem_bytes(BN * 128, MT, KMAX, OCC, 8) : halo_smem_bytes(BN, MT, KMAX);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<BN, MT, KMAX, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (OCC == 2)
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<BN, MT, KMAX, OCC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  conv_halo_kernel<BN, MT, KMAX, OCC><<<grid, CONV_THREADS, smem, st>>>(maps, grp);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

static void launch_swap_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  const int smem = halo_smem_bytes(128, 2, HALO_MAXK);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<128, 2, HALO_MAXK, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  conv_halo_kernel<128, 2, HALO_MAXK, 1, true><<<grid, CONV_THREADS, smem, st>>>(maps, grp);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

template <int BN, int MT, int KMAX, int OCC = 1>
static void launch_pair_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  const int smem = occ_smem_bytes(BN * 64, MT, KMAX, OCC, 12);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_pair_kernel<BN, MT, KMAX, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (OCC == 2)
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_pair_kernel<BN, MT, KMAX, OCC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(CONV_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FRCNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_pair_kernel<BN, MT, KMAX, OCC>, maps, grp));
}
template <int BN>
static void launch_pair_bres_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  const int smem = bres_smem_bytes(BN, grp.p[0].KH * grp.p[0].KW * grp.p[0].cchunks);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_pair_bres_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(CONV_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FRCNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_pair_bres_kernel<BN>, maps, grp));
}
static void launch_pair_key(int BN, int MT, int occ, const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  if (occ == 3) {   // weights resident (conv_pair_bres_kernel)
    switch (BN) {
      case 64: launch_pair_bres_cfg<64>(maps, grp, grid, st); break;
      case 128: launch_pair_bres_cfg<128>(maps, grp, grid, st); break;
      default: throw Error{FRCNN_E_INVALID, "conv (pair kernel, resident weights): BN = 64 or 128"};
    }
    return;
  }
  if (occ == 2) {
    switch (BN * 10 + MT) {
      case 641: launch_pair_cfg<64, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 642: launch_pair_cfg<64, 2, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 1281: launch_pair_cfg<128, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 1921: launch_pair_cfg<192, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 2561: launch_pair_cfg<256, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      default: throw Error{FRCNN_E_INVALID, "conv (pair kernel, 2 CTAs / SM): unsupported (BN, MT)"};
    }
    return;
  }
  switch (BN * 10 + MT) {
    case 641: launch_pair_cfg<64, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 642: launch_pair_cfg<64, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 1281: launch_pair_cfg<128, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 1282: launch_pair_cfg<128, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 1921: launch_pair_cfg<192, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 1922: launch_pair_cfg<192, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 2561: launch_pair_cfg<256, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 2562: launch_pair_cfg<256, 2, HALO_MAXK>(maps, grp, grid, st); break;
    default: throw Error{FRCNN_E_INVALID, "conv (pair kernel): unsupported (BN, MT)"};
  }
}

static void launch_halo_key(int BN, int MT, int occ, const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  if (occ == 2) {
    switch (BN * 10 + MT) {
      case 641: launch_halo_cfg<64, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 1281: launch_halo_cfg<128, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      case 1921: launch_halo_cfg<192, 1, HALO_MAXK, 2>(maps, grp, grid, st); break;
      default: throw Error{FRCNN_E_INVALID, "conv (halo kernel, 2 CTAs / SM): unsupported (BN, MT)"};
    }
    return;
  }
  switch (BN * 10 + MT) {
    case 641: launch_halo_cfg<64, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 642: launch_halo_cfg<64, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 1281: launch_halo_cfg<128, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 1282: launch_halo_cfg<128, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 1921: launch_halo_cfg<192, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 1922: launch_halo_cfg<192, 2, HALO_MAXK>(maps, grp, grid, st); break;
    case 2561: launch_halo_cfg<256, 1, HALO_MAXK>(maps, grp, grid, st); break;
    case 2562: launch_halo_cfg<256, 2, HALO_MAXK>(maps, grp, grid, st); break;
    default: throw Error{FRCNN_E_INVALID, "conv (halo kernel): unsupported (BN, MT)"};
  }
}

template <int BN, int T>
static void launch_wgrad_halo_cfg(const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  const int smem = wgh_smem_bytes(BN);
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_halo_kernel<BN, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  conv_wgrad_halo_kernel<BN, T><<<grid, CONV_THREADS, smem, st>>>(maps, grp);
  FRCNN_CUDA_TRY(cudaGetLastError());
}
static void launch_wgrad_halo_key(int BN, const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  switch (BN) {
    case 64: launch_wgrad_halo_cfg<64, 5>(maps, grp, grid, st); break;
    case 128: launch_wgrad_halo_cfg<128, 3>(maps, grp, grid, st); break;
    case 192: launch_wgrad_halo_cfg<192, 2>(maps, grp, grid, st); break;
    case 256: launch_wgrad_halo_cfg<256, 2>(maps, grp, grid, st); break;
    default: throw Error{FRCNN_E_INVALID, "wgrad (tap-group kernel): unsupported Cin tile"};
  }
}

static void launch_key(int BN, int MT, const ConvMaps& maps, const ConvGroup& grp, int grid, cudaStream_t st) {
  switch (BN * 10 + MT) {
    case 641: launch_cfg<64, 1>(maps, grp, grid, st); break;
    case 642: launch_cfg<64, 2>(maps, grp, grid, st); break;
    case 1281: launch_cfg<128, 1>(maps, grp, grid, st); break;
    case 1282: launch_cfg<128, 2>(maps, grp, grid, st); break;
    case 1921: launch_cfg<192, 1>(maps, grp, grid, st); break;
    case 2561: launch_cfg<256, 1>(maps, grp, grid, st); break;
    default: throw Error{FRCNN_E_INVALID, "conv: unsupported (BN, MT)"};
  }
}

void conv_launch(const ConvLaunch& L, cudaStream_t st) {
  FRCNN_REQUIRE(!L.p.f16 || (L.w_copies == 2 && !L.p.wgrad), FRCNN_E_STATE, "conv: fp16 operands need the two-copy weight pack");
  ConvGroup grp;
  grp.n = 1;
  grp.p[0] = L.p;
  grp.unit_end[0] = L.p.n_tiles_m * L.p.n_tiles_n * L.p.splits;
  for (int g = 1; g < MAX_GROUP; ++g) {
    grp.p[g] = L.p;
    grp.unit_end[g] = grp.unit_end[0];
  }
  if (L.first) {
    static DeviceOnce configured;
    if (first_use_on_device(configured)) {
      FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FIRST_SMEM));
    }
    FRCNN_REQUIRE(L.p.img != nullptr, FRCNN_E_STATE, "first-layer kernel: image pointer not set");
    if (L.p.first_tma && L.p.scale == 1.0f && (reinterpret_cast<uintptr_t>(L.p.img) & 15) == 0) {
      static DeviceOnce configured_tma;
      if (first_use_on_device(configured_tma)) {
        FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_first_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
      }
      // tensor map over the caller's fp32 NCHW frames (a pure host-side encode, redone per launch: the pointer changes)
      const ConvParams& p = L.p;
      CUtensorMap tmImg;
      cuuint64_t dims[4] = {(cuuint64_t)p.Win, (cuuint64_t)p.Hin, 3, (cuuint64_t)p.N};
      cuuint64_t strides[3] = {(cuuint64_t)p.Win * 4, (cuuint64_t)p.Hin * p.Win * 4, (cuuint64_t)3 * p.Hin * p.Win * 4};
      cuuint32_t box[4] = {FT_PATCH_W, FT_PATCH_H, 3, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = get_encode()(&tmImg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.img), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA, "cuTensorMapEncodeTiled(frame) failed, CUresult " + std::to_string((int)r));
      conv_first_tma_kernel<<<L.grid, FT_THREADS, FT_SMEM, st>>>(tmImg, grp, L.w_first);
      FRCNN_CUDA_TRY(cudaGetLastError());
      return;
    }
    conv_first_kernel<<<L.grid, FIRST_THREADS, FIRST_SMEM, st>>>(L.tmOut, grp, L.w_first);
    FRCNN_CUDA_TRY(cudaGetLastError());
    return;
  }
  ConvMaps maps;
  for (int g = 0; g < MAX_GROUP; ++g) {
    maps.a[g] = L.tmA;
    maps.b[g] = L.tmB;
    maps.o[g] = L.tmOut;
  }
  if (L.p.wgrad && L.p.halo) launch_wgrad_halo_key(L.BN, maps, grp, L.grid, st);
  else if (L.p.swap) launch_swap_cfg(maps, grp, L.grid, st);
  else if (L.p.pair) launch_pair_key(L.BN, L.p.MT, L.p.occ, maps, grp, L.grid, st);
  else if (L.p.halo) launch_halo_key(L.BN, L.p.MT, L.p.occ, maps, grp, L.grid, st);
  else launch_key(L.BN, L.p.MT, maps, grp, L.grid, st);
}

void conv_launch_group(const ConvLaunch* const* Ls, int n, int num_sms, cudaStream_t st) {
  FRCNN_REQUIRE(n >= 1 && n <= MAX_GROUP, FRCNN_E_INVALID, "conv group: 1..4 members");
  ConvGroup grp;
  ConvMaps maps;
  grp.n = n;
  int total = 0;
  for (int g = 0; g < MAX_GROUP; ++g) {
    const ConvLaunch& L = *Ls[g < n ? g : n - 1];
    FRCNN_REQUIRE(!L.first && !L.p.halo && L.p.MT == 1 && L.BN == Ls[0]->BN, FRCNN_E_INVALID, "conv group: members must share BN, MT = 1");
    FRCNN_REQUIRE(L.p.mode == EPI_F32_SLICES || L.p.mode == EPI_F32_REDUCE, FRCNN_E_INVALID, "conv group: fp32 epilogues only");
    grp.p[g] = L.p;
    maps.a[g] = L.tmA;
    maps.b[g] = L.tmB;
    maps.o[g] = L.tmOut;
    if (g < n) total += L.p.n_tiles_m * L.p.n_tiles_n * L.p.splits;
    grp.unit_end[g] = total;
  }
  launch_key(Ls[0]->BN, 1, maps, grp, total < num_sms ? total : num_sms, st);
}

// ---- fused anchor-network kernel (conv_head_kernel): plan + launch
// Fills P->grp / P->maps and the host-side unit schedule for `n_heads` anchor networks on N frames; the caller uploads
// units / cta_off, allocates the slices and counters (sizes returned) and stores the device pointers in P->sched.
void conv_head_plan(HeadPlan* P, const HeadDesc* heads, int n_heads, int N, int num_sms, std::vector<int4>* units_out,
                    std::vector<int>* cta_off_out, size_t slice_floats[MAX_GROUP], int counter_ints[MAX_GROUP]) {
  FRCNN_REQUIRE(n_heads >= 1 && n_heads <= MAX_GROUP && N >= 1 && N < 65536, FRCNN_E_INVALID, "anchor networks: 1..4 heads");
  P->grp = ConvGroup();
  P->grp.n = n_heads;
  P->flops = 0.0;
  P->fix = HeadFixArgs();
  struct U { int cost, head, img, tile, kh0, kh1, s, nsl; };
  std::vector<U> units;
  int tiles[MAX_GROUP] = {0, 0, 0, 0};
  for (int g = 0; g < n_heads; ++g) {
    const HeadDesc& h = heads[g];
    FRCNN_REQUIRE(h.Cin % 64 == 0 && h.K >= 1 && h.K <= HEAD_MAXK && h.Hin >= h.K && h.Win >= h.K, FRCNN_E_INVALID,
                  "anchor network: Cin % 64 == 0, kernel size <= 7, input at least as large as the kernel");
    ConvParams& p = P->grp.p[g];
    p = ConvParams();
    p.N = N; p.Hin = h.Hin; p.Win = h.Win; p.Cin = h.Cin; p.Cout = HEAD_CM; p.KH = h.K; p.KW = h.K;
    p.Hout = h.Hin - h.K + 1; p.Wout = h.Win - h.K + 1;
    p.cchunks = h.Cin / 64; p.k_iters = h.K * h.K * p.cchunks; p.mode = EPI_HEAD; p.scale = 1.f; p.MT = 1; p.splits = 1;
    // linear positions that carry a valid output: up to (Hout - 1) * Win + Wout - 1
    tiles[g] = ((p.Hout - 1) * h.Win + p.Wout + BLOCK_M - 1) / BLOCK_M;
    P->flops += 2.0 * N * p.Hout * p.Wout * (double)HEAD_CM * h.K * h.K * h.Cin;
    // the input map as a matrix [N * Hin * Win][Cin]: box {64 channels, 136 rows}
    cuuint64_t dims[2] = {(cuuint64_t)h.Cin, (cuuint64_t)N * h.Hin * h.Win};
    cuuint64_t strides[1] = {(cuuint64_t)h.Cin * 2};
    cuuint32_t box[2] = {BLOCK_K, HEADK_SLAB_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&P->maps.a[g], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(h.in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA, "cuTensorMapEncodeTiled(anchor network input) failed, CUresult " + std::to_string((int)r));
    make_tmap_weight(&P->maps.b[g], h.w_packed, HEAD_CM, h.K * h.K * h.Cin, HEAD_CM / 2, 2);   // half boxes: 128 filters per CTA of a pair
    P->maps.o[g] = P->maps.b[g];
  }
  for (int g = n_heads; g < MAX_GROUP; ++g) {
    P->grp.p[g] = P->grp.p[n_heads - 1];
    P->maps.a[g] = P->maps.a[n_heads - 1];
    P->maps.b[g] = P->maps.b[n_heads - 1];
    P->maps.o[g] = P->maps.o[n_heads - 1];
  }
  // units run on CTA pairs (two neighbouring tiles each).  A head's reduction is split by filter rows when one unsplit
  // tile pair would be longer than a pair's fair share of the launch: the smallest number of slices whose longest slice
  // stays within 1.25 fair shares (an uneven split -- 5 rows in 4 slices -- only adds fix-up work), else one per row
  const int n_pairs_max = std::max(1, num_sms / 2);
  long pair_cost = 0;
  // the split count of a head must not depend on the batch size (batched == per-image results, bit for bit): the fair
  // share is that of ONE frame on the whole machine
  for (int g = 0; g < n_heads; ++g) pair_cost += (long)((tiles[g] + 1) / 2) * P->grp.p[g].k_iters;
  const long fair = std::max<long>(16, (pair_cost + n_pairs_max - 1) / n_pairs_max);
  for (int g = 0; g < n_heads; ++g) {
    const ConvParams& p = P->grp.p[g];
    const int row_cost = p.KW * p.cchunks;
    int nsl = p.KH;
    for (int n = 1; n <= p.KH; ++n)
      if ((long)((p.KH + n - 1) / n) * row_cost * 4 <= fair * 5) { nsl = n; break; }
    // FRCNN_HEAD_TAIL=1 (measurement): the tail of EVERY head runs in head_fixup_kernel (a unit then ends with its slice store
    // instead of the ~15 us in-epilogue tail).  Measured slower with frames in flight (4 829 vs 4 975 img/s: the fix-up
    // kernel's 292 blocks cost 14.5 SM-us against 8 saved), so the default keeps the tail of unsplit units in their epilogue.
    static const int tail_kernel = getenv("FRCNN_HEAD_TAIL") ? atoi(getenv("FRCNN_HEAD_TAIL")) : 0;
    const bool sliced = nsl > 1 || tail_kernel;
    slice_floats[g] = sliced ? (size_t)N * tiles[g] * nsl * BLOCK_M * HEAD_CM : 0;
    counter_ints[g] = 0;
    P->sched.tiles[g] = tiles[g];
    if (sliced) {
      HeadFixArgs& fx = P->fix;
      fx.head[fx.n] = g;
      fx.nsl[fx.n] = nsl;
      fx.block_end[fx.n] = (fx.n ? fx.block_end[fx.n - 1] : 0) + N * tiles[g] * 4;
      ++fx.n;
    }
    for (int n = 0; n < N; ++n)
      for (int t = 0; t < (tiles[g] + 1) / 2; ++t)
        for (int s = 0; s < nsl; ++s) {
          const int kh0 = (int)((long)p.KH * s / nsl), kh1 = (int)((long)p.KH * (s + 1) / nsl);
          // every unit also pays for its epilogue (slice store, or bias + PReLU + 1 x 1 tail): ~8 reduction steps' worth
          units.push_back(U{(kh1 - kh0) * row_cost + 8, g | (sliced ? (1 << 24) : 0), n, t, kh0, kh1, s, nsl});
        }
  }
  for (int g = n_heads; g < MAX_GROUP; ++g) { slice_floats[g] = 0; counter_ints[g] = 0; P->sched.tiles[g] = 0; }
  // longest-processing-time schedule over the persistent pairs (deterministic: stable sort, lowest pair index on ties)
  std::stable_sort(units.begin(), units.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
  const int G = (int)std::min<size_t>((size_t)n_pairs_max, units.size());
  std::vector<long> load(G, 0);
  std::vector<std::vector<int>> mine(G);
  for (size_t i = 0; i < units.size(); ++i) {
    int best = 0;
    for (int b = 1; b < G; ++b)
      if (load[b] < load[best]) best = b;
    load[best] += units[i].cost;
    mine[best].push_back((int)i);
  }
  units_out->clear();
  cta_off_out->assign(G + 1, 0);
  for (int b = 0; b < G; ++b) {
    for (int i : mine[b]) {
      const U& u = units[i];
      units_out->push_back(make_int4(u.head | (u.s << 8) | (u.nsl << 16), u.img, u.tile, u.kh0 | (u.kh1 << 8)));
    }
    (*cta_off_out)[b + 1] = (int)units_out->size();
  }
  P->grid = 2 * G;
  P->n_units = (int)units.size();
}

void conv_launch_heads(const HeadPlan& P, cudaStream_t st) {
  static DeviceOnce configured;
  if (first_use_on_device(configured)) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADK_SMEM));
  }
  for (int g = 0; g < P.grp.n; ++g) {
    const ConvParams& p = P.grp.p[g];
    FRCNN_REQUIRE(p.w2 && p.b2 && p.bias && p.out, FRCNN_E_STATE, "anchor networks: tail parameters / outputs not set");
  }
  FRCNN_REQUIRE(P.sched.units && P.sched.cta_off, FRCNN_E_STATE, "anchor networks: schedule not uploaded");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P.grid, 1, 1);
  cfg.blockDim = dim3(CONV_THREADS, 1, 1);
  cfg.dynamicSmemBytes = HEADK_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FRCNN_CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_head_kernel, P.maps, P.grp, P.sched));
  if (P.fix.n > 0) {
    head_fixup_kernel<<<P.fix.block_end[P.fix.n - 1], 256, 0, st>>>(P.grp, P.sched, P.fix);
    FRCNN_CUDA_TRY(cudaGetLastError());
  }
}

void conv_launch_head_group(const ConvLaunch* const* Ls, int n, int num_sms, cudaStream_t st) {
  FRCNN_REQUIRE(n >= 1 && n <= MAX_GROUP, FRCNN_E_INVALID, "anchor head group: 1..4 members");
  ConvGroup grp;
  ConvMaps maps;
  grp.n = n;
  int total = 0;
  for (int g = 0; g < MAX_GROUP; ++g) {
    const ConvLaunch& L = *Ls[g < n ? g : n - 1];
    FRCNN_REQUIRE(L.p.mode == EPI_HEAD && L.p.halo && L.p.w2 && L.p.b2 && L.p.bias && L.p.out, FRCNN_E_INVALID,
                  "anchor head group: members must be prepared by conv_prepare_head with their tail parameters set");
    grp.p[g] = L.p;
    maps.a[g] = L.tmA;
    maps.b[g] = L.tmB;
    maps.o[g] = L.tmOut;
    if (g < n) total += L.p.n_tiles_m;
    grp.unit_end[g] = total;
  }
  launch_halo_cfg<HEAD_CM, 1, HEAD_MAXK>(maps, grp, total < num_sms ? total : num_sms, st);
}

}  // namespace frcnn
