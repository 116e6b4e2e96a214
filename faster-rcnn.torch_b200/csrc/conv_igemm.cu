// Implicit-GEMM convolution / GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
// Replaces nn.SpatialConvolution forward as pnet uses it (models/model_utilities.lua:8,31; stride 1, square or
// rectangular kernels, symmetric zero padding) and nn.Linear forward as cnet uses it (model_utilities.lua:82).
//
// Data layout in HBM
//   activations  NHWC bf16, C % 64 == 0          (a GEMM operand [rows][K] is the case H = 1, W = rows)
//   weights      [Cout][KH][KW][Cin] bf16        (K-major rows of K = KH*KW*Cin)
// GEMM view      M = N*Hout*Wout output pixels, N = Cout, K = KH*KW*Cin.
//
// One M tile is a BH x BW rectangle of 128 output pixels of one image, so that the A operand of filter tap
// (kh, kw) and channel chunk c is ONE 4-D TMA box {64 ch, BW, BH, 1} at (c, w0 + kw - padW, h0 + kh - padH, n):
// zero padding and image borders are the TMA unit's out-of-bounds zero fill, no im2col buffer exists.  The box
// lands in shared memory as 128 rows of 128 bytes with the 128-byte swizzle -- the canonical K-major UMMA
// operand layout -- and is consumed by four tcgen05.mma (M=128, N=BN, K=16) per 64-channel chunk.
//
// Kernel structure (persistent, warp-specialised, 256 threads, 1 CTA / SM):
//   warp 0   TMA producer   (one elected lane)        smem ring: full[s] / empty[s] mbarriers
//   warp 1   MMA issuer     (one elected lane)        TMEM accumulators double-buffered: tmem_full / tmem_empty
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld 32x32b -> bias + PReLU + scale -> bf16 NHWC   (or fp32 red.add for split-K)
#include <stdio.h>

#include "common.h"
#include "ptx.cuh"

namespace frcnn {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;                         // bf16 elements = 128 bytes = one swizzle row
static constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
static constexpr int SMEM_BUDGET = 200 * 1024;

__host__ __device__ constexpr int conv_stages(int BN) { return SMEM_BUDGET / (A_STAGE_BYTES + BN * BLOCK_K * 2); }
__host__ __device__ constexpr int tmem_cols(int BN) { return 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512); }

int conv_smem_bytes(int BN) { return conv_stages(BN) * (A_STAGE_BYTES + BN * BLOCK_K * 2) + 1024 + 256; }

struct TileCoord {
  int n_img, h0, w0, n0, k_begin, k_end, m0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile, int BN) {
  TileCoord t;
  int ks = tile % p.splits;
  int r = tile / p.splits;
  int nt = r % p.n_tiles_n;
  int mt = r / p.n_tiles_n;
  int tw = mt % p.tiles_w;
  int r2 = mt / p.tiles_w;
  int th = r2 % p.tiles_h;
  t.n_img = r2 / p.tiles_h;
  t.h0 = th * p.BH;
  t.w0 = tw * p.BW;
  t.n0 = nt * BN;
  t.k_begin = ks * p.k_per_split;
  t.k_end = min(p.k_iters, t.k_begin + p.k_per_split);
  t.m0 = t.w0;  // first GEMM row of the tile; only used with m_limit (GEMM use: BH == 1, N == 1, tiles_h == 1)
  return t;
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
    conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  constexpr int STAGES = conv_stages(BN);
  constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  constexpr uint32_t TX_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr uint32_t IDESC = ptx::make_idesc_bf16(BLOCK_M, BN);
  constexpr int TMEM_COLS = tmem_cols(BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles_m * p.n_tiles_n * p.splits;
  const int m_limit = p.m_limit ? *p.m_limit : 0x7fffffff;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_base_slot, TMEM_COLS);
    ptx::tmem_relinquish();
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileCoord t = decode_tile(p, tile, BN);
        if (t.m0 >= m_limit) continue;
        for (int k = t.k_begin; k < t.k_end; ++k) {
          int tap = k / p.cchunks;
          int cc = k - tap * p.cchunks;
          int kh = tap / p.KW;
          int kw = tap - kh * p.KW;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], TX_BYTES);
          ptx::tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], cc * BLOCK_K, t.w0 + kw - p.padW,
                           t.h0 + kh - p.padH, t.n_img);
          ptx::tma_load_2d(smem_b + stage * B_STAGE_BYTES, &tmB, &full_bar[stage], k * BLOCK_K, t.n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        TileCoord t = decode_tile(p, tile, BN);
        if (t.m0 >= m_limit) continue;
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int k = t.k_begin; k < t.k_end; ++k) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
          for (int j = 0; j < BLOCK_K / 16; ++j) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            ptx::mma_bf16_ss(d_tmem, da + 2 * j, db + 2 * j, IDESC, (k > t.k_begin || j > 0) ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (k == t.k_end - 1) ptx::mma_commit(&tmem_full[acc]);
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
    // ------------------------------------------------------------ epilogue (warps 4..7 <-> TMEM lanes 0..127)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dy = row >> p.bw_shift;
    const int dx = row & (p.BW - 1);
    const float slope = p.prelu ? __ldg(p.prelu) : 1.0f;
    const bool has_prelu = p.prelu != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      TileCoord t = decode_tile(p, tile, BN);
      if (t.m0 >= m_limit) continue;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int h = t.h0 + dy, w = t.w0 + dx;
      const bool valid = (h < p.Hout) && (w < p.Wout);
      const size_t pix = ((size_t)t.n_img * p.Hout + h) * p.Wout + w;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c0, v);
        ptx::tmem_ld_wait();
        const int cbase = t.n0 + c0;
        if (valid && cbase < p.Cout) {
          if (p.mode == EPI_BF16_NHWC) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x0 = __uint_as_float(v[2 * j]), x1 = __uint_as_float(v[2 * j + 1]);
              if (p.bias) {
                x0 += __ldg(p.bias + cbase + 2 * j);
                x1 += __ldg(p.bias + cbase + 2 * j + 1);
              }
              if (has_prelu) {
                x0 = x0 > 0.f ? x0 : x0 * slope;
                x1 = x1 > 0.f ? x1 : x1 * slope;
              }
              x0 *= p.scale;
              x1 *= p.scale;
              o[j] = ptx::pack_bf16x2(x0, x1);
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.Cout + cbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          } else {
            float* dst = p.out_f32 + pix * p.Cout + cbase;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
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
                    std::to_string(C) + "x" + std::to_string(W) + "x" + std::to_string(H) + "x" + std::to_string(N));
}

void make_tmap_weight(CUtensorMap* m, const bf16* base, int Cout, int K, int BN) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FRCNN_REQUIRE(r == CUDA_SUCCESS, FRCNN_E_CUDA,
                "cuTensorMapEncodeTiled(weight) failed, CUresult " + std::to_string((int)r));
}

// Pick the BW x BH (= 128) rectangle that wastes the fewest padded pixels.
void conv_choose_tile(int Hout, int Wout, int* BW, int* BH) {
  long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    int bh = 128 / bw;
    long padded = (long)((Wout + bw - 1) / bw) * bw * (long)((Hout + bh - 1) / bh) * bh;
    if (best < 0 || padded < best) {
      best = padded;
      *BW = bw;
      *BH = bh;
    }
  }
}

static int choose_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 192 == 0) return 192;
  if (Cout % 128 == 0) return 128;
  if (Cout <= 64) return 64;
  return 128;
}

void conv_prepare(ConvLaunch* L, const bf16* in, const bf16* w_packed, int N, int Hin, int Win, int Cin, int Cout,
                  int KH, int KW, int padH, int padW, int mode, int num_sms, int force_splits, int force_bn) {
  FRCNN_REQUIRE(Cin % 64 == 0, FRCNN_E_INVALID, "conv: Cin must be a multiple of 64");
  FRCNN_REQUIRE(Cout % 32 == 0, FRCNN_E_INVALID, "conv: Cout must be a multiple of 32");
  ConvParams& p = L->p;
  p = ConvParams();
  p.N = N; p.Hin = Hin; p.Win = Win; p.Cin = Cin;
  p.Hout = Hin + 2 * padH - KH + 1;
  p.Wout = Win + 2 * padW - KW + 1;
  FRCNN_REQUIRE(p.Hout > 0 && p.Wout > 0, FRCNN_E_INVALID, "conv: input smaller than the kernel");
  p.Cout = Cout; p.KH = KH; p.KW = KW; p.padH = padH; p.padW = padW;
  conv_choose_tile(p.Hout, p.Wout, &p.BW, &p.BH);
  p.bw_shift = 0;
  while ((1 << p.bw_shift) < p.BW) ++p.bw_shift;
  p.tiles_w = (p.Wout + p.BW - 1) / p.BW;
  p.tiles_h = (p.Hout + p.BH - 1) / p.BH;
  p.n_tiles_m = N * p.tiles_h * p.tiles_w;
  int BN = force_bn > 0 ? force_bn : choose_bn(Cout);
  L->BN = BN;
  p.n_tiles_n = (Cout + BN - 1) / BN;
  p.cchunks = Cin / 64;
  p.k_iters = KH * KW * p.cchunks;
  int splits = 1;
  if (mode == EPI_F32_ATOMIC) {
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
  p.mode = mode;
  p.scale = 1.0f;
  make_tmap_act(&L->tmA, in, N, Hin, Win, Cin, p.BW, p.BH);
  make_tmap_weight(&L->tmB, w_packed, Cout, KH * KW * Cin, BN);
  int total = p.n_tiles_m * p.n_tiles_n * p.splits;
  L->grid = total < num_sms ? total : num_sms;
}

template <int BN>
static void launch_bn(const ConvLaunch& L, cudaStream_t st) {
  static bool configured = false;
  int smem = conv_smem_bytes(BN);
  if (!configured) {
    FRCNN_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  conv_igemm_kernel<BN><<<L.grid, 256, smem, st>>>(L.tmA, L.tmB, L.p);
  FRCNN_CUDA_TRY(cudaGetLastError());
}

void conv_launch(const ConvLaunch& L, cudaStream_t st) {
  switch (L.BN) {
    case 64: launch_bn<64>(L, st); break;
    case 128: launch_bn<128>(L, st); break;
    case 192: launch_bn<192>(L, st); break;
    case 256: launch_bn<256>(L, st); break;
    default: throw Error{FRCNN_E_INVALID, "conv: unsupported BN"};
  }
}

}  // namespace frcnn
