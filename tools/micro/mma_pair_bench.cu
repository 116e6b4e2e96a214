// Microbenchmark (measurement only): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) from the leader's one
// thread, per N, operands = zeros resident in both CTAs' shared memory.  Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I faster-rcnn.torch_b200/csrc tools/micro/mma_pair_bench.cu -o /tmp/mma_pair_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include "ptx.cuh"

using namespace frcnn;

template <int N, int M>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(int iters, long long* out, int halo = 0, int fence = 0, int fill = 0, int rot = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                  // 16 KB A tile (this CTA's M / 2 rows); halo: a 24 KB (8+2) x (16+2)-pixel box
  uint8_t* smem_b = smem + 2 * 24576;      // 9 slots of N / 2 rows of 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (N / 2) * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = ptx::cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < (2 * 24576 + 9 * (N / 2) * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = fill ? (0x3C003C00u ^ ((uint32_t)i * 2654435761u & 0x03FF03FFu)) : 0u;   // fp16-ish bit patterns near 1.0
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(slot, 512);
    ptx::tmem_relinquish_2cta();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0 && leader) {
    const uint32_t idesc = ptx::make_idesc_bf16(M, N);
    const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a));
    const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint64_t a = da;
      if (halo) {   // the halo kernel's A operand: tap (kh, kw) = row offset into the box, 8-row groups 10 rows apart
        const int tap = it % 9, row0 = (tap / 3) * 10 + tap % 3;
        a = ptx::make_desc_k_sw128_sbo(ptx::smem_u32(smem_a) + row0 * 128, 1280u, 0u);
      }
      uint64_t b = db;
      if (rot) {
        b = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + (it % 9) * (N / 2) * 128));
        if (!halo) a = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + ((it / 9) & 1) * 24576));
      }
      if (fence) ptx::tc_fence_after();
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::mma_bf16_ss_2cta(tmem, a + 2 * j, b + 2 * j, idesc, 1u);
    }
    long long t1 = clock64();
    ptx::mma_commit_2cta(&bars[0], 1);
    ptx::mbar_wait(&bars[0], 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem, 512);
}

template <int N, int M>
void run(int halo = 0, int fence = 0, int grid = 2, int fill = 0, int rot = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 2 * 24576 + 9 * (N / 2) * 128 + 1024 + 256;
  cudaFuncSetAttribute(bench<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep) bench<N, M><<<grid, 128, smem>>>(iters, d, halo, fence, fill, rot);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("cta_group::2 rot=%d grid=%d fill=%d halo=%d fence=%d M=%d N=%3d: issue %.1f clk/MMA, issue+drain %.1f clk/MMA  %s\n", halo, fence, M, N,
         (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<256, 256>();
  run<128, 256>();
  run<64, 256>();
  run<256, 128>();
  run<128, 128>();
  run<128, 256>(1, 0);
  run<128, 256>(1, 1);
  run<256, 256>(1, 1);
  run<128, 256>(0, 1);
  run<128, 256>(1, 1, 148, 0);
  run<128, 256>(1, 1, 2, 1);
  run<128, 256>(1, 1, 148, 1);
  run<256, 256>(0, 0, 148, 1);
  run<128, 256>(0, 0, 148, 1);
  run<128, 256>(0, 0, 2, 1, 1);
  run<128, 256>(1, 0, 2, 1, 1);
  run<128, 256>(1, 1, 148, 1, 1);
  run<256, 256>(0, 0, 2, 1, 1);
  run<256, 256>(1, 1, 148, 1, 1);
  return 0;
}
