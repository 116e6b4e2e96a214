// Microbenchmark (measurement only): do tcgen05.mma streams issued by TWO threads of one CTA (different warps, different TMEM
// accumulators) overlap on the SM's tensor core?  One stream: ~130 clocks per M128 K16 step whatever N is.  Build on the box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I faster-rcnn.torch_b200/csrc tools/micro/mma_streams_bench.cu -o /tmp/mma_streams_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include "ptx.cuh"

using namespace frcnn;

template <int N>
__global__ void __launch_bounds__(256, 1) bench(int iters, int streams, long long* out, int fence = 0, int wait = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                 // 4 x 16 KB A tiles
  uint8_t* smem_b = smem + 4 * 16384;     // 4 x N * 128 B tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 4 * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  const int s = warp == 1 ? 0 : (warp == 3 ? 1 : (warp == 5 ? 2 : (warp == 7 ? 3 : -1)));
  if (s >= 0 && s < streams) {   // whole warp converged, one elected lane issues
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (wait) ptx::mbar_wait(&bars[4 + (it & 3)], 1);     // an already completed phase: returns at once (what a landed operand box looks like)
      if (fence) ptx::tc_fence_after();
      const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + ((it + s) & 3) * 16384));
      const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + ((it + s) & 3) * N * 128));
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::mma_bf16_ss_w(tmem + s * (512 / 4), da + 2 * j, db + 2 * j, idesc, 1u);
    }
    long long t1 = clock64();
    ptx::mma_commit_w(&bars[s]);
    ptx::mbar_wait(&bars[s], 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[2 * s] = t1 - t0; out[2 * s + 1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

template <int N>
void run(int streams, int fence = 0, int wait = 0) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  const int smem = 4 * 16384 + 4 * N * 128 + 1024 + 256;
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep) bench<N><<<1, 256, smem>>>(iters, streams, d, fence, wait);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8] = {0};
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("N=%3d streams=%d fence=%d wait=%d: per stream %.1f clk/MMA (issue+drain %.1f) -> aggregate %.1f clk/MMA  %s\n", N, streams, (double)h[0] / (iters * 4),
         (double)h[1] / (iters * 4), (double)h[1] / (iters * 4) / streams, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int st : {1, 2, 4}) {
    if (st <= 2) run<256>(st);   // 2 x 256 columns
    run<128>(st);
    if (st <= 4) run<64>(st);
  }
  run<128>(1, 1, 0);
  run<128>(1, 0, 1);
  run<128>(1, 1, 1);
  run<64>(1, 1, 1);
  run<256>(1, 1, 1);
  return 0;
}
