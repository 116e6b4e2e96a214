// Microbenchmark (measurement only): what does ONE thread issuing tcgen05.mma back to back sustain, per instruction shape?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I faster-rcnn.torch_b200/csrc tools/micro/mma_issue_bench.cu -o gpurun_out/mma_issue_bench
// Variants: N in {64, 128, 256}; accumulators = 1 (every MMA read-modify-writes the same TMEM tile), 2 or 4 (round robin);
// spinners = number of extra warps of the CTA spinning on an mbarrier (as the epilogue warps of the conv kernels do while
// the MMA warp works); fence = tcgen05.fence::after_thread_sync before every group of 4 MMAs.  Operands: zeros in shared
// memory (K-major SWIZZLE_128B tiles).  Reports clocks per MMA (clock64 around the issue loop + a final commit / wait).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include "ptx.cuh"

using namespace frcnn;

template <int N>
__global__ void __launch_bounds__(384, 1) bench(int iters, int n_acc, int spinners, int fence, int a_step, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                 // 4 x 16 KB A tiles
  uint8_t* smem_b = smem + 4 * 16384;     // 4 x N * 128 B tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 4 * N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);   // never completed: what the spinners wait on
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
  __shared__ volatile int done;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (fence) ptx::tc_fence_after();
      const uint64_t da = ptx::make_desc_k_sw128(ptx::smem_u32(smem_a + (a_step ? (it & 3) * 16384 : 0)));
      const uint64_t db = ptx::make_desc_k_sw128(ptx::smem_u32(smem_b + (a_step ? (it & 3) * N * 128 : 0)));
#pragma unroll
      for (int j = 0; j < 4; ++j) ptx::mma_bf16_ss(tmem + ((it * 4 + j) % n_acc) * N, da + 2 * j, db + 2 * j, idesc, 1u);
    }
    long long t1 = clock64();
    ptx::mma_commit(&bars[0]);
    ptx::mbar_wait(&bars[0], 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    done = 1;
  } else if (warp >= 4 && warp < 4 + spinners) {
    // spin like an epilogue warp waiting for its accumulator: mbarrier try_wait loop until the issuer is done
    while (!done) {
      ptx::mbar_try_wait(&bars[1], 0);
    }
  }
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem, 512);
}

template <int N>
void run(int n_acc, int spinners, int fence, int a_step) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 4 * 16384 + 4 * N * 128 + 1024 + 256;
  cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep) bench<N><<<1, 384, smem>>>(iters, n_acc, spinners, fence, a_step, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d acc=%d spinners=%d fence=%d rotate_operands=%d: issue %.1f clk/MMA, issue+drain %.1f clk/MMA (floor %d)  %s\n", N, n_acc, spinners, fence,
         a_step, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int spin : {0, 8}) {
    run<256>(1, spin, 0, 0);
    run<256>(2, spin, 0, 0);
    run<128>(1, spin, 0, 0);
    run<128>(2, spin, 0, 0);
    run<128>(4, spin, 0, 0);
    run<64>(1, spin, 0, 0);
    run<64>(4, spin, 0, 0);
  }
  run<128>(1, 8, 1, 0);
  run<128>(1, 8, 1, 1);
  run<256>(1, 8, 1, 1);
  run<256>(2, 8, 1, 1);
  return 0;
}
