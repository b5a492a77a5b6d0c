// Micro-benchmark: issue rate of tcgen05.mma (M=128, K=16, fp16) for several N, with and without a
// tcgen05.commit after every 4 MMAs.  One CTA per SM, one issuing thread.  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../metro_pose3d_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace metro;

template <int N, bool kCommit>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar + 1, 1); ptx::fence_mbar_init(); }
  if (warp == 2) { ptx::tmem_alloc(&s_tmem, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, N);
    const uint32_t sa = ptx::smem_u32(smem);
    const uint64_t da = ptx::make_sw128_kmajor_desc(sa), db = ptx::make_sw128_kmajor_desc(sa + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ptx::umma_f16(tmem, da + 2 * kk, db + 2 * kk, idesc, 1);
      if (kCommit) ptx::umma_commit(bar);
    }
    const long long t1 = clock64();
    ptx::umma_commit(bar + 1);
    ptx::mbar_wait(bar + 1, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int N, bool kCommit>
void run(int grid) {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  const int iters = 2000;
  const int sm = 16384 + N * 128 + 1024;
  cudaFuncSetAttribute(k<N, kCommit>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  for (int rep = 0; rep < 2; ++rep) {
    k<N, kCommit><<<grid, 128, sm>>>(d, iters);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d commit=%d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (err %s)\n", N, int(kCommit), grid,
         double(h[0]) / (4.0 * iters), double(h[1]) / (4.0 * iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, false>(grid); run<64, true>(grid);
    run<128, false>(grid); run<128, true>(grid);
    run<256, false>(grid); run<256, true>(grid);
  }
  return 0;
}
