// Micro-benchmarks of the tcgen05 tensor pipe on one SM / one CTA pair (M = 128 per CTA, K = 16, fp16):
//   * issue/completion rate of tcgen05.mma for several N, one CTA (cta_group::1) and CTA pairs (cta_group::2)
//   * the same while a second thread streams bulk copies (cp.async.bulk, the TMA engine) into other
//     shared-memory buffers, to see how much the shared-memory write traffic of the operand pipeline costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../metro_pose3d_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace metro;

constexpr int kRing = 4, kCopyBytes = 16384;

template <int N, int kPair, bool kCopy, bool kNoSw = false>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters, const unsigned char *src, size_t src_bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[2 + kRing];
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kB = (N / kPair) * 128;
  unsigned char *ring = smem + 16384 + ((kB + 1023) & ~1023);
  for (int i = threadIdx.x; i < (16384 + kB) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 + kRing; ++i) ptx::mbar_init(bar + i, 1);
    ptx::fence_mbar_init();
    s_stop = 0;
  }
  if (warp == 2) {
    if (kPair == 2) { ptx::tmem_alloc_pair(&s_tmem, 512); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(&s_tmem, 512); ptx::tmem_relinquish(); }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  if (kPair == 2) ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  const bool leader = kPair == 1 || ptx::cluster_ctarank() == 0;
  if (warp == 1 && lane == 0 && leader) {
    constexpr uint32_t idesc = ptx::make_idesc_f16(128 * kPair, N);
    const uint32_t sa = ptx::smem_u32(smem);
    uint64_t da = ptx::make_sw128_kmajor_desc(sa), db = ptx::make_sw128_kmajor_desc(sa + 16384);
    if (kNoSw) {   // un-swizzled K-major operands: A rows 16 B apart with overlapping K chunks (LBO 16, SBO 128), B dense
      da = uint64_t((sa >> 4) & 0x3FFF) | (uint64_t(16 >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
      db = uint64_t(((sa + 16384) >> 4) & 0x3FFF) | (uint64_t(1024 >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (uint64_t(1) << 46);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kPair == 2) ptx::umma_f16_pair(tmem, da + 2 * kk, db + 2 * kk, idesc, 1);
        else ptx::umma_f16(tmem, da + 2 * kk, db + 2 * kk, idesc, 1);
      }
      if (kPair == 2) ptx::umma_commit_pair(bar, 1); else ptx::umma_commit(bar);
    }
    if (kPair == 2) ptx::umma_commit_pair(bar + 1, 1); else ptx::umma_commit(bar + 1);
    ptx::mbar_wait(bar + 1, 0);
    const long long t2 = clock64();
    s_stop = 1;
    if (blockIdx.x == 0) out[0] = t2 - t0;
  }
  if (kCopy && warp == 0 && lane == 0) {
    // stream 16 KB bulk copies through a 4-slot ring until the MMA thread is done
    size_t off = size_t(blockIdx.x) * 65536;
    uint32_t phase = 0;
    long long n = 0;
    for (int i = 0; i < kRing; ++i) {
      ptx::mbar_arrive_expect_tx(bar + 2 + i, kCopyBytes);
      ptx::bulk_load_1d(ring + i * kCopyBytes, src + (off % src_bytes), kCopyBytes, bar + 2 + i);
      off += kCopyBytes;
    }
    const long long t0 = clock64();
    while (!(kPair == 2 && !leader ? n >= iters : s_stop)) {
      for (int i = 0; i < kRing; ++i) {
        ptx::mbar_wait(bar + 2 + i, phase);
        ptx::mbar_arrive_expect_tx(bar + 2 + i, kCopyBytes);
        ptx::bulk_load_1d(ring + i * kCopyBytes, src + (off % src_bytes), kCopyBytes, bar + 2 + i);
        off += kCopyBytes;
        ++n;
      }
      phase ^= 1;
    }
    for (int i = 0; i < kRing; ++i) ptx::mbar_wait(bar + 2 + i, phase);
    if (blockIdx.x == 0) { out[1] = n * kCopyBytes; out[2] = clock64() - t0; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kPair == 2) ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (kPair == 2) ptx::tmem_dealloc_pair(tmem, 512); else ptx::tmem_dealloc(tmem, 512);
  }
}

template <int N, int kPair, bool kCopy, bool kNoSw = false>
void run(int grid, const unsigned char *src, size_t src_bytes) {
  long long *d, h[3] = {0, 0, 0};
  cudaMalloc(&d, 24);
  cudaMemset(d, 0, 24);
  const int iters = 2000;
  const int sm = 16384 + N * 128 + 2048 + kRing * kCopyBytes;
  cudaFuncSetAttribute(k<N, kPair, kCopy, kNoSw>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = sm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kPair; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k<N, kPair, kCopy, kNoSw>, d, iters, src, src_bytes);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("N=%3d ctas/tile=%d copy=%d noswizzle=%d grid=%3d: %.1f cyc/MMA", N, kPair, int(kCopy), int(kNoSw), grid, double(h[0]) / (4.0 * iters));
  if (kCopy) printf("   copy stream %.1f B/cyc/SM", double(h[1]) / double(h[2] ? h[2] : 1));
  printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  unsigned char *src;
  const size_t src_bytes = size_t(64) << 20;      // 64 MB: L2 resident after the first pass
  cudaMalloc(&src, src_bytes + 65536 * 160);
  cudaMemset(src, 0, src_bytes + 65536 * 160);
  run<64, 1, false, true>(148, src, src_bytes);
  run<128, 1, false, true>(148, src, src_bytes);
  run<256, 1, false, true>(148, src, src_bytes);
  for (int grid : {148}) {
    run<64, 1, false>(grid, src, src_bytes); run<64, 1, true>(grid, src, src_bytes);
    run<128, 1, false>(grid, src, src_bytes); run<128, 1, true>(grid, src, src_bytes);
    run<256, 1, false>(grid, src, src_bytes); run<256, 1, true>(grid, src, src_bytes);
    run<64, 2, false>(grid, src, src_bytes); run<64, 2, true>(grid, src, src_bytes);
    run<128, 2, false>(grid, src, src_bytes); run<128, 2, true>(grid, src, src_bytes);
    run<256, 2, false>(grid, src, src_bytes); run<256, 2, true>(grid, src, src_bytes);
  }
  return 0;
}
