// Micro-benchmark: cost of TMA tensor loads (cp.async.bulk.tensor.2d, 128-byte rows, 128B swizzle) per SM:
// one thread issues `nbox` boxes of `rows` x 128 B back to back into a ring of shared memory and waits for
// all of them; reports issue cycles per box and bytes/cycle/SM with every SM doing the same (L2-resident source).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../metro_pose3d_b200/csrc tma_rate.cu -o tma_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace metro;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, long long *out, int rows, int nbox, int inflight,
                                            int total_rows, int issuers) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[128];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 128; ++i) ptx::mbar_init(bars + i, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    uint64_t *bar = bars + 32 * w;
    unsigned char *base = smem + size_t(w) * inflight * rows * 128;
    const int box_bytes = rows * 128;
    long long t_issue = 0;
    const long long t0 = clock64();
    int row = ((blockIdx.x * 4 + w) * 977) % (total_rows - rows);
    for (int i = 0; i < nbox; ++i) {
      const int slot = i % inflight;
      if (i >= inflight) ptx::mbar_wait(bar + slot, ((i / inflight) - 1) & 1);
      const long long a = clock64();
      ptx::mbar_arrive_expect_tx(bar + slot, box_bytes);
      ptx::tma_load_2d(base + slot * box_bytes, &map, bar + slot, 0, row);
      t_issue += clock64() - a;
      row += rows;
      if (row + rows > total_rows) row = 0;
    }
    for (int i = nbox - inflight; i < nbox; ++i) ptx::mbar_wait(bar + (i % inflight), (i / inflight) & 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && w == 0) { out[0] = t1 - t0; out[1] = t_issue; }
  }
}

int main() {
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = reinterpret_cast<Enc>(fp);
  const int total_rows = 1 << 19;                 // 64 MB of 128-byte rows
  void *src;
  cudaMalloc(&src, size_t(total_rows) * 128);
  cudaMemset(src, 0, size_t(total_rows) * 128);
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rows : {32, 64, 128, 256}) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {64, cuuint64_t(total_rows)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, cuuint32_t(rows)};
    const cuuint32_t es[2] = {1, 1};
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int issuers : {1, 2, 4})
    for (int inflight : {4}) {
      if (issuers * inflight * rows * 128 > 190 * 1024) continue;
      const int nbox = 512;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 128, issuers * inflight * rows * 128 + 1024>>>(map, d, rows, nbox, inflight, total_rows, issuers);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("box %3d rows (%5d B), %d issuing threads x %2d in flight: %.0f cyc/box per thread, issue %.0f cyc/box, %.1f B/cyc/SM  (%s)\n", rows, rows * 128,
             issuers, inflight, double(h[0]) / nbox, double(h[1]) / nbox, double(issuers) * double(nbox) * rows * 128 / double(h[0]),
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
