// Micro-benchmark: how much of a gap in the issuing thread the tensor pipe's queue hides.  One thread issues units of
// `per_unit` tcgen05.mma (M128 x N x K16) separated by `gap` cycles of unrelated work; reports cycles per unit against
// the gap-free figure.  (conv_halo2 / conv_umma / wgrad_umma issue 4-8 MMAs between two barrier waits.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan-codes_b200/csrc tools/ubench/mma_gap.cu -o /tmp/mma_gap
#include <cstdio>

#include "sm100_ptx.cuh"
using namespace dfgan;

__global__ void __launch_bounds__(128, 1) mma_gap_kernel(int N, int per_unit, int units, int gap, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B);
  const uint64_t bdesc = make_smem_desc(smem_u32(smem + 128 * 128), 16, 1024, SWZ_128B);
  if (warp == 0) {
    const bool leader = elect_one();
    long long t0 = clock64();
    if (leader) {
      for (int u = 0; u < units; ++u) {
        if (gap) {
          const long long w0 = clock64();
          while (clock64() - w0 < gap) {
          }
        }
        for (int k = 0; k < per_unit; ++k) umma_bf16(tm, adesc + 2u * (k & 3), bdesc + 2u * (k & 3), idesc, (u | k) != 0);
      }
      umma_commit(&done);
      while (!mbar_try_wait(&done, 0)) {
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(mma_gap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int units = 400;
  for (int N : {64, 128, 256})
    for (int per_unit : {4, 8})
      for (int gap : {0, 50, 100, 150, 200, 300}) {
        mma_gap_kernel<<<148, 128, 60 * 1024>>>(N, per_unit, units, gap, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("N=%3d %d MMAs per unit, gap %3d cycles: %7.1f cycles/unit (%s)\n", N, per_unit, gap,
               static_cast<double>(h) / units, cudaGetErrorString(e));
      }
  return 0;
}
