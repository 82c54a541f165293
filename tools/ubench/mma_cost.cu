// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32, K = 16) as a function of M and N, issued
// back-to-back into one accumulator, and the TMEM lane layout of an M = 64 accumulator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan-codes_b200/csrc tools/ubench/mma_cost.cu -o gpurun_out/mma_cost
#include <cstdio>
#include <vector>

#include "sm100_ptx.cuh"
using namespace dfgan;

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(int M, int N, int iters, long long* cycles, float* dump) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // A: 128 rows x 64 bf16 (SW128 K-major atom layout, value = row index), B: 256 rows x 64 bf16 (value 1 in k=0)
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 128 * 128);
  for (int i = threadIdx.x; i < 128 * 64; i += 128) A[i] = __float2bfloat16(0.f);
  for (int i = threadIdx.x; i < 256 * 64; i += 128) B[i] = __float2bfloat16(0.f);
  __syncthreads();
  // element (row r, k = 0) lives at byte r*128 + ((0 ^ (r & 7)) * 16) in the 128B-swizzled K-major layout
  for (int r = threadIdx.x; r < 128; r += 128) A[(r * 128 + ((r & 7) * 16)) / 2] = __float2bfloat16(static_cast<float>(r));
  for (int r = threadIdx.x; r < 256; r += 128) B[(r * 128 + ((r & 7) * 16)) / 2] = __float2bfloat16(1.f);
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t idesc = make_idesc_bf16(M, N, 0, 0);
  const uint64_t adesc = make_smem_desc(smem_u32(A), 16, 1024, SWZ_128B);
  const uint64_t bdesc = make_smem_desc(smem_u32(B), 16, 1024, SWZ_128B);
  if (warp == 0) {
    const bool leader = elect_one();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (leader) umma_bf16(tm, adesc, bdesc, idesc, i != 0);
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    while (!mbar_try_wait(&bar, 0)) {
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // dump column 0 of every lane this warp may read: D[row][0] = iters * row  ->  lane -> row map
  float v[16];
  tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16), v);
  dump[threadIdx.x] = v[0] / static_cast<float>(iters);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
  long long* cyc;
  float* dump;
  cudaMalloc(&cyc, 8);
  cudaMalloc(&dump, 128 * 4);
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int M : {128, 64})
    for (int N : {16, 32, 64, 128, 256}) {
      cudaMemset(dump, 0, 512);
      mma_cost_kernel<<<1, 128, 60 * 1024>>>(M, N, iters, cyc, dump);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0;
      std::vector<float> d(128);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(d.data(), dump, 512, cudaMemcpyDeviceToHost);
      printf("M=%3d N=%3d: %7.1f cycles/MMA (%s)\n", M, N, static_cast<double>(h) / iters, cudaGetErrorString(e));
      if (N == 256) {
        printf("  lane -> row (column 0): ");
        for (int l = 0; l < 128; l += 1) printf("%g%s", d[l], (l % 32 == 31) ? " | " : " ");
        printf("\n");
      }
    }
  return 0;
}
