// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) for the descriptor patterns of the halo kernels:
// shifted start rows inside a swizzled halo tile, stride-byte-offset = halo pitch (not a multiple of 1024 B),
// SWIZZLE_64B rows.  Operands are zeros; only the timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan-codes_b200/csrc tools/ubench/mma_desc_cost.cu -o tools/ubench/mma_desc_cost.bin
#include <cstdio>
#include <vector>

#include "sm100_ptx.cuh"
using namespace dfgan;

struct Cfg {
  int N, swz128, row_bytes, pitch_rows, n_taps, k_steps, shifted, iters;
};

__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
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
  const uint32_t idesc = make_idesc_bf16(128, c.N, 0, 0);
  const uint32_t swz = c.swz128 ? SWZ_128B : SWZ_64B;
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + 64 * 1024;
  const uint32_t sbo_a = c.pitch_rows * c.row_bytes;
  const uint32_t sbo_b = 8 * c.row_bytes;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint64_t adesc0 = make_smem_desc(a_base, 16, sbo_a, swz);
    const uint64_t bdesc0 = make_smem_desc(b_base, 16, sbo_b, swz);
    long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      const uint32_t d = tm + (i & 3) * c.N;
#pragma unroll 1
      for (int t = 0; t < c.n_taps; ++t) {
        // tap (dy, dx) = (t / 3, t % 3): start row dy * pitch + dx
        const uint32_t row0 = c.shifted ? ((t / 3) * c.pitch_rows + (t % 3)) : 0;
        const uint64_t adesc = adesc0 + ((row0 * c.row_bytes) >> 4);
        const uint64_t bdesc = bdesc0 + (((t & 1) * c.N * c.row_bytes) >> 4);
        for (int kk = 0; kk < c.k_steps; ++kk) {
          if (leader) umma_bf16(d, adesc + 2u * kk, bdesc + 2u * kk, idesc, (t | kk) != 0);
        }
      }
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
    while (!mbar_try_wait(&bar, 0)) {
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Named { const char* name; Cfg c; };
  std::vector<Named> cases;
  for (int N : {32, 64, 128, 256}) {
    cases.push_back({"SW128 aligned, pitch 8, 1 desc        ", {N, 1, 128, 8, 1, 1, 0, 2000}});
    cases.push_back({"SW128 aligned, pitch 8, 9 taps x 4 k  ", {N, 1, 128, 8, 9, 4, 0, 60}});
    cases.push_back({"SW128 aligned, pitch 10, 9 taps x 4 k ", {N, 1, 128, 10, 9, 4, 0, 60}});
    cases.push_back({"SW128 SHIFTED, pitch 10, 9 taps x 4 k ", {N, 1, 128, 10, 9, 4, 1, 60}});
    cases.push_back({"SW128 SHIFTED, pitch 16, 9 taps x 4 k ", {N, 1, 128, 16, 9, 4, 1, 60}});
    cases.push_back({"SW64  aligned, pitch 8, 9 taps x 2 k  ", {N, 0, 64, 8, 9, 2, 0, 60}});
    cases.push_back({"SW64  SHIFTED, pitch 10, 9 taps x 2 k ", {N, 0, 64, 10, 9, 2, 1, 60}});
    cases.push_back({"SW64  SHIFTED, pitch 16, 9 taps x 2 k ", {N, 0, 64, 16, 9, 2, 1, 60}});
  }
  for (auto& nc : cases) {
    k<<<1, 128, 180 * 1024>>>(nc.c, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n_mma = static_cast<double>(nc.c.iters) * nc.c.n_taps * nc.c.k_steps;
    printf("N=%3d %s: %7.1f cycles/MMA (%s)\n", nc.c.N, nc.name, static_cast<double>(h) / n_mma, cudaGetErrorString(e));
  }
  return 0;
}
