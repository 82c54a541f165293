// Micro-benchmark: cycles per tcgen05.mma when issued the way conv_halo_kernel issues them -- per 128-pixel tile
// 9 taps x (CIN / 16) K-steps, every tap its own A start row inside a halo tile (row pitch 8 or 10 pixels) and its own
// B tile, accumulators rotating over 4 TMEM stages, one commit per tile -- against the constant-descriptor loop of
// mma_cost.cu.  No TMA, no epilogue: isolates the tensor pipe + operand reads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gan-codes_b200/csrc tools/ubench/mma_pattern.cu -o /tmp/mma_pattern
#include <cstdio>
#include <vector>

#include "sm100_ptx.cuh"
using namespace dfgan;

template <int CIN>
__global__ void __launch_bounds__(128, 1) mma_pattern_kernel(int N, int pitch, int tiles, int vary_a, int vary_b, int commit_each,
                                                             int rotate, int delay, int issuers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  constexpr uint32_t kRow = CIN * 2u;
  uint8_t* sB = smem;                                   // 9 taps x N rows x kRow
  uint8_t* sA = smem + 9 * 128 * 128;                   // halo tile: 18 x pitch rows
  for (int i = threadIdx.x; i < (9 * 128 * 128 + 18 * 10 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    mbar_init(&done, issuers);
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
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  constexpr uint32_t swz = CIN == 64 ? SWZ_128B : SWZ_64B;
  const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, pitch * kRow, swz);
  const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, 8 * kRow, swz);
  const uint64_t b_step16 = (static_cast<uint32_t>(N) * kRow) >> 4;
  // `issuers` warps alternate tiles (each commits its own); `delay` = cycles of unrelated work per tile in the issuing
  // thread (stands for the barrier waits / descriptor set-up of the real kernel's MMA warp)
  if (warp < issuers) {
    const bool leader = elect_one();
    long long t0 = clock64();
    if (leader) {
      for (int tile = warp; tile < tiles; tile += issuers) {
        if (delay) {
          const long long w0 = clock64();
          while (clock64() - w0 < delay) {
          }
        }
        const int acc = rotate ? (tile & 3) : 0;
        const uint32_t d = tm + acc * N;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t ad = adesc0 + (vary_a ? ((static_cast<uint32_t>((t / 3) * pitch + (t % 3)) * kRow) >> 4) : 0u);
          const uint64_t bd = bdesc0 + (vary_b ? static_cast<uint64_t>(t) * b_step16 : 0u);
#pragma unroll
          for (int k = 0; k < CIN / 16; ++k) umma_bf16(d, ad + 2u * k, bd + 2u * k, idesc, (t | k) != 0);
        }
        if (commit_each) umma_commit(&bar[acc]);
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
  if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(mma_pattern_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_pattern_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int tiles = 200;
  struct Cfg { int cin, N, pitch, va, vb, ce, rot, grid; const char* what; int delay = 0, issuers = 1; };
  Cfg cfgs[] = {
      {64, 64, 8, 0, 0, 0, 0, 1, "C64 N64 constant descriptors, 1 SM"},
      {64, 64, 8, 0, 0, 0, 0, 148, "C64 N64 constant descriptors, 148 SMs"},
      {64, 64, 8, 1, 0, 0, 0, 148, "C64 N64 tap-shifted A (pitch 8)"},
      {64, 64, 10, 1, 0, 0, 0, 148, "C64 N64 tap-shifted A (pitch 10)"},
      {64, 64, 10, 1, 1, 0, 0, 148, "C64 N64 tap-shifted A + per-tap B"},
      {64, 64, 10, 1, 1, 1, 0, 148, "C64 N64 ... + commit per tile"},
      {64, 64, 10, 1, 1, 1, 1, 148, "C64 N64 ... + rotating accumulators"},
      {32, 32, 10, 1, 1, 1, 1, 148, "C32 N32 full pattern"},
      {32, 32, 8, 0, 0, 0, 0, 148, "C32 N32 constant descriptors"},
      {64, 128, 10, 1, 1, 1, 1, 148, "C64 N128 full pattern"},
      {64, 64, 10, 1, 1, 1, 1, 148, "C64 N64 full + 300-cycle gap per tile", 300, 1},
      {64, 64, 10, 1, 1, 1, 1, 148, "C64 N64 full + 700-cycle gap per tile", 700, 1},
      {64, 64, 10, 1, 1, 1, 1, 148, "C64 N64 full + 700-cycle gap, 2 issuing warps", 700, 2},
      {32, 32, 10, 1, 1, 1, 1, 148, "C32 N32 full + 500-cycle gap per tile", 500, 1},
      {32, 32, 10, 1, 1, 1, 1, 148, "C32 N32 full + 500-cycle gap, 2 issuing warps", 500, 2},
      {32, 32, 10, 1, 1, 1, 1, 148, "C32 N32 full + 500-cycle gap, 3 issuing warps", 500, 3},
  };
  for (const Cfg& c : cfgs) {
    if (c.cin == 64)
      mma_pattern_kernel<64><<<c.grid, 128, 180 * 1024>>>(c.N, c.pitch, tiles, c.va, c.vb, c.ce, c.rot, c.delay, c.issuers, cyc);
    else
      mma_pattern_kernel<32><<<c.grid, 128, 180 * 1024>>>(c.N, c.pitch, tiles, c.va, c.vb, c.ce, c.rot, c.delay, c.issuers, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const int mmas = tiles * 9 * (c.cin / 16);
    printf("%-48s: %7.1f cycles/MMA, %7.0f cycles/tile (%s)\n", c.what, static_cast<double>(h) / mmas,
           static_cast<double>(h) / tiles, cudaGetErrorString(e));
  }
  return 0;
}
