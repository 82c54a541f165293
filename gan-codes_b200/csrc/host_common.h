// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/dfgan_b200.h"

namespace dfgan {

extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional per-launch timing of the two GEMM engines (bench.py's roofline leg): when enabled, every
// dfgan_conv_run / dfgan_wgrad_run launch is bracketed by a CUDA event pair on its own stream.
struct ProfileScope {
  // label (optional): kernel family + shape, written per launch by dfgan_profile_read when DFGAN_PROFILE_DUMP=<file>
  ProfileScope(int engine, double flops, double bytes, cudaStream_t st, const char* label = nullptr);
  ~ProfileScope();
  int slot;
  cudaStream_t st;
};

// cuTensorMapEncodeTiled resolved through the runtime so the library carries no link-time
// dependency on libcuda (it must load on a CPU-only box for the export test).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

int num_sms();

// bf16 NHWC view -> 4-D tiled tensor map with box {bc, bw, bh, bn}; OOB elements read as zero.
int make_map_nhwc(CUtensorMap* m, const dfgan_view4& v, int bc, int bw, int bh, int bn);
// bf16 NHWC view -> 5-D map {8 ch, W, H, C/8, N} with box {8, bw, bh, C/8, 1}, no swizzle: one load delivers
// a halo tile as [channel slice][h][w][8 ch], the canonical no-swizzle K-major UMMA operand layout.
int make_map_halo(CUtensorMap* m, const dfgan_view4& v, int bw, int bh);
// bf16 row-major matrix [rows][cols] -> 2-D tiled tensor map with box {bc, brows}.
int make_map_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int bc, int brows);

// Launch with the programmatic-dependent-launch attribute (the kernel must call pdl_wait() before it touches
// global memory).  DFGAN_NO_PDL=1 falls back to plain stream-ordered launches.
bool pdl_enabled();
// cluster > 1 additionally launches thread-block clusters of that many CTAs along x (grid must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster), at[n].val.clusterDim.y = 1, at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at, cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

#define DFGAN_CUDA_CHECK_LAUNCH()                  \
  do {                                             \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return static_cast<int>(e__); \
  } while (0)

}  // namespace dfgan
