// Discriminator stem nn.Conv2d(3, 32, 3, 1, 1) (discriminator/model.py:13) as ONE kernel: the im2col tile of 128
// consecutive pixels (K = 27 padded to 32, bf16) is built in shared memory straight from the fp32 NCHW image in the
// canonical K-major 64-byte-swizzled UMMA layout, multiplied by the packed [32 x 32] weight matrix with two
// tcgen05.mma (M128 x N32 x K16, fp32 accumulation in TMEM) and finished by the thread that owns the pixel
// (tcgen05.ld, bias, bf16, two 256-bit stores).  Replaces the round-1 pair im2col3 (12 B in + 64 B out per pixel)
// + 1x1 GEMM on the 64-byte/pixel im2col tensor (64 B in + 64 B out): the fused form moves 12 + 64 bytes per pixel.
// The im2col tensor itself is only written on request (`xcol`, for the stem's weight-gradient GEMM).
// Arithmetic is identical to the pair it replaces: operands rounded to bf16 (RNE), fp32 accumulation, K order
// ci*9 + r*3 + s.  HBM-bound: no pipelining inside a CTA, several small CTAs per SM overlap each other's phases.
#include "host_common.h"
#include "sm100_ptx.cuh"

namespace dfgan {

constexpr int kStemThreads = 128;

__global__ void __launch_bounds__(kStemThreads, 6)
    stem_fwd_kernel(const float* __restrict__ img, const float* __restrict__ rowscale,
                    const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ xcol, int H, int W, long long P) {
  // A: 128 rows x 64 B (8 KB), B: 32 rows x 64 B (2 KB); 64-byte swizzle = XOR of address bits [4,6) with [7,9)
  __shared__ uint8_t raw[128 * 64 + 32 * 64 + 1024];
  __shared__ float sbias[32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // the swizzle is a function of absolute shared-memory address bits: align the operand tiles to 1024 bytes
  uint8_t* sA = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sB = sA + 128 * 64;
  const int t = threadIdx.x, warp = t >> 5;
  if (t == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 32);
    tmem_relinquish();
  }
  {
    // weight matrix: thread t moves the 16-byte chunk (row t / 4, chunk t % 4) to its swizzled position
    const int r = t >> 2, c = t & 3;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(w) + t);
    *reinterpret_cast<uint4*>(sB + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = v;
  }
  if (t < 32) sbias[t] = bias ? __ldg(bias + t) : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
  const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 512, SWZ_64B);
  const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 512, SWZ_64B);
  const long long HW = static_cast<long long>(H) * W;
  const long long n_tiles = (P + 127) / 128;
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long p = tile * 128 + t;
    const bool valid = p < P;
    uint32_t u[16];
    if (valid) {
      const int n = static_cast<int>(p / HW);
      const int q = static_cast<int>(p - n * HW);
      const int h = q / W, x = q - h * W;
      const float sc = rowscale ? __ldg(rowscale + n) : 1.f;
      const float* base = img + static_cast<long long>(n) * 3 * HW;
      float v[32];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            // unconditional loads from clamped coordinates (all 27 in flight), zeroed where the tap is padding
            const int hh = h + r - 1, ww = x + s - 1;
            const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
            const int hc = min(max(hh, 0), H - 1), wc = min(max(ww, 0), W - 1);
            const float a = __ldg(base + (static_cast<long long>(ci) * H + hc) * W + wc);
            v[ci * 9 + r * 3 + s] = in ? sc * a : 0.f;
          }
#pragma unroll
      for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) u[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) u[i] = 0u;
    }
    {
      const int sw = (t >> 1) & 3;
      uint8_t* row = sA + t * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
    }
    if (xcol && valid) {
      __nv_bfloat16* dst = xcol + p * 32;
      stg256(dst, make_uint4(u[0], u[1], u[2], u[3]), make_uint4(u[4], u[5], u[6], u[7]));
      stg256(dst + 16, make_uint4(u[8], u[9], u[10], u[11]), make_uint4(u[12], u[13], u[14], u[15]));
    }
    fence_proxy_async();          // generic-proxy writes of sA -> visible to the tensor core (async proxy)
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      umma_bf16(tmem, adesc, bdesc, idesc, 0);
      umma_bf16(tmem, adesc + 2u, bdesc + 2u, idesc, 1);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    float lo[16], hi[16];
    tmem_ld16(taddr, lo);
    tmem_ld16(taddr + 16, hi);
    if (valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i) lo[i] += sbias[i], hi[i] += sbias[16 + i];
      uint4 a0, a1, b0, b1;
      a0.x = pack_bf16x2(lo[0], lo[1]), a0.y = pack_bf16x2(lo[2], lo[3]), a0.z = pack_bf16x2(lo[4], lo[5]);
      a0.w = pack_bf16x2(lo[6], lo[7]), a1.x = pack_bf16x2(lo[8], lo[9]), a1.y = pack_bf16x2(lo[10], lo[11]);
      a1.z = pack_bf16x2(lo[12], lo[13]), a1.w = pack_bf16x2(lo[14], lo[15]);
      b0.x = pack_bf16x2(hi[0], hi[1]), b0.y = pack_bf16x2(hi[2], hi[3]), b0.z = pack_bf16x2(hi[4], hi[5]);
      b0.w = pack_bf16x2(hi[6], hi[7]), b1.x = pack_bf16x2(hi[8], hi[9]), b1.y = pack_bf16x2(hi[10], hi[11]);
      b1.z = pack_bf16x2(hi[12], hi[13]), b1.w = pack_bf16x2(hi[14], hi[15]);
      __nv_bfloat16* dst = out + p * 32;
      stg256(dst, a0, a1);
      stg256(dst + 16, b0, b1);
    }
    // the next tile overwrites sA and the accumulator: every warp must have finished its tcgen05.ld first
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace dfgan

using namespace dfgan;

extern "C" int dfgan_stem_fwd(const float* img, const float* rowscale, const void* w, const float* bias, void* out,
                              void* xcol, int N, int H, int W, dfgan_stream_t st) {
  if (!img || !w || !out || N < 1 || H < 1 || W < 1) return DFGAN_E_ARG;
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(xcol)) & 31) return DFGAN_E_ARG;
  if (reinterpret_cast<uintptr_t>(w) & 15) return DFGAN_E_ARG;
  const long long P = static_cast<long long>(N) * H * W;
  const long long tiles = (P + 127) / 128;
  long long grid = 6LL * num_sms();          // one resident wave (6 CTAs per SM: 60 KB of smem, 192 TMEM columns)
  if (grid > tiles) grid = tiles;
  stem_fwd_kernel<<<static_cast<int>(grid), kStemThreads, 0, static_cast<cudaStream_t>(st)>>>(
      img, rowscale, static_cast<const __nv_bfloat16*>(w), bias, static_cast<__nv_bfloat16*>(out),
      static_cast<__nv_bfloat16*>(xcol), H, W, P);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}
