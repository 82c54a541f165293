// Tap-list implicit-GEMM convolution (forward / data-gradient) and its weight gradient on the
// sm_100a tensor cores: TMA (tiled, zero-filled halos) -> shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue.  One persistent CTA per SM, three roles: warp 0 TMA producer, warp 1 MMA
// issuer (+ TMEM owner), warps 2..5 epilogue.  TMEM holds two accumulator stages so the epilogue
// of tile i overlaps the main loop of tile i+1.
//
// Geometry: an M tile is 128 output pixels = a BN x BH x BW brick of the NHWC output grid; for a
// tap (dy,dx) the A operand is the same brick of the source view shifted by (dy,dx), which a 4-D
// tiled tensor map delivers as 128 rows of BK channels in the canonical K-major swizzled layout
// (out-of-range rows/columns arrive as zeros = the convolution's zero padding).
#include "host_common.h"
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>

namespace dfgan {

constexpr int kConvThreads = 192;
constexpr uint32_t kSpinLimit = 1u << 27;

struct ConvMaps {
  CUtensorMap a[DFGAN_MAX_SRC];
  CUtensorMap b;
};

struct ConvK {
  int n_taps, n_chunks;
  int tap_src[DFGAN_MAX_TAPS], tap_dy[DFGAN_MAX_TAPS], tap_dx[DFGAN_MAX_TAPS],
      tap_wrow[DFGAN_MAX_TAPS];
  int lbw, lbh;  // log2(BW), log2(BH); BN = 128 >> (lbw + lbh)
  int tiles_w, tiles_h, tiles_n;
  int n_tiles_n, total_tiles;
  int n_tile;
  int N, H, W, Cout;
  uint32_t idesc;
  int n_stages;
  uint32_t a_bytes, b_bytes, tmem_cols;
  __nv_bfloat16* out;
  long long out_sw, out_sh, out_sn;
  const float* bias;
  const float* scale_dev;
  float scale_imm;
  int lrelu_pre;
  float slope;
  __nv_bfloat16* aux;
  const __nv_bfloat16* addend;
  long long add_sw, add_sh, add_sn;
  int add_shift;
  float add_scale;
  const __nv_bfloat16* mask_src;
  __nv_bfloat16* out2;
  int post_mode;  // 0 none, 1 lrelu, 2 affine-lrelu-affine-lrelu
  const float *tab_ga, *tab_ba, *tab_gb, *tab_bb;
  long long tab_stride;
  float* out_nchw;  // fp32 [N][co_valid][H][W] store (image boundary), optional tanh
  int co_valid, tanh_out;
  int phase_cout;   // > 0: columns are 4 output phases x phase_cout channels of a 2x-upsampled output grid
  int col_split;    // halo kernels: both epilogue warpgroups work on EVERY tile (alternate 16-column chunks)
                    // instead of alternating tiles -- for launches with only 1-2 tiles per CTA
  int k_splits;     // generic kernel: > 1 = split-K, each CTA runs a slice of the (tap, chunk) loop and adds its fp32
  float* ws;        // partial tile to ws[pixel][Cout]; conv_splitk_epilogue_kernel applies the fused epilogue afterwards
  int dbg;          // DFGAN_DBG bit mask (conv_halo_kernel only): 1 skip epilogue body, 2 skip MMAs, 4 skip TMA loads,
                    // 8 skip tcgen05.ld -- bottleneck isolation, results are garbage
  // generic kernel, weight tiles as ONE linear bulk copy per k-iteration from the operand packed as the swizzled
  // shared-memory image (no per-row TMA cost): b_img != nullptr; byte strides as in Halo2K
  const uint8_t* b_img;
  long long w_tap_stride, w_chunk_stride;
  // generic kernel, on-chip split-K: a thread-block cluster of `cluster` CTAs shares one output tile, CTA r runs
  // slice r of the (tap, chunk) loop; the fp32 partial tiles are exchanged through distributed shared memory
  // (CTA r receives columns [r * n_tile / cluster, ...) of every partial into stg[src][row][cols]), summed in rank
  // order and finished by the fused epilogue -- no global workspace, no second pass.  grid == total_tiles * cluster.
  int cluster;
  uint32_t stg_off;   // byte offset of the staging area inside the (1024-aligned) dynamic shared memory
  uint32_t epi;       // EPI_* feature mask the epilogue loop is instantiated for (EPI_GENERIC: run-time tests)
  int ltw, lth;       // log2(tiles_w), log2(tiles_h) when both are powers of two, else -1
};

__device__ __forceinline__ void wait_bounded(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n_stages) {
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// pack 16 floats to bf16 (round-to-nearest-even), store them, and replace v by the stored (rounded) values:
// bf16 -> fp32 is a 16-bit shift / mask of the packed word, no second conversion
__device__ __forceinline__ void store_round_bf16x16(__nv_bfloat16* dst, float* v) {
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  stg256(dst, make_uint4(u[0], u[1], u[2], u[3]), make_uint4(u[4], u[5], u[6], u[7]));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* dst, const float* v) {
  uint4 s0, s1;
  s0.x = pack_bf16x2(v[0], v[1]), s0.y = pack_bf16x2(v[2], v[3]);
  s0.z = pack_bf16x2(v[4], v[5]), s0.w = pack_bf16x2(v[6], v[7]);
  s1.x = pack_bf16x2(v[8], v[9]), s1.y = pack_bf16x2(v[10], v[11]);
  s1.z = pack_bf16x2(v[12], v[13]), s1.w = pack_bf16x2(v[14], v[15]);
  stg256(dst, s0, s1);
}
__device__ __forceinline__ void load_bf16x16(const __nv_bfloat16* src, float* v) {
  const uint4* sp = reinterpret_cast<const uint4*>(src);
  const uint4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
  const uint32_t u[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) unpack_bf16x2(u[i], v[2 * i], v[2 * i + 1]);
}
// 16 consecutive fp32 (64-byte aligned: bias vectors and gamma/beta table rows at 16-column chunks)
__device__ __forceinline__ void load_f32x16(const float* src, float* v) {
  const float4* sp = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = __ldg(sp + i);
    v[4 * i] = f.x, v[4 * i + 1] = f.y, v[4 * i + 2] = f.z, v[4 * i + 3] = f.w;
  }
}
// the same for 32-byte aligned rows (gamma/beta tables; dfgan_conv_run checks the alignment): two 256-bit loads
__device__ __forceinline__ void load_f32x16_a32(const float* src, float* v) {
  uint4 q[4];
  ldg256(src, q[0], q[1]);
  ldg256(src + 8, q[2], q[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[4 * i] = __uint_as_float(q[i].x), v[4 * i + 1] = __uint_as_float(q[i].y);
    v[4 * i + 2] = __uint_as_float(q[i].z), v[4 * i + 3] = __uint_as_float(q[i].w);
  }
}
__device__ __forceinline__ float round_bf16(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

// ---------------------------------------------------------------------------------------------
// Fused epilogue, specialised at compile time on its feature set.  The epilogue warps are latency-bound on their own
// instruction stream (ncu source page, round 2: 384 instructions per 128-pixel tile of a C_out = 32 layer for 32
// outputs per thread, a third of the samples in instruction-fetch / branch stalls while every feature is a run-time
// `if (p.x)` on a constant-bank load), so every kernel switches ONCE to a loop instantiated for the launch's feature
// mask; masks outside the curated list run the EPI_GENERIC instantiation (all features tested at run time).
enum : uint32_t {
  EPI_BIAS = 1, EPI_LRELU = 2, EPI_MASK = 4, EPI_AUX = 8, EPI_ADDEND = 16, EPI_OUT = 32, EPI_POST1 = 64, EPI_POST2 = 128,
  EPI_NCHW = 256, EPI_PHASE = 512, EPI_GENERIC = 0x80000000u
};
template <uint32_t F>
struct Epi {
  static constexpr bool G = (F & EPI_GENERIC) != 0;
  __device__ __forceinline__ static bool bias(const ConvK& p) { return G ? p.bias != nullptr : (F & EPI_BIAS) != 0; }
  __device__ __forceinline__ static bool lrelu(const ConvK& p) { return G ? p.lrelu_pre != 0 : (F & EPI_LRELU) != 0; }
  __device__ __forceinline__ static bool mask(const ConvK& p) { return G ? p.mask_src != nullptr : (F & EPI_MASK) != 0; }
  __device__ __forceinline__ static bool aux(const ConvK& p) { return G ? p.aux != nullptr : (F & EPI_AUX) != 0; }
  __device__ __forceinline__ static bool addend(const ConvK& p) { return G ? p.addend != nullptr : (F & EPI_ADDEND) != 0; }
  __device__ __forceinline__ static bool out(const ConvK& p) { return G ? p.out != nullptr : (F & EPI_OUT) != 0; }
  __device__ __forceinline__ static bool post(const ConvK& p) { return G ? p.post_mode != 0 : (F & (EPI_POST1 | EPI_POST2)) != 0; }
  __device__ __forceinline__ static bool post2(const ConvK& p) { return G ? p.post_mode == 2 : (F & EPI_POST2) != 0; }
  __device__ __forceinline__ static bool nchw(const ConvK& p) { return G ? p.out_nchw != nullptr : (F & EPI_NCHW) != 0; }
  __device__ __forceinline__ static bool phase(const ConvK& p) { return G ? p.phase_cout != 0 : (F & EPI_PHASE) != 0; }
};
// the masks the two networks launch (tools: DFGAN_DBG=64 prints them); everything else -> EPI_GENERIC
#define DFGAN_EPI_LIST(X) X(32u) X(33u) X(34u) X(36u) X(48u) X(58u) X(60u) X(137u) X(177u) X(113u) X(560u) X(649u) X(256u) X(257u)
__host__ __device__ constexpr bool epi_listed(uint32_t f) {
#define DFGAN_EPI_CMP(V) if (f == V) return true;
  DFGAN_EPI_LIST(DFGAN_EPI_CMP)
#undef DFGAN_EPI_CMP
  return false;
}
// EPI_DISPATCH(mask, CALL): CALL(F) expands to the statement run with the compile-time mask F
#define DFGAN_EPI_CASE_(V) case V: DFGAN_EPI_CALL_(V); break;
#define EPI_DISPATCH(mask)                    \
  switch (mask) {                             \
    DFGAN_EPI_LIST(DFGAN_EPI_CASE_)           \
    default: DFGAN_EPI_CALL_(EPI_GENERIC); break; \
  }

// One 16-column chunk of the fused epilogue for one output pixel (see dfgan_conv_run in the header).
// mask / addend rows of a chunk, requested ahead of use (their DRAM latency was the largest marginal cost of the
// fused epilogues, tools/prof_epilogue.py): chunk 0 before the accumulator-ready wait, chunk c+1 before chunk c
struct EpiPre {
  uint4 m0, m1, a0, a1;
};
// request the mask / addend rows of the chunk at tile column c (output column gcol) for this thread's pixel
template <uint32_t F>
__device__ __forceinline__ void epi_pre_issue(const ConvK& p, EpiPre& e, bool valid, int gcol, int c, int n, int h,
                                              int w, long long o_off, long long a_off) {
  if (!(Epi<F>::mask(p) || Epi<F>::addend(p))) return;
  if (!valid || c >= p.n_tile || gcol >= p.Cout) return;
  int col = gcol;
  long long oo = o_off;
  if (Epi<F>::phase(p)) {
    const int ph = gcol / p.phase_cout;
    col = gcol - ph * p.phase_cout;
    oo = n * p.out_sn + (2 * h + (ph >> 1)) * p.out_sh + (2 * w + (ph & 1)) * p.out_sw;
  }
  if (Epi<F>::mask(p)) ldg256(p.mask_src + oo + col, e.m0, e.m1);
  if (Epi<F>::addend(p)) ldg256(p.addend + a_off + col, e.a0, e.a1);
}
__device__ __forceinline__ void unpack_bf16x16(const uint4& s0, const uint4& s1, float* v) {
  const uint32_t u[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) unpack_bf16x2(u[i], v[2 * i], v[2 * i + 1]);
}

template <uint32_t F>
__device__ __forceinline__ void epilogue_chunk(const ConvK& p, float (&v)[16], int col, int n, int h, int w,
                                               long long o_off, long long a_off, float scale, const EpiPre& pre) {
    using E = Epi<F>;
    if (E::bias(p)) {
      float b[16];
      load_f32x16(p.bias + col, b);
#pragma unroll
      for (int i = 0; i < 16; i += 2) add2(v[i], v[i + 1], v[i], v[i + 1], b[i], b[i + 1]);
    }
    if (E::lrelu(p)) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) lrelu2(v[i], v[i + 1], p.slope);
    }
    if (E::mask(p)) {
      float m[16];
      unpack_bf16x16(pre.m0, pre.m1, m);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= (m[i] > 0.f ? 1.f : p.slope);
    }
    if (E::aux(p)) {
      // continue from the stored (bf16) value: what the backward pass reads back is exactly what
      // the rest of the forward used
      store_round_bf16x16(p.aux + o_off + col, v);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) mul2(v[i], v[i + 1], v[i], v[i + 1], scale, scale);
    if (E::addend(p)) {
      float a[16];
      unpack_bf16x16(pre.a0, pre.a1, a);
#pragma unroll
      for (int i = 0; i < 16; i += 2) fma2(v[i], v[i + 1], a[i], a[i + 1], p.add_scale, p.add_scale, v[i], v[i + 1]);
    }
    if (E::nchw(p)) {
      // image boundary: fp32 planar store, lanes of a warp hold consecutive w -> coalesced
      const long long hw = static_cast<long long>(p.H) * p.W;
      float* dst = p.out_nchw + (static_cast<long long>(n) * p.co_valid) * hw +
                   static_cast<long long>(h) * p.W + w;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (col + i < p.co_valid) dst[(col + i) * hw] = p.tanh_out ? tanhf(v[i]) : v[i];
      }
    }
    if (E::out(p)) {
      if (E::post(p)) store_round_bf16x16(p.out + o_off + col, v);
      else store_bf16x16(p.out + o_off + col, v);
    }
    if (E::post(p)) {
      // the follow-up pointwise chain is evaluated on the bf16-rounded value that `out`/`aux`
      // hold, so a backward pass that recomputes it from the stored tensor sees identical masks
      if (!E::out(p)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = round_bf16(v[i]);
      }
      if (E::post2(p)) {
        const long long t_off = static_cast<long long>(n) * p.tab_stride + col;
        float ga[16], ba[16];
        load_f32x16_a32(p.tab_ga + t_off, ga);
        load_f32x16_a32(p.tab_ba + t_off, ba);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          fma2(v[i], v[i + 1], ga[i], ga[i + 1], v[i], v[i + 1], ba[i], ba[i + 1]);
          lrelu2(v[i], v[i + 1], p.slope);
        }
        load_f32x16_a32(p.tab_gb + t_off, ga);
        load_f32x16_a32(p.tab_bb + t_off, ba);
#pragma unroll
        for (int i = 0; i < 16; i += 2) fma2(v[i], v[i + 1], ga[i], ga[i + 1], v[i], v[i + 1], ba[i], ba[i + 1]);
      }
#pragma unroll
      for (int i = 0; i < 16; i += 2) lrelu2(v[i], v[i + 1], p.slope);
      store_bf16x16(p.out2 + o_off + col, v);
    }
}
// chunk at output column `col` of pixel (n, h, w): resolves the output phase of merged-phase launches
template <uint32_t F>
__device__ __forceinline__ void epilogue_at(const ConvK& p, float (&v)[16], int col, int n, int h, int w, long long o_off,
                                            long long a_off, float scale, const EpiPre& pre) {
  if (Epi<F>::phase(p)) {
    const int ph = col / p.phase_cout;
    const long long oo = n * p.out_sn + (2 * h + (ph >> 1)) * p.out_sh + (2 * w + (ph & 1)) * p.out_sw;
    epilogue_chunk<F>(p, v, col - ph * p.phase_cout, n, h, w, oo, a_off, scale, pre);
  } else {
    epilogue_chunk<F>(p, v, col, n, h, w, o_off, a_off, scale, pre);
  }
}

// Prefetch depth of the mask / addend rows (16-column chunks requested ahead of use).  The epilogue warps are bound by
// the latency of these loads, not by their instruction stream (round 2: with ONE chunk in flight per thread a
// C_out = 64 dgrad + mask launch sat at 3700 cycles per 128-pixel tile = 4 chunks x ~one DRAM latency, 53 % of the HBM
// roofline, nothing saturated in ncu): all chunks of a tile (up to kEpiDepth) are requested BEFORE the accumulator
// wait, so the latency is exposed once per tile and overlaps the MMAs of that tile.  Register budget: 8 registers per
// tensor and chunk; launches that read both tensors keep two chunks in flight.
template <uint32_t F>
struct EpiDepth {
  static constexpr bool both = Epi<F>::G || ((F & EPI_MASK) && (F & EPI_ADDEND));
  static constexpr bool any = Epi<F>::G || (F & (EPI_MASK | EPI_ADDEND));
  static constexpr int D = any ? (both ? 2 : 4) : 1;
};
// requests chunks c0, c0 + cstep, ... (D of them) of a tile whose columns are [0, n_cols) at output column colbase + c
template <uint32_t F>
__device__ __forceinline__ void epi_prefetch_head(const ConvK& p, EpiPre (&pre)[EpiDepth<F>::D], bool valid, int colbase,
                                                  int c0, int cstep, int n_cols, int n, int h, int w, long long o_off,
                                                  long long a_off) {
  constexpr int D = EpiDepth<F>::D;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const int c = c0 + j * cstep;
    epi_pre_issue<F>(p, pre[j], valid, colbase + c, c < n_cols ? 0 : p.n_tile, n, h, w, o_off, a_off);
  }
}
// processes the chunks of a tile from TMEM address taddr (+ column), refilling the prefetch ring as it goes
template <uint32_t F>
__device__ __forceinline__ void epi_process_tile(const ConvK& p, EpiPre (&pre)[EpiDepth<F>::D], uint32_t taddr, bool valid,
                                                 int colbase, int c0, int cstep, int n_cols, int n, int h, int w,
                                                 long long o_off, long long a_off, float scale) {
  constexpr int D = EpiDepth<F>::D;
  for (int cb = c0; cb < n_cols; cb += D * cstep) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const int c = cb + j * cstep;
      if (c < n_cols) {
        const int col = colbase + c;
        float v[16];
        if (!(p.dbg & 8)) {
          tmem_ld16(taddr + c, v);
          if (valid && col < p.Cout && !(p.dbg & 1)) epilogue_at<F>(p, v, col, n, h, w, o_off, a_off, scale, pre[j]);
        }
        const int cn = c + D * cstep;
        epi_pre_issue<F>(p, pre[j], valid, colbase + cn, cn < n_cols ? 0 : p.n_tile, n, h, w, o_off, a_off);
      }
    }
  }
}

// tile index -> (tile column, tile row, image) of the 16 x 8 brick grid; powers of two (every layer of the two
// networks) take the shift / mask path instead of two integer divisions per tile and thread
__device__ __forceinline__ void brick_coords(const ConvK& p, int mt, int& tw, int& th, int& n) {
  if (p.ltw >= 0 && p.lth >= 0) {
    tw = mt & (p.tiles_w - 1);
    const int r = mt >> p.ltw;
    th = r & (p.tiles_h - 1);
    n = r >> p.lth;
  } else {
    tw = mt % p.tiles_w;
    const int r = mt / p.tiles_w;
    th = r % p.tiles_h;
    n = r / p.tiles_h;
  }
}

// Epilogue loop of the two halo kernels (16 x 8 pixel bricks; thread = pixel): two warpgroups alternate tiles over
// n_acc TMEM accumulator stages, or (col_split) both work on every tile on alternating 16-column chunks.
template <uint32_t F>
__device__ __forceinline__ void brick_epilogue(const ConvK& p, int n_acc, uint64_t* tfull, uint64_t* tempty,
                                               uint32_t tmem_base, int warp, int lane) {
  const int group = (warp - 2) >> 2;       // 0 or 1
  const int q = warp & 3;                   // TMEM lane quadrant this warp may read
  const int row = q * 32 + lane;
  const int bw = row & 7, bh = row >> 3;
  const float scale = p.scale_imm * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
  int acc = 0, i = 0;
  uint32_t aphase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++i) {
    if (p.col_split || (i & 1) == group) {
      int nt = 0, mt = tile;
      if (p.n_tiles_n > 1) {
        nt = tile % p.n_tiles_n;
        mt = tile / p.n_tiles_n;
      }
      int tw, th, n;
      brick_coords(p, mt, tw, th, n);
      const int w = (tw << 3) + bw, h = (th << 4) + bh;
      const bool valid = (w < p.W) && (h < p.H);
      const long long o_off = n * p.out_sn + h * p.out_sh + w * p.out_sw;
      const long long a_off = Epi<F>::addend(p) ? (n * p.add_sn + (h >> p.add_shift) * p.add_sh +
                                                   (w >> p.add_shift) * p.add_sw)
                                                : 0;
      const int c0 = p.col_split ? group * 16 : 0, cstep = p.col_split ? 32 : 16;
      const int colbase = nt * p.n_tile;
      EpiPre pre[EpiDepth<F>::D];
      epi_prefetch_head<F>(p, pre, valid, colbase, c0, cstep, p.n_tile, n, h, w, o_off, a_off);
      wait_bounded(&tfull[acc], aphase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.n_tile);
      epi_process_tile<F>(p, pre, taddr, valid, colbase, c0, cstep, p.n_tile, n, h, w, o_off, a_off, scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (++acc == n_acc) {
      acc = 0;
      aphase ^= 1;
    }
  }
}

// Epilogue of the generic kernel: thread = pixel row of a BN x BH x BW brick.
template <uint32_t F>
__device__ __forceinline__ void generic_epilogue(const ConvK& p, uint8_t* smem, uint64_t* tfull, uint64_t* tempty,
                                                 uint32_t tmem_base, int warp, int lane) {
  const int BW = 1 << p.lbw, BH = 1 << p.lbh;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int bw = row & (BW - 1);
  const int bh = (row >> p.lbw) & (BH - 1);
  const int bn = row >> (p.lbw + p.lbh);
  const float scale = p.scale_imm * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
  if (p.cluster > 1) {
    // ---- on-chip split-K (one work item per CTA: item == blockIdx.x, slice == cluster rank)
    const uint32_t S = static_cast<uint32_t>(p.cluster), rank = cluster_ctarank();
    const int cw = p.n_tile / p.cluster;                    // columns this CTA finishes (multiple of 16)
    const int tile = static_cast<int>(blockIdx.x) / p.k_splits;
    const int nt = tile % p.n_tiles_n;
    int mt = tile / p.n_tiles_n;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    const int tn = mt / p.tiles_h;
    const int w = (tw << p.lbw) + bw, h = (th << p.lbh) + bh;
    const int n = tn * (128 >> (p.lbw + p.lbh)) + bn;
    const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
    const long long o_off = n * p.out_sn + h * p.out_sh + w * p.out_sw;
    const long long a_off = Epi<F>::addend(p) ? (n * p.add_sn + (h >> p.add_shift) * p.add_sh + (w >> p.add_shift) * p.add_sw) : 0;
    float* stg = reinterpret_cast<float*>(smem + p.stg_off);          // [S][cw / 4][128 rows][4] fp32
    const int col0 = nt * p.n_tile + static_cast<int>(rank) * cw;      // first output column of this CTA's share
    EpiPre nxt{};
    epi_pre_issue<F>(p, nxt, valid, col0, 0, n, h, w, o_off, a_off);
    wait_bounded(&tfull[0], 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // staging layout [src][column quad][row][4 floats]: the 32 lanes of a warp (consecutive rows) write / read 512
    // consecutive bytes per instruction -- conflict-free on both sides of the exchange
    const int cq = cw >> 2;
    const uint32_t my_base = smem_u32(stg + (static_cast<size_t>(rank) * cq * 128 + row) * 4);
    for (uint32_t d = 0; d < S; ++d) {
      const uint32_t dst = mapa_cluster(my_base, d);
      for (int c = 0; c < cw; c += 16) {
        float v[16];
        tmem_ld16(taddr + d * cw + c, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          st_cluster_f32x4(dst + static_cast<uint32_t>(((c + i) >> 2) * 128 * 16), v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    for (int c = 0; c < cw; c += 16) {
      const EpiPre cur = nxt;
      const int col = col0 + c;
      epi_pre_issue<F>(p, nxt, valid, col + 16, c + 16 < cw ? 0 : p.n_tile, n, h, w, o_off, a_off);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
      for (uint32_t s2 = 0; s2 < S; ++s2) {                  // fixed rank order: the sum is reproducible
        const float4* src = reinterpret_cast<const float4*>(stg) + (static_cast<size_t>(s2) * cq + (c >> 2)) * 128 + row;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 f = src[i * 128];
          v[4 * i] += f.x, v[4 * i + 1] += f.y, v[4 * i + 2] += f.z, v[4 * i + 3] += f.w;
        }
      }
      if (valid && col < p.Cout) epilogue_at<F>(p, v, col, n, h, w, o_off, a_off, scale, cur);
    }
    return;
  }
  int as = 0;
  uint32_t aphase = 0;
  for (int item = blockIdx.x; item < p.total_tiles * p.k_splits; item += gridDim.x) {
    const int tile = item / p.k_splits;
    const int nt = tile % p.n_tiles_n;
    int mt = tile / p.n_tiles_n;
    const int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int th = mt % p.tiles_h;
    const int tn = mt / p.tiles_h;
    const int w = (tw << p.lbw) + bw, h = (th << p.lbh) + bh;
    const int n = tn * (128 >> (p.lbw + p.lbh)) + bn;
    const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
    const long long o_off = n * p.out_sn + h * p.out_sh + w * p.out_sw;
    const long long a_off = Epi<F>::addend(p) ? (n * p.add_sn + (h >> p.add_shift) * p.add_sh +
                                                 (w >> p.add_shift) * p.add_sw)
                                              : 0;
    if (p.k_splits > 1) {
      // split-K through global memory: add this slice's partial sums to the fp32 workspace; the fused epilogue runs
      // in a second pass
      wait_bounded(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.n_tile);
      float* wrow = p.ws + ((static_cast<long long>(n) * p.H + h) * p.W + w) * p.Cout + nt * p.n_tile;
      for (int c = 0; c < p.n_tile; c += 16) {
        float v[16];
        tmem_ld16(taddr + c, v);
        if (valid && nt * p.n_tile + c < p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(wrow + c + i, v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
      continue;
    }
    EpiPre pre[EpiDepth<F>::D];
    epi_prefetch_head<F>(p, pre, valid, nt * p.n_tile, 0, 16, p.n_tile, n, h, w, o_off, a_off);
    wait_bounded(&tfull[as], aphase);
    tc_fence_after();
    const uint32_t taddr =
        tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.n_tile);
    epi_process_tile<F>(p, pre, taddr, valid, nt * p.n_tile, 0, 16, p.n_tile, n, h, w, o_off, a_off, scale);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[as]);
    as ^= 1;
    if (as == 0) aphase ^= 1;
  }
}

template <int BK>
__global__ void __launch_bounds__(kConvThreads, 1)
    conv_umma_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvK p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.n_stages;
  uint64_t* tfull = bars + 2 * p.n_stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DFGAN_MAX_SRC; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // distributed shared memory may only be touched once every CTA of the cluster has started executing
  if (p.cluster > 1) cluster_sync_all();
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on its results are needed.  The successor may start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();

  const int BW = 1 << p.lbw, BH = 1 << p.lbh;
  const int k_iters = p.n_taps * p.n_chunks;

  if (warp == 0) {
    // whole warp runs the loop (coordinates stay warp-uniform -> uniform registers), one elected lane issues
    const bool leader = elect_one();
    PipeState ps;
    // work item = (tile, k-split): split s runs the slice [it0, it1) of the flattened (tap, chunk) loop
    for (int item = blockIdx.x; item < p.total_tiles * p.k_splits; item += gridDim.x) {
      const int tile = item / p.k_splits, ks = item - tile * p.k_splits;
      const int it0 = (k_iters * ks) / p.k_splits, it1 = (k_iters * (ks + 1)) / p.k_splits;
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int tn = mt / p.tiles_h;
      const int w0 = tw << p.lbw, h0 = th << p.lbh, n0 = tn * (128 >> (p.lbw + p.lbh));
      for (int it = it0; it < it1; ++it) {
        const int t = it / p.n_chunks, c = it - t * p.n_chunks;
        const CUtensorMap* ma = &maps.a[p.tap_src[t]];
        const int wy = h0 + p.tap_dy[t], wx = w0 + p.tap_dx[t];
        const int brow = p.tap_wrow[t] + nt * p.n_tile;
        wait_bounded(&empty[ps.stage], ps.phase ^ 1);
        uint8_t* sa = smem + static_cast<size_t>(ps.stage) * stage_bytes;
        if (leader) {
          mbar_expect_tx(&full[ps.stage], stage_bytes);
          tma_load_4d(sa, ma, &full[ps.stage], c * BK, wx, wy, n0);
          if (p.b_img)
            bulk_load(sa + p.a_bytes,
                      p.b_img + static_cast<long long>(p.tap_wrow[t]) * p.w_tap_stride + c * p.w_chunk_stride +
                          static_cast<long long>(nt) * p.n_tile * 128,
                      p.b_bytes, &full[ps.stage]);
          else
            tma_load_2d(sa + p.a_bytes, &maps.b, &full[ps.stage], c * BK, brow);
        }
        __syncwarp();
        ps.advance(p.n_stages);
      }
    }
    if (p.cluster > 1) {
      __syncwarp();
      cluster_sync_all();
    }
  } else if (warp == 1) {
    PipeState ps;
    int as = 0;
    uint32_t aphase = 0;
    constexpr uint32_t swz = (BK == 64) ? SWZ_128B : SWZ_64B;
    constexpr uint32_t sbo = 8u * BK * 2u;
    for (int item = blockIdx.x; item < p.total_tiles * p.k_splits; item += gridDim.x) {
      const int ks = item % p.k_splits;
      const int it0 = (k_iters * ks) / p.k_splits, it1 = (k_iters * (ks + 1)) / p.k_splits;
      wait_bounded(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.n_tile);
      // descriptors are computed warp-uniformly (uniform registers); only the tcgen05 instructions are
      // predicated on the elected lane -- see the note in conv_halo_kernel
      const bool leader = elect_one();
      for (int it = it0; it < it1; ++it) {
        wait_bounded(&full[ps.stage], ps.phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(ps.stage) * stage_bytes);
        const uint64_t adesc = make_smem_desc(a_addr, 16, sbo, swz);
        const uint64_t bdesc = make_smem_desc(a_addr + p.a_bytes, 16, sbo, swz);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if (leader) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, ((it - it0) | k) != 0);
        }
        if (leader) umma_commit(&empty[ps.stage]);
        __syncwarp();
        ps.advance(p.n_stages);
      }
      if (leader) umma_commit(&tfull[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (p.cluster > 1) {
      __syncwarp();
      cluster_sync_all();
    }
  } else {
#define DFGAN_EPI_CALL_(V) generic_epilogue<V>(p, smem, tfull, tempty, tmem_base, warp, lane)
    EPI_DISPATCH(p.epi)
#undef DFGAN_EPI_CALL_
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Second pass of a split-K launch: one thread per (pixel, 16-column chunk) reads the summed fp32 tile from the
// workspace (and zeroes it for the next launch) and applies the same fused epilogue as the single-pass kernels.
__global__ void __launch_bounds__(256) conv_splitk_epilogue_kernel(const __grid_constant__ ConvK p) {
  const int chunks = p.Cout / 16;
  const long long total = static_cast<long long>(p.N) * p.H * p.W * chunks;
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % chunks);
  const long long pix = i / chunks;
  const int w = static_cast<int>(pix % p.W), h = static_cast<int>((pix / p.W) % p.H);
  const int n = static_cast<int>(pix / (static_cast<long long>(p.W) * p.H));
  float* src = p.ws + pix * p.Cout + c * 16;
  float v[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 f = __ldcg(reinterpret_cast<const float4*>(src) + j);
    v[4 * j] = f.x, v[4 * j + 1] = f.y, v[4 * j + 2] = f.z, v[4 * j + 3] = f.w;
    reinterpret_cast<float4*>(src)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float scale = p.scale_imm * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
  const long long o_off = n * p.out_sn + h * p.out_sh + w * p.out_sw;
  const long long a_off = p.addend ? (n * p.add_sn + (h >> p.add_shift) * p.add_sh + (w >> p.add_shift) * p.add_sw) : 0;
  EpiPre pre{};
  epi_pre_issue<EPI_GENERIC>(p, pre, true, c * 16, 0, n, h, w, o_off, a_off);
  epilogue_at<EPI_GENERIC>(p, v, c * 16, n, h, w, o_off, a_off, scale, pre);
}

// ---------------------------------------------------------------------------------------------
// Weight gradient.  D[(tap,ci) x co] += sum_pixels X_shifted[pixel,(tap,ci)] * G[pixel,co]; both
// operands are MN-major (channels contiguous, the pixel index is the GEMM K), split-K over pixel
// tiles, fp32 atomics into the gradient tensor.
struct WgradMaps {
  CUtensorMap a[DFGAN_MAX_SRC];
  CUtensorMap g;
};

struct WgradK {
  int n_taps;
  int tap_src[DFGAN_MAX_TAPS], tap_dy[DFGAN_MAX_TAPS], tap_dx[DFGAN_MAX_TAPS],
      tap_id[DFGAN_MAX_TAPS];
  int lbw, lbh;
  int tiles_w, tiles_h, tiles_n, pix_tiles;
  int cb, lcb;          // channels per A row block (64 or 32), log2
  int blocks_per_tap;   // Cin / cb
  int row_blocks;       // n_taps * blocks_per_tap
  int cbg;              // channels per B block
  int n_tile, n_blocks_b;
  int n_tiles_m, n_tiles_n, k_splits, total_items;
  int Cin, Cout;
  uint32_t idesc;
  int n_stages;
  uint32_t a_bytes, b_bytes, a_blk_bytes, b_blk_bytes, tmem_cols;
  float* dw;
  long long s_co, s_ci, s_tap;
  const float* scale_dev;
  float scale_imm;
  int ci_valid, co_valid;
};

__global__ void __launch_bounds__(kConvThreads, 1)
    wgrad_umma_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradK p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.n_stages;
  uint64_t* tfull = bars + 2 * p.n_stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DFGAN_MAX_SRC; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.g);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on its results are needed.  The successor may start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();
  const int blocks_per_tile = 128 >> p.lcb;

  if (warp == 0) {
    const bool leader = elect_one();
    {
      PipeState ps;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int ks = item % p.k_splits;
        int r = item / p.k_splits;
        const int nt = r % p.n_tiles_n;
        const int mt = r / p.n_tiles_n;
        const int pt0 = static_cast<int>((static_cast<long long>(p.pix_tiles) * ks) / p.k_splits);
        const int pt1 =
            static_cast<int>((static_cast<long long>(p.pix_tiles) * (ks + 1)) / p.k_splits);
        for (int pt = pt0; pt < pt1; ++pt) {
          int m = pt;
          const int tw = m % p.tiles_w;
          m /= p.tiles_w;
          const int th = m % p.tiles_h;
          const int tn = m / p.tiles_h;
          const int w0 = tw << p.lbw, h0 = th << p.lbh, n0 = tn * (128 >> (p.lbw + p.lbh));
          wait_bounded(&empty[ps.stage], ps.phase ^ 1);
          uint8_t* sa = smem + static_cast<size_t>(ps.stage) * stage_bytes;
          if (leader) mbar_expect_tx(&full[ps.stage], stage_bytes);
          for (int b = 0; b < blocks_per_tile; ++b) {
            int rb = mt * blocks_per_tile + b;
            if (rb >= p.row_blocks) rb = mt * blocks_per_tile;  // padding block: reload a valid one
            const int t = rb / p.blocks_per_tap;
            const int c0 = (rb % p.blocks_per_tap) << p.lcb;
            if (leader)
              tma_load_4d(sa + b * p.a_blk_bytes, &maps.a[p.tap_src[t]], &full[ps.stage], c0,
                          w0 + p.tap_dx[t], h0 + p.tap_dy[t], n0);
          }
          for (int b = 0; b < p.n_blocks_b; ++b) {
            if (leader)
              tma_load_4d(sa + p.a_bytes + b * p.b_blk_bytes, &maps.g, &full[ps.stage],
                          nt * p.n_tile + b * p.cbg, w0, h0, n0);
          }
          __syncwarp();
          ps.advance(p.n_stages);
        }
      }
    }
  } else if (warp == 1) {
    PipeState ps;
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t swz_a = (p.cb == 64) ? SWZ_128B : SWZ_64B;
    const uint32_t swz_b = (p.cbg == 64) ? SWZ_128B : SWZ_64B;
    const uint32_t sbo_a = 8u * p.cb * 2u, sbo_b = 8u * p.cbg * 2u;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int ks = item % p.k_splits;
      const int pt0 = static_cast<int>((static_cast<long long>(p.pix_tiles) * ks) / p.k_splits);
      const int pt1 =
          static_cast<int>((static_cast<long long>(p.pix_tiles) * (ks + 1)) / p.k_splits);
      wait_bounded(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.n_tile);
      const bool leader = elect_one();   // descriptors stay warp-uniform, see conv_halo_kernel
      for (int pt = pt0; pt < pt1; ++pt) {
        wait_bounded(&full[ps.stage], ps.phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(ps.stage) * stage_bytes);
        const uint64_t adesc = make_smem_desc(a_addr, p.a_blk_bytes, sbo_a, swz_a);
        const uint64_t bdesc = make_smem_desc(a_addr + p.a_bytes, p.b_blk_bytes, sbo_b, swz_b);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 pixels per MMA = two 8-pixel swizzle atoms
          if (leader)
            umma_bf16(d_tmem, adesc + static_cast<uint64_t>((2u * sbo_a * k) >> 4),
                      bdesc + static_cast<uint64_t>((2u * sbo_b * k) >> 4), p.idesc,
                      (pt > pt0) || (k != 0));
        }
        if (leader) umma_commit(&empty[ps.stage]);
        __syncwarp();
        ps.advance(p.n_stages);
      }
      if (leader) umma_commit(&tfull[as]);
      __syncwarp();
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const float scale = p.scale_imm * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int r = item / p.k_splits;
      const int nt = r % p.n_tiles_n;
      const int mt = r / p.n_tiles_n;
      const int rb = mt * blocks_per_tile + (row >> p.lcb);
      const int ci = ((rb % p.blocks_per_tap) << p.lcb) + (row & (p.cb - 1));
      const bool valid = rb < p.row_blocks && ci < p.ci_valid;
      const int t = rb < p.row_blocks ? rb / p.blocks_per_tap : 0;
      float* dst = p.dw + ci * p.s_ci + p.tap_id[t] * p.s_tap;
      wait_bounded(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * p.n_tile);
      for (int c = 0; c < p.n_tile; c += 16) {
        float v[16];
        tmem_ld16(taddr + c, v);
        const int col = nt * p.n_tile + c;
        if (valid && col < p.Cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < p.co_valid) atomicAdd(dst + (col + i) * p.s_co, v[i] * scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

constexpr size_t kSmemBudget = 200 * 1024;

// ---------------------------------------------------------------------------------------------
// Halo variant for the high-resolution, narrow layers (C_in <= 64), where the tap-list kernel above is bound
// by the TMA engine: it delivers about one box row (<= 128 B) per ~4.5 cycles per SM whatever the row size
// (measured: 1152 rows per 128-pixel tile for a 3x3 layer).  Here an M tile is a 16(h) x 8(w) pixel brick and
// the brick PLUS ITS HALO is loaded once per source view as pixel rows of C_in*2 bytes in the swizzled
// K-major layout; a filter tap (dy,dx) is then only a different start row of the UMMA descriptor:
//   mode 0: one halo tile per dx (row pitch 8 pixels): every tap start is swizzle-atom aligned      (432 rows)
//   mode 1: one halo tile, row pitch 16 pixels: dy shifts are atom aligned, dx shifts use the          (288 rows)
//           descriptor's base-offset field
//   mode 2: one halo tile, row pitch 8 + halo: arbitrary row shifts, base-offset field per tap         (180 rows)
// All taps of the packed weight matrix stay resident in shared memory for the life of the CTA.
// Roles: warp 0 TMA producer, warps 1 and 10 MMA issuers (alternate tiles), warps 2..9 two epilogue warpgroups alternating tiles over
// up to four TMEM accumulator stages.
constexpr int kHaloThreads = 320;
// conv_halo_kernel runs a SECOND MMA-issuing warp (warp 10): tcgen05.mma issue is effectively synchronous with the
// tensor pipe (tools/ubench/mma_pattern.cu: every cycle the issuing thread spends between two tiles -- barrier waits,
// fences, descriptor set-up, commits; ~700 cycles in this kernel -- is a cycle the pipe idles: 2450 instead of 1750
// cycles per 36-MMA tile), and two warps that alternate tiles hide each other's gaps completely.  Tiles have their
// own accumulators, so the two issue streams need no ordering between them.
constexpr int kHaloNarrowThreads = 352;

// Tap geometries whose A-descriptor offsets are compile-time constants (taps sorted by source, dy, dx).  The MMA
// issue loop is bound by moving descriptor words from vector to uniform registers (R2UR, ~20 cycles each, 11 per tap
// when the offsets are runtime values added in the vector datapath); with immediates the whole tile is one R2UR'd base
// plus UIADD3.64 per MMA.  GEOM_RUNTIME keeps host-computed offsets (any tap list).
enum : int { GEOM_RUNTIME = 0, GEOM_1 = 1, GEOM_3X3 = 2, GEOM_2X2X4 = 3, GEOM_POOL4 = 4, GEOM_2X2 = 5 };
__host__ __device__ constexpr int geom_taps(int geom) {
  return geom == GEOM_1 ? 1 : (geom == GEOM_3X3 ? 9 : (geom == GEOM_2X2X4 ? 16 : (geom == GEOM_RUNTIME ? 0 : 4)));
}
// (halo-tile part + start row of tap t) >> 4 for channel count cin
__host__ __device__ constexpr uint32_t geom_off16(int geom, int cin, int t) {
  const uint32_t row = static_cast<uint32_t>(cin) * 2u;
  if (geom == GEOM_3X3) return (static_cast<uint32_t>((t / 3) * 10 + (t % 3)) * row) >> 4;
  if (geom == GEOM_2X2) return (static_cast<uint32_t>((t / 2) * 9 + (t % 2)) * row) >> 4;
  if (geom == GEOM_POOL4) return (static_cast<uint32_t>(t) * 16u * 8u * row) >> 4;
  if (geom == GEOM_2X2X4) {
    const uint32_t part = (17u * 9u * row + 1023u) & ~1023u;
    return (static_cast<uint32_t>(t / 4) * part + static_cast<uint32_t>(((t % 4) / 2) * 9 + (t % 2)) * row) >> 4;
  }
  return 0;
}

struct HaloK {
  int dy_min[DFGAN_MAX_SRC], dx_min[DFGAN_MAX_SRC];
  int n_parts[DFGAN_MAX_SRC];  // halo tiles per source (mode 0: one per dx)
  int part_base[DFGAN_MAX_SRC];
  int n_src, n_parts_total, mode;
  int pitch, hal_h;            // pixels per halo row in shared memory, halo rows
  uint32_t a_part_bytes;       // hal_h * pitch * C_in * 2, padded to 1024 B
  uint32_t b_tap_bytes;        // n_tile * C_in * 2
  int n_acc;                   // TMEM accumulator stages (2 or 4)
  int n_a_stages;              // halo tiles in flight (TMA latency ~2 us: 2 is not enough)
  int use_bo;
  uint32_t a_tap_off16[DFGAN_MAX_TAPS];   // (halo-tile part + start row of the tap) >> 4, precomputed on the host
};

__device__ __forceinline__ uint64_t with_base_offset(uint64_t desc, uint32_t saddr) {
  return desc | (static_cast<uint64_t>((saddr >> 7) & 7u) << 49);
}

template <int CIN, int GEOM>
__global__ void __launch_bounds__(kHaloNarrowThreads, 1)
    conv_halo_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvK p,
                     const __grid_constant__ HaloK hp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int kAStages = hp.n_a_stages;
  constexpr uint32_t kRowBytes = CIN * 2u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_taps) * hp.b_tap_bytes;
  const uint32_t a_stage_bytes = static_cast<uint32_t>(hp.n_parts_total) * hp.a_part_bytes;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kAStages * a_stage_bytes);
  uint64_t* full = bars;               // [8]
  uint64_t* empty = bars + 8;          // [8]
  uint64_t* tfull = bars + 16;         // [4]
  uint64_t* tempty = bars + 20;        // [4]
  uint64_t* bfull = bars + 24;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = hp.n_acc * p.n_tile <= 32 ? 32u : (hp.n_acc * p.n_tile <= 64 ? 64u :
                             (hp.n_acc * p.n_tile <= 128 ? 128u : (hp.n_acc * p.n_tile <= 256 ? 256u : 512u)));

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DFGAN_MAX_SRC; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < kAStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], p.col_split ? 8 : 4);
    }
    mbar_init(bfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on its results are needed.  The successor may start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) mbar_expect_tx(bfull, b_bytes);
    for (int t = 0; t < p.n_taps; ++t) {
      if (leader) tma_load_2d(sB + static_cast<size_t>(t) * hp.b_tap_bytes, &maps.b, bfull, 0, p.tap_wrow[t]);
    }
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = static_cast<uint32_t>(hp.n_parts_total) * hp.hal_h * hp.pitch * kRowBytes;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int tw = tile % p.tiles_w;
      const int r = tile / p.tiles_w;
      const int th = r % p.tiles_h;
      const int n = r / p.tiles_h;
      wait_bounded(&empty[stage], phase ^ 1);
      if (p.dbg & 4) {
        if (leader) mbar_arrive(&full[stage]);
        __syncwarp();
        if (++stage == kAStages) {
          stage = 0;
          phase ^= 1;
        }
        continue;
      }
      if (leader) mbar_expect_tx(&full[stage], tx);
      uint8_t* dst = sA + static_cast<size_t>(stage) * a_stage_bytes;
      for (int s = 0; s < hp.n_src; ++s)
        for (int q = 0; q < hp.n_parts[s]; ++q) {
          if (leader)
            tma_load_4d(dst + static_cast<size_t>(hp.part_base[s] + q) * hp.a_part_bytes, &maps.a[s], &full[stage],
                        0, (tw << 3) + hp.dx_min[s] + q, (th << 4) + hp.dy_min[s], n);
        }
      __syncwarp();
      if (++stage == kAStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 || warp == 10) {
    constexpr uint32_t swz = (CIN == 64) ? SWZ_128B : SWZ_64B;
    constexpr uint32_t sbo_b = 8u * kRowBytes;
    const uint32_t sbo_a = static_cast<uint32_t>(hp.pitch) * kRowBytes;
    // Two issuing warps alternate tiles (see kHaloNarrowThreads); both walk the whole tile sequence so that the
    // stage / accumulator counters stay those of the single-issuer loop.
    // Everything the MMA needs (descriptors, TMEM address) is computed by the WHOLE warp so that it is
    // warp-uniform and lives in uniform registers; only the tcgen05 instructions themselves sit under the
    // elected-lane predicate.  (Computing descriptors inside `if (lane == 0)` makes the compiler wrap every
    // UTCHMMA in an ELECT/R2UR waterfall loop -- ~150 cycles per MMA, the former bound of this kernel.)
    const int mine = (warp == 1) ? 0 : 1;
    int stage = 0, acc = 0, it = 0;
    uint32_t phase = 0, aphase = 0;
    wait_bounded(bfull, 0);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, sbo_b, swz);
    const uint64_t b_step16 = hp.b_tap_bytes >> 4;
    const bool leader = elect_one();
    const bool dual = !(p.dbg & 16);     // DFGAN_DBG bit 16: single issuing warp (A/B measurement)
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      if (dual ? ((it & 1) == mine) : (mine == 0)) {
        wait_bounded(&tempty[acc], aphase ^ 1);
        wait_bounded(&full[stage], phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
        const uint32_t a_base = smem_u32(sA + static_cast<size_t>(stage) * a_stage_bytes);
        // The issue loop must stay a handful of uniform-datapath instructions per MMA (an N=32 MMA is ~40 cycles,
        // tools/ubench/mma_cost.cu): tap offsets come precomputed from the host, the tap loop is unrolled at
        // compile time (NT) so they are constant-bank immediates.
        const uint64_t adesc0 = make_smem_desc(a_base, 16, sbo_a, swz);
        if (leader && !(p.dbg & 2)) {
          // ONE elected-lane region for the whole tile; the two bases pass through an empty asm so that nothing
          // derived from them is hoisted out of the tile loop into vector registers (each use would cost two R2UR)
          uint64_t ad0 = adesc0, bd0 = bdesc0;
          asm volatile("" : "+l"(ad0), "+l"(bd0));
          constexpr int NT = geom_taps(GEOM);
          if (NT > 0) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              const uint64_t adesc = ad0 + geom_off16(GEOM, CIN, t);
              const uint64_t bdesc = bd0 + static_cast<uint64_t>(t) * b_step16;
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (t | k) != 0);
            }
          } else {
            for (int t = 0; t < p.n_taps; ++t) {
              const uint64_t adesc = ad0 + hp.a_tap_off16[t];
              const uint64_t bdesc = bd0 + static_cast<uint64_t>(t) * b_step16;
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (t | k) != 0);
            }
          }
        }
        if (leader) {
          umma_commit(&empty[stage]);
          umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
      if (++stage == kAStages) {
        stage = 0;
        phase ^= 1;
      }
      if (++acc == hp.n_acc) {
        acc = 0;
        aphase ^= 1;
      }
    }
  } else {
#define DFGAN_EPI_CALL_(V) brick_epilogue<V>(p, hp.n_acc, tfull, tempty, tmem_base, warp, lane)
    EPI_DISPATCH(p.epi)
#undef DFGAN_EPI_CALL_
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Decides whether a launch can take the halo path and fills its geometry; returns false to fall back.
static bool plan_halo(const dfgan_conv_desc* d, HaloK& hp, int& smem_bytes) {
  if (d->Cin != 32 && d->Cin != 64) return false;
  if (d->Cout % 16 || d->Cout > 128) return false;
  if (d->W < 8 || d->H < 16 || static_cast<long long>(d->W) * d->H < 64 * 64) return false;
  // mode 2 is the default: tcgen05 applies the swizzle XOR to absolute shared-memory address bits, so a
  // descriptor may start at any 16-byte-aligned row of a TMA-written swizzled tile (verified on B200 by
  // tools/bringup_conv.py halo for all three modes)
  static const int mode_env = getenv("DFGAN_HALO_MODE") ? atoi(getenv("DFGAN_HALO_MODE")) : 2;
  int dy_max[DFGAN_MAX_SRC], dx_max[DFGAN_MAX_SRC];
  bool used[DFGAN_MAX_SRC] = {false, false, false, false};
  for (int s = 0; s < DFGAN_MAX_SRC; ++s) hp.dy_min[s] = hp.dx_min[s] = dy_max[s] = dx_max[s] = 0;
  for (int t = 0; t < d->n_taps; ++t) {
    const int s = d->taps[t].src;
    if (s < 0 || s >= d->n_src) return false;
    if (!used[s]) {
      used[s] = true;
      hp.dy_min[s] = dy_max[s] = d->taps[t].dy;
      hp.dx_min[s] = dx_max[s] = d->taps[t].dx;
    } else {
      hp.dy_min[s] = d->taps[t].dy < hp.dy_min[s] ? d->taps[t].dy : hp.dy_min[s];
      dy_max[s] = d->taps[t].dy > dy_max[s] ? d->taps[t].dy : dy_max[s];
      hp.dx_min[s] = d->taps[t].dx < hp.dx_min[s] ? d->taps[t].dx : hp.dx_min[s];
      dx_max[s] = d->taps[t].dx > dx_max[s] ? d->taps[t].dx : dx_max[s];
    }
  }
  int ext_h = 0, ext_w = 0;
  for (int s = 0; s < d->n_src; ++s) {
    if (!used[s]) return false;
    ext_h = dy_max[s] - hp.dy_min[s] > ext_h ? dy_max[s] - hp.dy_min[s] : ext_h;
    ext_w = dx_max[s] - hp.dx_min[s] > ext_w ? dx_max[s] - hp.dx_min[s] : ext_w;
  }
  if (ext_h > 2 || ext_w > 2) return false;
  hp.mode = ext_w == 0 ? 0 : mode_env;   // no dx shift: a single aligned tile is all three modes at once
  hp.n_src = d->n_src;
  hp.hal_h = 16 + ext_h;
  hp.pitch = hp.mode == 0 ? 8 : (hp.mode == 1 ? 16 : 8 + ext_w);
  int total = 0;
  for (int s = 0; s < d->n_src; ++s) {
    hp.part_base[s] = total;
    hp.n_parts[s] = hp.mode == 0 ? (dx_max[s] - hp.dx_min[s] + 1) : 1;
    total += hp.n_parts[s];
  }
  hp.n_parts_total = total;
  hp.a_part_bytes = (static_cast<uint32_t>(hp.hal_h * hp.pitch * d->Cin * 2) + 1023u) & ~1023u;
  hp.b_tap_bytes = static_cast<uint32_t>(d->Cout) * d->Cin * 2u;
  hp.n_acc = 4;
  hp.use_bo = 0;
  const size_t fixed = static_cast<size_t>(d->n_taps) * hp.b_tap_bytes + 1024 + 512;
  const size_t per_stage = static_cast<size_t>(total) * hp.a_part_bytes;
  const size_t cap = 220 * 1024;
  if (fixed + 2 * per_stage > cap) return false;
  int stages = static_cast<int>((cap - fixed) / per_stage);
  hp.n_a_stages = stages > 8 ? 8 : stages;
  smem_bytes = static_cast<int>(fixed + hp.n_a_stages * per_stage);
  for (int t = 0; t < d->n_taps; ++t) {
    const int s = d->taps[t].src;
    const int ddx = d->taps[t].dx - hp.dx_min[s], ddy = d->taps[t].dy - hp.dy_min[s];
    const uint32_t part = static_cast<uint32_t>(hp.part_base[s] + (hp.mode == 0 ? ddx : 0));
    const uint32_t row0 = static_cast<uint32_t>(ddy * hp.pitch + (hp.mode == 0 ? 0 : ddx));
    hp.a_tap_off16[t] = (part * hp.a_part_bytes + row0 * static_cast<uint32_t>(d->Cin) * 2u) >> 4;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Halo variant of the weight gradient for the high-resolution, narrow layers (C_in, C_out <= 64).  Per
// 16 x 8 pixel brick ONE halo tile per source view and ONE tile of the output gradient are loaded (the generic
// kernel re-loads both once per 128-row block of (tap, ci): ~6x the TMA rows).  The pixel index is the GEMM K.
//   A (M = 128 rows) = up to 128 / C_in filter taps that are horizontal neighbours: tap dx+1 is the same halo
//       tile one pixel row (C_in * 2 bytes) further, i.e. the descriptor's leading-dimension offset;
//   B (N = C_out)    = the output-gradient brick;      D[group] lives in its own TMEM columns.
// A CTA accumulates ALL its pixel bricks (a contiguous slice of the image batch) in TMEM and adds the result
// to the fp32 gradient once at the end.
constexpr int kWHaloGroupsMax = 8;

struct WHaloK {
  int n_src, n_groups, slots;            // slots = 128 / C_in taps per MMA
  int dy_min[DFGAN_MAX_SRC], dx_min[DFGAN_MAX_SRC];
  int part_base[DFGAN_MAX_SRC];          // halo tile index of each source inside a stage (== source index)
  int grp_src[kWHaloGroupsMax], grp_row0[kWHaloGroupsMax];   // start row of slot 0 inside the source's tile
  uint32_t grp_off16[kWHaloGroupsMax];                      // (source tile + start row) >> 4, precomputed
  int grp_tap[kWHaloGroupsMax][4];       // output tap id per slot, -1 = padding slot (rows discarded)
  int pitch, hal_h;
  uint32_t a_part_bytes, g_bytes, stage_bytes;
  int n_stages;
  int tiles_w, tiles_h, pix_tiles;
  int Cin, Cout;
  uint32_t idesc, tmem_cols;
  float* dw;
  long long s_co, s_ci, s_tap;
  const float* scale_dev;
  float scale_imm;
  int ci_valid, co_valid;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
    wgrad_halo_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WHaloK p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr uint32_t kRowA = CIN * 2u, kRowG = COUT * 2u;
  constexpr uint32_t swz_a = (CIN == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t swz_g = (COUT == 64) ? SWZ_128B : SWZ_64B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.n_stages) * p.stage_bytes);
  uint64_t* full = bars;        // [8]
  uint64_t* empty = bars + 8;   // [8]
  uint64_t* done = bars + 16;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DFGAN_MAX_SRC; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.g);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on its results are needed.  The successor may start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();

  // contiguous slice of pixel bricks for this CTA
  const int t0 = static_cast<int>((static_cast<long long>(p.pix_tiles) * blockIdx.x) / gridDim.x);
  const int t1 = static_cast<int>((static_cast<long long>(p.pix_tiles) * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0) {
    const bool leader = elect_one();
    PipeState ps;
    const uint32_t tx = static_cast<uint32_t>(p.n_src) * p.hal_h * p.pitch * kRowA + 128u * kRowG;
    for (int tile = t0; tile < t1; ++tile) {
      const int tw = tile % p.tiles_w;
      const int r = tile / p.tiles_w;
      const int th = r % p.tiles_h;
      const int n = r / p.tiles_h;
      wait_bounded(&empty[ps.stage], ps.phase ^ 1);
      uint8_t* dst = smem + static_cast<size_t>(ps.stage) * p.stage_bytes;
      if (leader) mbar_expect_tx(&full[ps.stage], tx);
      for (int s = 0; s < p.n_src; ++s) {
        if (leader)
          tma_load_4d(dst + static_cast<size_t>(s) * p.a_part_bytes, &maps.a[s], &full[ps.stage], 0,
                      (tw << 3) + p.dx_min[s], (th << 4) + p.dy_min[s], n);
      }
      if (leader)
        tma_load_4d(dst + static_cast<size_t>(p.n_src) * p.a_part_bytes, &maps.g, &full[ps.stage], 0, tw << 3,
                    th << 4, n);
      __syncwarp();
      ps.advance(p.n_stages);
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    PipeState ps;
    const uint32_t sbo_a = static_cast<uint32_t>(p.pitch) * kRowA;   // next 8-pixel K group = next image row
    constexpr uint32_t sbo_g = 8u * kRowG;
    const uint64_t a_kstep16 = (2u * sbo_a) >> 4;
    constexpr uint64_t g_kstep16 = (2u * sbo_g) >> 4;
    for (int tile = t0; tile < t1; ++tile) {
      wait_bounded(&full[ps.stage], ps.phase);
      tc_fence_after();
      const uint32_t st_base = smem_u32(smem + static_cast<size_t>(ps.stage) * p.stage_bytes);
      const uint32_t g_base = st_base + static_cast<uint32_t>(p.n_src) * p.a_part_bytes;
      // ONE elected-lane region per brick; the bases pass through an empty asm so nothing derived from them is
      // hoisted into vector registers (see conv_halo_kernel): the issue loop is UIADD3.64 + UTCHMMA only
      const uint64_t adesc0 = make_smem_desc(st_base, kRowA, sbo_a, swz_a);
      const uint64_t bdesc0 = make_smem_desc(g_base, kRowG, sbo_g, swz_g);
      const uint32_t acc_flag = tile > t0;
      if (leader) {
        uint64_t ad0 = adesc0, bd0 = bdesc0, ks = a_kstep16;
        asm volatile("" : "+l"(ad0), "+l"(bd0), "+l"(ks));
#pragma unroll
        for (int g = 0; g < kWHaloGroupsMax; ++g) {
          if (g < p.n_groups) {
            uint64_t ag = ad0 + p.grp_off16[g];
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * COUT);
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 16 pixels per MMA: image rows 2k and 2k+1 of the brick
              umma_bf16(d_tmem, ag, bd0 + static_cast<uint64_t>(k) * g_kstep16, p.idesc, acc_flag | (k != 0));
              ag += ks;                     // running add: one uniform add per MMA instead of hoisted multiples
            }
          }
        }
      }
      if (leader) umma_commit(&empty[ps.stage]);
      __syncwarp();
      ps.advance(p.n_stages);
    }
    if (leader) umma_commit(done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int slot = row / CIN, ci = row % CIN;
    const float scale = p.scale_imm * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
    if (t1 > t0) {
      wait_bounded(done, 0);
      tc_fence_after();
      for (int g = 0; g < p.n_groups; ++g) {
        const int tap = p.grp_tap[g][slot];
        const bool valid = tap >= 0 && ci < p.ci_valid;
        float* dst = p.dw + ci * p.s_ci + (tap >= 0 ? tap : 0) * p.s_tap;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * COUT);
#pragma unroll
        for (int c = 0; c < COUT; c += 16) {
          float v[16];
          tmem_ld16(taddr + c, v);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < p.co_valid) atomicAdd(dst + (c + i) * p.s_co, v[i] * scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Plans the halo weight-gradient launch; returns false when the shape must take the generic kernel.
static bool plan_wgrad_halo(const dfgan_wgrad_desc* d, WHaloK& k, int& smem_bytes) {
  if ((d->Cin != 32 && d->Cin != 64) || (d->Cout != 32 && d->Cout != 64)) return false;
  const int W = d->g.W, H = d->g.H;
  if (W < 8 || H < 16 || static_cast<long long>(W) * H < 64 * 64) return false;
  int dy_max[DFGAN_MAX_SRC], dx_max[DFGAN_MAX_SRC];
  bool used[DFGAN_MAX_SRC] = {false, false, false, false};
  for (int s = 0; s < DFGAN_MAX_SRC; ++s) k.dy_min[s] = k.dx_min[s] = dy_max[s] = dx_max[s] = 0;
  for (int t = 0; t < d->n_taps; ++t) {
    const int s = d->taps[t].src;
    if (s < 0 || s >= d->n_src) return false;
    const int dy = d->taps[t].dy, dx = d->taps[t].dx;
    if (!used[s]) {
      used[s] = true;
      k.dy_min[s] = dy_max[s] = dy;
      k.dx_min[s] = dx_max[s] = dx;
    } else {
      k.dy_min[s] = dy < k.dy_min[s] ? dy : k.dy_min[s];
      dy_max[s] = dy > dy_max[s] ? dy : dy_max[s];
      k.dx_min[s] = dx < k.dx_min[s] ? dx : k.dx_min[s];
      dx_max[s] = dx > dx_max[s] ? dx : dx_max[s];
    }
  }
  int ext_h = 0, ext_w = 0;
  for (int s = 0; s < d->n_src; ++s) {
    if (!used[s]) return false;
    ext_h = dy_max[s] - k.dy_min[s] > ext_h ? dy_max[s] - k.dy_min[s] : ext_h;
    ext_w = dx_max[s] - k.dx_min[s] > ext_w ? dx_max[s] - k.dx_min[s] : ext_w;
  }
  if (ext_h > 2 || ext_w > 2) return false;
  k.slots = 128 / d->Cin;
  k.n_src = d->n_src;
  // groups: per source, per dy, runs of `slots` horizontally adjacent taps
  int ng = 0;
  for (int s = 0; s < d->n_src; ++s) {
    k.part_base[s] = s;
    for (int dy = k.dy_min[s]; dy <= dy_max[s]; ++dy)
      for (int dx0 = k.dx_min[s]; dx0 <= dx_max[s]; dx0 += k.slots) {
        int ids[4] = {-1, -1, -1, -1};
        bool any = false;
        for (int j = 0; j < k.slots; ++j)
          for (int t = 0; t < d->n_taps; ++t)
            if (d->taps[t].src == s && d->taps[t].dy == dy && d->taps[t].dx == dx0 + j) {
              ids[j] = d->taps[t].wrow;
              any = true;
            }
        if (!any) continue;
        if (ng >= kWHaloGroupsMax) return false;
        k.grp_src[ng] = s;
        for (int j = 0; j < 4; ++j) k.grp_tap[ng][j] = ids[j];
        ++ng;
      }
  }
  if (ng * d->Cout > 512) return false;
  k.n_groups = ng;
  {
    // padding slots read past the window: the last slot of the last run of a row reads pixel
    // 7 + (run start) + slots - 1 relative to dx_min (runs start at multiples of `slots`)
    const int runs = (ext_w + 1 + k.slots - 1) / k.slots;
    const int need = 8 + runs * k.slots - 1;
    k.pitch = need > 8 + ext_w ? need : 8 + ext_w;
  }
  k.hal_h = 16 + ext_h;
  int gi = 0;
  for (int s = 0; s < d->n_src; ++s)
    for (int dy = k.dy_min[s]; dy <= dy_max[s]; ++dy)
      for (int dx0 = k.dx_min[s]; dx0 <= dx_max[s]; dx0 += k.slots) {
        if (gi < ng && k.grp_src[gi] == s) {
          bool match = false;
          for (int t = 0; t < d->n_taps && !match; ++t)
            match = d->taps[t].src == s && d->taps[t].dy == dy && d->taps[t].dx >= dx0 && d->taps[t].dx < dx0 + k.slots;
          if (!match) continue;
          k.grp_row0[gi] = (dy - k.dy_min[s]) * k.pitch + (dx0 - k.dx_min[s]);
          ++gi;
        }
      }
  if (gi != ng) return false;
  k.a_part_bytes = (static_cast<uint32_t>(k.hal_h * k.pitch * d->Cin * 2) + 1023u) & ~1023u;
  for (int g = 0; g < ng; ++g)
    k.grp_off16[g] = (static_cast<uint32_t>(k.grp_src[g]) * k.a_part_bytes +
                      static_cast<uint32_t>(k.grp_row0[g]) * static_cast<uint32_t>(d->Cin) * 2u) >> 4;
  k.g_bytes = (128u * d->Cout * 2u + 1023u) & ~1023u;
  k.stage_bytes = static_cast<uint32_t>(d->n_src) * k.a_part_bytes + k.g_bytes;
  int stages = static_cast<int>((215u * 1024u) / k.stage_bytes);
  if (stages < 2) return false;
  k.n_stages = stages > 8 ? 8 : stages;
  smem_bytes = static_cast<int>(static_cast<size_t>(k.n_stages) * k.stage_bytes + 1024 + 512);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(ng * d->Cout)) cols <<= 1;
  k.tmem_cols = cols;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Halo variant for the WIDE layers (C_in a multiple of 64, >= 128) at >= 16x8 output grids.  The generic kernel
// is bound by the TMA engine (one box row per ~4.5 cycles: 128 A rows + n_tile B rows per k-iteration against
// 4 MMAs of n_tile/2 cycles).  Here, per 64-channel chunk,
//   * the A brick + halo is ONE tensor load per source view (180 rows for all 9 taps instead of 9 x 128);
//   * the B tile of every tap is ONE linear bulk copy (cp.async.bulk) from weights pre-packed as the swizzled
//     shared-memory image (dfgan_pack_weights image=1) -- no per-row cost at all.
// Two rings: A (halo tiles, one per chunk) and B (weight tiles, one per chunk and tap).
struct Halo2K {
  int dy_min[DFGAN_MAX_SRC], dx_min[DFGAN_MAX_SRC];
  int n_src, pitch, hal_h, n_chunks;
  uint32_t a_part_bytes, a_stage_bytes, b_bytes;
  int n_a_stages, n_b_stages, n_acc;
  int pair;                 // 1: a CTA works on TWO pixel tiles per pipeline step; both reuse every weight tile it loads
  uint32_t a_tile_bytes;    // halo tiles of one pixel tile inside an A stage (a_stage_bytes = (1 + pair) * a_tile_bytes)
  uint32_t a_tap_off16[DFGAN_MAX_TAPS];
  long long w_tap_stride, w_chunk_stride;   // bytes between taps / 64-channel chunks of the packed image
  const uint8_t* wimg;
};

__global__ void __launch_bounds__(kHaloThreads, 1)
    conv_halo2_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ ConvK p,
                      const __grid_constant__ Halo2K hp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + static_cast<size_t>(hp.n_a_stages) * hp.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(hp.n_b_stages) * hp.b_bytes);
  uint64_t* fullA = bars;          // [4]
  uint64_t* emptyA = bars + 4;     // [4]
  uint64_t* fullB = bars + 8;      // [8]
  uint64_t* emptyB = bars + 16;    // [8]
  uint64_t* tfull = bars + 24;     // [4]
  uint64_t* tempty = bars + 28;    // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = hp.n_acc * p.n_tile <= 128 ? 128u : (hp.n_acc * p.n_tile <= 256 ? 256u : 512u);
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DFGAN_MAX_SRC; ++i) tma_prefetch_desc(&maps.a[i]);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&fullA[i], 1);
      mbar_init(&emptyA[i], 1);
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], p.col_split ? 8 : 4);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&fullB[i], 1);
      mbar_init(&emptyB[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel; from here on its results are needed.  The successor may start its own prologue right away.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    const bool leader = elect_one();
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t txa = static_cast<uint32_t>(hp.n_src) * hp.hal_h * hp.pitch * 128u;
    // pair mode: the CTA's tiles k and k+1 of its sequence (tile, tile + gridDim.x) form one pipeline step
    const int step = (hp.pair ? 2 : 1) * static_cast<int>(gridDim.x);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += step) {
      const int n_in_step = (hp.pair && tile + static_cast<int>(gridDim.x) < p.total_tiles) ? 2 : 1;
      const int nt = tile % p.n_tiles_n;
      const uint8_t* wbase = hp.wimg + static_cast<long long>(nt) * p.n_tile * 128;
      for (int c = 0; c < hp.n_chunks; ++c) {
        wait_bounded(&emptyA[sa], pa ^ 1);
        if (leader) mbar_expect_tx(&fullA[sa], txa * n_in_step);
        for (int j = 0; j < n_in_step; ++j) {
          int mt = (tile + j * static_cast<int>(gridDim.x)) / p.n_tiles_n;
          const int tw = mt % p.tiles_w;
          mt /= p.tiles_w;
          const int th = mt % p.tiles_h;
          const int n = mt / p.tiles_h;
          uint8_t* dst = sA + static_cast<size_t>(sa) * hp.a_stage_bytes + static_cast<size_t>(j) * hp.a_tile_bytes;
          for (int s = 0; s < hp.n_src; ++s) {
            if (leader)
              tma_load_4d(dst + static_cast<size_t>(s) * hp.a_part_bytes, &maps.a[s], &fullA[sa], c * 64,
                          (tw << 3) + hp.dx_min[s], (th << 4) + hp.dy_min[s], n);
          }
        }
        __syncwarp();
        if (++sa == hp.n_a_stages) sa = 0, pa ^= 1;
        for (int t = 0; t < p.n_taps; ++t) {
          wait_bounded(&emptyB[sb], pb ^ 1);
          if (leader) {
            mbar_expect_tx(&fullB[sb], hp.b_bytes);
            bulk_load(sB + static_cast<size_t>(sb) * hp.b_bytes,
                      wbase + static_cast<long long>(p.tap_wrow[t]) * hp.w_tap_stride + c * hp.w_chunk_stride,
                      hp.b_bytes, &fullB[sb]);
          }
          __syncwarp();
          if (++sb == hp.n_b_stages) sb = 0, pb ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    // The whole issue loop runs in the ONE elected lane (waits included): no per-tap warp synchronisation, and
    // every descriptor derives from per-stage bases with uniform adds (see conv_halo_kernel)
    if (leader) {
      const uint32_t sbo_a = static_cast<uint32_t>(hp.pitch) * 128u;
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      int sa = 0, sb = 0, acc = 0;
      uint32_t pa = 0, pb = 0, aphase = 0;
      const int step = (hp.pair ? 2 : 1) * static_cast<int>(gridDim.x);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += step) {
        const bool two = hp.pair && tile + static_cast<int>(gridDim.x) < p.total_tiles;
        // accumulator slots follow the epilogue's per-tile rotation: the second tile of a pair takes the next slot
        // (n_acc is even, so a pair never wraps between its two slots)
        wait_bounded(&tempty[acc], aphase ^ 1);
        if (two) wait_bounded(&tempty[acc + 1], aphase ^ 1);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
        const uint32_t d_tmem1 = d_tmem + static_cast<uint32_t>(p.n_tile);
        for (int c = 0; c < hp.n_chunks; ++c) {
          wait_bounded(&fullA[sa], pa);
          uint64_t ad0 = make_smem_desc(sA_u + static_cast<uint32_t>(sa) * hp.a_stage_bytes, 16, sbo_a, SWZ_128B);
          asm volatile("" : "+l"(ad0));
          const uint64_t ad1 = ad0 + (hp.a_tile_bytes >> 4);
          for (int t = 0; t < p.n_taps; ++t) {
            wait_bounded(&fullB[sb], pb);
            tc_fence_after();
            const uint64_t adesc = ad0 + hp.a_tap_off16[t];
            uint64_t bdesc = make_smem_desc(sB_u + static_cast<uint32_t>(sb) * hp.b_bytes, 16, 1024, SWZ_128B);
            asm volatile("" : "+l"(bdesc));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (c | t | k) != 0);
            if (two) {
              const uint64_t adesc1 = ad1 + hp.a_tap_off16[t];
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(d_tmem1, adesc1 + 2u * k, bdesc + 2u * k, p.idesc, (c | t | k) != 0);
            }
            umma_commit(&emptyB[sb]);
            if (++sb == hp.n_b_stages) sb = 0, pb ^= 1;
          }
          umma_commit(&emptyA[sa]);
          if (++sa == hp.n_a_stages) sa = 0, pa ^= 1;
        }
        umma_commit(&tfull[acc]);
        if (++acc == hp.n_acc) acc = 0, aphase ^= 1;
        if (two) {
          umma_commit(&tfull[acc]);
          if (++acc == hp.n_acc) acc = 0, aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
#define DFGAN_EPI_CALL_(V) brick_epilogue<V>(p, hp.n_acc, tfull, tempty, tmem_base, warp, lane)
    EPI_DISPATCH(p.epi)
#undef DFGAN_EPI_CALL_
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// Plans the wide-layer halo launch (needs the image-packed weights); false = use the generic kernel.
static bool plan_halo2(const dfgan_conv_desc* d, Halo2K& hp, int& n_tile, int& smem_bytes) {
  if (!d->wpack_img || d->Cin % 64 || d->Cin < 128 || d->Cout % 16) return false;
  if (d->W < 8 || d->H < 16) return false;
  int dy_max[DFGAN_MAX_SRC], dx_max[DFGAN_MAX_SRC];
  bool used[DFGAN_MAX_SRC] = {false, false, false, false};
  for (int s = 0; s < DFGAN_MAX_SRC; ++s) hp.dy_min[s] = hp.dx_min[s] = dy_max[s] = dx_max[s] = 0;
  for (int t = 0; t < d->n_taps; ++t) {
    const int s = d->taps[t].src;
    if (s < 0 || s >= d->n_src) return false;
    const int dy = d->taps[t].dy, dx = d->taps[t].dx;
    if (!used[s]) {
      used[s] = true;
      hp.dy_min[s] = dy_max[s] = dy;
      hp.dx_min[s] = dx_max[s] = dx;
    } else {
      hp.dy_min[s] = dy < hp.dy_min[s] ? dy : hp.dy_min[s];
      dy_max[s] = dy > dy_max[s] ? dy : dy_max[s];
      hp.dx_min[s] = dx < hp.dx_min[s] ? dx : hp.dx_min[s];
      dx_max[s] = dx > dx_max[s] ? dx : dx_max[s];
    }
  }
  int ext_h = 0, ext_w = 0;
  for (int s = 0; s < d->n_src; ++s) {
    if (!used[s]) return false;
    ext_h = dy_max[s] - hp.dy_min[s] > ext_h ? dy_max[s] - hp.dy_min[s] : ext_h;
    ext_w = dx_max[s] - hp.dx_min[s] > ext_w ? dx_max[s] - hp.dx_min[s] : ext_w;
  }
  if (ext_h > 2 || ext_w > 2) return false;
  n_tile = d->Cout <= 256 ? d->Cout : 256;
  if (d->Cout % n_tile || n_tile % 8) return false;
  hp.n_src = d->n_src, hp.n_chunks = d->Cin / 64;
  hp.pitch = 8 + ext_w, hp.hal_h = 16 + ext_h;
  hp.a_part_bytes = (static_cast<uint32_t>(hp.hal_h * hp.pitch * 128) + 1023u) & ~1023u;
  hp.a_tile_bytes = hp.a_part_bytes * static_cast<uint32_t>(d->n_src);
  hp.b_bytes = static_cast<uint32_t>(n_tile) * 128u;
  hp.n_acc = n_tile <= 128 ? 4 : 2;
  const size_t cap = 220 * 1024, fixed = 1024 + 512;
  // Pair mode: per 128-pixel tile a CTA streams taps * chunks weight tiles of n_tile x 64 (1.2 MB for a 3x3 C256
  // layer against 92 KB of halo tiles), which makes the wide layers L2 -> shared-memory bandwidth bound (~6 TB/s,
  // measured 730 TFLOP/s).  Two pixel tiles per pipeline step reuse every weight tile: half the weight traffic.
  // Used when every CTA has at least two tiles anyway and the doubled halo stage still leaves 2 + 3 stages.
  static const bool pair_off = getenv("DFGAN_NO_PAIR") != nullptr;
  const long long tiles = static_cast<long long>((d->W + 7) / 8) * ((d->H + 15) / 16) * d->N * (d->Cout / n_tile);
  // (only with four accumulator slots, n_tile <= 128: with two slots a pair occupies both and the epilogue no longer
  // overlaps the next pair's MMAs -- measured 1166 -> 1125 TFLOP/s on 3x3 C256, against 917 -> 1049 on C128)
  hp.pair = (!pair_off && d->Cout == n_tile && n_tile <= 128 && tiles >= num_sms() &&
             fixed + 2 * (2 * hp.a_tile_bytes) + 3 * hp.b_bytes <= cap) ? 1 : 0;
  hp.a_stage_bytes = hp.a_tile_bytes * (hp.pair ? 2u : 1u);
  // at least 2 halo stages and 3 weight stages; prefer a third halo stage, then as many weight stages as fit
  if (fixed + 2 * hp.a_stage_bytes + 3 * hp.b_bytes > cap) return false;
  hp.n_a_stages = (fixed + 3 * hp.a_stage_bytes + 4 * hp.b_bytes <= cap) ? 3 : 2;
  int nb = static_cast<int>((cap - fixed - hp.n_a_stages * hp.a_stage_bytes) / hp.b_bytes);
  hp.n_b_stages = nb > 8 ? 8 : nb;
  smem_bytes = static_cast<int>(fixed + hp.n_a_stages * hp.a_stage_bytes + hp.n_b_stages * hp.b_bytes);
  for (int t = 0; t < d->n_taps; ++t) {
    const int s = d->taps[t].src;
    const uint32_t row0 = static_cast<uint32_t>((d->taps[t].dy - hp.dy_min[s]) * hp.pitch + (d->taps[t].dx - hp.dx_min[s]));
    hp.a_tap_off16[t] = (static_cast<uint32_t>(s) * hp.a_part_bytes + row0 * 128u) >> 4;
  }
  // image layout: [slot][chunk][w_rows_per_slot][64]; a tap's wrow is slot * rows_per_slot (rows, not bytes)
  hp.w_chunk_stride = static_cast<long long>(d->img_rows_per_slot) * 128;
  hp.w_tap_stride = static_cast<long long>(hp.n_chunks) * 128;   // per ROW of wrow: slot stride / rows_per_slot
  hp.wimg = static_cast<const uint8_t*>(d->wpack_img);
  return true;
}

static void brick_geometry(int W, int H, int& lbw, int& lbh) {
  int bw = 1;
  while (bw < W && bw < 16) bw <<= 1;
  int bh = 1;
  while (bh < H && bw * bh < 128) bh <<= 1;
  lbw = ilog2(bw);
  lbh = ilog2(bh);
}

static uint32_t pow2_cols(int c) {
  uint32_t v = 32;
  while (v < static_cast<uint32_t>(c)) v <<= 1;
  return v;
}

}  // namespace dfgan

using namespace dfgan;

// feature set of a launch's fused epilogue (bit mask, see EPI_* in the kernel section)
static uint32_t epi_flags_of(const dfgan_conv_desc* d) {
  uint32_t f = 0;
  if (d->bias) f |= 1u;
  if (d->lrelu_pre) f |= 2u;
  if (d->mask_src) f |= 4u;
  if (d->aux) f |= 8u;
  if (d->addend) f |= 16u;
  if (d->out) f |= 32u;
  if (d->post_mode == 1) f |= 64u;
  if (d->post_mode == 2) f |= 128u;
  if (d->out_nchw) f |= 256u;
  if (d->phase_cout) f |= 512u;
  return f;
}
static void dbg_print_launch(const char* family, int cin, int geom, const dfgan_conv_desc* d, int n_tile) {
  static const bool on = getenv("DFGAN_DBG") && (atoi(getenv("DFGAN_DBG")) & 64);
  if (on)
    fprintf(stderr, "launch %s cin %d geom %d epi %u | N%d %dx%d Cout%d taps%d n_tile %d add_shift %d tanh %d\n", family, cin,
            geom, epi_flags_of(d), d->N, d->H, d->W, d->Cout, d->n_taps, n_tile, d->add_shift, d->tanh_out);
}
// per-launch label for the profile dump (host_common: DFGAN_PROFILE_DUMP)
struct LaunchLabel {
  char s[120];
  LaunchLabel(const char* family, int variant, const dfgan_conv_desc* d, int n_tile, int grid) {
    snprintf(s, sizeof(s), "%s/%d N%d %dx%d Cin%d Cout%d taps%d src%d epi%u ntile%d grid%d", family, variant, d->N, d->H,
             d->W, d->Cin, d->Cout, d->n_taps, d->n_src, epi_flags_of(d), n_tile, grid);
  }
  LaunchLabel(const char* family, const dfgan_wgrad_desc* d, int ks, int grid) {
    snprintf(s, sizeof(s), "%s N%d %dx%d Cin%d Cout%d taps%d src%d ks%d grid%d", family, d->g.N, d->g.H, d->g.W, d->Cin,
             d->Cout, d->n_taps, d->n_src, ks, grid);
  }
};
static void fill_epilogue(ConvK& k, const dfgan_conv_desc* d) {
  k.out = static_cast<__nv_bfloat16*>(d->out);
  k.out_sw = d->out_sw, k.out_sh = d->out_sh, k.out_sn = d->out_sn;
  k.bias = d->bias, k.scale_dev = d->scale_dev, k.scale_imm = d->scale_imm;
  k.lrelu_pre = d->lrelu_pre, k.slope = d->slope;
  k.aux = static_cast<__nv_bfloat16*>(d->aux);
  k.addend = static_cast<const __nv_bfloat16*>(d->addend);
  k.add_sw = d->add_sw, k.add_sh = d->add_sh, k.add_sn = d->add_sn;
  k.add_shift = d->add_shift, k.add_scale = d->add_scale;
  k.mask_src = static_cast<const __nv_bfloat16*>(d->mask_src);
  k.out2 = static_cast<__nv_bfloat16*>(d->out2);
  k.post_mode = d->post_mode;
  k.tab_ga = d->tab_ga, k.tab_ba = d->tab_ba, k.tab_gb = d->tab_gb, k.tab_bb = d->tab_bb;
  k.tab_stride = d->tab_stride;
  k.out_nchw = d->out_nchw, k.co_valid = d->co_valid, k.tanh_out = d->tanh_out;
  k.phase_cout = d->phase_cout;
  static const bool epi_generic = getenv("DFGAN_EPI_GENERIC") != nullptr;   // A/B switch: run-time feature tests
  const uint32_t f = epi_flags_of(d);
  k.epi = (!epi_generic && epi_listed(f)) ? f : EPI_GENERIC;
}
// after tiles_w / tiles_h are known: shift / mask tile decoding when both are powers of two
static void fill_tile_logs(ConvK& k) {
  const bool pw = (k.tiles_w & (k.tiles_w - 1)) == 0, ph = (k.tiles_h & (k.tiles_h - 1)) == 0;
  k.ltw = (pw && ph) ? ilog2(k.tiles_w) : -1;
  k.lth = (pw && ph) ? ilog2(k.tiles_h) : -1;
}

// executed work of a launch: real output pixels x Cout x (taps * Cin) MACs; bytes = every source view read once
// + weights + the outputs written (the algorithmic minimum for this launch)
static double launch_bytes(const dfgan_conv_desc* d) {
  const double pix = static_cast<double>(d->N) * d->H * d->W;
  double bytes = 2.0 * d->w_rows * d->Cin;
  for (int s = 0; s < d->n_src; ++s)
    bytes += 2.0 * d->src[s].N * d->src[s].H * d->src[s].W * d->src[s].C;
  bytes += pix * d->Cout * 2.0 * ((d->out != nullptr) + (d->aux != nullptr) + (d->out2 != nullptr)) +
           (d->addend ? pix * d->Cout * 2.0 / (1 << (2 * d->add_shift)) : 0.0) +
           (d->mask_src ? pix * d->Cout * 2.0 : 0.0) + (d->out_nchw ? pix * d->co_valid * 4.0 : 0.0);
  return bytes;
}

// Generic-kernel tiling: brick geometry, N tile, tile count and the number of K splits an under-filled launch wants.
struct GenericPlan {
  int lbw, lbh, n_tile, tiles_w, tiles_h, tiles_n, n_tiles_n, total_tiles, k_iters, k_splits;
  int cluster;    // > 1: on-chip split-K over a thread-block cluster of that many CTAs
  int b_image;    // weight tiles come from the swizzled-image copy of the operand (one bulk copy per k-iteration)
};
// environment switches, read once per process (the same cached value everywhere they are consulted)
static bool env_off(const char* name) { return getenv(name) != nullptr && atoi(getenv(name)) != 0; }
static bool halo_disabled() { static const bool v = env_off("DFGAN_NO_HALO"); return v; }
static bool halo2_disabled() { static const bool v = env_off("DFGAN_NO_HALO2"); return v; }
// (measured round 2, tools/prof_lowres.py: the bulk-copied weight tile is 5-10 % SLOWER than the tiled load in the
// generic kernel -- these launches are bound by the ~60 B/cycle one SM can pull in, not by TMA box rows -- so the
// image path of the generic kernel is opt-in: DFGAN_BIMG=1)
static bool bimg_disabled() { static const bool v = !env_off("DFGAN_BIMG"); return v; }
static bool cluster_disabled() { static const bool v = env_off("DFGAN_NO_CLUSTER"); return v; }
static bool generic_b_image(const dfgan_conv_desc* d) {
  return !bimg_disabled() && d->wpack_img && d->Cin % 64 == 0 && d->img_rows_per_slot > 0;
}
static bool plan_generic(const dfgan_conv_desc* d, GenericPlan& g) {
  brick_geometry(d->W, d->H, g.lbw, g.lbh);
  const int BW = 1 << g.lbw, BH = 1 << g.lbh, BN = 128 / (BW * BH);
  int n_tile = d->n_tile_hint;
  if (n_tile <= 0) {
    // widest N tile that still gives every SM a tile: small grids (low resolutions) trade A-tile re-reads
    // from L2 for occupancy of the 148 SMs
    const long long m_tiles = static_cast<long long>((d->W + BW - 1) / BW) * ((d->H + BH - 1) / BH) *
                              ((d->N + BN - 1) / BN);
    n_tile = d->Cout <= 256 ? d->Cout : 256;
    // (halving N doubles the A-tile traffic and halves the tensor-pipe efficiency of each MMA, so only do it
    // while fewer than half of the SMs would have work)
    static const int n_tile_min = getenv("DFGAN_NTILE_MIN") ? atoi(getenv("DFGAN_NTILE_MIN")) : 64;
    while (n_tile > n_tile_min && n_tile % 32 == 0 && d->Cout % (n_tile / 2) == 0 &&
           2 * m_tiles * ((d->Cout + n_tile - 1) / n_tile) < num_sms())
      n_tile /= 2;
  }
  const int BK0 = (d->Cin % 64 == 0) ? 64 : 32;
  const int k_iters0 = d->n_taps * (d->Cin / BK0);
  g.cluster = 1;
  g.b_image = generic_b_image(d) ? 1 : 0;
  if (!cluster_disabled() && d->n_tile_hint == 0 && !d->out_nchw && k_iters0 >= 8) {
    // Low-resolution launches: a handful of output tiles with K loops of 16..128 iterations, each bound by what ONE SM
    // can pull in (TMA: one box row per ~4.5 cycles; L2 -> SM ~64 B/cycle).  A cluster of S CTAs per tile splits the K
    // loop S ways and reduces on chip.  N tile <= 128 keeps the fp32 staging area at 64 KB.
    const long long m_tiles = static_cast<long long>((d->W + BW - 1) / BW) * ((d->H + BH - 1) / BH) *
                              ((d->N + BN - 1) / BN);
    int nt_c = d->Cout <= 128 ? d->Cout : 128;
    if (d->Cout % nt_c == 0 && nt_c % 16 == 0) {
      const long long tiles_c = m_tiles * (d->Cout / nt_c);
      static const int sp_max = getenv("DFGAN_CLUSTER_MAX") ? atoi(getenv("DFGAN_CLUSTER_MAX")) : 4;
      int sp = sp_max;
      while (sp > 1 && (tiles_c * sp > num_sms() || k_iters0 / sp < 2 || (nt_c / sp) % 16 != 0)) sp >>= 1;
      if (sp > 1) {
        n_tile = nt_c;
        g.cluster = sp;
      }
    }
  }
  if (n_tile > 256 || n_tile % 16 || n_tile > d->Cout) return false;
  g.n_tile = n_tile;
  g.tiles_w = (d->W + BW - 1) / BW;
  g.tiles_h = (d->H + BH - 1) / BH;
  g.tiles_n = (d->N + BN - 1) / BN;
  g.n_tiles_n = (d->Cout + n_tile - 1) / n_tile;
  g.total_tiles = g.tiles_w * g.tiles_h * g.tiles_n * g.n_tiles_n;
  const int BK = (d->Cin % 64 == 0) ? 64 : 32;
  g.k_iters = d->n_taps * (d->Cin / BK);
  // split-K: an under-filled launch (fewer tiles than half the SMs) with a long K loop is bound by the load
  // bandwidth of the few SMs it runs on (~60 B/cycle each); slices of the (tap, chunk) loop spread it over all SMs
  // Measured on the training iteration: 12.02 -> 11.98 ms (the second pass and the atomics eat most of the gain, and
  // these launches already overlap with the weight gradients on side streams), so it is opt-in: DFGAN_SPLITK=1.
  static const bool splitk_off = getenv("DFGAN_SPLITK") == nullptr || atoi(getenv("DFGAN_SPLITK")) == 0;
  g.k_splits = 1;
  if (g.cluster > 1) {
    g.k_splits = g.cluster;
    return true;
  }
  if (!splitk_off && d->n_tile_hint == 0 && !d->out_nchw && 2 * g.total_tiles <= num_sms() && g.k_iters >= 16) {
    int sp = num_sms() / g.total_tiles;
    if (sp > g.k_iters / 8) sp = g.k_iters / 8;
    if (sp > 16) sp = 16;
    if (sp >= 2) g.k_splits = sp;
  }
  return true;
}
static bool takes_halo_path(const dfgan_conv_desc* d) {
  dfgan_conv_desc ds = *d;   // plan_halo is order-independent in its accept / reject decision
  HaloK hp{};
  int smem = 0;
  if (!halo_disabled() && d->n_tile_hint == 0 && plan_halo(&ds, hp, smem)) return true;
  Halo2K h2{};
  int nt2 = 0;
  return !halo2_disabled() && d->n_tile_hint == 0 && plan_halo2(d, h2, nt2, smem);
}
extern "C" int64_t dfgan_conv_splitk_bytes(const dfgan_conv_desc* d) {
  if (!d || d->Cin % 32 || d->Cout % 16 || d->N < 1 || d->H < 1 || d->W < 1 || d->n_taps < 1) return 0;
  if (takes_halo_path(d)) return 0;
  GenericPlan g{};
  if (!plan_generic(d, g) || g.k_splits < 2 || g.cluster > 1) return 0;   // cluster split-K needs no workspace
  return static_cast<int64_t>(d->N) * d->H * d->W * d->Cout * 4;
}

// 1 when dfgan_conv_run would take the wide-layer halo kernel (reads wpack_img), 0 when it reads wpack.
extern "C" int dfgan_conv_uses_image(const dfgan_conv_desc* d) {
  if (!d) return 0;
  // the narrow-layer halo kernel (C_in <= 64) never has an image; the wide-layer halo kernel and the generic kernel
  // (weight tiles by bulk copy) both read it whenever one is offered -- the same cached switches as dfgan_conv_run
  {
    dfgan_conv_desc ds = *d;
    HaloK hp{};
    int smem = 0;
    if (!halo_disabled() && d->n_tile_hint == 0 && plan_halo(&ds, hp, smem)) return 0;
  }
  Halo2K h2{};
  int smem2 = 0, nt2 = 0;
  if (!halo2_disabled() && d->n_tile_hint == 0 && plan_halo2(d, h2, nt2, smem2)) return 1;
  return generic_b_image(d) ? 1 : 0;
}

extern "C" int dfgan_conv_run(const dfgan_conv_desc* d, dfgan_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !d->wpack) return DFGAN_E_ARG;
  if (!d->out && !d->aux && !d->out2 && !d->out_nchw) return DFGAN_E_ARG;
  if (d->post_mode < 0 || d->post_mode > 2 || (d->post_mode && !d->out2)) return DFGAN_E_ARG;
  if (d->post_mode == 2 && !(d->tab_ga && d->tab_ba && d->tab_gb && d->tab_bb)) return DFGAN_E_ARG;
  // 256-bit accesses: table rows and every bf16 tensor the epilogue touches must be 32-byte aligned per 16 columns
  if (d->post_mode == 2 && (((reinterpret_cast<uintptr_t>(d->tab_ga) | reinterpret_cast<uintptr_t>(d->tab_ba) |
                              reinterpret_cast<uintptr_t>(d->tab_gb) | reinterpret_cast<uintptr_t>(d->tab_bb)) & 31) ||
                            (d->tab_stride & 7)))
    return DFGAN_E_ARG;
  if (((reinterpret_cast<uintptr_t>(d->out) | reinterpret_cast<uintptr_t>(d->aux) | reinterpret_cast<uintptr_t>(d->out2) |
        reinterpret_cast<uintptr_t>(d->mask_src) | reinterpret_cast<uintptr_t>(d->addend)) & 31) ||
      ((d->out_sw | d->out_sh | d->out_sn | d->add_sw | d->add_sh | d->add_sn) & 15))
    return DFGAN_E_ARG;
  if (d->out_nchw && (d->co_valid < 1 || d->co_valid > d->Cout)) return DFGAN_E_ARG;
  if (d->phase_cout && (d->phase_cout % 16 || d->Cout != 4 * d->phase_cout || d->out_nchw)) return DFGAN_E_ARG;
  if (d->n_src < 1 || d->n_src > DFGAN_MAX_SRC || d->n_taps < 1 || d->n_taps > DFGAN_MAX_TAPS)
    return DFGAN_E_ARG;
  if (d->Cin % 32 || d->Cout % 16 || d->N < 1 || d->H < 1 || d->W < 1) return DFGAN_E_SHAPE;
  const int BK = (d->Cin % 64 == 0) ? 64 : 32;
  {
    HaloK hp{};
    int smem_halo = 0;
    const bool halo_off = halo_disabled();
    // taps sorted by (source, dy, dx): the sum does not depend on the order, and the standard geometries then have
    // compile-time descriptor offsets (geom_off16)
    dfgan_conv_desc ds = *d;
    for (int i = 1; i < ds.n_taps && i < DFGAN_MAX_TAPS; ++i)
      for (int j = i; j > 0; --j) {
        const dfgan_tap &a = ds.taps[j - 1], &b = ds.taps[j];
        const bool gt = a.src != b.src ? a.src > b.src : (a.dy != b.dy ? a.dy > b.dy : a.dx > b.dx);
        if (!gt) break;
        const dfgan_tap tmp = ds.taps[j - 1];
        ds.taps[j - 1] = ds.taps[j];
        ds.taps[j] = tmp;
      }
    if (!halo_off && d->n_tile_hint == 0 && plan_halo(&ds, hp, smem_halo)) {
      d = &ds;
      ConvK k{};
      ConvMaps maps{};
      k.lbw = 3, k.lbh = 4;
      k.tiles_w = (d->W + 7) / 8, k.tiles_h = (d->H + 15) / 16, k.tiles_n = d->N;
      k.n_tiles_n = 1, k.total_tiles = k.tiles_w * k.tiles_h * k.tiles_n;
      k.n_tile = d->Cout;
      k.N = d->N, k.H = d->H, k.W = d->W, k.Cout = d->Cout;
      k.n_taps = d->n_taps, k.n_chunks = 1;
      for (int t = 0; t < d->n_taps; ++t) {
        k.tap_src[t] = d->taps[t].src, k.tap_dy[t] = d->taps[t].dy, k.tap_dx[t] = d->taps[t].dx;
        k.tap_wrow[t] = d->taps[t].wrow;
      }
      k.idesc = make_idesc_bf16(128, k.n_tile, 0, 0);
      fill_epilogue(k, d);
      fill_tile_logs(k);
      for (int s = 0; s < DFGAN_MAX_SRC; ++s) {
        const dfgan_view4& v = d->src[s < d->n_src ? s : 0];
        if (v.C != d->Cin) return DFGAN_E_SHAPE;
        int rc = make_map_nhwc(&maps.a[s], v, d->Cin, hp.pitch, hp.hal_h, 1);
        if (rc) return rc;
      }
      int rc = make_map_2d(&maps.b, d->wpack, d->w_rows, d->Cin, d->Cin, k.n_tile);
      if (rc) return rc;
      int grid = k.total_tiles;
      const int cap = d->max_ctas > 0 ? d->max_ctas : num_sms();
      if (grid > cap) grid = cap;
      k.col_split = (k.total_tiles <= 3 * grid && k.n_tile >= 32) ? 1 : 0;
      k.dbg = getenv("DFGAN_DBG") ? atoi(getenv("DFGAN_DBG")) : 0;
      // standard geometry?  (host-computed offsets must equal the compile-time table)
      int geom = GEOM_RUNTIME;
      static const bool geom_off = getenv("DFGAN_NO_GEOM") != nullptr;
      for (int g = GEOM_1; g <= GEOM_2X2 && !geom_off; ++g) {
        if (geom_taps(g) != d->n_taps) continue;
        bool same = true;
        for (int t = 0; t < d->n_taps; ++t) same = same && hp.a_tap_off16[t] == geom_off16(g, d->Cin, t);
        if (same) {
          geom = g;
          break;
        }
      }
      const double pix = static_cast<double>(d->N) * d->H * d->W;
      const LaunchLabel lab("halo", geom, d, k.n_tile, grid);
      ProfileScope prof(0, 2.0 * pix * d->Cout * d->n_taps * d->Cin * (d->phase_cout ? 16.0 / 36.0 : 1.0),
                        launch_bytes(d), stream, lab.s);
#define DFGAN_LAUNCH_HALO(CI, GEOM)                                                                       \
  do {                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<CI, GEOM>,                                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_halo);         \
    if (e != cudaSuccess) return static_cast<int>(e);                                                     \
    e = launch_pdl(conv_halo_kernel<CI, GEOM>, grid, kHaloNarrowThreads, smem_halo, stream, 1, maps, k, hp);       \
    if (e != cudaSuccess) return static_cast<int>(e);                                                     \
  } while (0)
#define DFGAN_LAUNCH_HALO_C(CI)                                                                           \
  do {                                                                                                    \
    switch (geom) {                                                                                       \
      case GEOM_1: DFGAN_LAUNCH_HALO(CI, GEOM_1); break;                                                  \
      case GEOM_3X3: DFGAN_LAUNCH_HALO(CI, GEOM_3X3); break;                                              \
      case GEOM_2X2X4: DFGAN_LAUNCH_HALO(CI, GEOM_2X2X4); break;                                          \
      case GEOM_POOL4: DFGAN_LAUNCH_HALO(CI, GEOM_POOL4); break;                                          \
      case GEOM_2X2: DFGAN_LAUNCH_HALO(CI, GEOM_2X2); break;                                              \
      default: DFGAN_LAUNCH_HALO(CI, GEOM_RUNTIME); break;                                                \
    }                                                                                                     \
  } while (0)
      dbg_print_launch("halo", d->Cin, geom, d, k.n_tile);
      if (d->Cin == 64) DFGAN_LAUNCH_HALO_C(64);
      else DFGAN_LAUNCH_HALO_C(32);
#undef DFGAN_LAUNCH_HALO_C
#undef DFGAN_LAUNCH_HALO
      count_launch();
      DFGAN_CUDA_CHECK_LAUNCH();
      return DFGAN_OK;
    }
  }

  {
    Halo2K h2{};
    int smem2 = 0, nt2 = 0;
    const bool halo2_off = halo2_disabled();
    if (!halo2_off && d->n_tile_hint == 0 && plan_halo2(d, h2, nt2, smem2)) {
      ConvK k{};
      ConvMaps maps{};
      k.lbw = 3, k.lbh = 4;
      k.tiles_w = (d->W + 7) / 8, k.tiles_h = (d->H + 15) / 16, k.tiles_n = d->N;
      k.n_tile = nt2;
      k.n_tiles_n = d->Cout / nt2;
      k.total_tiles = k.tiles_w * k.tiles_h * k.tiles_n * k.n_tiles_n;
      k.N = d->N, k.H = d->H, k.W = d->W, k.Cout = d->Cout;
      k.n_taps = d->n_taps, k.n_chunks = h2.n_chunks;
      for (int t = 0; t < d->n_taps; ++t) {
        k.tap_src[t] = d->taps[t].src, k.tap_dy[t] = d->taps[t].dy, k.tap_dx[t] = d->taps[t].dx;
        k.tap_wrow[t] = d->taps[t].wrow;
      }
      k.idesc = make_idesc_bf16(128, k.n_tile, 0, 0);
      fill_epilogue(k, d);
      fill_tile_logs(k);
      for (int s = 0; s < DFGAN_MAX_SRC; ++s) {
        const dfgan_view4& v = d->src[s < d->n_src ? s : 0];
        if (v.C != d->Cin) return DFGAN_E_SHAPE;
        int rc = make_map_nhwc(&maps.a[s], v, 64, h2.pitch, h2.hal_h, 1);
        if (rc) return rc;
      }
      int grid = h2.pair ? (k.total_tiles + 1) / 2 : k.total_tiles;
      const int cap = d->max_ctas > 0 ? d->max_ctas : num_sms();
      if (grid > cap) grid = cap;
      k.col_split = (k.total_tiles <= 3 * grid) ? 1 : 0;
      const double pix = static_cast<double>(d->N) * d->H * d->W;
      const LaunchLabel lab("halo2", h2.pair, d, k.n_tile, grid);
      ProfileScope prof(0, 2.0 * pix * d->Cout * d->n_taps * d->Cin * (d->phase_cout ? 16.0 / 36.0 : 1.0),
                        launch_bytes(d), stream, lab.s);
      dbg_print_launch("halo2", d->Cin, h2.pair, d, k.n_tile);
      cudaError_t e = cudaFuncSetAttribute(conv_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
      if (e != cudaSuccess) return static_cast<int>(e);
      e = launch_pdl(conv_halo2_kernel, grid, kHaloThreads, smem2, stream, 1, maps, k, h2);
      if (e != cudaSuccess) return static_cast<int>(e);
      count_launch();
      DFGAN_CUDA_CHECK_LAUNCH();
      return DFGAN_OK;
    }
  }

  ConvK k{};
  ConvMaps maps{};
  GenericPlan gp{};
  if (!plan_generic(d, gp)) return DFGAN_E_SHAPE;
  k.lbw = gp.lbw, k.lbh = gp.lbh;
  const int BW = 1 << k.lbw, BH = 1 << k.lbh, BN = 128 / (BW * BH);
  const int n_tile = gp.n_tile;
  k.tiles_w = gp.tiles_w, k.tiles_h = gp.tiles_h, k.tiles_n = gp.tiles_n;
  k.n_tiles_n = gp.n_tiles_n, k.total_tiles = gp.total_tiles;
  k.n_tile = n_tile;
  k.k_splits = 1;
  k.dbg = getenv("DFGAN_DBG") ? atoi(getenv("DFGAN_DBG")) : 0;
  k.cluster = gp.cluster;
  if (gp.cluster > 1) {
    k.k_splits = gp.cluster;
  } else if (gp.k_splits > 1 && d->workspace &&
             d->ws_bytes >= static_cast<int64_t>(d->N) * d->H * d->W * d->Cout * 4) {
    k.k_splits = gp.k_splits;
    k.ws = static_cast<float*>(d->workspace);
  }
  k.N = d->N, k.H = d->H, k.W = d->W, k.Cout = d->Cout;
  k.n_taps = d->n_taps;
  k.n_chunks = d->Cin / BK;
  for (int t = 0; t < d->n_taps; ++t) {
    if (d->taps[t].src < 0 || d->taps[t].src >= d->n_src) return DFGAN_E_ARG;
    k.tap_src[t] = d->taps[t].src;
    k.tap_dy[t] = d->taps[t].dy;
    k.tap_dx[t] = d->taps[t].dx;
    k.tap_wrow[t] = d->taps[t].wrow;
  }
  k.idesc = make_idesc_bf16(128, n_tile, 0, 0);
  k.a_bytes = 128u * BK * 2u;
  k.b_bytes = static_cast<uint32_t>(n_tile) * BK * 2u;
  // cluster split-K: [cluster][128][n_tile / cluster] fp32 staging behind the pipeline stages and barriers
  const size_t stg_bytes = k.cluster > 1 ? static_cast<size_t>(128) * n_tile * 4 : 0;
  int stages = static_cast<int>((kSmemBudget - stg_bytes) / (k.a_bytes + k.b_bytes));
  if (stages > 8) stages = 8;
  if (k.cluster > 1) {
    const int need = (gp.k_iters + k.cluster - 1) / k.cluster;   // a slice never has more iterations in flight
    if (stages > need) stages = need < 2 ? 2 : need;
  }
  if (stages < 2) return DFGAN_E_SHAPE;
  k.n_stages = stages;
  k.tmem_cols = pow2_cols(2 * n_tile);
  fill_epilogue(k, d);
  if (gp.b_image && BK == 64) {
    // image layout [slot][chunk][rows_per_slot][64] (dfgan_pack_weights image=1); a tap's wrow counts rows
    k.b_img = static_cast<const uint8_t*>(d->wpack_img);
    k.w_chunk_stride = static_cast<long long>(d->img_rows_per_slot) * 128;
    k.w_tap_stride = static_cast<long long>(d->Cin / 64) * 128;
  }

  for (int s = 0; s < DFGAN_MAX_SRC; ++s) {
    const dfgan_view4& v = d->src[s < d->n_src ? s : 0];
    if (v.C != d->Cin) return DFGAN_E_SHAPE;
    int rc = make_map_nhwc(&maps.a[s], v, BK, BW, BH, BN);
    if (rc) return rc;
  }
  int rc = make_map_2d(&maps.b, d->wpack, d->w_rows, d->Cin, BK, n_tile);
  if (rc) return rc;

  // stages | barriers + TMEM slot (256 B) | staging (16-byte aligned) ; +1024 for the in-kernel alignment
  k.stg_off = static_cast<uint32_t>(static_cast<size_t>(stages) * (k.a_bytes + k.b_bytes) + 256);
  const size_t smem = static_cast<size_t>(stages) * (k.a_bytes + k.b_bytes) + 1024 + 256 + stg_bytes;
  int grid = k.total_tiles * k.k_splits;
  const int cap = d->max_ctas > 0 ? d->max_ctas : num_sms();
  if (grid > cap && k.cluster <= 1) grid = cap;       // a cluster launch is one CTA per work item
  if (k.dbg & 32)
    fprintf(stderr, "generic: N%d %dx%d Cin%d Cout%d taps%d | n_tile %d tiles %d k_iters %d k_splits %d cluster %d stages %d "
            "a %u b %u stg_off %u smem %zu grid %d bimg %p\n", d->N, d->H, d->W, d->Cin, d->Cout, d->n_taps, n_tile,
            k.total_tiles, gp.k_iters, k.k_splits, k.cluster, stages, k.a_bytes, k.b_bytes, k.stg_off, smem, grid,
            (const void*)k.b_img);
  const double pix = static_cast<double>(d->N) * d->H * d->W;
  const LaunchLabel lab("generic", k.cluster, d, n_tile, grid);
  ProfileScope prof(0, 2.0 * pix * d->Cout * d->n_taps * d->Cin * (d->phase_cout ? 16.0 / 36.0 : 1.0),
                    launch_bytes(d), stream, lab.s);
  dbg_print_launch("generic", d->Cin, k.cluster, d, n_tile);
  cudaError_t e;
  if (BK == 64) {
    e = cudaFuncSetAttribute(conv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    e = launch_pdl(conv_umma_kernel<64>, grid, kConvThreads, smem, stream, k.cluster, maps, k);
    if (e != cudaSuccess) return static_cast<int>(e);
  } else {
    e = cudaFuncSetAttribute(conv_umma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    e = launch_pdl(conv_umma_kernel<32>, grid, kConvThreads, smem, stream, k.cluster, maps, k);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  count_launch();
  if (k.k_splits > 1 && k.cluster <= 1) {
    const long long total = static_cast<long long>(d->N) * d->H * d->W * (d->Cout / 16);
    conv_splitk_epilogue_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream>>>(k);
    count_launch();
  }
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

extern "C" int dfgan_wgrad_run(const dfgan_wgrad_desc* d, dfgan_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !d->dw || !d->g.ptr) return DFGAN_E_ARG;
  if (d->n_src < 1 || d->n_src > DFGAN_MAX_SRC || d->n_taps < 1 || d->n_taps > DFGAN_MAX_TAPS)
    return DFGAN_E_ARG;
  if (d->Cin % 32 || d->Cout % 32 || d->g.C != d->Cout) return DFGAN_E_SHAPE;

  {
    WHaloK hk{};
    int smem_halo = 0;
    static const bool whalo_off = getenv("DFGAN_NO_WGRAD_HALO") != nullptr;
    if (!whalo_off && d->k_splits == 0 && plan_wgrad_halo(d, hk, smem_halo)) {
      WgradMaps maps{};
      const int W = d->g.W, H = d->g.H, N = d->g.N;
      hk.tiles_w = (W + 7) / 8, hk.tiles_h = (H + 15) / 16;
      hk.pix_tiles = hk.tiles_w * hk.tiles_h * N;
      hk.Cin = d->Cin, hk.Cout = d->Cout;
      hk.idesc = make_idesc_bf16(128, d->Cout, 1, 1);
      hk.dw = d->dw, hk.s_co = d->s_co, hk.s_ci = d->s_ci, hk.s_tap = d->s_tap;
      hk.scale_dev = d->scale_dev, hk.scale_imm = d->scale_imm;
      hk.ci_valid = d->ci_valid > 0 ? d->ci_valid : d->Cin;
      hk.co_valid = d->co_valid > 0 ? d->co_valid : d->Cout;
      for (int s = 0; s < DFGAN_MAX_SRC; ++s) {
        const dfgan_view4& v = d->src[s < d->n_src ? s : 0];
        if (v.C != d->Cin) return DFGAN_E_SHAPE;
        int rc = make_map_nhwc(&maps.a[s], v, d->Cin, hk.pitch, hk.hal_h, 1);
        if (rc) return rc;
      }
      int rc = make_map_nhwc(&maps.g, d->g, d->Cout, 8, 16, 1);
      if (rc) return rc;
      int grid = hk.pix_tiles < num_sms() ? hk.pix_tiles : num_sms();
      const double pix = static_cast<double>(N) * H * W;
      double bytes = 2.0 * pix * d->Cout + 4.0 * d->n_taps * d->Cin * d->Cout;
      for (int s = 0; s < d->n_src; ++s)
        bytes += 2.0 * d->src[s].N * d->src[s].H * d->src[s].W * d->src[s].C;
      const LaunchLabel lab("wgrad_halo", d, 0, grid);
      ProfileScope prof(1, 2.0 * pix * d->Cout * d->n_taps * d->Cin, bytes, stream, lab.s);
#define DFGAN_LAUNCH_WHALO(CI, CO)                                                                        \
  do {                                                                                                    \
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel<CI, CO>,                                       \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_halo);         \
    if (e != cudaSuccess) return static_cast<int>(e);                                                     \
    e = launch_pdl(wgrad_halo_kernel<CI, CO>, grid, kConvThreads, smem_halo, stream, 1, maps, hk);           \
    if (e != cudaSuccess) return static_cast<int>(e);                                                     \
  } while (0)
      if (d->Cin == 32 && d->Cout == 32) DFGAN_LAUNCH_WHALO(32, 32);
      else if (d->Cin == 32 && d->Cout == 64) DFGAN_LAUNCH_WHALO(32, 64);
      else if (d->Cin == 64 && d->Cout == 32) DFGAN_LAUNCH_WHALO(64, 32);
      else DFGAN_LAUNCH_WHALO(64, 64);
#undef DFGAN_LAUNCH_WHALO
      count_launch();
      DFGAN_CUDA_CHECK_LAUNCH();
      return DFGAN_OK;
    }
  }

  WgradK k{};
  WgradMaps maps{};
  const int W = d->g.W, H = d->g.H, N = d->g.N;
  brick_geometry(W, H, k.lbw, k.lbh);
  const int BW = 1 << k.lbw, BH = 1 << k.lbh, BN = 128 / (BW * BH);
  k.tiles_w = (W + BW - 1) / BW;
  k.tiles_h = (H + BH - 1) / BH;
  k.tiles_n = (N + BN - 1) / BN;
  k.pix_tiles = k.tiles_w * k.tiles_h * k.tiles_n;
  k.n_taps = d->n_taps;
  for (int t = 0; t < d->n_taps; ++t) {
    if (d->taps[t].src < 0 || d->taps[t].src >= d->n_src) return DFGAN_E_ARG;
    k.tap_src[t] = d->taps[t].src;
    k.tap_dy[t] = d->taps[t].dy;
    k.tap_dx[t] = d->taps[t].dx;
    k.tap_id[t] = d->taps[t].wrow;
  }
  k.cb = (d->Cin % 64 == 0) ? 64 : 32;
  k.lcb = ilog2(k.cb);
  k.blocks_per_tap = d->Cin / k.cb;
  k.row_blocks = d->n_taps * k.blocks_per_tap;
  k.cbg = (d->Cout % 64 == 0) ? 64 : 32;
  k.n_tile = d->Cout <= 256 ? d->Cout : 256;
  if (k.n_tile % k.cbg) return DFGAN_E_SHAPE;
  k.n_blocks_b = k.n_tile / k.cbg;
  const int blocks_per_tile = 128 / k.cb;
  k.n_tiles_m = (k.row_blocks + blocks_per_tile - 1) / blocks_per_tile;
  k.n_tiles_n = (d->Cout + k.n_tile - 1) / k.n_tile;
  const int base_items = k.n_tiles_m * k.n_tiles_n;
  int ks = d->k_splits;
  if (ks <= 0) {
    // exactly one wave of work items: every split ends in 128 x n_tile fp32 atomics, and the measured optimum
    // (tools/prof_wgrad.py) is items ~= #SMs for the wide layers; narrow layers (few row tiles) keep more
    // splits because their main loop, not the atomics, dominates
    static const int waves = getenv("DFGAN_WGRAD_WAVES") ? atoi(getenv("DFGAN_WGRAD_WAVES")) : 1;
    ks = base_items >= 8 ? waves * num_sms() / base_items : (2 * num_sms() + base_items - 1) / base_items;
    // keep at least 4 pixel tiles per split so the accumulate chain amortises the epilogue
    int max_ks = k.pix_tiles / 4;
    if (max_ks < 1) max_ks = 1;
    if (ks > max_ks) ks = max_ks;
  }
  if (ks > k.pix_tiles) ks = k.pix_tiles;
  if (ks < 1) ks = 1;
  k.k_splits = ks;
  k.total_items = base_items * ks;
  k.Cin = d->Cin, k.Cout = d->Cout;
  k.idesc = make_idesc_bf16(128, k.n_tile, 1, 1);
  k.a_blk_bytes = 128u * k.cb * 2u;
  k.a_bytes = k.a_blk_bytes * blocks_per_tile;  // always 32 KiB
  k.b_blk_bytes = 128u * k.cbg * 2u;
  k.b_bytes = k.b_blk_bytes * k.n_blocks_b;
  int stages = static_cast<int>(kSmemBudget / (k.a_bytes + k.b_bytes));
  if (stages > 6) stages = 6;
  if (stages < 2) return DFGAN_E_SHAPE;
  k.n_stages = stages;
  k.tmem_cols = pow2_cols(2 * k.n_tile);
  k.dw = d->dw;
  k.s_co = d->s_co, k.s_ci = d->s_ci, k.s_tap = d->s_tap;
  k.scale_dev = d->scale_dev, k.scale_imm = d->scale_imm;
  k.ci_valid = d->ci_valid > 0 ? d->ci_valid : d->Cin;
  k.co_valid = d->co_valid > 0 ? d->co_valid : d->Cout;

  for (int s = 0; s < DFGAN_MAX_SRC; ++s) {
    const dfgan_view4& v = d->src[s < d->n_src ? s : 0];
    if (v.C != d->Cin) return DFGAN_E_SHAPE;
    int rc = make_map_nhwc(&maps.a[s], v, k.cb, BW, BH, BN);
    if (rc) return rc;
  }
  int rc = make_map_nhwc(&maps.g, d->g, k.cbg, BW, BH, BN);
  if (rc) return rc;

  const size_t smem = static_cast<size_t>(stages) * (k.a_bytes + k.b_bytes) + 1024 + 256;
  int grid = k.total_items;
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  const double pix = static_cast<double>(N) * H * W;
  double bytes = 2.0 * pix * d->Cout + 4.0 * d->n_taps * d->Cin * d->Cout;
  for (int s = 0; s < d->n_src; ++s)
    bytes += 2.0 * d->src[s].N * d->src[s].H * d->src[s].W * d->src[s].C;
  const LaunchLabel lab("wgrad_umma", d, ks, grid);
  ProfileScope prof(1, 2.0 * pix * d->Cout * d->n_taps * d->Cin, bytes, stream, lab.s);
  cudaError_t e = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = launch_pdl(wgrad_umma_kernel, grid, kConvThreads, smem, stream, 1, maps, k);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}
