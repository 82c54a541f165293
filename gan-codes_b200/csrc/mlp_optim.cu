// (1) The 56 gamma/beta MLPs of the generator's AffineBlocks (generator/fusion_block.py:8-18,34-35) and
//     linear_in (generator/model.py:13,34) as ONE batched launch per layer over a device-side work table,
//     forward and backward.  These are weight-bandwidth-bound GEMV-like problems (batch 32..256 rows against
//     6.5 M fp32 parameters): fp32 CUDA-core math straight from the master weights, no packing.
// (2) Weight packing fp32 OIHW -> bf16 GEMM operands (incl. the sub-pixel folding of upsample->conv3x3 and
//     the un-folding of its weight gradient).
// (3) Fused global-norm clip + Adam over flat parameter buffers (deep_fusion_gan/model.py:42-43,195-197).
#include "host_common.h"
#include "sm100_ptx.cuh"

#include <cstddef>

namespace dfgan {

constexpr int kLinThreads = 256;
constexpr int kLinBT = 32;  // batch rows per block tile

// out[b, o] = act(sum_i in[b, i] * w[o, i] + bias[o]),  o in [o0, o0 + 64).
// The 64 x K weight tile is staged in shared memory with coalesced loads that stay in flight together (the first
// version read the weights with one broadcast 4-byte load per FMA: 96 us of pure load latency for 30 MB); a warp then
// accumulates 8 outputs at a time for its 32 batch rows (lane = row) from shared memory.
__global__ void __launch_bounds__(kLinThreads)
    linear_fwd_kernel(const dfgan_linear_item* __restrict__ items, int B) {
  extern __shared__ float sm[];
  const dfgan_linear_item it = items[blockIdx.x];
  const int b0 = blockIdx.y * kLinBT;
  const int K = it.k;
  const int ld = K + 1;
  float* s_x = sm;                 // [32][ld]
  float* s_w = sm + kLinBT * ld;   // [64][ld]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows of K contiguous floats: a warp streams one row at a time, up to 8 loads of 128 bytes in flight per lane
  for (int row = warp; row < kLinBT + 64; row += kLinThreads / 32) {
    const bool is_x = row < kLinBT;
    const int r = is_x ? row : row - kLinBT;
    const bool ok = is_x ? (b0 + r < B) : (it.o0 + r < it.n_out);
    const float* src = is_x ? it.in + static_cast<long long>(b0 + (ok ? r : 0)) * it.in_stride
                            : it.w + static_cast<long long>(it.o0 + (ok ? r : 0)) * K;
    float* dst = (is_x ? s_x : s_w) + r * ld;
    float v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = lane + 32 * m;
      v[m] = __ldg(src + (k < K ? k : 0));
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = lane + 32 * m;
      if (k < K) dst[k] = ok ? v[m] : 0.f;
    }
  }
  __syncthreads();
  const float* xrow = s_x + lane * ld;
  const int o_base = warp * 8;   // 8 warps x 8 outputs = 64
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float xv = xrow[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, s_w[(o_base + j) * ld + k], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int oo = it.o0 + o_base + j;
    if (oo < it.n_out && b0 + lane < B) {
      float a = acc[j] + (it.bias ? __ldg(it.bias + oo) : 0.f);
      if (it.relu) a = fmaxf(a, 0.f);
      it.out[static_cast<long long>(b0 + lane) * it.out_stride + oo] = a;
    }
  }
}

// din[b, k] (+)= mask(b,k) * sum_o dout[b, o] * w[o, k];  k tile = 64 columns per item (o0 reused as k0)
__global__ void __launch_bounds__(kLinThreads)
    linear_bwd_data_kernel(const dfgan_linear_item* __restrict__ items, int B) {
  extern __shared__ float sm[];
  const dfgan_linear_item it = items[blockIdx.x];
  const int b0 = blockIdx.y * kLinBT;
  const int NO = it.n_out;
  // stage dout tile [32][NO]
  for (int i = threadIdx.x; i < kLinBT * NO; i += kLinThreads) {
    const int b = i / NO, o = i - b * NO;
    sm[b * NO + o] = (b0 + b < B) ? __ldg(it.out + static_cast<long long>(b0 + b) * it.out_stride + o) : 0.f;
  }
  __syncthreads();
  // thread -> (k = o0 + tid % 64, batch group = tid / 64 -> 8 rows)
  const int k = it.o0 + (threadIdx.x & 63);
  const int bg = threadIdx.x >> 6;
  if (k >= it.k) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // sixteen weight rows requested together (one load per FMA group left the kernel waiting on DRAM 256 times; with
  // eight in flight a 32-CTA launch -- one generator block's MLPs -- still spent 45 us on 32 dependent round trips)
  constexpr int kRows = 16;
  for (int o = 0; o < NO; o += kRows) {
    float w[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) w[u] = __ldg(it.w + static_cast<long long>(o + u < NO ? o + u : o) * it.k + k);
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      if (o + u < NO) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(sm[(bg * 8 + j) * NO + o + u], w[u], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int b = b0 + bg * 8 + j;
    if (b < B) {
      const long long off = static_cast<long long>(b) * it.in_stride + k;
      // relu backward: `mask` holds the forward activation (same layout as din)
      const_cast<float*>(it.in)[off] = (it.mask && !(__ldg(it.mask + off) > 0.f)) ? 0.f : acc[j];
    }
  }
}

// dw[o, k] += sum_b dout[b, o] * in[b, k] ; dbias[o] += sum_b dout[b, o];  64 outputs per item
__global__ void __launch_bounds__(kLinThreads)
    linear_bwd_weight_kernel(const dfgan_linear_item* __restrict__ items, int B) {
  extern __shared__ float sm[];  // dout tile [B chunk 32][64] + in tile [32][K]
  const dfgan_linear_item it = items[blockIdx.x];
  const int K = it.k;
  float* s_d = sm;
  float* s_x = sm + kLinBT * 64;
  // thread -> column k = tid % K_pad ... each thread owns (o subset) x (k): loop layout below
  const int n_o = min(64, it.n_out - it.o0);
  for (int b0 = 0; b0 < B; b0 += kLinBT) {
    __syncthreads();
    {
      float v[8];   // kLinBT * 64 / kLinThreads = 8 elements per thread, all requested before the first use
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = threadIdx.x + u * kLinThreads, b = i >> 6, o = i & 63;
        const bool ok = b0 + b < B && o < n_o;
        v[u] = __ldg(ok ? it.out + static_cast<long long>(b0 + b) * it.out_stride + it.o0 + o : it.out);
        if (!ok) v[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s_d[threadIdx.x + u * kLinThreads] = v[u];
    }
    for (int i0 = threadIdx.x; i0 < kLinBT * K; i0 += 8 * kLinThreads) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * kLinThreads;
        const int ic = i < kLinBT * K ? i : i0;
        const int b = ic / K, k = ic - b * K;
        const bool ok = i < kLinBT * K && b0 + b < B;
        v[u] = __ldg(ok ? it.in + static_cast<long long>(b0 + b) * it.in_stride + k : it.in);
        if (!ok) v[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * kLinThreads < kLinBT * K) s_x[i0 + u * kLinThreads] = v[u];
    }
    __syncthreads();
    // four (o, k) elements per pass: their current gradient values are requested together, ahead of the sums
    for (int e0 = threadIdx.x; e0 < n_o * K; e0 += 4 * kLinThreads) {
      float* dst[4];
      float cur[4], acc[4];
      int oo[4], kk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * kLinThreads;
        const int ec = e < n_o * K ? e : e0;
        oo[u] = ec / K, kk[u] = ec - oo[u] * K;
        dst[u] = const_cast<float*>(it.w) + static_cast<long long>(it.o0 + oo[u]) * K + kk[u];
        cur[u] = *dst[u];
        acc[u] = 0.f;
      }
#pragma unroll 8
      for (int b = 0; b < kLinBT; ++b) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(s_d[b * 64 + oo[u]], s_x[b * K + kk[u]], acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u * kLinThreads < n_o * K) *dst[u] = cur[u] + acc[u];  // each (o,k) is owned by exactly one thread
    }
    if (it.bias && threadIdx.x < n_o) {
      float acc = 0.f;
      for (int b = 0; b < kLinBT; ++b) acc += s_d[b * 64 + threadIdx.x];
      const_cast<float*>(it.bias)[it.o0 + threadIdx.x] += acc;
    }
  }
}

// ------------------------------------------------------------------------------------ weight packing
constexpr int kPackSlots = 36;   // 9 shifts x 4 output phases of the merged sub-pixel operands
struct PackP {
  const float* w;  // [Co][Ci][kh*kw] fp32
  __nv_bfloat16* out;
  int Co, Ci, KK;
  int transpose;  // 0: rows = co, cols = ci ; 1: rows = ci, cols = co
  int rows_pad, cols_pad, n_slots;
  float scale;
  int image;  // G >= 1: [slot / G][64-column chunk][(slot % G) * rows_pad + row][64] with the 16-byte pieces of every row XOR-swizzled by (row & 7):
              // the exact SWIZZLE_128B K-major shared-memory image, fetched by ONE linear bulk copy per tile
  int8_t taps[kPackSlots][4];  // per slot: source tap indices summed (-1 = unused)
};
__device__ __forceinline__ long long pack_dst(const PackP& p, int slot, int row, int col, long long linear) {
  if (!p.image) return linear;
  // image = G >= 1: G consecutive slots form one tile (merged phases: G = 4), rows (slot % G) * rows_pad + row
  const int G = p.image;
  const int slot_i = slot / G, row_i = (slot - slot_i * G) * p.rows_pad + row, rows_i = G * p.rows_pad;
  const int nc = p.cols_pad >> 6, c = col >> 6, j = (col & 63) >> 3, e = col & 7;
  return ((static_cast<long long>(slot_i) * nc + c) * rows_i + row_i) * 64 + (((j ^ (row_i & 7)) << 3) + e);
}
// One 256-thread block per 32 x 32 tile of (co, ci) pairs.  Phase 1 stages the tile's fp32 weights (32 runs of
// 32*KK contiguous floats) in shared memory with fully coalesced loads; phase 2 writes every slot of the operand with
// 16-byte stores (8 consecutive columns per thread), reading the tile transposed when the operand is.  (History: one
// thread per output element = 1.4 ms for the discriminator; one thread per (row, col) with scalar loads at a 36-byte
// stride and 2-byte stores = 0.24 ms at 0.5 TB/s; the traffic is ~0.1 ms.)
constexpr int kPackTile = 32;
constexpr int kPackSmemFloats = kPackTile * (kPackTile * 16 + 1);
__device__ __forceinline__ void pack_tile(const PackP& p, int tile, float* sm) {
  const int tiles_c = (p.cols_pad + kPackTile - 1) / kPackTile;
  const int row0 = (tile / tiles_c) * kPackTile, col0 = (tile % tiles_c) * kPackTile;
  const int co0 = p.transpose ? col0 : row0, ci0 = p.transpose ? row0 : col0;
  const int KK = p.KK, run = kPackTile * KK, stride = run + 1;
  const int n_co = p.Co - co0, n_ci = p.Ci - ci0;   // valid extent of the tile (may be <= 0 in the padding)
  // a warp streams one co row (32*KK contiguous floats) at a time, KK loads of 128 bytes kept in flight together
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lim = n_ci * KK;
#pragma unroll
    for (int j = 0; j < kPackTile / 8; ++j) {
      const int co_l = warp + 8 * j;
      const float* src = p.w + (static_cast<long long>(co0 + co_l) * p.Ci + ci0) * KK;
      // all 16 loads are issued unconditionally (invalid ones re-read p.w[0]) so that they stay in flight together:
      // with predicated loads the compiler serialised load -> use, one DRAM latency per 128 bytes
      float v[16];
      uint32_t ok = 0;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int r = lane + 32 * m;
        const bool in = (m < KK) && (co_l < n_co) && (r < lim);
        ok |= static_cast<uint32_t>(in) << m;
        v[m] = __ldg(in ? src + r : p.w);
      }
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (m < KK) sm[co_l * stride + lane + 32 * m] = ((ok >> m) & 1u) ? v[m] * p.scale : 0.f;
    }
  }
  __syncthreads();
  const int item = threadIdx.x & 127, row_l = item >> 2, cg = item & 3;
  const int row = row0 + row_l, col = col0 + cg * 8;
  if (row >= p.rows_pad || col >= p.cols_pad) return;
  const long long per_slot = static_cast<long long>(p.rows_pad) * p.cols_pad;
  for (int slot = threadIdx.x >> 7; slot < p.n_slots; slot += 2) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = p.taps[slot][j];   // block-uniform
      if (t >= 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int co_l = p.transpose ? cg * 8 + e : row_l, ci_l = p.transpose ? row_l : cg * 8 + e;
          v[e] += sm[co_l * stride + ci_l * KK + t];
        }
      }
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]), o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]), o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]), o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]), o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(p.out + pack_dst(p, slot, row, col, slot * per_slot + static_cast<long long>(row) * p.cols_pad + col)) = o;
  }
}
static int pack_blocks(int rows_pad, int cols_pad) {
  return ((rows_pad + kPackTile - 1) / kPackTile) * ((cols_pad + kPackTile - 1) / kPackTile);
}
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackP p) {
  extern __shared__ float pack_sm[];
  pack_tile(p, blockIdx.x, pack_sm);
}
// All weight tensors of a network in one launch: entries[e] = PackP + first block; a block finds its entry by
// scanning the (<= 128-entry) table.
struct PackEntry {
  PackP p;
  int block_start;
};
__global__ void __launch_bounds__(256) pack_batched_kernel(const PackEntry* __restrict__ entries, int n_entries) {
  extern __shared__ float pack_sm[];
  __shared__ __align__(16) PackP sp;
  // every thread runs the same binary search (broadcast loads; a linear scan by one thread cost more than the tile's
  // work), then the block copies the entry cooperatively
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (static_cast<int>(blockIdx.x) >= __ldg(&entries[mid].block_start)) lo = mid;
    else hi = mid - 1;
  }
  const int start = __ldg(&entries[lo].block_start);
  static_assert(sizeof(PackP) % 4 == 0, "PackP is copied word by word");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&entries[lo].p);
  for (int i = threadIdx.x; i < static_cast<int>(sizeof(PackP) / 4); i += 256) reinterpret_cast<uint32_t*>(&sp)[i] = __ldg(src + i);
  __syncthreads();
  pack_tile(sp, blockIdx.x - start, pack_sm);
}

// ------------------------------------------------------------------------------------ weight-gradient un-pack
// The weight-gradient GEMMs accumulate with fp32 atomics; into the reference's OIHW layout (tap fastest) the 32
// lanes of a warp (consecutive ci) scatter over 36..64-byte strides, into a tap-major scratch [tap][co][ci] they are
// one coalesced 128-byte line (measured: 61 -> 43 us for a 3x3 C256 layer).  This kernel folds the scratch back:
//   dst[co][ci][t] += src[t][co][ci] ;  src[t][co][ci] = 0   (ready for the next sweep)
// one 256-thread block per 32 x 32 tile of (co, ci) pairs, both sides coalesced through shared memory.
struct UnpackEntry {
  float* src;   // [KK][Co][Ci]
  float* dst;   // [Co][Ci][KK]
  int Co, Ci, KK, block_start;
};
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const UnpackEntry* __restrict__ entries, int n_entries) {
  extern __shared__ float us[];
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (static_cast<int>(blockIdx.x) >= __ldg(&entries[mid].block_start)) lo = mid;
    else hi = mid - 1;
  }
  const UnpackEntry e = entries[lo];
  const int tile = blockIdx.x - e.block_start;
  const int tiles_c = (e.Ci + 31) / 32;
  const int co0 = (tile / tiles_c) * 32, ci0 = (tile % tiles_c) * 32;
  const int KK = e.KK, run = 32 * KK, stride = run + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long plane = static_cast<long long>(e.Co) * e.Ci;
  const bool ci_ok = ci0 + lane < e.Ci;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co_l = warp + 8 * j;
    const bool ok = ci_ok && co0 + co_l < e.Co;
    float* src = e.src + static_cast<long long>(ok ? co0 + co_l : 0) * e.Ci + (ok ? ci0 + lane : 0);
    float v[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = __ldcg(src + (t < KK ? t : 0) * plane);   // unconditional: all in flight
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      if (t < KK) {
        us[co_l * stride + lane * KK + t] = ok ? v[t] : 0.f;
        if (ok) src[t * plane] = 0.f;
      }
    }
  }
  __syncthreads();
  const int lim = (e.Ci - ci0 < 32 ? e.Ci - ci0 : 32) * KK;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co_l = warp + 8 * j;
    if (co0 + co_l >= e.Co) continue;
    float* dst = e.dst + (static_cast<long long>(co0 + co_l) * e.Ci + ci0) * KK;
    float cur[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int r = lane + 32 * m;
      cur[m] = __ldcg(dst + ((m < KK && r < lim) ? r : 0));
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int r = lane + 32 * m;
      if (m < KK && r < lim) dst[r] = cur[m] + us[co_l * stride + r];
    }
  }
}

// dw[co][ci][r*3+s] += sum_{a,c} folded[(a*2+c)*4 + dyi_a(r)*2 + dxi_c(s)][co][ci]   (sub-pixel un-fold)
__global__ void unfold_kernel(const float* __restrict__ folded, float* __restrict__ dw, int Co, int Ci) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(Co) * Ci * 9;
  if (i >= total) return;
  const int t = i % 9;
  const long long cc = i / 9;  // co*Ci + ci
  const int r = t / 3, s = t % 3;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int yi = a == 0 ? (r == 0 ? 0 : 1) : (r == 2 ? 1 : 0);
      const int xi = c == 0 ? (s == 0 ? 0 : 1) : (s == 2 ? 1 : 0);
      acc += __ldg(folded + static_cast<long long>((a * 2 + c) * 4 + yi * 2 + xi) * Co * Ci + cc);
    }
  dw[i] += acc;
}

// ------------------------------------------------------------------------------------ clip + Adam
// Deterministic variant: every block writes its partial sum (fixed in-block order) to scratch[blockIdx.x]; the block
// that finishes last adds all partials in index order.  The result does not depend on the order in which blocks
// run, so every data-parallel rank derives bit-identical clip coefficients from the all-reduced gradient (with
// float atomics the ranks' parameters drifted apart by ulps whenever the clip was active).
// scratch: gridDim.x floats + one zero-initialised uint counter (reset by the last block).
__global__ void __launch_bounds__(256) sqnorm_det_kernel(const float* __restrict__ g, long long n, float* out,
                                                         float* scratch) {
  __shared__ float red[8];
  __shared__ bool last;
  // four independent 16-byte loads in flight per thread and iteration; every partial sum is formed in a fixed order,
  // so the result does not depend on scheduling
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const long long stride = gridDim.x * 1024LL;
  long long i = (blockIdx.x * 256LL + threadIdx.x) * 4;
  for (; i + 3 * stride + 3 < n; i += 4 * stride) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(g + i));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(g + i + stride));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(g + i + 2 * stride));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(g + i + 3 * stride));
    a0 += v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
    a1 += v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
    a2 += v2.x * v2.x + v2.y * v2.y + v2.z * v2.z + v2.w * v2.w;
    a3 += v3.x * v3.x + v3.y * v3.y + v3.z * v3.z + v3.w * v3.w;
  }
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      a0 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) a0 += g[j] * g[j];
    }
  }
  float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + gridDim.x);
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    scratch[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    // the last block folds the per-block partials: strided per-thread sums, then the same fixed tree as above
    __threadfence();
    float s = 0.f;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += 256) s += __ldcg(scratch + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[k];
      *out += t;
      *counter = 0u;
    }
  }
}
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, float* out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (blockIdx.x * 256LL + threadIdx.x) * 4; i < n; i += gridDim.x * 1024LL) {
    if (i + 3 < n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) acc += g[j] * g[j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(out, s);
  }
}
// torch.nn.utils.clip_grad_norm_(max_norm) followed by torch.optim.Adam.step():
//   g' = g * gscale * min(1, max_norm / (gscale*sqrt(sumsq) + 1e-6))
//   m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, long long n, const float* sumsq, const int* step, float gscale,
                float max_norm, float lr, float b1, float b2, float eps) {
  const float t = static_cast<float>(*step);
  const float norm = gscale * sqrtf(__ldg(sumsq));
  const float coef = gscale * (max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f);
  const float c1 = 1.f - powf(b1, t), c2s = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / c1;
  // beta1 == 0 (the reference's betas = (0.0, 0.9), deep_fusion_gan/model.py:42-43): exp_avg is the clipped gradient
  // itself -- it is still written (optimiser state_dict), but the old value is not read: 24 instead of 28 B per element
  const bool keep_m = b1 != 0.f;
  // 16-byte accesses (the flat buffers are 16-byte aligned and padded to multiples of four elements); a scalar
  // tail covers any other caller
  const long long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) ? 0 : n / 4;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = keep_m ? reinterpret_cast<float4*>(m)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    const float gs[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
    float* mm = reinterpret_cast<float*>(&m4);
    float* vv = reinterpret_cast<float*>(&v4);
    float* pp = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = b1 * mm[j] + (1.f - b1) * gs[j];
      vv[j] = b2 * vv[j] + (1.f - b2) * gs[j] * gs[j];
      pp[j] -= step_size * mm[j] / (sqrtf(vv[j]) / c2s + eps);
    }
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = n4 * 4 + blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const float gi = g[i] * coef;
    const float mi = b1 * (keep_m ? m[i] : 0.f) + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) / c2s + eps);
  }
}

}  // namespace dfgan

using namespace dfgan;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" {

int dfgan_linear_batched(const dfgan_linear_item* items_dev, int n_items, int B, int k_max, int n_out_max,
                         int mode, dfgan_stream_t st) {
  if (!items_dev || n_items < 1 || B < 1 || k_max < 1 || k_max > 256) return DFGAN_E_ARG;
  const int by = (B + kLinBT - 1) / kLinBT;
  if (mode == 0) {
    const size_t smem = static_cast<size_t>(kLinBT + 64) * (k_max + 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    linear_fwd_kernel<<<dim3(n_items, by), kLinThreads, smem, STREAM(st)>>>(items_dev, B);
  } else if (mode == 1) {
    if (n_out_max < 1 || n_out_max > 256) return DFGAN_E_ARG;
    const size_t smem = static_cast<size_t>(kLinBT) * n_out_max * sizeof(float);
    linear_bwd_data_kernel<<<dim3(n_items, by), kLinThreads, smem, STREAM(st)>>>(items_dev, B);
  } else if (mode == 2) {
    const size_t smem = static_cast<size_t>(kLinBT) * (64 + k_max) * sizeof(float);
    linear_bwd_weight_kernel<<<n_items, kLinThreads, smem, STREAM(st)>>>(items_dev, B);
  } else {
    return DFGAN_E_ARG;
  }
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_pack_weights(const float* w, void* out, int Co, int Ci, int KK, int transpose, int rows_pad,
                       int cols_pad, int n_slots, const int8_t* slot_taps /* host [n_slots][4] */,
                       float scale, int image, dfgan_stream_t st) {
  if (!w || !out || !slot_taps || n_slots < 1 || n_slots > kPackSlots) return DFGAN_E_ARG;
  if (rows_pad < (transpose ? Ci : Co) || cols_pad < (transpose ? Co : Ci)) return DFGAN_E_SHAPE;
  PackP p{};
  p.w = w, p.out = static_cast<__nv_bfloat16*>(out);
  p.Co = Co, p.Ci = Ci, p.KK = KK, p.transpose = transpose;
  if (image && (cols_pad & 63)) return DFGAN_E_SHAPE;
  p.rows_pad = rows_pad, p.cols_pad = cols_pad, p.n_slots = n_slots, p.scale = scale, p.image = image;
  for (int s = 0; s < kPackSlots; ++s)
    for (int j = 0; j < 4; ++j) p.taps[s][j] = s < n_slots ? slot_taps[s * 4 + j] : -1;
  if (KK > 16 || (cols_pad & 7)) return DFGAN_E_SHAPE;
  const int smem = static_cast<int>(kPackSmemFloats * sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  // three 66 KB blocks per SM need the large shared-memory carveout (the default heuristic gave one block per SM)
  cudaFuncSetAttribute(pack_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  pack_kernel<<<pack_blocks(rows_pad, cols_pad), 256, smem, STREAM(st)>>>(p);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_pack_entry_size(void) { return static_cast<int>(sizeof(PackEntry)); }
int dfgan_pack_entry_bs_offset(void) { return static_cast<int>(offsetof(PackEntry, block_start)); }

/* Fills one host-side table entry (entry_out points at dfgan_pack_entry_size() bytes); returns the number of
 * 256-thread blocks the entry needs, or a negative error.  block_start is the running block offset. */
int dfgan_pack_entry_fill(void* entry_out, const float* w, void* out, int Co, int Ci, int KK, int transpose,
                          int rows_pad, int cols_pad, int n_slots, const int8_t* slot_taps, float scale,
                          int image, int block_start) {
  if (!entry_out || !w || !out || !slot_taps || n_slots < 1 || n_slots > kPackSlots) return DFGAN_E_ARG;
  if (rows_pad < (transpose ? Ci : Co) || cols_pad < (transpose ? Co : Ci)) return DFGAN_E_SHAPE;
  PackEntry e{};
  e.p.w = w, e.p.out = static_cast<__nv_bfloat16*>(out);
  e.p.Co = Co, e.p.Ci = Ci, e.p.KK = KK, e.p.transpose = transpose;
  if (image && (cols_pad & 63)) return DFGAN_E_SHAPE;
  e.p.rows_pad = rows_pad, e.p.cols_pad = cols_pad, e.p.n_slots = n_slots, e.p.scale = scale, e.p.image = image;
  for (int s = 0; s < kPackSlots; ++s)
    for (int j = 0; j < 4; ++j) e.p.taps[s][j] = s < n_slots ? slot_taps[s * 4 + j] : -1;
  e.block_start = block_start;
  if (KK > 16 || (cols_pad & 7)) return DFGAN_E_SHAPE;
  *static_cast<PackEntry*>(entry_out) = e;
  return pack_blocks(rows_pad, cols_pad);
}

int dfgan_pack_batched(const void* entries_dev, int n_entries, int total_blocks, dfgan_stream_t st) {
  if (!entries_dev || n_entries < 1 || total_blocks < 1) return DFGAN_E_ARG;
  const int smem = static_cast<int>(kPackSmemFloats * sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(pack_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaFuncSetAttribute(pack_batched_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  pack_batched_kernel<<<total_blocks, 256, smem, STREAM(st)>>>(static_cast<const PackEntry*>(entries_dev), n_entries);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_unpack_entry_size(void) { return static_cast<int>(sizeof(UnpackEntry)); }
/* Fills one host-side table entry; returns the number of blocks the entry needs (or a negative error). */
int dfgan_unpack_entry_fill(void* entry_out, float* src, float* dst, int Co, int Ci, int KK, int block_start) {
  if (!entry_out || !src || !dst || Co < 1 || Ci < 1 || KK < 1 || KK > 16) return DFGAN_E_ARG;
  UnpackEntry e{src, dst, Co, Ci, KK, block_start};
  *static_cast<UnpackEntry*>(entry_out) = e;
  return ((Co + 31) / 32) * ((Ci + 31) / 32);
}
int dfgan_wgrad_unpack(const void* entries_dev, int n_entries, int total_blocks, dfgan_stream_t st) {
  if (!entries_dev || n_entries < 1 || total_blocks < 1) return DFGAN_E_ARG;
  const int smem = static_cast<int>(kPackSmemFloats * sizeof(float));
  cudaError_t e = cudaFuncSetAttribute(wgrad_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  wgrad_unpack_kernel<<<total_blocks, 256, smem, STREAM(st)>>>(static_cast<const UnpackEntry*>(entries_dev), n_entries);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_unfold_subpixel_wgrad(const float* folded, float* dw, int Co, int Ci, dfgan_stream_t st) {
  if (!folded || !dw) return DFGAN_E_ARG;
  const long long total = static_cast<long long>(Co) * Ci * 9;
  unfold_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, STREAM(st)>>>(folded, dw, Co, Ci);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_sqnorm(const float* g, int64_t n, float* out, dfgan_stream_t st) {
  if (!g || !out || n < 1 || (reinterpret_cast<uintptr_t>(g) & 15)) return DFGAN_E_ARG;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  sqnorm_kernel<<<static_cast<int>(blocks), 256, 0, STREAM(st)>>>(g, n, out);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_sqnorm_det(const float* g, int64_t n, float* out, float* scratch, int scratch_floats, dfgan_stream_t st) {
  if (!g || !out || !scratch || n < 1 || (reinterpret_cast<uintptr_t>(g) & 15)) return DFGAN_E_ARG;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();   // 4 x 148 blocks x 4 loads in flight per thread
  if (scratch_floats < blocks + 1) return DFGAN_E_ARG;
  sqnorm_det_kernel<<<static_cast<int>(blocks), 256, 0, STREAM(st)>>>(g, n, out, scratch);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                    const int* step, float gscale, float max_norm, float lr, float b1, float b2, float eps,
                    dfgan_stream_t st) {
  if (!p || !g || !m || !v || !sumsq || !step || n < 1) return DFGAN_E_ARG;
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  adam_kernel<<<static_cast<int>(blocks), 256, 0, STREAM(st)>>>(p, g, m, v, n, sumsq, step, gscale, max_norm,
                                                               lr, b1, b2, eps);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

}  // extern "C"
