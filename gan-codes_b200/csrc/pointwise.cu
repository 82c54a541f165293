// HBM-bound pointwise / reduction kernels around the tensor-core GEMMs: the AffineBlock chain and its
// backward (with the per-(sample,channel) reductions for the gamma/beta MLPs), 2x2 pooling, the residual
// gate backward of the discriminator, layout changes at the image boundary and the logit head.
// All activations are contiguous NHWC bf16 [N][HW][C] with C % 8 == 0; every thread moves 16-byte vectors
// (8 channels), a warp covers consecutive channels/pixels so global accesses are fully coalesced.
#include "host_common.h"

#include <cstdlib>
#include "sm100_ptx.cuh"

namespace dfgan {

constexpr int kPwThreads = 256;

struct bf8 {
  float v[8];
};
__device__ __forceinline__ bf8 ld8(const __nv_bfloat16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  bf8 r;
  unpack_bf16x2(u.x, r.v[0], r.v[1]);
  unpack_bf16x2(u.y, r.v[2], r.v[3]);
  unpack_bf16x2(u.z, r.v[4], r.v[5]);
  unpack_bf16x2(u.w, r.v[6], r.v[7]);
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]);
  u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// 16 consecutive channels of one pixel as ONE 256-bit store (a full 32-byte sector per instruction: kernels in which
// every thread owns a whole pixel row would otherwise write each sector in two 16-byte halves)
__device__ __forceinline__ void st16(__nv_bfloat16* p, const float* v) {
  uint4 lo, hi;
  lo.x = pack_bf16x2(v[0], v[1]), lo.y = pack_bf16x2(v[2], v[3]), lo.z = pack_bf16x2(v[4], v[5]), lo.w = pack_bf16x2(v[6], v[7]);
  hi.x = pack_bf16x2(v[8], v[9]), hi.y = pack_bf16x2(v[10], v[11]), hi.z = pack_bf16x2(v[12], v[13]), hi.w = pack_bf16x2(v[14], v[15]);
  stg256(p, lo, hi);
}
__device__ __forceinline__ bf8 ldf8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  return bf8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ float dlrelu(float pre, float slope) { return pre > 0.f ? 1.f : slope; }

// 4-channel (8-byte) variants for the register-heavy backward kernel: half the per-thread state of the 8-wide
// form, which is what lets four 256-thread blocks share an SM (a streaming kernel needs the occupancy).
struct bf4 {
  float v[4];
};
__device__ __forceinline__ bf4 ld4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  bf4 r;
  unpack_bf16x2(u.x, r.v[0], r.v[1]);
  unpack_bf16x2(u.y, r.v[2], r.v[3]);
  return r;
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const bf4& r) {
  uint2 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ bf4 ldf4(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  return bf4{{a.x, a.y, a.z, a.w}};
}
// block_reduce8 for 4-wide accumulators (tpp = C / 4 threads per pixel); red: shared float[kPwThreads * 4]
__device__ __forceinline__ void block_reduce4(bf4 acc, float* red, int tpp, float* dst) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_warp = tpp < 32 ? tpp : 32;
  for (int o = 16; o >= tpp; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc.v[i] += __shfl_xor_sync(0xffffffffu, acc.v[i], o);
  }
  __syncthreads();
  if (lane < per_warp) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[(warp * per_warp + lane) * 4 + i] = acc.v[i];
  }
  __syncthreads();
  if (tid < tpp) {
    const int n_entries = (kPwThreads / 32) * per_warp;
    float s[4] = {0, 0, 0, 0};
    for (int e = tid; e < n_entries; e += tpp) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] += red[e * 4 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(dst + i, s[i]);
  }
}

// Sum `acc` over all threads of the block that share (threadIdx.x % tpp) and atomically add the result to
// dst[8].  Stage 1: warp shuffles across the lanes with equal (lane % tpp) (tpp < 32 only); stage 2: one
// partial per (warp, channel group) through shared memory, at most 8 partials per group are summed serially.
// (A first version let tpp threads walk 256/tpp partials each: for C = 32 that serial tail cost more than the
// block's whole streaming loop.)   red: shared float[kPwThreads * 8].
__device__ __forceinline__ void block_reduce8(bf8 acc, float* red, int tpp, float* dst, float scale) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_warp = tpp < 32 ? tpp : 32;
  for (int o = 16; o >= tpp; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] += __shfl_xor_sync(0xffffffffu, acc.v[i], o);
  }
  __syncthreads();
  if (lane < per_warp) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(warp * per_warp + lane) * 8 + i] = acc.v[i];
  }
  __syncthreads();
  if (tid < tpp) {
    const int n_entries = (kPwThreads / 32) * per_warp;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = tid; e < n_entries; e += tpp) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += red[e * 8 + i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dst + i, s[i] * scale);
  }
}
// Sum a scalar over the block and atomically add it to *dst.
__device__ __forceinline__ void block_reduce1(float x, float* red, float* dst, float scale) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kPwThreads / 32; ++i) s += red[i];
    atomicAdd(dst, s * scale);
  }
}

// ------------------------------------------------------------------------------------ AffineBlock pair
// out = lrelu(gb * lrelu(ga * x + ba) + bb)      generator/residual_block.py:36-39 / 43-46
__global__ void __launch_bounds__(kPwThreads)
    affine2_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                       const float* ga, const float* ba, const float* gb, const float* bb,
                       long long tab_stride, int HW, int C, float slope) {
  const int n = blockIdx.y;
  const int tpp = C >> 3;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp, ppi = kPwThreads / tpp;
  const long long t_off = n * tab_stride + cg * 8;
  const bf8 g1 = ldf8(ga + t_off), b1 = ldf8(ba + t_off), g2 = ldf8(gb + t_off), b2 = ldf8(bb + t_off);
  const long long base = static_cast<long long>(n) * HW * C;
  for (int p = blockIdx.x * ppi + pl; p < HW; p += gridDim.x * ppi) {
    const long long off = base + static_cast<long long>(p) * C + cg * 8;
    bf8 v = ld8(x + off);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v.v[i] = lrelu(fmaf(g2.v[i], lrelu(fmaf(g1.v[i], v.v[i], b1.v[i]), slope), b2.v[i]), slope);
    st8(out + off, v);
  }
}

// Backward of the pair (masks recomputed from the stored bf16 input x):
//   g   = g_a * (gscale ? *gscale : 1)
//   u1  = ga*x+ba, l1 = lrelu(u1), u2 = gb*l1+bb
//   dot += sum g_a * lrelu(u2)                         (optional: residual-gate gradient)
//   gu2 = g*lrelu'(u2) ; d_gb += sum gu2*l1 ; d_bb += sum gu2
//   gu1 = gu2*gb*lrelu'(u1) ; d_ga += sum gu1*x ; d_ba += sum gu1
//   gx  = gu1*ga (+ g_add) ; colsum += sum gx (optional: bias gradient of the producing conv)
// three blocks per SM (85 registers): at four (64 registers) the compiler spilled inside the streaming loop and the
// local-memory traffic was 2.4x the kernel's real traffic (ncu: 30 M local sectors against 13 M global)
__global__ void __launch_bounds__(kPwThreads, 3)
    affine2_bwd_kernel(const __nv_bfloat16* __restrict__ g_a, const __nv_bfloat16* __restrict__ x,
                       const __nv_bfloat16* __restrict__ g_add, __nv_bfloat16* __restrict__ gx,
                       const float* ga, const float* ba, const float* gb, const float* bb,
                       float* d_ga, float* d_ba, float* d_gb, float* d_bb, long long tab_stride,
                       const float* gscale, float* dot, float* colsum, int HW, int C, int pix_per_block,
                       float slope) {
  __shared__ float red[kPwThreads * 4];
  const int n = blockIdx.y;
  const int tpp = C >> 2;   // 4 channels per thread
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp, ppi = kPwThreads / tpp;
  const long long t_off = n * tab_stride + cg * 4;
  const bf4 g1 = ldf4(ga + t_off), b1 = ldf4(ba + t_off), g2 = ldf4(gb + t_off), b2 = ldf4(bb + t_off);
  const float gs = gscale ? __ldg(gscale) : 1.f;
  const long long base = static_cast<long long>(n) * HW * C;
  bf4 a_ga{}, a_ba{}, a_gb{}, a_bb{}, a_cs{};
  float a_dot = 0.f;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  constexpr int U = 2;   // pixels in flight per thread: both batches' loads are issued before any math
  for (int pb = p0 + pl; pb < p1; pb += U * ppi) {
    bf4 gin[U], xv[U], o[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p < p1) {
        const long long off = base + static_cast<long long>(p) * C + cg * 4;
        gin[u] = ld4(g_a + off);
        xv[u] = ld4(x + off);
        if (g_add) o[u] = ld4(g_add + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p < p1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float u1 = fmaf(g1.v[i], xv[u].v[i], b1.v[i]);
          const float l1 = lrelu(u1, slope);
          const float u2 = fmaf(g2.v[i], l1, b2.v[i]);
          a_dot += gin[u].v[i] * lrelu(u2, slope);
          const float gu2 = gin[u].v[i] * gs * dlrelu(u2, slope);
          a_gb.v[i] += gu2 * l1;
          a_bb.v[i] += gu2;
          const float gu1 = gu2 * g2.v[i] * dlrelu(u1, slope);
          a_ga.v[i] += gu1 * xv[u].v[i];
          a_ba.v[i] += gu1;
          const float r = gu1 * g1.v[i] + (g_add ? o[u].v[i] : 0.f);
          o[u].v[i] = r;
          a_cs.v[i] += r;
        }
        st4(gx + base + static_cast<long long>(p) * C + cg * 4, o[u]);
      }
    }
  }
  block_reduce4(a_ga, red, tpp, d_ga + t_off);
  block_reduce4(a_ba, red, tpp, d_ba + t_off);
  block_reduce4(a_gb, red, tpp, d_gb + t_off);
  block_reduce4(a_bb, red, tpp, d_bb + t_off);
  if (colsum) block_reduce4(a_cs, red, tpp, colsum + cg * 4);
  if (dot) block_reduce1(a_dot, red, dot, 1.f);
}

// ------------------------------------------------------------------------------------ 2x2 pooling
// out[n,i,j,c] = scale * sum_{a,b < f} x[n, f*i+a, f*j+b, c]  (f = 1 << shift);  colsum[c] += cs * sum x
// avg_pool2d (discriminator/residual_block.py:27) with scale .25; backward of F.interpolate
// (generator/model.py:40) with scale 1; shift 0 degenerates to a copy + column sum.
// SHIFT is a template parameter so that the copy + column-sum form (SHIFT = 0, one load per pixel) can keep four
// pixels in flight per thread without the register cost of the four-tap form.
template <int SHIFT, int U>
__global__ void __launch_bounds__(kPwThreads, 4)
    pool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, float* colsum,
                const float* cs_dev, float cs_imm, int Ho, int Wo, int C, float scale, int pix_per_block) {
  constexpr int shift = SHIFT;
  constexpr int NQ = SHIFT ? 4 : 1;
  __shared__ float red[kPwThreads * 8];
  const int n = blockIdx.y;
  const int tpp = C >> 3;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp, ppi = kPwThreads / tpp;
  constexpr int f = 1 << shift;
  const int Wi = Wo << shift;
  const long long ibase = static_cast<long long>(n) * (Ho << shift) * Wi * C;
  const long long obase = static_cast<long long>(n) * Ho * Wo * C;
  bf8 acc{};
  const int HW = Ho * Wo;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  for (int pb = p0 + pl; pb < p1; pb += U * ppi) {
    bf8 v[U][NQ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p < p1) {
        const int i = p / Wo, j = p - i * Wo;
#pragma unroll
        for (int a = 0; a < f; ++a)
#pragma unroll
          for (int b = 0; b < f; ++b)
            v[u][a * f + b] = ld8(x + ibase + (static_cast<long long>(f * i + a) * Wi + (f * j + b)) * C + cg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p >= p1) continue;
      bf8 s{};
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s.v[k] += v[u][q].v[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc.v[k] += s.v[k];
        s.v[k] *= scale;
      }
      if (out) st8(out + obase + static_cast<long long>(p) * C + cg * 8, s);
    }
  }
  if (colsum) block_reduce8(acc, red, tpp, colsum + cg * 8, cs_imm * (cs_dev ? __ldg(cs_dev) : 1.f));
}

// ------------------------------------------------------------------------------------ D residual gate
// y = s + gamma * r, r = lrelu(pre)  (discriminator/residual_block.py:30).  Backward wrt the residual:
//   g_pre = g_y * gamma * lrelu'(r) ; dgamma += <g_y, r> ; colsum[c] += sum g_y (scale_conv bias grad)
__global__ void __launch_bounds__(kPwThreads, 4)
    dres_bwd_kernel(const __nv_bfloat16* __restrict__ g_y, const __nv_bfloat16* __restrict__ r,
                    __nv_bfloat16* __restrict__ g_pre, const float* gamma, float* dgamma, float* colsum,
                    int HW, int C, int pix_per_block, float slope) {
  __shared__ float red[kPwThreads * 8];
  const int n = blockIdx.y;
  const int tpp = C >> 3;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp, ppi = kPwThreads / tpp;
  const float gm = __ldg(gamma);
  const long long base = static_cast<long long>(n) * HW * C;
  bf8 acc{};
  float dot = 0.f;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  constexpr int U = 2;
  for (int pb = p0 + pl; pb < p1; pb += U * ppi) {
    bf8 g[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p < p1) {
        const long long off = base + static_cast<long long>(p) * C + cg * 8;
        g[u] = ld8(g_y + off);
        rv[u] = ld8(r + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ppi;
      if (p < p1) {
        bf8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          dot += g[u].v[i] * rv[u].v[i];
          acc.v[i] += g[u].v[i];
          o.v[i] = g[u].v[i] * gm * dlrelu(rv[u].v[i], slope);
        }
        st8(g_pre + base + static_cast<long long>(p) * C + cg * 8, o);
      }
    }
  }
  if (colsum) block_reduce8(acc, red, tpp, colsum + cg * 8, 1.f);
  if (dgamma) block_reduce1(dot, red, dgamma, 1.f);
}

// out += scale * <a, b> over n8*8 bf16 elements
__global__ void __launch_bounds__(kPwThreads)
    dot_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n8,
               float* out, float scale) {
  __shared__ float red[kPwThreads / 32];
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(kPwThreads) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * kPwThreads) {
    const bf8 x = ld8(a + i * 8), y = ld8(b + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x.v[k] * y.v[k];
  }
  block_reduce1(acc, red, out, scale);
}

// ------------------------------------------------------------------------------------ image boundary
// Stem im2col (discriminator/model.py:13): out[n,h,w, ci*9+r*3+s] = rowscale[n] * img[n,ci,h+r-1,w+s-1],
// channels 27..31 zero.  One thread per pixel; reads are coalesced along w, each thread writes 64 bytes.
__global__ void __launch_bounds__(kPwThreads)
    im2col3_kernel(const float* __restrict__ img, const float* rowscale, __nv_bfloat16* __restrict__ out,
                   int H, int W) {
  const int n = blockIdx.y;
  const int p = blockIdx.x * kPwThreads + threadIdx.x;
  if (p >= H * W) return;
  const int h = p / W, w = p - h * W;
  const float sc = rowscale ? __ldg(rowscale + n) : 1.f;
  const float* base = img + static_cast<long long>(n) * 3 * H * W;
  float v[32];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        // unconditional loads from clamped coordinates (all 27 stay in flight), zeroed afterwards where the tap
        // falls into the padding
        const int hh = h + r - 1, ww = w + s - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const int hc = min(max(hh, 0), H - 1), wc = min(max(ww, 0), W - 1);
        const float x = __ldg(base + (static_cast<long long>(ci) * H + hc) * W + wc);
        v[ci * 9 + r * 3 + s] = in ? sc * x : 0.f;
      }
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
  __nv_bfloat16* dst = out + (static_cast<long long>(n) * H * W + p) * 32;
  st16(dst, v);
  st16(dst + 16, v + 16);
}

// Backward of tanh at the generator output (generator/model.py:29): g_t = g_img * (1 - img^2), packed as
// NHWC bf16 with 32 channels (3 real); dbias[c] += sum g_t.
__global__ void __launch_bounds__(kPwThreads)
    tanh_bwd_pack_kernel(const float* __restrict__ g_img, const float* __restrict__ img,
                         __nv_bfloat16* __restrict__ out, float* dbias, int HW) {
  __shared__ float red[kPwThreads / 32];
  const int n = blockIdx.y;
  const int p = blockIdx.x * kPwThreads + threadIdx.x;
  float g[3] = {0.f, 0.f, 0.f};
  if (p < HW) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long o = (static_cast<long long>(n) * 3 + c) * HW + p;
      const float y = __ldg(img + o);
      g[c] = __ldg(g_img + o) * (1.f - y * y);
    }
    __nv_bfloat16* dst = out + (static_cast<long long>(n) * HW + p) * 32;
    float t[16] = {g[0], g[1], g[2], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    st16(dst, t);
    const float z[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    st16(dst + 16, z);
  }
  if (dbias) {
    // the bias gradient uses the bf16-rounded values the weight-gradient GEMM will see
    for (int c = 0; c < 3; ++c)
      block_reduce1(__bfloat162float(__float2bfloat16_rn(g[c])), red, dbias + c, 1.f);
  }
}

// fp32 [N][C][HW] <-> bf16 [N][HW][C]   (linear_in.view at generator/model.py:37, embeds at the
// build_embeds / get_logits API boundary).  Small tensors only: one element per thread.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C,
                                    int HW, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = i % C;
  const long long t = i / C;
  const int p = t % HW;
  const long long n = t / HW;
  out[i] = __float2bfloat16_rn(__ldg(in + (n * C + c) * HW + p));
}
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int C,
                                    int HW, long long total, int accumulate) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int p = i % HW;
  const long long t = i / HW;
  const int c = t % C;
  const long long n = t / C;
  const float v = __bfloat162float(in[(n * HW + p) * C + c]);
  out[i] = accumulate ? out[i] + v : v;
}

// per-row sum of squares: out[n] += sum_i x[n][i]^2   (deep_fusion_gan/model.py:82)
__global__ void __launch_bounds__(kPwThreads)
    rowsq_kernel(const float* __restrict__ x, long long row_len, float* out) {
  __shared__ float red[kPwThreads / 32];
  const int n = blockIdx.y;
  const float* row = x + n * row_len;
  float acc = 0.f;
  for (long long i = (blockIdx.x * static_cast<long long>(kPwThreads) + threadIdx.x) * 4; i < row_len;
       i += static_cast<long long>(gridDim.x) * kPwThreads * 4) {
    if (i + 3 < row_len) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + i));
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < row_len; ++j) acc += row[j] * row[j];
    }
  }
  block_reduce1(acc, red, out + n, 1.f);
}

// ------------------------------------------------------------------------------------ logit head
// cat[b,p,0:Ce] = embed[b,:,p] (fp32 NCHW), cat[b,p,Ce:Ce+Cs] = sent[b,:]*sscale[b]  (discriminator/model.py:49-54)
__global__ void head_cat_kernel(const float* __restrict__ embed, const float* __restrict__ sent,
                                const float* sscale, __nv_bfloat16* __restrict__ cat, int Ce, int Cs,
                                long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Ct = Ce + Cs;
  const int c = i % Ct;
  const long long t = i / Ct;
  const int p = t % 16;
  const long long b = t / 16;
  float v;
  if (c < Ce)
    v = __ldg(embed + (b * Ce + c) * 16 + p);
  else
    v = __ldg(sent + b * Cs + (c - Ce)) * (sscale ? __ldg(sscale + b) : 1.f);
  cat[i] = __float2bfloat16_rn(v);
}
// g_embed[b,c,p] = g_cat[b,p,c] ; g_sent[b,c] = sum_p g_cat[b,p,Ce+c]
__global__ void head_cat_bwd_kernel(const __nv_bfloat16* __restrict__ g_cat, float* __restrict__ g_embed,
                                    float* __restrict__ g_sent, int Ce, int Cs, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Ct = Ce + Cs;
  const long long per_b = static_cast<long long>(Ce) * 16 + Cs;
  const long long b = i / per_b;
  const long long r = i - b * per_b;
  const __nv_bfloat16* src = g_cat + b * 16 * Ct;
  if (r < static_cast<long long>(Ce) * 16) {
    const int c = r / 16, p = r % 16;
    g_embed[(b * Ce + c) * 16 + p] = __bfloat162float(src[p * Ct + c]);
  } else if (g_sent) {
    const int c = r - static_cast<long long>(Ce) * 16;
    float s = 0.f;
    for (int p = 0; p < 16; ++p) s += __bfloat162float(src[p * Ct + Ce + c]);
    g_sent[b * Cs + c] = s;
  }
}
// logits[b] = sum_{p,c} hid[b,p,c] * w2[0,c,p]       (discriminator/model.py:33; hid is post-LReLU)
__global__ void __launch_bounds__(kPwThreads)
    head_out_kernel(const __nv_bfloat16* __restrict__ hid, const float* __restrict__ w2, float* logits,
                    int Ch) {
  __shared__ float red[kPwThreads / 32];
  const int b = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < 16 * Ch; i += kPwThreads) {
    const int p = i / Ch, c = i - p * Ch;
    acc += __bfloat162float(hid[static_cast<long long>(b) * 16 * Ch + i]) * __ldg(w2 + c * 16 + p);
  }
  if (threadIdx.x == 0) logits[b] = 0.f;
  block_reduce1(acc, red, logits + b, 1.f);
}
// g_hid[b,p,c] = g_logit[b] * w2[c,p] * lrelu'(hid[b,p,c]) ;  dw2[c,p] += sum_b g_logit[b] * act[b,p,c]
__global__ void head_out_bwd_kernel(const float* __restrict__ g_logit, const __nv_bfloat16* __restrict__ hid,
                                    const __nv_bfloat16* __restrict__ act, const float* __restrict__ w2,
                                    __nv_bfloat16* __restrict__ g_hid, float* dw2, int B, int Ch, float slope) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 16*Ch
  if (i >= 16 * Ch) return;
  const int p = i / Ch, c = i - p * Ch;
  const float w = __ldg(w2 + c * 16 + p);
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const long long o = static_cast<long long>(b) * 16 * Ch + i;
    const float gl = g_logit ? __ldg(g_logit + b) : 1.f;
    if (g_hid) g_hid[o] = __float2bfloat16_rn(gl * w * dlrelu(__bfloat162float(hid[o]), slope));
    if (dw2) acc += gl * __bfloat162float(act[o]);
  }
  if (dw2) atomicAdd(dw2 + c * 16 + p, acc);
}

static inline int blocks_for(long long total, int threads) {
  return static_cast<int>((total + threads - 1) / threads);
}
static inline bool c_ok(int C) { return C >= 8 && C % 8 == 0 && (kPwThreads % (C / 8)) == 0; }
static inline bool c4_ok(int C) { return C >= 4 && C % 4 == 0 && C / 4 <= kPwThreads && (kPwThreads % (C / 4)) == 0; }
// Pixels per block such that the grid is an exact number of resident waves: `bps` blocks per SM, rounded DOWN per
// image -- a grid that spills a few blocks into an extra wave costs a whole extra pass (measured with
// tools/prof_pointwise.py: pool 46 -> 36 us and 49 -> 27 us, dres_bwd 74 -> 58 us at 32x128x128x64 once the grid fits).
// Per kernel: bps = its occupancy (one wave of fat blocks) or twice that; DFGAN_PW_BPS overrides for experiments.
static inline int pick_ppb(int N, int HW, int C, int bps) {
  const int ppi = kPwThreads / (C / 8);
  static const int bps_env = getenv("DFGAN_PW_BPS") ? atoi(getenv("DFGAN_PW_BPS")) : 0;
  long long want_blocks = static_cast<long long>(bps_env > 0 ? bps_env : bps) * num_sms();
  long long per_n = want_blocks / N;   // floor: the grid must not spill a few blocks into an extra wave
  if (per_n < 1) per_n = 1;
  int ppb = static_cast<int>((HW + per_n - 1) / per_n);
  ppb = ((ppb + ppi - 1) / ppi) * ppi;
  if (ppb < ppi * 4) ppb = ppi * 4;
  return ppb;
}


// ------------------------------------------------------------------------------------ small fused scalar math
// (each of these replaced 5-25 two-microsecond torch launches that sat on the critical path of the CUDA graph)
//
// Generator block, small parameter gradients from the per-channel sums cs[c] = sum_pixels g_z[., c]
// (generator/residual_block.py:47,59 and :29-33):  d conv_2.bias += gamma * cs ;  d scale_conv.bias += cs ;
// d gamma += <cs, conv_2.bias>   (the bias part of d/dgamma <g_z, gamma * (conv_2(h) + bias)>)
__global__ void __launch_bounds__(256)
    gblock_small_grads_kernel(const float* __restrict__ cs, const float* __restrict__ gamma,
                              const float* __restrict__ bias2, float* g_bias2, float* g_gamma, float* g_sbias, int C) {
  __shared__ float red[8];
  float part = 0.f;
  const float gm = __ldg(gamma);
  for (int c = threadIdx.x; c < C; c += 256) {
    const float v = __ldg(cs + c);
    g_bias2[c] += gm * v;
    if (g_sbias) g_sbias[c] += v;
    part += v * __ldg(bias2 + c);
  }
  block_reduce1(part, red, g_gamma, 1.f);
}

// Discriminator hinge loss and its logit cotangents (deep_fusion_gan/model.py:175-189) from the concatenated
// logits [real (B) ; mismatched (B-1) ; fake (B)]:
//   loss = mean relu(1 - real) + (mean relu(1 + fake) + mean relu(1 + mis)) / 2
//   g[real] = -(real < 1) / B ; g[mis] = (mis > -1) / (2 (B-1)) ; g[fake] = (fake > -1) / (2 B)
__global__ void __launch_bounds__(256) d_hinge_kernel(const float* __restrict__ lg, float* loss, float* g, int B) {
  __shared__ float red[8];
  float part = 0.f;
  const int n = 3 * B - 1;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = __ldg(lg + i);
    if (i < B) {
      part += fmaxf(1.f - v, 0.f) / B;
      g[i] = v < 1.f ? -1.f / B : 0.f;
    } else if (i < 2 * B - 1) {
      part += 0.5f * fmaxf(1.f + v, 0.f) / (B - 1);
      g[i] = v > -1.f ? 0.5f / (B - 1) : 0.f;
    } else {
      part += 0.5f * fmaxf(1.f + v, 0.f) / B;
      g[i] = v > -1.f ? 0.5f / B : 0.f;
    }
  }
  if (threadIdx.x == 0) *loss = 0.f;
  block_reduce1(part, red, loss, 1.f);
}

// MA-GP loss and the per-sample coefficient of its gradient (deep_fusion_gan/model.py:82-85, 203):
//   n_b = sqrt(sq_b + 1e-8) ; loss = 2 * mean(clamp(n_b, 0, 10)^6) ; coef_b = (12 / B) n_b^4 [n_b <= 10]
__global__ void __launch_bounds__(256) gp_coef_kernel(const float* __restrict__ sq, float* loss, float* coef, int B) {
  __shared__ float red[8];
  float part = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float n = sqrtf(__ldg(sq + i) + 1e-8f);
    const float nc = fminf(fmaxf(n, 0.f), 10.f);
    const float n2 = nc * nc;
    part += 2.f * n2 * n2 * n2 / B;
    const float m2 = n * n;
    coef[i] = n <= 10.f ? (12.f / B) * m2 * m2 : 0.f;
  }
  if (threadIdx.x == 0) *loss = 0.f;
  block_reduce1(part, red, loss, 1.f);
}

// Text-image loss the trainer logs every iteration (damsm/loss.py:4-26 via deep_fusion_gan/model.py:87-89, 220-223;
// it never enters a gradient): 1 - mean_b cos(pad(mean_hw(img_b)), sent_b), where pad places the C channel means in
// the first C of the E sentence dimensions (F.linear with eye(E, C), :17-18) and cos is the product of the two
// F.normalize'd vectors (eps 1e-12, :20-23).
// Stage 1: partial[(plane * kTxtSplit + s)] = sum of slice s of plane (b, c) of the fp32 NCHW image.  Every partial
// is written by exactly one block and summed in index order by stage 2, so the result is bit-reproducible.
constexpr int kTxtSplit = 4;
__global__ void __launch_bounds__(kPwThreads)
    plane_sum_kernel(const float* __restrict__ x, long long plane_len, float* __restrict__ partial) {
  __shared__ float red[kPwThreads / 32];
  const long long slice = plane_len / kTxtSplit;          // plane_len is a multiple of 4 * kTxtSplit (host check)
  const float* src = x + blockIdx.y * plane_len + blockIdx.x * slice;
  float acc = 0.f;
  for (long long i = threadIdx.x * 4ll; i < slice; i += kPwThreads * 4ll) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    acc += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kPwThreads / 32; ++i) s += red[i];
    partial[blockIdx.y * kTxtSplit + blockIdx.x] = s;
  }
}
// Stage 2: one block; warp w owns samples w, w + 8, ...
__global__ void __launch_bounds__(256)
    txtimg_loss_kernel(const float* __restrict__ partial, float inv_plane_len, int C, const float* __restrict__ sent,
                       int E, int B, float* loss) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nc = C < E ? C : E;
  float wsum = 0.f;
  for (int b = w; b < B; b += 8) {
    const float* s = sent + static_cast<long long>(b) * E;
    float ss = 0.f;
    for (int e = lane; e < E; e += 32) {
      const float v = __ldg(s + e);
      ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float mm = 0.f, ms = 0.f;
    for (int c = 0; c < nc; ++c) {
      const float* pp = partial + (static_cast<long long>(b) * C + c) * kTxtSplit;
      float m = 0.f;
      for (int k = 0; k < kTxtSplit; ++k) m += __ldg(pp + k);
      m *= inv_plane_len;
      mm += m * m;
      ms += m * __ldg(s + c);
    }
    wsum += ms / (fmaxf(sqrtf(mm), 1e-12f) * fmaxf(sqrtf(ss), 1e-12f));
  }
  if (lane == 0) red[w] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    *loss = 1.f - t / B;
  }
}

}  // namespace dfgan

using namespace dfgan;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" {

int dfgan_affine2_fwd(const void* x, void* out, const float* ga, const float* ba, const float* gb,
                      const float* bb, int64_t tab_stride, int N, int HW, int C, dfgan_stream_t st) {
  if (!x || !out || !ga || !ba || !gb || !bb || !c_ok(C) || N < 1 || HW < 1) return DFGAN_E_ARG;
  const int ppi = kPwThreads / (C / 8);
  int gx = (HW + ppi - 1) / ppi;
  const int cap = (8 * num_sms() + N - 1) / N;
  if (gx > cap) gx = cap;
  affine2_fwd_kernel<<<dim3(gx, N), kPwThreads, 0, STREAM(st)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), ga, ba, gb, bb, tab_stride, HW,
      C, 0.2f);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_affine2_bwd(const void* g_a, const void* x, const void* g_add, void* gx, const float* ga,
                      const float* ba, const float* gb, const float* bb, float* d_ga, float* d_ba,
                      float* d_gb, float* d_bb, int64_t tab_stride, const float* gscale, float* dot,
                      float* colsum, int N, int HW, int C, dfgan_stream_t st) {
  if (!g_a || !x || !gx || !ga || !ba || !gb || !bb || !d_ga || !d_ba || !d_gb || !d_bb || !c4_ok(C))
    return DFGAN_E_ARG;
  const int ppb = pick_ppb(N, HW, 2 * C, 6);   // C/4 threads per pixel -> 1024/C pixels per iteration; occupancy 3
  affine2_bwd_kernel<<<dim3((HW + ppb - 1) / ppb, N), kPwThreads, 0, STREAM(st)>>>(
      static_cast<const __nv_bfloat16*>(g_a), static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(g_add), static_cast<__nv_bfloat16*>(gx), ga, ba, gb, bb, d_ga, d_ba,
      d_gb, d_bb, tab_stride, gscale, dot, colsum, HW, C, ppb, 0.2f);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_pool(const void* x, void* out, float* colsum, const float* cs_dev, float cs_imm, int N, int Ho,
               int Wo, int C, int shift, float scale, dfgan_stream_t st) {
  if (!x || (!out && !colsum) || !c_ok(C) || shift < 0 || shift > 1) return DFGAN_E_ARG;
  const int ppb = pick_ppb(N, Ho * Wo, C, 3);
  const dim3 grid((Ho * Wo + ppb - 1) / ppb, N);
  if (shift)
    pool_kernel<1, 1><<<grid, kPwThreads, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(x),
                                                           static_cast<__nv_bfloat16*>(out), colsum, cs_dev, cs_imm,
                                                           Ho, Wo, C, scale, ppb);
  else
    pool_kernel<0, 4><<<grid, kPwThreads, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(x),
                                                           static_cast<__nv_bfloat16*>(out), colsum, cs_dev, cs_imm,
                                                           Ho, Wo, C, scale, ppb);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_dres_bwd(const void* g_y, const void* r, void* g_pre, const float* gamma, float* dgamma,
                   float* colsum, int N, int HW, int C, dfgan_stream_t st) {
  if (!g_y || !r || !g_pre || !gamma || !c_ok(C)) return DFGAN_E_ARG;
  const int ppb = pick_ppb(N, HW, C, 4);
  dres_bwd_kernel<<<dim3((HW + ppb - 1) / ppb, N), kPwThreads, 0, STREAM(st)>>>(
      static_cast<const __nv_bfloat16*>(g_y), static_cast<const __nv_bfloat16*>(r),
      static_cast<__nv_bfloat16*>(g_pre), gamma, dgamma, colsum, HW, C, ppb, 0.2f);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_dot(const void* a, const void* b, int64_t n, float* out, float scale, dfgan_stream_t st) {
  if (!a || !b || !out || n < 8 || n % 8) return DFGAN_E_ARG;
  int g = blocks_for(n / 8, kPwThreads);
  if (g > 4 * num_sms()) g = 4 * num_sms();
  dot_kernel<<<g, kPwThreads, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(a),
                                               static_cast<const __nv_bfloat16*>(b), n / 8, out, scale);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_im2col3(const float* img, const float* rowscale, void* out, int N, int H, int W, dfgan_stream_t st) {
  if (!img || !out) return DFGAN_E_ARG;
  im2col3_kernel<<<dim3(blocks_for(static_cast<long long>(H) * W, kPwThreads), N), kPwThreads, 0, STREAM(st)>>>(
      img, rowscale, static_cast<__nv_bfloat16*>(out), H, W);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_tanh_bwd_pack(const float* g_img, const float* img, void* out, float* dbias, int N, int HW,
                        dfgan_stream_t st) {
  if (!g_img || !img || !out) return DFGAN_E_ARG;
  tanh_bwd_pack_kernel<<<dim3(blocks_for(HW, kPwThreads), N), kPwThreads, 0, STREAM(st)>>>(
      g_img, img, static_cast<__nv_bfloat16*>(out), dbias, HW);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_nchw_to_nhwc(const float* in, void* out, int N, int C, int HW, dfgan_stream_t st) {
  if (!in || !out) return DFGAN_E_ARG;
  const long long total = static_cast<long long>(N) * C * HW;
  nchw_to_nhwc_kernel<<<blocks_for(total, 256), 256, 0, STREAM(st)>>>(in, static_cast<__nv_bfloat16*>(out), C,
                                                                      HW, total);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_nhwc_to_nchw(const void* in, float* out, int N, int C, int HW, int accumulate, dfgan_stream_t st) {
  if (!in || !out) return DFGAN_E_ARG;
  const long long total = static_cast<long long>(N) * C * HW;
  nhwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(in),
                                                                      out, C, HW, total, accumulate);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_rowsq(const float* x, int N, int64_t row_len, float* out, dfgan_stream_t st) {
  if (!x || !out || (row_len & 3)) return DFGAN_E_ARG;
  int gx = blocks_for(row_len / 4, kPwThreads);
  const int cap = (4 * num_sms() + N - 1) / N;
  if (gx > cap) gx = cap;
  rowsq_kernel<<<dim3(gx, N), kPwThreads, 0, STREAM(st)>>>(x, row_len, out);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_head_cat(const float* embed, const float* sent, const float* sscale, void* cat, int B, int Ce,
                   int Cs, dfgan_stream_t st) {
  if (!embed || !sent || !cat) return DFGAN_E_ARG;
  const long long total = static_cast<long long>(B) * 16 * (Ce + Cs);
  head_cat_kernel<<<blocks_for(total, 256), 256, 0, STREAM(st)>>>(embed, sent, sscale,
                                                                  static_cast<__nv_bfloat16*>(cat), Ce, Cs, total);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_head_cat_bwd(const void* g_cat, float* g_embed, float* g_sent, int B, int Ce, int Cs,
                       dfgan_stream_t st) {
  if (!g_cat || !g_embed) return DFGAN_E_ARG;
  const long long total = static_cast<long long>(B) * (static_cast<long long>(Ce) * 16 + Cs);
  head_cat_bwd_kernel<<<blocks_for(total, 256), 256, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(g_cat),
                                                                      g_embed, g_sent, Ce, Cs, total);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_head_out(const void* hid, const float* w2, float* logits, int B, int Ch, dfgan_stream_t st) {
  if (!hid || !w2 || !logits || B < 1) return DFGAN_E_ARG;
  head_out_kernel<<<B, kPwThreads, 0, STREAM(st)>>>(static_cast<const __nv_bfloat16*>(hid), w2, logits, Ch);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_head_out_bwd(const float* g_logit, const void* hid, const void* act, const float* w2, void* g_hid,
                       float* dw2, int B, int Ch, dfgan_stream_t st) {
  if (!w2 || (!g_hid && !dw2) || (g_hid && !hid) || (dw2 && !act)) return DFGAN_E_ARG;
  head_out_bwd_kernel<<<blocks_for(16 * Ch, 128), 128, 0, STREAM(st)>>>(
      g_logit, static_cast<const __nv_bfloat16*>(hid), static_cast<const __nv_bfloat16*>(act), w2,
      static_cast<__nv_bfloat16*>(g_hid), dw2, B, Ch, 0.2f);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}


int dfgan_gblock_small_grads(const float* cs, const float* gamma, const float* bias2, float* g_bias2, float* g_gamma,
                             float* g_sbias, int C, dfgan_stream_t st) {
  if (!cs || !gamma || !bias2 || !g_bias2 || !g_gamma || C < 1) return DFGAN_E_ARG;
  gblock_small_grads_kernel<<<1, 256, 0, STREAM(st)>>>(cs, gamma, bias2, g_bias2, g_gamma, g_sbias, C);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_d_hinge(const float* logits, float* loss, float* g_logits, int B, dfgan_stream_t st) {
  if (!logits || !loss || !g_logits || B < 2) return DFGAN_E_ARG;
  d_hinge_kernel<<<1, 256, 0, STREAM(st)>>>(logits, loss, g_logits, B);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_gp_coef(const float* sq, float* loss, float* coef, int B, dfgan_stream_t st) {
  if (!sq || !loss || !coef || B < 1) return DFGAN_E_ARG;
  gp_coef_kernel<<<1, 256, 0, STREAM(st)>>>(sq, loss, coef, B);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_txtimg_loss(const float* img, int B, int C, int64_t plane_len, const float* sent, int E, float* partial,
                      int partial_floats, float* loss, dfgan_stream_t st) {
  if (!img || !sent || !partial || !loss || B < 1 || C < 1 || E < 1 || plane_len < 1) return DFGAN_E_ARG;
  if (plane_len % (4 * kTxtSplit) || (reinterpret_cast<uintptr_t>(img) & 15)) return DFGAN_E_ARG;
  if (static_cast<int64_t>(partial_floats) < static_cast<int64_t>(B) * C * kTxtSplit) return DFGAN_E_ARG;
  if (static_cast<int64_t>(B) * C > 65535) return DFGAN_E_SHAPE;        // planes are the grid's y dimension
  plane_sum_kernel<<<dim3(kTxtSplit, B * C), kPwThreads, 0, STREAM(st)>>>(img, plane_len, partial);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  txtimg_loss_kernel<<<1, 256, 0, STREAM(st)>>>(partial, 1.f / static_cast<float>(plane_len), C, sent, E, B, loss);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

}  // extern "C"
