// Frozen text encoder: RNNEncoder.forward of text_encoder/model.py:33-56 in eval mode --
// embedding lookup, one bidirectional LSTM layer over a length-packed batch, sentence embedding = the two final
// hidden states side by side.  It sits on the critical path of every training iteration (deep_fusion_gan/model.py:171)
// and is tiny (0.5 GFLOP at B=32, T=18): 18 dependent time steps, so it is latency- not throughput-bound.  Two launches:
//
//   lstm_inproj_kernel : gates_x[dir][b*T+t][4H] = emb[token[b,t]] . W_ih[dir]^T + b_ih[dir] + b_hh[dir] for every
//                        (b, t) at once -- the part of the LSTM without a time dependence, as one gathered fp32 GEMM;
//   lstm_recur_kernel  : one CTA per (direction, sentence), one thread per gate row.  The thread keeps its W_hh row
//                        on chip for the whole sequence (64 weights in registers, 64 in shared memory), so a time
//                        step is 128 FMAs against the broadcast hidden state + two block barriers; nothing but the
//                        pre-computed input gates (prefetched one step ahead) is read from global memory in the loop.
//
// fp32 throughout with precise expf / tanhf: the encoder is frozen and its output feeds every loss, so it is kept
// at the reference's precision (parity 1e-5 against the fp32 oracle, tests in tools/bringup_text.py).
// Gate order i, f, g, o and  c' = f c + i g,  h' = o tanh(c')  as in torch.nn.LSTM.  Packed-sequence semantics
// (text_encoder/model.py:39-44): direction 0 runs t = 0 .. len-1, direction 1 runs t = len-1 .. 0, both from a
// zero state; positions >= len are never touched.
#include "host_common.h"

namespace dfgan {

constexpr int kH = 128;        // hidden units per direction (RNNEncoder.load builds nhidden=256 over two directions)
constexpr int kG = 4 * kH;     // gate rows per direction
constexpr int kRegK = 64;      // leading part of a W_hh row held in registers; the rest lives in shared memory
constexpr int kPM = 32, kPN = 64, kPK = 32;   // input-projection tile: (b,t) rows x gate columns x K slice

__global__ void __launch_bounds__(256)
    lstm_inproj_kernel(const long long* __restrict__ tokens, const float* __restrict__ emb,
                       const float* __restrict__ wih_f, const float* __restrict__ wih_r,
                       const float* __restrict__ bih_f, const float* __restrict__ bhh_f,
                       const float* __restrict__ bih_r, const float* __restrict__ bhh_r, float* __restrict__ gx,
                       int M, int I, long long ntoken) {
  __shared__ float s_a[kPM][kPK + 1];
  __shared__ float s_w[kPK][kPN + 1];
  __shared__ long long s_tok[kPM];
  const int tid = threadIdx.x, lane = tid % 32, wy = tid / 32;
  const int m0 = blockIdx.x * kPM, n0 = blockIdx.y * kPN;
  const int dir = n0 / kG, c0 = n0 % kG;
  const float* __restrict__ w = dir ? wih_r : wih_f;
  if (tid < kPM) {
    const long long tok = (m0 + tid < M) ? tokens[m0 + tid] : -1;
    s_tok[tid] = (tok >= 0 && tok < ntoken) ? tok : -1;      // out-of-range ids read as a zero embedding
  }
  __syncthreads();
  float acc[4][2] = {};
  for (int k0 = 0; k0 < I; k0 += kPK) {
#pragma unroll
    for (int j = 0; j < kPM * kPK / 256; ++j) {
      const int idx = tid + j * 256, r = idx / kPK, k = idx % kPK;
      const long long tok = s_tok[r];
      s_a[r][k] = (tok >= 0 && k0 + k < I) ? __ldg(emb + tok * I + k0 + k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kPN * kPK / 256; ++j) {
      const int idx = tid + j * 256, c = idx / kPK, k = idx % kPK;
      s_w[k][c] = (k0 + k < I) ? __ldg(w + static_cast<long long>(c0 + c) * I + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPK; ++k) {
      const float w0 = s_w[k][lane], w1 = s_w[k][lane + 32];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = s_a[wy * 4 + r][k];
        acc[r][0] = fmaf(a, w0, acc[r][0]);
        acc[r][1] = fmaf(a, w1, acc[r][1]);
      }
    }
    __syncthreads();
  }
  const float* __restrict__ bi = dir ? bih_r : bih_f;
  const float* __restrict__ bh = dir ? bhh_r : bhh_f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = c0 + lane + 32 * j;
    const float bias = __ldg(bi + c) + __ldg(bh + c);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + wy * 4 + r;
      if (m < M) gx[(static_cast<long long>(dir) * M + m) * kG + c] = acc[r][j] + bias;
    }
  }
}

__global__ void __launch_bounds__(kG, 1)
    lstm_recur_kernel(const float* __restrict__ gx, const float* __restrict__ whh_f, const float* __restrict__ whh_r,
                      const long long* __restrict__ lens, float* __restrict__ sent, int B, int T) {
  extern __shared__ float4 s_w4[];               // [(kH - kRegK) / 4][kG]: thread `row` reads s_w4[q * kG + row]
  __shared__ __align__(16) float s_h[kH];
  __shared__ float s_act[kG];
  const int dir = blockIdx.x / B, b = blockIdx.x % B, row = threadIdx.x;
  const float* __restrict__ w = (dir ? whh_r : whh_f) + static_cast<long long>(row) * kH;
  float wr[kRegK];
#pragma unroll
  for (int k = 0; k < kRegK; k += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + k));
    wr[k] = t.x, wr[k + 1] = t.y, wr[k + 2] = t.z, wr[k + 3] = t.w;
  }
#pragma unroll
  for (int q = 0; q < (kH - kRegK) / 4; ++q) s_w4[q * kG + row] = __ldg(reinterpret_cast<const float4*>(w + kRegK) + q);
  if (row < kH) s_h[row] = 0.f;
  long long len64 = lens[b];
  const int len = static_cast<int>(len64 < 0 ? 0 : (len64 > T ? T : len64));
  const float* __restrict__ gxb = gx + (static_cast<long long>(dir) * B * T + static_cast<long long>(b) * T) * kG + row;
  float c = 0.f, h = 0.f;
  float pre_next = len > 0 ? __ldg(gxb + static_cast<long long>(dir ? len - 1 : 0) * kG) : 0.f;
  __syncthreads();
  const bool is_g = row >= 2 * kH && row < 3 * kH;     // warp-uniform: the cell-candidate rows take tanh
  for (int s = 0; s < len; ++s) {
    const int t = dir ? len - 1 - s : s;
    const float pre = pre_next;
    if (s + 1 < len) pre_next = __ldg(gxb + static_cast<long long>(dir ? t - 1 : t + 1) * kG);
    float a0 = pre, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float4* h4 = reinterpret_cast<const float4*>(s_h);
#pragma unroll
    for (int k = 0; k < kRegK; k += 4) {
      const float4 hv = h4[k / 4];
      a0 = fmaf(wr[k], hv.x, a0), a1 = fmaf(wr[k + 1], hv.y, a1);
      a2 = fmaf(wr[k + 2], hv.z, a2), a3 = fmaf(wr[k + 3], hv.w, a3);
    }
#pragma unroll
    for (int q = 0; q < (kH - kRegK) / 4; ++q) {
      const float4 wv = s_w4[q * kG + row], hv = h4[kRegK / 4 + q];
      a0 = fmaf(wv.x, hv.x, a0), a1 = fmaf(wv.y, hv.y, a1);
      a2 = fmaf(wv.z, hv.z, a2), a3 = fmaf(wv.w, hv.w, a3);
    }
    const float z = (a0 + a1) + (a2 + a3);
    s_act[row] = is_g ? tanhf(z) : 1.f / (1.f + expf(-z));
    __syncthreads();
    if (row < kH) {
      c = fmaf(s_act[kH + row], c, s_act[row] * s_act[2 * kH + row]);
      h = s_act[3 * kH + row] * tanhf(c);
      s_h[row] = h;
    }
    __syncthreads();
  }
  if (row < kH) sent[static_cast<long long>(b) * 2 * kH + dir * kH + row] = h;
}

}  // namespace dfgan

using namespace dfgan;

extern "C" {

size_t dfgan_text_encoder_workspace_bytes(int B, int T) {
  if (B < 1 || T < 1) return 0;
  return sizeof(float) * 2 * static_cast<size_t>(B) * T * kG;
}

int dfgan_text_encoder_fwd(const int64_t* tokens, const int64_t* lens, const float* emb, int64_t ntoken,
                           const float* const* w_ih, const float* const* w_hh, const float* const* b_ih,
                           const float* const* b_hh, float* sent, void* workspace, size_t ws_bytes, int B, int T,
                           int I, int H, dfgan_stream_t st) {
  if (!tokens || !lens || !emb || !w_ih || !w_hh || !b_ih || !b_hh || !sent || !workspace) return DFGAN_E_ARG;
  if (B < 1 || T < 1 || I < 1 || ntoken < 1) return DFGAN_E_ARG;
  if (H != kH) return DFGAN_E_SHAPE;                 // one hidden width: the reference's only configuration
  if (ws_bytes < dfgan_text_encoder_workspace_bytes(B, T)) return DFGAN_E_ARG;
  for (int d = 0; d < 2; ++d) {
    if (!w_ih[d] || !w_hh[d] || !b_ih[d] || !b_hh[d]) return DFGAN_E_ARG;
    if (reinterpret_cast<uintptr_t>(w_hh[d]) & 15) return DFGAN_E_ARG;       // rows are read as float4
  }
  float* gx = static_cast<float*>(workspace);
  const int M = B * T;
  const dim3 grid((M + kPM - 1) / kPM, 2 * kG / kPN);
  lstm_inproj_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(
      reinterpret_cast<const long long*>(tokens), emb, w_ih[0], w_ih[1], b_ih[0], b_hh[0], b_ih[1], b_hh[1], gx, M, I,
      static_cast<long long>(ntoken));
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  constexpr int smem = (kH - kRegK) / 4 * kG * static_cast<int>(sizeof(float4));
  cudaError_t e = cudaFuncSetAttribute(lstm_recur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  lstm_recur_kernel<<<2 * B, kG, smem, static_cast<cudaStream_t>(st)>>>(
      gx, w_hh[0], w_hh[1], reinterpret_cast<const long long*>(lens), sent, B, T);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

}  // extern "C"
