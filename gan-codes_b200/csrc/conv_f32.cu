// fp32 precision tier: plain CUDA-core convolution (forward, data gradient, weight gradient) on NCHW fp32
// tensors with OIHW weights -- the layouts of the reference's nn.Conv2d (generator/residual_block.py:16,21,25;
// discriminator/residual_block.py:11,13,19; discriminator/model.py:13,31,33; generator/model.py:13,28).  No
// tensor cores, no reduced-precision storage: every product and sum is fp32 FMA, so results agree with the
// reference's fp32 (TF32-off) path to rounding-order noise (north_star's 1e-4 tier).  This is the verification
// tier, not the fast path: the three operators are closed under differentiation (each one's gradient is another
// one of the three), which is what the MA-GP double backward needs.
#include "host_common.h"

namespace dfgan {

constexpr int kT = 16;       // output tile edge (threads per block = kT * kT)
constexpr int kCoT = 16;     // output channels per block (register accumulators per thread)
constexpr int kCiT = 8;      // input channels staged per iteration

struct ConvF32 {
  int N, Cin, H, W, Cout, KH, KW, stride, pad, OH, OW;
};

// y[n,co,oh,ow] = bias[co] + sum_{ci,r,s} x[n,ci,oh*st-pad+r,ow*st-pad+s] * w[co,ci,r,s]
__global__ void __launch_bounds__(kT * kT)
    conv_f32_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ y, ConvF32 p) {
  extern __shared__ float sm[];
  const int PH = (kT - 1) * p.stride + p.KH, PW = (kT - 1) * p.stride + p.KW;
  float* s_x = sm;                                   // [kCiT][PH][PW]
  float* s_w = sm + kCiT * PH * PW;                  // [kCiT][KH*KW][kCoT]
  const int tiles_w = (p.OW + kT - 1) / kT;
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
  const int co0 = blockIdx.y * kCoT, n = blockIdx.z;
  const int tx = threadIdx.x % kT, ty = threadIdx.x / kT;
  const int oh = th * kT + ty, ow = tw * kT + tx;
  const int ih0 = th * kT * p.stride - p.pad, iw0 = tw * kT * p.stride - p.pad;
  const int KK = p.KH * p.KW;
  float acc[kCoT];
#pragma unroll
  for (int j = 0; j < kCoT; ++j) acc[j] = 0.f;
  for (int ci0 = 0; ci0 < p.Cin; ci0 += kCiT) {
    __syncthreads();
    for (int i = threadIdx.x; i < kCiT * PH * PW; i += kT * kT) {
      const int c = i / (PH * PW), r = (i / PW) % PH, s = i % PW;
      const int ci = ci0 + c, ih = ih0 + r, iw = iw0 + s;
      s_x[i] = (ci < p.Cin && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
                   ? __ldg(x + ((static_cast<long long>(n) * p.Cin + ci) * p.H + ih) * p.W + iw)
                   : 0.f;
    }
    for (int i = threadIdx.x; i < kCiT * KK * kCoT; i += kT * kT) {
      const int j = i % kCoT, k = (i / kCoT) % KK, c = i / (kCoT * KK);
      const int ci = ci0 + c, co = co0 + j;
      s_w[i] = (ci < p.Cin && co < p.Cout) ? __ldg(w + (static_cast<long long>(co) * p.Cin + ci) * KK + k) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < kCiT; ++c)
      for (int r = 0; r < p.KH; ++r)
        for (int s = 0; s < p.KW; ++s) {
          const float xv = s_x[(c * PH + ty * p.stride + r) * PW + tx * p.stride + s];
          const float* wp = s_w + (c * KK + r * p.KW + s) * kCoT;
#pragma unroll
          for (int j = 0; j < kCoT; ++j) acc[j] = fmaf(xv, wp[j], acc[j]);
        }
  }
  if (oh < p.OH && ow < p.OW) {
#pragma unroll
    for (int j = 0; j < kCoT; ++j) {
      const int co = co0 + j;
      if (co < p.Cout)
        y[((static_cast<long long>(n) * p.Cout + co) * p.OH + oh) * p.OW + ow] = acc[j] + (bias ? __ldg(bias + co) : 0.f);
    }
  }
}

// gx[n,ci,ih,iw] = sum_{co,r,s : ih+pad-r = st*oh, iw+pad-s = st*ow} gy[n,co,oh,ow] * w[co,ci,r,s]
__global__ void __launch_bounds__(kT * kT)
    conv_f32_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx, ConvF32 p) {
  extern __shared__ float sm[];
  // output-gradient rows/cols that can reach a kT x kT tile of input pixels
  const int PH = (kT + p.KH - 2) / p.stride + 2, PW = (kT + p.KW - 2) / p.stride + 2;
  float* s_g = sm;                                   // [kCiT (co chunk)][PH][PW]
  float* s_w = sm + kCiT * PH * PW;                  // [kCiT (co)][KH*KW][kCoT (ci)]
  const int tiles_w = (p.W + kT - 1) / kT;
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
  const int ci0 = blockIdx.y * kCoT, n = blockIdx.z;
  const int tx = threadIdx.x % kT, ty = threadIdx.x / kT;
  const int ih = th * kT + ty, iw = tw * kT + tx;
  // first output row/col staged: floor((tile origin + pad - (K-1)) / stride), floor towards -inf
  const int a0 = th * kT + p.pad - (p.KH - 1), b0 = tw * kT + p.pad - (p.KW - 1);
  const int oh_base = a0 >= 0 ? a0 / p.stride : -((-a0 + p.stride - 1) / p.stride);
  const int ow_base = b0 >= 0 ? b0 / p.stride : -((-b0 + p.stride - 1) / p.stride);
  const int KK = p.KH * p.KW;
  float acc[kCoT];
#pragma unroll
  for (int j = 0; j < kCoT; ++j) acc[j] = 0.f;
  for (int co0 = 0; co0 < p.Cout; co0 += kCiT) {
    __syncthreads();
    for (int i = threadIdx.x; i < kCiT * PH * PW; i += kT * kT) {
      const int c = i / (PH * PW), r = (i / PW) % PH, s = i % PW;
      const int co = co0 + c, oh = oh_base + r, ow = ow_base + s;
      s_g[i] = (co < p.Cout && oh >= 0 && oh < p.OH && ow >= 0 && ow < p.OW)
                   ? __ldg(gy + ((static_cast<long long>(n) * p.Cout + co) * p.OH + oh) * p.OW + ow)
                   : 0.f;
    }
    for (int i = threadIdx.x; i < kCiT * KK * kCoT; i += kT * kT) {
      const int j = i % kCoT, k = (i / kCoT) % KK, c = i / (kCoT * KK);
      const int co = co0 + c, ci = ci0 + j;
      s_w[i] = (ci < p.Cin && co < p.Cout) ? __ldg(w + (static_cast<long long>(co) * p.Cin + ci) * KK + k) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < p.KH; ++r) {
      const int a = ih + p.pad - r;
      if (a < 0 || a % p.stride) continue;
      const int oh = a / p.stride - oh_base;
      if (oh < 0 || oh >= PH) continue;
      for (int s = 0; s < p.KW; ++s) {
        const int b = iw + p.pad - s;
        if (b < 0 || b % p.stride) continue;
        const int ow = b / p.stride - ow_base;
        if (ow < 0 || ow >= PW) continue;
        for (int c = 0; c < kCiT; ++c) {
          const float gv = s_g[(c * PH + oh) * PW + ow];
          const float* wp = s_w + (c * KK + r * p.KW + s) * kCoT;
#pragma unroll
          for (int j = 0; j < kCoT; ++j) acc[j] = fmaf(gv, wp[j], acc[j]);
        }
      }
    }
  }
  if (ih < p.H && iw < p.W) {
#pragma unroll
    for (int j = 0; j < kCoT; ++j) {
      const int ci = ci0 + j;
      if (ci < p.Cin) gx[((static_cast<long long>(n) * p.Cin + ci) * p.H + ih) * p.W + iw] = acc[j];
    }
  }
}

// gw[co,ci,r,s] = sum_{n,oh,ow} gy[n,co,oh,ow] * x[n,ci,oh*st-pad+r,ow*st-pad+s]; one block per (co, ci) pair,
// fixed reduction order (deterministic)
__global__ void __launch_bounds__(256)
    conv_f32_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, ConvF32 p) {
  __shared__ float red[8][16];
  const int ci = blockIdx.x, co = blockIdx.y;
  const int KK = p.KH * p.KW;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  const long long pix = static_cast<long long>(p.N) * p.OH * p.OW;
  for (long long i = threadIdx.x; i < pix; i += 256) {
    const int ow = static_cast<int>(i % p.OW), oh = static_cast<int>((i / p.OW) % p.OH);
    const int n = static_cast<int>(i / (static_cast<long long>(p.OW) * p.OH));
    const float g = __ldg(gy + ((static_cast<long long>(n) * p.Cout + co) * p.OH + oh) * p.OW + ow);
    const float* xp = x + (static_cast<long long>(n) * p.Cin + ci) * p.H * p.W;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < KK) {
        const int ih = oh * p.stride - p.pad + k / p.KW, iw = ow * p.stride - p.pad + k % p.KW;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) acc[k] = fmaf(g, __ldg(xp + static_cast<long long>(ih) * p.W + iw), acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < KK) {
    float v = 0.f;
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    gw[(static_cast<long long>(co) * p.Cin + ci) * KK + threadIdx.x] = v;
  }
}

static int check_f32(const ConvF32& p) {
  if (p.N < 1 || p.Cin < 1 || p.Cout < 1 || p.H < 1 || p.W < 1 || p.KH < 1 || p.KW < 1 || p.KH * p.KW > 16 ||
      p.stride < 1 || p.stride > 2 || p.pad < 0)
    return DFGAN_E_ARG;
  if (p.OH != (p.H + 2 * p.pad - p.KH) / p.stride + 1 || p.OW != (p.W + 2 * p.pad - p.KW) / p.stride + 1)
    return DFGAN_E_SHAPE;
  if (p.N > 65535 || (p.Cout + kCoT - 1) / kCoT > 65535 || (p.Cin + kCoT - 1) / kCoT > 65535 || p.Cout > 65535)
    return DFGAN_E_SHAPE;
  return DFGAN_OK;
}

}  // namespace dfgan

using namespace dfgan;

extern "C" {

int dfgan_conv2d_f32_fprop(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int H, int W,
                           int Cout, int KH, int KW, int stride, int pad, int OH, int OW, dfgan_stream_t st) {
  const ConvF32 p{N, Cin, H, W, Cout, KH, KW, stride, pad, OH, OW};
  if (!x || !w || !y) return DFGAN_E_ARG;
  if (int rc = check_f32(p)) return rc;
  const int PH = (kT - 1) * stride + KH, PW = (kT - 1) * stride + KW;
  const size_t smem = sizeof(float) * (static_cast<size_t>(kCiT) * PH * PW + static_cast<size_t>(kCiT) * KH * KW * kCoT);
  cudaError_t e = cudaFuncSetAttribute(conv_f32_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const dim3 grid(((OW + kT - 1) / kT) * ((OH + kT - 1) / kT), (Cout + kCoT - 1) / kCoT, N);
  conv_f32_fprop_kernel<<<grid, kT * kT, smem, static_cast<cudaStream_t>(st)>>>(x, w, bias, y, p);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_conv2d_f32_dgrad(const float* gy, const float* w, float* gx, int N, int Cin, int H, int W, int Cout, int KH,
                           int KW, int stride, int pad, int OH, int OW, dfgan_stream_t st) {
  const ConvF32 p{N, Cin, H, W, Cout, KH, KW, stride, pad, OH, OW};
  if (!gy || !w || !gx) return DFGAN_E_ARG;
  if (int rc = check_f32(p)) return rc;
  const int PH = (kT + KH - 2) / stride + 2, PW = (kT + KW - 2) / stride + 2;
  const size_t smem = sizeof(float) * (static_cast<size_t>(kCiT) * PH * PW + static_cast<size_t>(kCiT) * KH * KW * kCoT);
  cudaError_t e = cudaFuncSetAttribute(conv_f32_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const dim3 grid(((W + kT - 1) / kT) * ((H + kT - 1) / kT), (Cin + kCoT - 1) / kCoT, N);
  conv_f32_dgrad_kernel<<<grid, kT * kT, smem, static_cast<cudaStream_t>(st)>>>(gy, w, gx, p);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

int dfgan_conv2d_f32_wgrad(const float* x, const float* gy, float* gw, int N, int Cin, int H, int W, int Cout, int KH,
                           int KW, int stride, int pad, int OH, int OW, dfgan_stream_t st) {
  const ConvF32 p{N, Cin, H, W, Cout, KH, KW, stride, pad, OH, OW};
  if (!x || !gy || !gw) return DFGAN_E_ARG;
  if (int rc = check_f32(p)) return rc;
  const dim3 grid(Cin, Cout);
  conv_f32_wgrad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(x, gy, gw, p);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

}  // extern "C"
