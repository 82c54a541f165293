/*
 * dfgan_b200 -- C ABI of the B200 (sm_100a) kernels behind the DF-GAN training-step hot path.
 *
 * The reference (Nandan-Upadhyaya/GAN-codes, Deep_Fusion_GAN/src) has no FFI: its hot path bottoms out in
 * torch.nn call sites.  Every entry point below replaces one family of those call sites; the citation
 * after "replaces:" is the reference file:line (relative to Deep_Fusion_GAN/src/) whose arithmetic the
 * kernel performs.  All pointers are DEVICE pointers unless stated otherwise, activations are NHWC bf16,
 * parameters and gradients are fp32 in the reference's own (OIHW / [out,in]) layout.  No entry point
 * allocates, synchronises with the host or keeps global state besides a lazily resolved driver entry
 * point.  Return value: 0 = OK, negative = invalid argument / unsupported shape (DFGAN_E_*), positive =
 * cudaError_t of the failing runtime call.
 */
#ifndef DFGAN_B200_H_
#define DFGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFGAN_OK 0
#define DFGAN_E_ARG (-1)
#define DFGAN_E_SHAPE (-2)
#define DFGAN_E_DRIVER (-3)

#define DFGAN_MAX_SRC 4
#define DFGAN_MAX_TAPS 16

typedef void* dfgan_stream_t; /* cudaStream_t */

/* NHWC bf16 tensor view, channel stride 1; sw/sh/sn are ELEMENT strides of w, h, n. */
typedef struct dfgan_view4 {
  const void* ptr;
  int32_t C, W, H, N;
  int64_t sw, sh, sn;
} dfgan_view4;

/* One filter tap of an implicit GEMM: reads src[src] at pixel (h+dy, w+dx) (zero outside the view) and
 * multiplies with rows [wrow, wrow+Cout) of the packed weight matrix. */
typedef struct dfgan_tap {
  int32_t src, dy, dx, wrow;
} dfgan_tap;

/*
 * dfgan_conv_run -- tap-list implicit GEMM on tcgen05/TMEM, operands staged by TMA.
 *   acc[n,h,w,co] = sum_t sum_ci src[t.src][n, h+t.dy, w+t.dx, ci] * wpack[t.wrow + co][ci]
 *   v   = acc + bias[co]                      (bias optional)
 *   v   = lrelu_pre ? leaky_relu(v, slope) : v
 *   aux[n,h,w,co] = v                         (aux optional, same strides as out)
 *   v   = v * scale_imm * (scale_dev ? *scale_dev : 1)
 *   v  += add_scale * addend[n, h>>add_shift, w>>add_shift, co]      (addend optional)
 *   v  *= (mask_src[n,h,w,co] > 0 ? 1 : slope)                       (mask_src optional, out strides)
 *   out[n,h,w,co] = bf16(v)   at out + n*out_sn + h*out_sh + w*out_sw + co
 * replaces: nn.Conv2d forward and the data-gradient half of aten::convolution_backward at
 *   generator/residual_block.py:16,21,25,40,47  discriminator/residual_block.py:11,13,19
 *   discriminator/model.py:31 ; the stride-2 conv is expressed with four parity views of the input,
 *   its data gradient with four output phases (strided out_s*), the fused epilogue covers
 *   generator/residual_block.py:59, discriminator/residual_block.py:12,14,30 and F.leaky_relu backward.
 */
typedef struct dfgan_conv_desc {
  dfgan_view4 src[DFGAN_MAX_SRC];
  int32_t n_src, n_taps;
  dfgan_tap taps[DFGAN_MAX_TAPS];
  const void* wpack; /* bf16 [w_rows][Cin], K-major */
  int32_t w_rows, Cin, Cout;
  int32_t N, H, W; /* output grid */
  void* out;
  int64_t out_sw, out_sh, out_sn;
  const float* bias;
  const float* scale_dev;
  float scale_imm;
  int32_t lrelu_pre;
  float slope;
  void* aux;
  const void* addend;
  int64_t add_sw, add_sh, add_sn;
  int32_t add_shift;
  float add_scale;
  const void* mask_src;
  int32_t n_tile_hint; /* 0 = auto */
  int32_t max_ctas;    /* 0 = one per SM */
} dfgan_conv_desc;

int dfgan_conv_run(const dfgan_conv_desc* d, dfgan_stream_t stream);

/*
 * dfgan_wgrad_run -- weight gradient of the same tap-list convolution (split-K over pixels, fp32 atomics):
 *   dw[co*s_co + ci*s_ci + t.wrow*s_tap] += scale * sum_{n,h,w} g[n,h,w,co] * src[t.src][n,h+t.dy,w+t.dx,ci]
 * replaces: the weight-gradient half of aten::convolution_backward for the call sites listed above, and
 *   the `gW = conv(ggI, gO)` half of aten::_convolution_double_backward (deep_fusion_gan/model.py:207).
 */
typedef struct dfgan_wgrad_desc {
  dfgan_view4 src[DFGAN_MAX_SRC];
  int32_t n_src, n_taps;
  dfgan_tap taps[DFGAN_MAX_TAPS]; /* wrow = tap id in the output addressing */
  dfgan_view4 g;                  /* [N,H,W,Cout] */
  int32_t Cin, Cout;
  float* dw;
  int64_t s_co, s_ci, s_tap;
  const float* scale_dev;
  float scale_imm;
  int32_t k_splits; /* 0 = auto */
} dfgan_wgrad_desc;

int dfgan_wgrad_run(const dfgan_wgrad_desc* d, dfgan_stream_t stream);

/* Library identification: version = major*10000 + minor*100 + patch; arch string is "sm_100a". */
int dfgan_version(void);
const char* dfgan_arch(void);
/* Number of kernels launched by this library since load (host counter; bench.py's gpu_launches). */
uint64_t dfgan_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DFGAN_B200_H_ */
