// Image side of the input pipeline for a batch of DECODED images, fused into one pass:
//   bounding-box crop          objects/dataset.py:116-126
//   Resize(int(imsize*76/64))  utils.py:16   -- torchvision F.resize on a PIL image = Pillow ImagingResample, BILINEAR
//   RandomCrop(imsize)         utils.py:17
//   RandomHorizontalFlip       utils.py:18
//   ToTensor + Normalize(.5)   objects/dataset.py:28-31
// The reference does this per image on one Python thread (DataLoader num_workers=0, utils.py:24).  Here the random
// decisions stay on the host (drawn from the torch generator in the reference's order, gan-codes_b200/augment.py) and
// the pixel work is two launches for the whole batch, BIT-EXACT with Pillow's integer arithmetic:
//
//   aug_coeffs_kernel : Pillow's precompute_coeffs + normalize_coeffs_8bpc for exactly the imsize output columns and
//                       rows the random crop keeps (flip folded into the column order): triangle filter whose support
//                       grows with the down-scaling factor, weights normalised in double precision (explicit
//                       round-to-nearest intrinsics -- no FMA contraction, so the doubles equal the C library's),
//                       rounded to 22-bit fixed point;
//   augment_kernel    : per output pixel the vertical pass over horizontally resampled uint8 values (each pass
//                       accumulates in int32 from 1 << 21 and clips to uint8, like ImagingResampleHorizontal/
//                       Vertical_8bpc), then the ToTensor/Normalize value from a 256-entry table computed by the
//                       caller with the reference's own fp32 expression.  Only the kept 256x256 window of the 304-ish
//                       resized image is ever computed (30 % less work than resize-then-crop), nothing intermediate
//                       is written to memory.
#include "host_common.h"

namespace dfgan {

constexpr int kPrecisionBits = 32 - 8 - 2;     // Pillow PRECISION_BITS
constexpr int kAugKMax = 24;                   // filter taps per axis: down-scaling factors up to 11.5

struct AugTables {        // per (image, axis): bounds and fixed-point weights, [x][imsize] so a warp reads them coalesced
  int* lo;                // [B][2][imsize]
  int* cnt;               // [B][2][imsize]
  int* k;                 // [B][2][kAugKMax][imsize]
};

__device__ __forceinline__ AugTables aug_tables(void* ws, int B, int S) {
  int* p = static_cast<int*>(ws);
  return AugTables{p, p + static_cast<size_t>(B) * 2 * S, p + static_cast<size_t>(B) * 4 * S};
}

__global__ void aug_coeffs_kernel(const dfgan_aug_item* __restrict__ items, void* ws, int B, int S) {
  const int b = blockIdx.x, axis = blockIdx.y;
  const dfgan_aug_item it = items[b];
  const AugTables t = aug_tables(ws, B, S);
  const int in_size = axis ? it.ch : it.cw, out_size = axis ? it.rh : it.rw;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;            // triangle filter support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int o = threadIdx.x; o < S; o += blockDim.x) {
    const int xx = axis ? it.top + o : it.left + (it.flip ? S - 1 - o : o);
    const double center = __dmul_rn(static_cast<double>(xx) + 0.5, scale);
    int lo = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (lo < 0) lo = 0;
    int hi = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (hi > in_size) hi = in_size;
    int n = hi - lo;
    if (n > kAugKMax) n = kAugKMax;              // excluded by the host-side check
    double w[kAugKMax];
    double ww = 0.0;
#pragma unroll 1
    for (int x = 0; x < n; ++x) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + lo), center), 0.5), ss);
      if (a < 0.0) a = -a;
      const double v = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      w[x] = v;
      ww = __dadd_rn(ww, v);
    }
    const size_t base = (static_cast<size_t>(b) * 2 + axis) * S + o;
    t.lo[base] = lo;
    t.cnt[base] = n;
#pragma unroll 1
    for (int x = 0; x < kAugKMax; ++x) {
      int kv = 0;
      if (x < n) {
        const double v = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
        const double f = __dmul_rn(v, static_cast<double>(1 << kPrecisionBits));
        kv = v < 0.0 ? __double2int_rz(__dadd_rn(-0.5, f)) : __double2int_rz(__dadd_rn(0.5, f));
      }
      t.k[((static_cast<size_t>(b) * 2 + axis) * kAugKMax + x) * S + o] = kv;
    }
  }
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256)
    augment_kernel(const dfgan_aug_item* __restrict__ items, const void* ws, const float* __restrict__ lut,
                   float* __restrict__ out, int B, int S) {
  __shared__ float s_lut[256];
  s_lut[threadIdx.x] = __ldg(lut + threadIdx.x);
  __syncthreads();
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= S * S) return;
  const int oy = p / S, ox = p % S;
  const dfgan_aug_item it = items[b];
  const AugTables t = aug_tables(const_cast<void*>(ws), B, S);
  const size_t hb = (static_cast<size_t>(b) * 2 + 0) * S + ox, vb = (static_cast<size_t>(b) * 2 + 1) * S + oy;
  const int xlo = t.lo[hb], xn = t.cnt[hb], ylo = t.lo[vb], yn = t.cnt[vb];
  const int* __restrict__ kh = t.k + (static_cast<size_t>(b) * 2 + 0) * kAugKMax * S + ox;
  const int* __restrict__ kv = t.k + (static_cast<size_t>(b) * 2 + 1) * kAugKMax * S + oy;
  int v0 = 1 << (kPrecisionBits - 1), v1 = v0, v2 = v0;
  for (int y = 0; y < yn; ++y) {
    const uint8_t* __restrict__ row = it.src + static_cast<size_t>(it.y1 + ylo + y) * it.pitch +
                                      static_cast<size_t>(it.x1 + xlo) * 3;
    int h0 = 1 << (kPrecisionBits - 1), h1 = h0, h2 = h0;
    for (int x = 0; x < xn; ++x) {
      const int k = kh[static_cast<size_t>(x) * S];
      h0 += static_cast<int>(__ldg(row + 3 * x)) * k;
      h1 += static_cast<int>(__ldg(row + 3 * x + 1)) * k;
      h2 += static_cast<int>(__ldg(row + 3 * x + 2)) * k;
    }
    const int k = kv[static_cast<size_t>(y) * S];
    v0 += clip8(h0) * k, v1 += clip8(h1) * k, v2 += clip8(h2) * k;
  }
  float* o = out + (static_cast<size_t>(b) * 3 * S + oy) * S + ox;
  const size_t plane = static_cast<size_t>(S) * S;
  o[0] = s_lut[clip8(v0)];
  o[plane] = s_lut[clip8(v1)];
  o[2 * plane] = s_lut[clip8(v2)];
}

}  // namespace dfgan

using namespace dfgan;

extern "C" {

size_t dfgan_augment_workspace_bytes(int B, int imsize) {
  if (B < 1 || imsize < 1) return 0;
  return sizeof(int) * static_cast<size_t>(B) * 2 * imsize * (2 + kAugKMax);
}

int dfgan_augment_check(const dfgan_aug_item* it, int imsize) {
  if (!it || imsize < 1) return DFGAN_E_ARG;
  if (!it->src || it->src_w < 1 || it->src_h < 1 || it->pitch < static_cast<int64_t>(it->src_w) * 3) return DFGAN_E_ARG;
  if (it->cw < 1 || it->ch < 1 || it->x1 < 0 || it->y1 < 0 || it->x1 + it->cw > it->src_w || it->y1 + it->ch > it->src_h)
    return DFGAN_E_ARG;                                           // crop rectangle outside the image
  if (it->rw < imsize || it->rh < imsize) return DFGAN_E_SHAPE;   // RandomCrop larger than the resized image
  if (it->left < 0 || it->top < 0 || it->left + imsize > it->rw || it->top + imsize > it->rh) return DFGAN_E_ARG;
  for (int a = 0; a < 2; ++a) {
    const double scale = static_cast<double>(a ? it->ch : it->cw) / static_cast<double>(a ? it->rh : it->rw);
    const double support = scale < 1.0 ? 1.0 : scale;
    int ksize = static_cast<int>(support);
    if (static_cast<double>(ksize) < support) ++ksize;
    if (ksize * 2 + 1 > kAugKMax) return DFGAN_E_SHAPE;           // down-scaling factor beyond the tap budget
  }
  return DFGAN_OK;
}

int dfgan_augment(const dfgan_aug_item* items_host, const dfgan_aug_item* items_dev, const float* lut256, float* out,
                  void* workspace, size_t ws_bytes, int B, int imsize, dfgan_stream_t st) {
  if (!items_dev || !lut256 || !out || !workspace || B < 1 || imsize < 1) return DFGAN_E_ARG;
  if (ws_bytes < dfgan_augment_workspace_bytes(B, imsize)) return DFGAN_E_ARG;
  if (items_host)
    for (int b = 0; b < B; ++b)
      if (int rc = dfgan_augment_check(items_host + b, imsize)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(st);
  aug_coeffs_kernel<<<dim3(B, 2), 128, 0, s>>>(items_dev, workspace, B, imsize);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  augment_kernel<<<dim3((imsize * imsize + 255) / 256, B), 256, 0, s>>>(items_dev, workspace, lut256, out, B, imsize);
  count_launch();
  DFGAN_CUDA_CHECK_LAUNCH();
  return DFGAN_OK;
}

}  // extern "C"
