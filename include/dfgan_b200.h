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
 *   v  *= (mask_src[n,h,w,co] > 0 ? 1 : slope)                       (mask_src optional, out strides)
 *   aux[n,h,w,co] = bf16(v) ; v = bf16(v)     (aux optional, out strides)
 *   v   = v * scale_imm * (scale_dev ? *scale_dev : 1)
 *   v  += add_scale * addend[n, h>>add_shift, w>>add_shift, co]      (addend optional)
 *   out_nchw[n][co][h][w] = tanh_out ? tanh(v) : v   for co < co_valid   (fp32 planar, optional)
 *   out[n,h,w,co] = bf16(v)   at out + n*out_sn + h*out_sh + w*out_sw + co   (optional)
 *   out2[n,h,w,co] = post(bf16(v))            (optional, out strides); post_mode 1: leaky_relu,
 *        2: leaky_relu(gb*leaky_relu(ga*v+ba)+bb) with ga.. = tab_*[n*tab_stride + co]  (AffineBlock pair)
 * At least one of out / aux / out2 / out_nchw must be given.
 * phase_cout > 0 merges the four phase convolutions of a 2x-upsampling layer (upsample->conv3x3, and the data gradient
 * of a 4x4 stride-2 conv) into ONE launch over the 9 distinct input shifts: N = 4*phase_cout, zero weight blocks where
 * a phase does not use a shift -- an MMA costs the same for any N <= 256, so this is 9 instead of 16 MMA groups.
 * replaces: nn.Conv2d forward and the data-gradient half of aten::convolution_backward at
 *   generator/residual_block.py:16,21,25,40,47  discriminator/residual_block.py:11,13,19
 *   discriminator/model.py:31 ; the stride-2 conv is expressed with four parity views of the input,
 *   its data gradient with four output phases (strided out_s*), the fused epilogue covers
 *   generator/residual_block.py:59, discriminator/residual_block.py:12,14,30 and F.leaky_relu backward;
 *   post_mode 2 covers generator/residual_block.py:36-39,43-46 + generator/fusion_block.py:37-39 of the
 *   NEXT convolution's input, out_nchw+tanh covers generator/model.py:28-29.
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
  void* out2;
  int32_t post_mode;
  const float *tab_ga, *tab_ba, *tab_gb, *tab_bb; /* fp32 [N][tab_stride] gamma/beta tables */
  int64_t tab_stride;
  float* out_nchw;
  int32_t co_valid, tanh_out;
  const void* wpack_img;      /* optional: the same operand packed with image=1 (dfgan_pack_weights); enables the
                                 wide-layer halo kernel (C_in % 64 == 0, C_in >= 128) */
  int32_t img_rows_per_slot;  /* rows_pad the image was packed with (rows between consecutive 64-column chunks) */
  int32_t phase_cout; /* > 0: Cout == 4*phase_cout; column block p=(a,c) is written to pixel (2h+a, 2w+c) of the
                         outputs (merged sub-pixel phases), bias / tables / addend are indexed by co % phase_cout */
  void* workspace;    /* optional ZEROED fp32 scratch of dfgan_conv_splitk_bytes(d) bytes: lets an under-filled
                         launch (low resolutions: fewer than #SMs/2 tiles, long K loop) split the (tap, chunk) loop
                         over several CTAs per tile; a second pass applies the epilogue and re-zeroes the scratch */
  int64_t ws_bytes;
} dfgan_conv_desc;

int dfgan_conv_run(const dfgan_conv_desc* d, dfgan_stream_t stream);
/* 1 when dfgan_conv_run(d) reads d->wpack_img (wide-layer halo kernel), 0 when it reads d->wpack: lets a caller
 * that re-packs weights every optimiser step skip the operand copy it never uses. */
int dfgan_conv_uses_image(const dfgan_conv_desc* d);
/* bytes of zeroed workspace with which dfgan_conv_run(d) would run split-K (0: the launch does not split) */
int64_t dfgan_conv_splitk_bytes(const dfgan_conv_desc* d);

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
  int32_t ci_valid, co_valid; /* 0 = all; rows/cols at or beyond are padding and are not written */
} dfgan_wgrad_desc;

int dfgan_wgrad_run(const dfgan_wgrad_desc* d, dfgan_stream_t stream);


/* ------------------------------------------------------------------------------------------------------
 * Pointwise / reduction kernels (csrc/pointwise.cu).  Activations: contiguous NHWC bf16 [N][HW][C], C % 8 == 0.
 * gamma/beta tables: fp32 [N][tab_stride], the pointer already offset to the block's first column.
 */
/* out = lrelu(gb*lrelu(ga*x+ba)+bb).  replaces: generator/residual_block.py:36-39,43-46 + fusion_block.py:37-39 */
int dfgan_affine2_fwd(const void* x, void* out, const float* ga, const float* ba, const float* gb,
                      const float* bb, int64_t tab_stride, int N, int HW, int C, dfgan_stream_t st);
/* Backward of the pair incl. the per-(n,c) reductions d_ga.. (atomically accumulated), optional gradient scale
 * (*gscale), optional residual-gate dot product dot += <g_a, lrelu(u2)>, optional added gradient g_add and
 * optional column sum of the result.  replaces: autograd of the same lines (mul/add/leaky_relu backward). */
int dfgan_affine2_bwd(const void* g_a, const void* x, const void* g_add, void* gx, const float* ga,
                      const float* ba, const float* gb, const float* bb, float* d_ga, float* d_ba,
                      float* d_gb, float* d_bb, int64_t tab_stride, const float* gscale, float* dot,
                      float* colsum, int N, int HW, int C, dfgan_stream_t st);
/* out[n,i,j,c] = scale * sum over the (1<<shift)^2 window of x ; colsum[c] += cs_imm*(*cs_dev)*sum x.
 * replaces: F.avg_pool2d (discriminator/residual_block.py:27), upsample_nearest2d backward (generator/model.py:40),
 * bias-gradient column sums of aten::convolution_backward. */
int dfgan_pool(const void* x, void* out, float* colsum, const float* cs_dev, float cs_imm, int N, int Ho,
               int Wo, int C, int shift, float scale, dfgan_stream_t st);
/* g_pre = g_y * *gamma * lrelu'(r) ; *dgamma += <g_y, r> ; colsum[c] += sum g_y.
 * replaces: backward of `shortcut + gamma * residual` and the trailing LeakyReLU, discriminator/residual_block.py:14,30 */
int dfgan_dres_bwd(const void* g_y, const void* r, void* g_pre, const float* gamma, float* dgamma,
                   float* colsum, int N, int HW, int C, dfgan_stream_t st);
/* *out += scale * <a, b>  (n bf16 elements, n % 8 == 0) */
int dfgan_dot(const void* a, const void* b, int64_t n, float* out, float scale, dfgan_stream_t st);
/* Stem im2col: NCHW fp32 image [N,3,H,W] (optionally scaled per sample) -> NHWC bf16 [N,H,W,32], channel
 * ci*9+r*3+s = img[ci][h+r-1][w+s-1].  replaces: the input gather of nn.Conv2d(3,n_c,3,1,1), discriminator/model.py:13 */
int dfgan_im2col3(const float* img, const float* rowscale, void* out, int N, int H, int W, dfgan_stream_t st);
/* Fused stem: out[n,h,w,co] = bf16( bias[co] + sum_k bf16(rowscale[n] * img[n,ci,h+r-1,w+s-1]) * w[co][k] ), k = ci*9+r*3+s,
 * out NHWC bf16 [N,H,W,32]; w = packed bf16 [32][32] (rows C_out, K padded 27 -> 32 with zeros); bias, rowscale and
 * xcol optional.  xcol != NULL additionally writes the im2col tensor of dfgan_im2col3 (operand of the stem's weight
 * gradient).  One kernel: im2col tile built in shared memory, tcgen05.mma, per-pixel epilogue.
 * replaces: nn.Conv2d(3, n_c, 3, 1, 1) forward, discriminator/model.py:13 (called from build_embeds, :36-41) */
int dfgan_stem_fwd(const float* img, const float* rowscale, const void* w, const float* bias, void* out, void* xcol,
                   int N, int H, int W, dfgan_stream_t st);
/* g_t = g_img*(1-img^2) packed NHWC bf16 [N,HW,32] (3 real channels); dbias[0..2] += sum.  replaces: tanh backward, generator/model.py:29 */
int dfgan_tanh_bwd_pack(const float* g_img, const float* img, void* out, float* dbias, int N, int HW,
                        dfgan_stream_t st);
/* fp32 [N][C][HW] <-> bf16 [N][HW][C] (API-boundary layout changes: generator/model.py:37, discriminator/model.py:36-41) */
int dfgan_nchw_to_nhwc(const float* in, void* out, int N, int C, int HW, dfgan_stream_t st);
int dfgan_nhwc_to_nchw(const void* in, float* out, int N, int C, int HW, int accumulate, dfgan_stream_t st);
/* out[n] += sum_i x[n][i]^2   (deep_fusion_gan/model.py:82) */
int dfgan_rowsq(const float* x, int N, int64_t row_len, float* out, dfgan_stream_t st);
/* Logit head (discriminator/model.py:43-59): concat builder, its backward, and the final 4x4 valid conv. */
int dfgan_head_cat(const float* embed, const float* sent, const float* sscale, void* cat, int B, int Ce,
                   int Cs, dfgan_stream_t st);
int dfgan_head_cat_bwd(const void* g_cat, float* g_embed, float* g_sent, int B, int Ce, int Cs,
                       dfgan_stream_t st);
int dfgan_head_out(const void* hid, const float* w2, float* logits, int B, int Ch, dfgan_stream_t st);
int dfgan_head_out_bwd(const float* g_logit, const void* hid, const void* act, const float* w2, void* g_hid,
                       float* dw2, int B, int Ch, dfgan_stream_t st);

/* ------------------------------------------------------------------------------------------------------
 * Batched small linears (csrc/mlp_optim.cu): one launch over a DEVICE table of 64-output work items.
 *   mode 0: out[b,o] = act(sum_k in[b,k]*w[o,k] + bias[o])            o in [o0, o0+64)
 *   mode 1: in[b,k]  = relu'(mask[b,k]) * sum_o out[b,o]*w[o,k]        k in [o0, o0+64)   (`in` is written)
 *   mode 2: w[o,k]  += sum_b out[b,o]*in[b,k] ; bias[o] += sum_b out[b,o]   (`w`,`bias` are written)
 * replaces: the 113 aten::addmm of Generator.forward (generator/fusion_block.py:8-18,34-35; generator/model.py:34)
 * and their autograd.
 */
typedef struct dfgan_linear_item {
  const float* in;
  int64_t in_stride;
  const float* w;
  const float* bias;
  float* out;
  int64_t out_stride;
  const float* mask;
  int32_t k, n_out, o0, relu;
} dfgan_linear_item;
int dfgan_linear_batched(const dfgan_linear_item* items_dev, int n_items, int B, int k_max, int n_out_max,
                         int mode, dfgan_stream_t st);

/* fp32 OIHW weights -> bf16 GEMM operand [n_slots][rows_pad][cols_pad]; slot s = scale * sum of the source
 * taps slot_taps[s][0..3] (-1 = unused; HOST array); transpose 0: rows=co, cols=ci; 1: rows=ci, cols=co.
 * image=1 writes the operand as [slot][64-column chunk][rows_pad][64] with every row's 16-byte pieces XOR-swizzled by
 * (row & 7), i.e. the SWIZZLE_128B shared-memory image, so a weight tile is ONE contiguous bulk copy (wpack_img). */
int dfgan_pack_weights(const float* w, void* out, int Co, int Ci, int KK, int transpose, int rows_pad,
                       int cols_pad, int n_slots, const int8_t* slot_taps, float scale, int image, dfgan_stream_t st);
/* The same for every weight tensor of a network in ONE launch: a host table of dfgan_pack_entry_size()-byte
 * entries is filled with dfgan_pack_entry_fill (returns the entry's block count; block_start = running sum),
 * copied to the device once, and replayed after every optimiser step by dfgan_pack_batched. */
int dfgan_pack_entry_size(void);
int dfgan_pack_entry_bs_offset(void); /* byte offset of the int32 block_start inside an entry (host-side re-linking) */
int dfgan_pack_entry_fill(void* entry_out, const float* w, void* out, int Co, int Ci, int KK, int transpose,
                          int rows_pad, int cols_pad, int n_slots, const int8_t* slot_taps, float scale,
                          int image, int block_start);
int dfgan_pack_batched(const void* entries_dev, int n_entries, int total_blocks, dfgan_stream_t st);
/* dw[co][ci][3][3] += un-folded sub-pixel weight gradient folded[16][co][ci]. */
/* Tap-major weight-gradient scratch -> OIHW gradient: dst[co][ci][t] += src[t][co][ci]; src = 0.  dfgan_wgrad_run
 * accumulates with fp32 atomics, which are coalesced only when ci is the fastest index; all scratch tensors of a
 * network are folded back in ONE launch over a device table of dfgan_unpack_entry_size()-byte entries. */
int dfgan_unpack_entry_size(void);
int dfgan_unpack_entry_fill(void* entry_out, float* src, float* dst, int Co, int Ci, int KK, int block_start);
int dfgan_wgrad_unpack(const void* entries_dev, int n_entries, int total_blocks, dfgan_stream_t st);

int dfgan_unfold_subpixel_wgrad(const float* folded, float* dw, int Co, int Ci, dfgan_stream_t st);

/* *out += sum g^2 ; then clip (global L2 norm, max_norm) + Adam over a flat buffer.  step: device int (>= 1).
 * replaces: clip_grad_norm_ + Adam.step, deep_fusion_gan/model.py:42-43,195-197,208-210,226-228 */
int dfgan_sqnorm(const float* g, int64_t n, float* out, dfgan_stream_t st);
/* The same with a result that does not depend on block scheduling (per-block partials in scratch, summed in index
 * order by the last block): data-parallel ranks must derive bit-identical clip coefficients.  scratch: at least
 * 4 * #SMs + 1 floats, the last used word (a counter) zero before the first call. */
int dfgan_sqnorm_det(const float* g, int64_t n, float* out, float* scratch, int scratch_floats, dfgan_stream_t st);
int dfgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq,
                    const int* step, float gscale, float max_norm, float lr, float b1, float b2, float eps,
                    dfgan_stream_t st);

/* Small fused scalar math of the training step (each replaces a chain of tiny elementwise launches).
 * dfgan_gblock_small_grads: generator block, from cs[c] = sum over pixels of the block's output cotangent:
 *   g_bias2 += gamma * cs (conv_2.bias), g_sbias += cs (scale_conv.bias, may be NULL), g_gamma += <cs, bias2>
 *   (generator/residual_block.py:29-33,47,59).
 * dfgan_d_hinge: logits = [real (B) ; mismatched (B-1) ; fake (B)] -> hinge loss and d loss / d logits
 *   (deep_fusion_gan/model.py:175-189).
 * dfgan_gp_coef: sq[b] = |d logit / d (image, sentence)|^2 -> MA-GP loss 2 mean(clamp(sqrt(sq + 1e-8), 0, 10)^6) and
 *   coef[b] = (12 / B) n_b^4 [n_b <= 10]  (deep_fusion_gan/model.py:82-85, 203).
 * dfgan_txtimg_loss: the text-image loss the reference's trainer logs every iteration without ever differentiating
 *   it (damsm/loss.py:4-26, called at deep_fusion_gan/model.py:87-89, 220-223): img fp32 [B,C,plane_len] (NCHW
 *   planes, 16-byte aligned, plane_len a multiple of 16), sent fp32 [B,E] (the sentence embeddings the iteration
 *   already holds -- the reference re-runs the frozen text encoder for them, damsm/loss.py:14) ->
 *   loss = 1 - mean_b <normalize(pad_E(mean over the plane of img_b)), normalize(sent_b)>.  Two launches, partial
 *   sums in `partial` (>= 4*B*C floats, caller-owned), bit-reproducible (no atomics). */
int dfgan_gblock_small_grads(const float* cs, const float* gamma, const float* bias2, float* g_bias2, float* g_gamma,
                             float* g_sbias, int C, dfgan_stream_t st);
int dfgan_d_hinge(const float* logits, float* loss, float* g_logits, int B, dfgan_stream_t st);
int dfgan_gp_coef(const float* sq, float* loss, float* coef, int B, dfgan_stream_t st);
int dfgan_txtimg_loss(const float* img, int B, int C, int64_t plane_len, const float* sent, int E, float* partial,
                      int partial_floats, float* loss, dfgan_stream_t st);

/* ---------------------------------------------------------------------------------------------------------
 * fp32 precision tier (north_star: "1e-4 for fp32 / tf32-off"): plain CUDA-core convolution on the reference's own
 * layouts -- x [N,Cin,H,W], w [Cout,Cin,KH,KW], y [N,Cout,OH,OW], all fp32 contiguous, stride 1 or 2, KH*KW <= 16.
 * Replaces nn.Conv2d / nn.Linear (as 1x1 on a 1x1 grid) of generator/residual_block.py:16,21,25,
 * generator/model.py:13,28, generator/fusion_block.py:8-18, discriminator/residual_block.py:11,13,19,
 * discriminator/model.py:13,31,33 and their autograd (aten::convolution_backward, _convolution_double_backward):
 * the three operators are closed under differentiation.  Verification tier, not the fast path. */
int dfgan_conv2d_f32_fprop(const float* x, const float* w, const float* bias /* may be NULL */, float* y, int N,
                           int Cin, int H, int W, int Cout, int KH, int KW, int stride, int pad, int OH, int OW,
                           dfgan_stream_t st);
/* gx[N,Cin,H,W] = d/dx <gy, conv(x, w)> */
int dfgan_conv2d_f32_dgrad(const float* gy, const float* w, float* gx, int N, int Cin, int H, int W, int Cout, int KH,
                           int KW, int stride, int pad, int OH, int OW, dfgan_stream_t st);
/* gw[Cout,Cin,KH,KW] = d/dw <gy, conv(x, w)>  (overwrites gw; deterministic reduction order) */
int dfgan_conv2d_f32_wgrad(const float* x, const float* gy, float* gw, int N, int Cin, int H, int W, int Cout, int KH,
                           int KW, int stride, int pad, int OH, int OW, dfgan_stream_t st);

/* ---------------------------------------------------------------------------------------------------------
 * Frozen text encoder: RNNEncoder.forward (text_encoder/model.py:33-56) in eval mode -- nn.Embedding lookup,
 * one bidirectional nn.LSTM layer over the length-packed batch (pack_padded_sequence, :39-44), sentence embedding
 * = [final h of the forward direction | final h of the reverse direction] (:54-56).  Called once per training
 * iteration (deep_fusion_gan/model.py:171) and again inside damsm/loss.py:14.
 *   tokens [B,T] int64 (padded captions), lens [B] int64 (any order; values are clamped to [0,T]; 0 gives a zero
 *   embedding), emb [ntoken,I] fp32, w_ih[d] [4H,I], w_hh[d] [4H,H], b_ih[d] / b_hh[d] [4H] fp32 for d = 0 (forward,
 *   rnn.*_l0) and d = 1 (rnn.*_l0_reverse), gate order i,f,g,o; the four pointer arrays are HOST arrays of two
 *   device pointers.  sent [B,2H] fp32.  H must be 128 (RNNEncoder.load: nhidden=256 over two directions).
 *   workspace: dfgan_text_encoder_workspace_bytes(B,T) bytes of device memory (the input-gate pre-activations).
 * Token ids outside [0,ntoken) read as a zero embedding (torch raises / device-asserts there).  Two launches, no
 * host synchronisation (the reference's cap_lens.tolist() is one). */
size_t dfgan_text_encoder_workspace_bytes(int B, int T);
int dfgan_text_encoder_fwd(const int64_t* tokens, const int64_t* lens, const float* emb, int64_t ntoken,
                           const float* const* w_ih, const float* const* w_hh, const float* const* b_ih,
                           const float* const* b_hh, float* sent, void* workspace, size_t ws_bytes, int B, int T,
                           int I, int H, dfgan_stream_t st);

/* ---------------------------------------------------------------------------------------------------------
 * Image side of the input pipeline for a batch of decoded RGB images, fused: the bounding-box crop of
 * DFGANDataset._get_image (objects/dataset.py:116-126), create_loader's Resize(int(imsize*76/64)) -> RandomCrop(imsize)
 * -> RandomHorizontalFlip (utils.py:15-19; Resize = Pillow's antialiased BILINEAR ImagingResample, restated bit-exactly)
 * and ToTensor + Normalize(0.5, 0.5) (objects/dataset.py:28-31).  The random decisions are inputs (drawn by the caller).
 *   item.src: uint8 [src_h][pitch bytes] with 3 interleaved channels per pixel (device memory);
 *   (x1, y1, cw, ch): the crop rectangle inside src; (rw, rh): the size Resize produces for a cw x ch image;
 *   (top, left): RandomCrop offset inside the resized image; flip: RandomHorizontalFlip decision.
 *   lut256: the fp32 value ToTensor + Normalize gives each uint8 level (device, 256 floats);
 *   out: fp32 [B][3][imsize][imsize]; workspace: dfgan_augment_workspace_bytes(B, imsize) bytes (device).
 * items_host (optional, same content as items_dev) is validated with dfgan_augment_check before anything is
 * launched: E_ARG for rectangles outside their image, E_SHAPE for a resized image smaller than imsize (torchvision:
 * ValueError) or a down-scaling factor above 11.5.  Two launches, no host synchronisation. */
typedef struct dfgan_aug_item {
  const uint8_t* src;
  int64_t pitch;
  int32_t src_w, src_h, x1, y1, cw, ch, rw, rh, top, left, flip, reserved;
} dfgan_aug_item;
size_t dfgan_augment_workspace_bytes(int B, int imsize);
int dfgan_augment_check(const dfgan_aug_item* item, int imsize);
int dfgan_augment(const dfgan_aug_item* items_host, const dfgan_aug_item* items_dev, const float* lut256, float* out,
                  void* workspace, size_t ws_bytes, int B, int imsize, dfgan_stream_t st);

/* ---------------------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory (batch sharded over the GPUs of one box, the
 * flat G / D gradients summed before clip_grad_norm_ + Adam -- the reference's optimiser steps are
 * deep_fusion_gan/model.py:195-197, 208-210, 226-228; the reference itself is single-device).  One process per GPU.
 *   dfgan_peer_alloc: zeroed cudaMalloc memory on the current device (the gradient buffer, the flag block of
 *     dfgan_peer_flag_bytes() bytes); dfgan_peer_export writes its dfgan_peer_handle_bytes()-byte IPC handle, which
 *     the caller passes to the other processes by any means; dfgan_peer_import maps a peer's handle (peer access is
 *     enabled by the mapping); dfgan_peer_close / dfgan_peer_free undo them.
 *   dfgan_peer_allreduce_f32: ONE kernel, in place: bufs[p] = start of the range inside rank p's buffer (bufs[rank] is
 *     the local one, every address 16-byte aligned), flags[p] = rank p's flag block, count = fp32 elements (multiple
 *     of 4).  Slice r of the range is summed by rank r in rank order and written to every rank, so all ranks end with
 *     bit-identical sums.  Every rank must issue the same sequence of calls per channel (0..5; calls on different
 *     channels may run concurrently on different streams); no host synchronisation, capturable in a CUDA graph.
 *     ctas: grid size (<= 128, 0 = 32); threads: block size (<= 256, 0 = 256; <= 128 selects the 40-register shape
 *     that fits on an SM next to a resident conv CTA, for an exchange overlapped with other kernels); a wait on a
 *     peer that exceeds about timeout_ms (0 = 20000; counted in polls of ~0.5-1 us each, not with a device timer) prints
 *     the waiting rank / CTA / call and traps the kernel. */
size_t dfgan_peer_flag_bytes(void);
int dfgan_peer_handle_bytes(void);
int dfgan_peer_alloc(size_t bytes, void** ptr);
int dfgan_peer_free(void* ptr);
int dfgan_peer_export(void* ptr, void* handle_out);
int dfgan_peer_import(const void* handle, void** ptr);
int dfgan_peer_close(void* ptr);
int dfgan_peer_allreduce_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count, int channel,
                             int ctas, int threads, int timeout_ms, dfgan_stream_t stream);
/* Sharded optimiser step: the three pieces of reduce-scatter -> clip + Adam on the owned slice -> all-gather.  Slice r of
 * a range of n4 = count / 4 16-byte elements is [per * r, min(per * (r + 1), n4)) with per = ceil(n4 / world), in all three.
 *   dfgan_peer_reduce_scatter_f32: like dfgan_peer_allreduce_f32 without the push: rank r ends with the reduced slice r
 *     in its own buffer (the other slices keep the local values); START rendezvous only.
 *   dfgan_peer_sum_scalar_f32: total = sum over ranks, in rank order (bit-identical everywhere), of *partial; slots[p] =
 *     rank p's peer-mapped array of 8 floats; one 32-thread CTA; also the rendezvous after which no rank still reads
 *     another rank's gradient buffer.
 *   dfgan_peer_allgather_f32: rank r pushes slice r of ITS buffer (bufs[rank]) into every peer's buffer; END rendezvous
 *     (on return of the kernel every slice of every peer has landed locally). */
int dfgan_peer_reduce_scatter_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count,
                                  int channel, int ctas, int threads, int timeout_ms, dfgan_stream_t stream);
int dfgan_peer_sum_scalar_f32(const float* partial, void* const* slots, float* total, void* const* flags, int world,
                              int rank, int channel, int timeout_ms, dfgan_stream_t stream);
int dfgan_peer_allgather_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count, int channel,
                             int ctas, int timeout_ms, dfgan_stream_t stream);
/* The same exchange reduced inside the NVSwitch: mc_range = the MULTICAST address of the range, i.e. the buffers of all
 * ranks are bound to one NVLS multicast object at equal offsets (mapped by the caller, e.g. torch's symmetric-memory
 * allocator: handle.multicast_ptr + byte offset).  Rank r issues one multimem.ld_reduce.add.f32 and one multimem.st per
 * 16 bytes of slice r; flags, channels, ctas / threads / timeout_ms as above (the flag blocks stay dfgan_peer_* memory). */
int dfgan_peer_allreduce_mc_f32(void* mc_range, void* const* flags, int world, int rank, int64_t count, int channel,
                                int ctas, int threads, int timeout_ms, dfgan_stream_t stream);

/* Per-launch timing of the two GEMM engines with CUDA events on the launching stream (bench.py roofline leg).
 * dfgan_profile_read: out[e*4 + {0: summed ms, 1: executed FLOPs, 2: algorithmic bytes, 3: launches}], e = 0 for
 * dfgan_conv_run, 1 for dfgan_wgrad_run; synchronises the device and clears the records. */
void dfgan_profile_enable(int on);
int dfgan_profile_read(double* out);

/* Library identification: version = major*10000 + minor*100 + patch; arch string is "sm_100a". */
int dfgan_version(void);
const char* dfgan_arch(void);
/* Number of kernels launched by this library since load (host counter; bench.py's gpu_launches). */
uint64_t dfgan_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DFGAN_B200_H_ */
