"""fp32 precision tier of the drop-in modules (north_star: results "within ... 1e-4 for fp32/tf32-off").

`Generator(...).set_precision("fp32")` / `Discriminator(...).set_precision("fp32")` route every nn.Conv2d and
nn.Linear of the reference graph through the CUDA-core fp32 kernels of csrc/conv_f32.cu (dfgan_conv2d_f32_*);
the pointwise glue (broadcast affine, LeakyReLU, nearest upsample, average pool, concat, tanh) stays on torch
tensor ops so that torch autograd differentiates it -- including twice, for the MA-GP step
(deep_fusion_gan/model.py:59-85, 202-210).  The three convolution operators are closed under differentiation:

    F(x, w) = conv(x, w)          dF: gx = D(gy, w),  gw = W(x, gy)
    D(gy, w) = d<gy,F>/dx         dD: g_gy = F(c, w), g_w = W(c, gy)
    W(x, gy) = d<gy,F>/dw         dW: g_x = D(gy, c), g_gy = F(x, c)

so `autograd.grad(..., create_graph=True)` followed by `.backward()` never leaves these kernels.  This is the
verification tier: it shows the kernels' math equals the reference's to fp32 rounding; the bf16 tensor-core
programs (gnet / dnet / engine) are the fast path.

Reference graph restated here: generator/model.py:32-47, generator/residual_block.py:29-59,
generator/fusion_block.py:33-39, discriminator/model.py:36-59, discriminator/residual_block.py:23-30.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F_

from . import _lib

F32 = torch.float32


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _geom(x_shape, w_shape, stride: int, pad: int):
    n, cin, h, w = x_shape
    cout, cin_w, kh, kw = w_shape
    if cin != cin_w:
        raise RuntimeError(f"Given groups=1, weight of size {list(w_shape)}, expected input{list(x_shape)} to have "
                           f"{cin_w} channels, but got {cin} channels instead")
    oh, ow = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    return (n, cin, h, w, cout, kh, kw, stride, pad, oh, ow)


def _c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


def _fprop(x, w, g):
    y = torch.empty(g[0], g[4], g[9], g[10], device=x.device, dtype=F32)
    _lib.check(_lib.lib().dfgan_conv2d_f32_fprop(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), *g, _s()),
               "dfgan_conv2d_f32_fprop")
    return y


def _dgrad(gy, w, g):
    gx = torch.empty(g[0], g[1], g[2], g[3], device=gy.device, dtype=F32)
    _lib.check(_lib.lib().dfgan_conv2d_f32_dgrad(gy.data_ptr(), w.data_ptr(), gx.data_ptr(), *g, _s()),
               "dfgan_conv2d_f32_dgrad")
    return gx


def _wgrad(x, gy, g):
    gw = torch.empty(g[4], g[1], g[5], g[6], device=x.device, dtype=F32)
    _lib.check(_lib.lib().dfgan_conv2d_f32_wgrad(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), *g, _s()),
               "dfgan_conv2d_f32_wgrad")
    return gw


class ConvF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, g):
        ctx.g = g
        ctx.save_for_backward(x, w)
        return _fprop(_c(x), _c(w), g)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = ConvD.apply(gy, w, ctx.g) if ctx.needs_input_grad[0] else None
        gw = ConvW.apply(x, gy, ctx.g) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class ConvD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gy, w, g):
        ctx.g = g
        ctx.save_for_backward(gy, w)
        return _dgrad(_c(gy), _c(w), g)

    @staticmethod
    def backward(ctx, c):
        gy, w = ctx.saved_tensors
        g_gy = ConvF.apply(c, w, ctx.g) if ctx.needs_input_grad[0] else None
        g_w = ConvW.apply(c, gy, ctx.g) if ctx.needs_input_grad[1] else None
        return g_gy, g_w, None


class ConvW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gy, g):
        ctx.g = g
        ctx.save_for_backward(x, gy)
        return _wgrad(_c(x), _c(gy), g)

    @staticmethod
    def backward(ctx, c):
        x, gy = ctx.saved_tensors
        g_x = ConvD.apply(gy, c, ctx.g) if ctx.needs_input_grad[0] else None
        g_gy = ConvF.apply(x, c, ctx.g) if ctx.needs_input_grad[1] else None
        return g_x, g_gy, None


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1,
           padding: int = 0) -> torch.Tensor:
    """torch.nn.functional.conv2d(x, weight, bias, stride, padding) on the fp32 CUDA-core kernels."""
    if not x.is_cuda:
        raise RuntimeError("gan_codes_b200.fp32.conv2d: CUDA tensors only (no CPU path)")
    y = ConvF.apply(x, weight, _geom(x.shape, weight.shape, stride, padding))
    return y if bias is None else y + bias.view(1, -1, 1, 1)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """torch.nn.functional.linear as a 1x1 convolution on a 1x1 grid (same kernels, no library GEMM)."""
    if x.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and "
                           f"{tuple(weight.t().shape)})")
    y = conv2d(x.view(x.shape[0], x.shape[1], 1, 1), weight.view(weight.shape[0], weight.shape[1], 1, 1), None)
    y = y.view(x.shape[0], weight.shape[0])
    return y if bias is None else y + bias


# ------------------------------------------------------------------------------------------------ generator
def _mlp(seq, sent):
    h = torch.relu(linear(sent, seq[0].weight, seq[0].bias))
    return linear(h, seq[2].weight, seq[2].bias)


def _affine(block, x, sent):
    """generator/fusion_block.py:33-39."""
    gamma, beta = _mlp(block.gamma_mlp, sent), _mlp(block.beta_mlp, sent)
    return gamma[:, :, None, None] * x + beta[:, :, None, None]


def _res_block_g(blk, x, sent):
    """generator/residual_block.py:29-59."""
    s = x if blk.scale_conv is None else conv2d(x, blk.scale_conv.weight, blk.scale_conv.bias)
    h = F_.leaky_relu(_affine(blk.fusion_block_1, x, sent), 0.2)
    h = F_.leaky_relu(_affine(blk.fusion_block_2, h, sent), 0.2)
    h = conv2d(h, blk.conv_1.weight, blk.conv_1.bias, 1, 1)
    h = F_.leaky_relu(_affine(blk.fusion_block_3, h, sent), 0.2)
    h = F_.leaky_relu(_affine(blk.fusion_block_4, h, sent), 0.2)
    h = conv2d(h, blk.conv_2.weight, blk.conv_2.bias, 1, 1)
    return s + blk.gamma * h


def generator_forward(gen, noise: torch.Tensor, sent: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """generator/model.py:32-47.  `taps` (optional dict) receives per-block activations for layer-wise parity."""
    noise, sent = noise.to(F32), sent.to(F32)
    if noise.dim() != 2 or noise.shape[1] != gen.linear_in.weight.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: noise {tuple(noise.shape)}")
    out = linear(noise, gen.linear_in.weight, gen.linear_in.bias)
    c0 = gen.linear_in.weight.shape[0] // 16
    out = out.view(-1, c0, 4, 4)
    for i, blk in enumerate(gen.res_blocks):
        out = _res_block_g(blk, out, sent)
        if taps is not None:
            taps[f"res_blocks.{i}"] = out
        out = F_.interpolate(out, scale_factor=2)
    out = _res_block_g(gen.res_block_out, out, sent)
    if taps is not None:
        taps["res_block_out"] = out
    conv = gen.conv_out[1]
    return torch.tanh(conv2d(F_.leaky_relu(out, 0.2), conv.weight, conv.bias, 1, 1))


# ------------------------------------------------------------------------------------------------ discriminator
def _res_block_d(blk, x):
    """discriminator/residual_block.py:23-30."""
    s = x if blk.scale_conv is None else conv2d(x, blk.scale_conv.weight, blk.scale_conv.bias)
    s = F_.avg_pool2d(s, 2)
    r = F_.leaky_relu(conv2d(x, blk.residual_conv[0].weight, None, 2, 1), 0.2)
    r = F_.leaky_relu(conv2d(r, blk.residual_conv[2].weight, None, 1, 1), 0.2)
    return s + blk.gamma * r


def build_embeds(disc, image: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """discriminator/model.py:36-41."""
    n_blocks = len(disc.img_forward) - 1
    if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != image.shape[3] or image.shape[2] != 4 << n_blocks:
        raise RuntimeError(f"Sizes of tensors must match: image {tuple(image.shape)} does not reduce to a 4x4 embed "
                           f"grid through {n_blocks} down-blocks")
    stem = disc.img_forward[0]
    x = conv2d(image.to(F32), stem.weight, stem.bias, 1, 1)
    for k in range(1, n_blocks + 1):
        x = _res_block_d(disc.img_forward[k], x)
        if taps is not None:
            taps[f"img_forward.{k}"] = x
    return x


def get_logits(disc, embed: torch.Tensor, sent: torch.Tensor) -> torch.Tensor:
    """discriminator/model.py:43-59."""
    w0, w2 = disc.img_sentence_forward[0].weight, disc.img_sentence_forward[2].weight
    sdim = w0.shape[1] - embed.shape[1]
    if embed.dim() != 4 or tuple(embed.shape[2:]) != (4, 4) or sent.shape[-1] != sdim or sent.shape[0] != embed.shape[0]:
        raise RuntimeError(f"Sizes of tensors must match: embed {tuple(embed.shape)} / sentence {tuple(sent.shape)} "
                           "(discriminator/model.py:49-54)")
    s = sent.to(F32).view(-1, sdim, 1, 1).repeat(1, 1, 4, 4)
    h = torch.cat((embed.to(F32), s), dim=1)
    h = F_.leaky_relu(conv2d(h, w0, None, 1, 1), 0.2)
    return conv2d(h, w2, None, 1, 0)
