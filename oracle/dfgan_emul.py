"""TEST INFRASTRUCTURE -- not a product path (see oracle/dfgan_oracle.py for the import rule).

Storage-format emulation of the CUDA pipeline on top of the fp32 oracle: the same arithmetic as
oracle/dfgan_oracle.py, evaluated in the pipeline's operator order (pool-before-1x1 shortcut, sub-pixel
folded upsample->conv3x3, App. C-1/C-2/C-4) with a bf16 rounding at every point where the kernels store a
tensor (weights, activations, gradient tensors).  With `Rounding(enabled=False)` it must agree with the fp32
oracle to fp32 re-association error (tests/test_oracle_golden.py checks that on the CPU); with rounding
enabled it is what a correct bf16 implementation of the reference computes, so the GPU parity tests can use
a tight tolerance that is not polluted by LeakyReLU mask flips (a bf16-vs-fp32 effect every bf16 pipeline
has; the tests report that deviation separately against the fp32 oracle).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F

from . import dfgan_oracle as O

SLOPE = O.SLOPE


class _RoundST(torch.autograd.Function):
    """bf16 rounding in the forward (straight-through) and of the incoming gradient in the backward."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.bfloat16().to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().to(g.dtype) if ctx.bwd else g), None, None


class Rounding:
    def __init__(self, enabled: bool = True):
        self.enabled = enabled

    def w(self, t):      # GEMM weight operand
        return _RoundST.apply(t, True, False) if self.enabled else t

    def a(self, t):      # stored activation: value and its gradient tensor are both bf16
        return _RoundST.apply(t, True, True) if self.enabled else t

    def g(self, t):      # only the gradient tensor is stored
        return _RoundST.apply(t, False, True) if self.enabled else t

    def v(self, t):      # only the value is stored in bf16 (gradient stays fp32)
        return _RoundST.apply(t, True, False) if self.enabled else t


# Teacher forcing: `force` maps tensor names to the values the CUDA pipeline stored (NCHW).  A forced tensor
# keeps the emulation's gradient path but takes the pipeline's VALUE, and LeakyReLU masks are taken from the
# sign of the forced post-activation -- so the backward below is the exact Jacobian-transpose at the point the
# kernels evaluated, and a gradient comparison is not polluted by mask flips caused by last-bit forward
# differences (a near-zero pre-activation whose sign differs changes a LeakyReLU slope from 1 to 0.2).
def _forced(t, force, name):
    if force is None or name not in force:
        return t
    if "__computed__" in force:      # what the emulation computes from the pipeline's own inputs
        force["__computed__"][name] = t.detach()
    return t + (force[name].to(t.dtype) - t).detach()


def _lrelu(pre, force, name):
    if force is None or name not in force:
        return F.leaky_relu(pre, SLOPE)
    if "__computed__" in force:
        force["__computed__"][name] = F.leaky_relu(pre, SLOPE).detach()
    post = force[name].to(pre.dtype)
    out = pre * torch.where(post > 0, 1.0, SLOPE).to(pre.dtype)
    return out + (post - out).detach()


# ------------------------------------------------------------------------------------------------ discriminator
def d_block(sd, p, x, q: Rounding, tap: Optional[Callable] = None, force: Optional[dict] = None):
    pooled = q.g(F.avg_pool2d(x, 2))
    if p + ".scale_conv.weight" in sd:
        s = q.a(F.conv2d(pooled, q.w(sd[p + ".scale_conv.weight"]), sd[p + ".scale_conv.bias"]))
    else:
        s = q.v(pooled)
    p1 = q.g(F.conv2d(x, q.w(sd[p + ".residual_conv.0.weight"]), None, stride=2, padding=1))
    c1 = q.v(_lrelu(p1, force, p + ".c1"))
    p2 = q.g(F.conv2d(c1, q.w(sd[p + ".residual_conv.2.weight"]), None, padding=1))
    r = q.v(_lrelu(p2, force, p + ".r"))
    y = _forced(q.a(s + sd[p + ".gamma"] * r), force, p + ".y")
    if tap:
        tap(p + ".c1", c1)
        tap(p + ".r", r)
        tap(p + ".y", y)
    return y


def discriminator_embeds(sd, image, q: Rounding, tap: Optional[Callable] = None, force: Optional[dict] = None):
    x = q.a(F.conv2d(q.v(image), q.w(sd["img_forward.0.weight"]), sd["img_forward.0.bias"], padding=1))
    x = _forced(x, force, "img_forward.0")
    if tap:
        tap("img_forward.0", x)
    k = 1
    while f"img_forward.{k}.gamma" in sd:
        x = d_block(sd, f"img_forward.{k}", x, q, tap, force)
        k += 1
    return x


def discriminator_logits(sd, embed, sent, q: Rounding, force: Optional[dict] = None):
    tiled = sent.reshape(-1, O.SENT_DIM, 1, 1).expand(-1, -1, 4, 4)
    h = q.a(torch.cat((embed, tiled), dim=1))
    pre = q.g(F.conv2d(h, q.w(sd["img_sentence_forward.0.weight"]), None, padding=1))
    h = q.v(_lrelu(pre, force, "head.hid"))
    return F.conv2d(h, sd["img_sentence_forward.2.weight"], None)


def force_from_trunk_ctx(ctx: dict, head_ctx: Optional[dict] = None) -> dict:
    """Pipeline activations (dnet.trunk_forward / head_forward ctx, NHWC bf16) -> forcing dict (NCHW fp32)."""
    nchw = lambda t: t.permute(0, 3, 1, 2).float()  # noqa: E731
    out = {}
    saved = ctx["saved"]
    for k, (x, c1, r) in enumerate(saved):
        p = f"img_forward.{k + 1}"
        out["img_forward.0" if k == 0 else f"img_forward.{k}.y"] = nchw(x)
        out[p + ".c1"], out[p + ".r"] = nchw(c1), nchw(r)
    if "emb" in ctx:
        out[f"img_forward.{len(saved)}.y"] = nchw(ctx["emb"])
    if head_ctx is not None:
        out["head.hid"] = nchw(head_ctx["hid"])
    return out


# ------------------------------------------------------------------------------------------------ generator
_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}


def subpixel_conv(x, w, b, q: Rounding):
    """conv3x3(pad 1) of nearest-2x-upsampled x, as four 2x2 phase convolutions with folded weights."""
    n, _, h, wd = x.shape
    out = x.new_zeros(n, w.shape[0], 2 * h, 2 * wd)
    for a in (0, 1):
        for c in (0, 1):
            k = torch.stack([torch.stack([sum(w[:, :, r, s] for r in rs for s in ss) for (_, ss) in _ROWS[c]], -1)
                             for (_, rs) in _ROWS[a]], -2)                  # [co, ci, 2, 2]
            pad = (1 if c == 0 else 0, 0 if c == 0 else 1, 1 if a == 0 else 0, 0 if a == 0 else 1)
            out[:, :, a::2, c::2] = F.conv2d(F.pad(x, pad), q.w(k), b)
    return out


def _affine2(sd, p, first, x, sent):
    h = F.leaky_relu(O.affine(sd, f"{p}.fusion_block_{first}", x, sent), SLOPE)
    return F.leaky_relu(O.affine(sd, f"{p}.fusion_block_{first + 1}", h, sent), SLOPE)


def generator_forward(sd, noise, sent, q: Rounding, tap: Optional[Callable] = None,
                      force: Optional[dict] = None):
    n_up = len({k.split(".")[1] for k in sd if k.startswith("res_blocks.")})
    c0 = sd["linear_in.weight"].shape[0] // 16
    z = q.a(F.linear(noise, sd["linear_in.weight"], sd["linear_in.bias"]).view(noise.shape[0], c0, 4, 4))
    z = _forced(z, force, "z_in")
    names = [f"res_blocks.{i}" for i in range(n_up)] + ["res_block_out"]
    a12 = _forced(q.a(_affine2(sd, names[0], 1, z, sent)), force, names[0] + ".a12")
    for i, p in enumerate(names):
        zs = q.g(z)
        if p + ".scale_conv.weight" in sd:
            s = q.a(F.conv2d(zs, q.w(sd[p + ".scale_conv.weight"]), sd[p + ".scale_conv.bias"]))
        else:
            s = zs
        if i == 0:
            h1 = F.conv2d(a12, q.w(sd[p + ".conv_1.weight"]), sd[p + ".conv_1.bias"], padding=1)
        else:
            h1 = subpixel_conv(a12, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"], q)
            s = F.interpolate(s, scale_factor=2)
        h1 = _forced(q.a(h1), force, p + ".h1")
        if tap:
            tap(p + ".conv_1", h1)
        a34 = _forced(q.a(_affine2(sd, p, 3, h1, sent)), force, p + ".a34")
        h2 = F.conv2d(a34, q.w(sd[p + ".conv_2.weight"]), sd[p + ".conv_2.bias"], padding=1)
        z = _forced(q.a(s + sd[p + ".gamma"] * h2), force, p + ".z")
        if tap:
            tap(p, z)
        if i + 1 < len(names):
            a12 = _forced(q.a(_affine2(sd, names[i + 1], 1, z, sent)), force, names[i + 1] + ".a12")
    act = _forced(q.a(F.leaky_relu(z, SLOPE)), force, "act")
    pre = q.g(F.conv2d(act, q.w(sd["conv_out.1.weight"]), sd["conv_out.1.bias"], padding=1))
    return _forced(torch.tanh(pre), force, "img")


def force_from_g_ctx(names, ctx: dict) -> dict:
    """gnet.generator_forward ctx (NHWC bf16) -> forcing dict (NCHW fp32).  names: block prefixes in order."""
    nchw = lambda t: t.permute(0, 3, 1, 2).float()  # noqa: E731
    out = {}
    saved = ctx["saved"]
    for i, (p, (z_prev, a12, h1, a34)) in enumerate(zip(names, saved)):
        out["z_in" if i == 0 else names[i - 1] + ".z"] = nchw(z_prev)
        out[p + ".a12"], out[p + ".h1"], out[p + ".a34"] = nchw(a12), nchw(h1), nchw(a34)
    out[names[-1] + ".z"] = nchw(ctx["z_last"])
    out["act"] = nchw(ctx["act_last"])
    out["img"] = ctx["img"].float()
    return out


# ------------------------------------------------------------------------------------------------ losses
def d_hinge_loss(d_sd, images, fake_images, sent, q: Rounding):
    b = images.shape[0]
    real_emb = discriminator_embeds(d_sd, images, q)
    loss_real = torch.relu(1.0 - discriminator_logits(d_sd, real_emb, sent, q)).mean()
    loss_mis = torch.relu(1.0 + discriminator_logits(d_sd, real_emb[:b - 1], sent[1:b], q)).mean()
    fake_emb = discriminator_embeds(d_sd, fake_images, q)
    loss_fake = torch.relu(1.0 + discriminator_logits(d_sd, fake_emb, sent, q)).mean()
    return loss_real + (loss_fake + loss_mis) / 2.0


def g_hinge_loss(d_sd, fake_images, sent, q: Rounding, force: Optional[dict] = None):
    return -discriminator_logits(d_sd, discriminator_embeds(d_sd, fake_images, q, None, force), sent, q,
                                 force).mean()


def magp_first_order(d_sd, images, sent, q: Rounding):
    """First-order quantities of the MA-GP (gx, gs, clamped norms) with emulated storage rounding."""
    x = images.detach().clone().requires_grad_()
    s = sent.detach().clone().requires_grad_()
    logits = discriminator_logits(d_sd, discriminator_embeds(d_sd, x, q), s, q)
    gx, gs = torch.autograd.grad(logits, (x, s), torch.ones_like(logits))
    g = torch.cat((gx.reshape(x.shape[0], -1), gs.reshape(x.shape[0], -1)), dim=1)
    return torch.clamp(torch.sqrt((g ** 2).sum(dim=1) + 1e-8), 0, 10), gx, gs


def magp_loss(d_sd, images, sent, q: Rounding, probe: Optional[dict] = None, force: Optional[dict] = None):
    """2 * mean(n^6) with the double backward available (create_graph), emulated storage rounding."""
    x = images.detach().clone().requires_grad_()
    s = sent.detach().clone().requires_grad_()
    logits = discriminator_logits(d_sd, discriminator_embeds(d_sd, x, q, None, force), s, q, force)
    gx, gs = torch.autograd.grad(logits, (x, s), torch.ones_like(logits), retain_graph=True, create_graph=True)
    g = torch.cat((gx.reshape(x.shape[0], -1), gs.reshape(x.shape[0], -1)), dim=1)
    n = torch.clamp(torch.sqrt((g ** 2).sum(dim=1) + 1e-8), 0, 10)
    if probe is not None:
        probe.update(gp_n=n.detach(), gp_gx=gx.detach(), gp_gs=gs.detach())
    return 2.0 * torch.mean(n ** 6)
