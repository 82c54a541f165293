"""Kernel programs of the discriminator: trunk (`build_embeds`), logit head (`get_logits`), their first-order
backward and the tangent-forward sweep that implements the MA-GP double backward.

Reference arithmetic (Deep_Fusion_GAN/src/): discriminator/model.py:11-59, discriminator/residual_block.py:7-30,
deep_fusion_gan/model.py:59-85 (autograd.grad with create_graph) and :207 (its backward).

Everything here only sequences C-ABI kernel launches on the current stream; tensors are NHWC bf16 between
kernels, fp32 NCHW at the API boundary (images, embeds) and fp32 for parameters / gradients.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import convops as co
from . import kernels as K

BF16, F32 = torch.bfloat16, torch.float32
SENT_DIM = 256
MERGE_MAX = 64   # see gnet.MERGE_MAX
LATE_BLOCKS = 3  # blocks (from the end of the trunk) whose operands are re-packed by the deferred table
import os as _os
DEFER_PACK = _os.environ.get("DFGAN_NO_DEFER_PACK", "0") != "1"     # A/B switches for the two side-stream placements
EARLY_FLUSH = _os.environ.get("DFGAN_NO_EARLY_FLUSH", "0") != "1"
FUSED_STEM = _os.environ.get("DFGAN_NO_FUSED_STEM", "0") != "1"    # A/B: round-1 pair im2col3 + 1x1 GEMM


def d_plan(params: Dict[str, torch.Tensor]) -> List[tuple]:
    """(prefix, C_in, C_out, has_scale_conv) per ResidualBlockD, read off the parameter shapes."""
    plan, k = [], 1
    while f"img_forward.{k}.gamma" in params:
        w = params[f"img_forward.{k}.residual_conv.0.weight"]
        plan.append((f"img_forward.{k}", w.shape[1], w.shape[0], f"img_forward.{k}.scale_conv.weight" in params))
        k += 1
    return plan


TAPS3_ID = [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
TAPS4_ID = [(src, dy, dx, i) for i, (src, dy, dx, _) in enumerate(co.taps_4x4s2(0))]
TAPS_POOL = [(i, 0, 0, 0) for i in range(4)]


class DPacked:
    """bf16 GEMM operands of one discriminator parameter set; `refresh` re-packs after an optimiser step."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.plan = d_plan(params)
        dev = params["img_forward.0.weight"].device
        self.n_c = params["img_forward.0.weight"].shape[0]
        if self.n_c != 32:
            raise ValueError("stem kernel is built for n_c == 32 (reference: Discriminator(n_c=32), "
                             "deep_fusion_gan/model.py:29)")
        e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
        self.w0_f, self.w0_d = e(32, 32), e(9 * 16, 32)
        self.blocks = []
        for (_, cin, cout, sc) in self.plan:
            self.blocks.append(dict(w4_f=e(16 * cout, cin), w4_d=e((36 if cin <= MERGE_MAX else 16) * cin, cout),
                                    w3_f=e(9 * cout, cout),
                                    w3_d=e(9 * cout, cout), ws_f=e(cout, cin) if sc else None,
                                    ws_d=e(cin, cout) if sc else None))
        wh = params["img_sentence_forward.0.weight"]
        self.ch, self.ccat = wh.shape[0], wh.shape[1]
        self.wh_f, self.wh_d = e(9 * self.ch, self.ccat), e(9 * self.ccat, self.ch)
        self._ptrs = None
        # tap-major scratch for the weight gradients of every 4x4 / 3x3 convolution (coalesced fp32 atomics),
        # folded back into the OIHW gradients once per sweep (K.WgradScratch)
        shapes = {}
        for (p, _, _, _) in self.plan:
            for nm in (p + ".residual_conv.0.weight", p + ".residual_conv.2.weight"):
                shapes[nm] = tuple(params[nm].shape)
        shapes["img_sentence_forward.0.weight"] = tuple(wh.shape)
        self.ws = K.WgradScratch(shapes, dev)
        self.refresh(params)

    def _build(self, params: Dict[str, torch.Tensor]) -> None:
        """Two re-pack tables: `_pp` (stem and the high-resolution blocks: few parameters, consumed first) and
        `_pp_late` (the last LATE_BLOCKS blocks and the head: > 90 % of the parameters, first consumed several hundred
        microseconds into a forward sweep) -- the trainer runs the second one next to the start of the next sweep."""
        dev = params["img_forward.0.weight"].device
        early, late = K.PackPlan(dev), K.PackPlan(dev)
        self.first_late = max(len(self.plan) - LATE_BLOCKS, 0)
        pp = early
        pp.add(params["img_forward.0.weight"], self.w0_f, as_matrix=True, cols_pad=32)
        pp.add(params["img_forward.0.weight"], self.w0_d, transpose=True, rows_pad=16)
        for i, ((p, cin, cout, sc), b) in enumerate(zip(self.plan, self.blocks)):
            pp = late if i >= self.first_late else early
            w4, w3 = params[p + ".residual_conv.0.weight"], params[p + ".residual_conv.2.weight"]
            pp.add_with_image(w4, b["w4_f"])
            if cin <= MERGE_MAX:
                pp.add_with_image(w4, b["w4_d"], transpose=True, slots=co.s2dgrad_merged_slots(), group=4)
            else:
                pp.add_with_image(w4, b["w4_d"], transpose=True)
            pp.add_with_image(w3, b["w3_f"])
            pp.add_with_image(w3, b["w3_d"], transpose=True)
            if sc:
                pp.add(params[p + ".scale_conv.weight"], b["ws_f"], scale=0.25)
                pp.add(params[p + ".scale_conv.weight"], b["ws_d"], transpose=True)
        pp = late
        pp.add(params["img_sentence_forward.0.weight"], self.wh_f)
        pp.add(params["img_sentence_forward.0.weight"], self.wh_d, transpose=True)
        self._pp, self._pp_late = early, late
        self._late_pending = False

    def refresh(self, params: Dict[str, torch.Tensor], defer_late: bool = False) -> None:
        """Re-pack every GEMM operand from the fp32 master weights: two kernel launches.  defer_late: the second
        (large) table runs on a side stream; every consumer of a late operand calls `sync_late` first."""
        ptrs = tuple(t.data_ptr() for t in params.values())
        if ptrs != self._ptrs:
            self._build(params)
            self._ptrs = ptrs
        self.sync_late()
        self._pp.run()
        if defer_late and co.SIDE_ENABLED and DEFER_PACK:
            dev = params["img_forward.0.weight"].device
            with co.side(0, dev, co.POOL_AUX):
                self._pp_late.run()
            self._late_pending = True
        else:
            self._pp_late.run()

    def sync_late(self) -> None:
        if self._late_pending:
            self._late_pending = False
            co.join(self._pp_late.device, co.POOL_AUX)

    def prune(self, used: set) -> None:
        self._pp.prune(used)
        self._pp_late.prune(used)


def _use_operand(t: torch.Tensor, what: str) -> None:
    """Packed operands consumed outside convops.conv_run take part in the same usage tracking / pruning protocol."""
    if co.TRACK_USAGE:
        co.USED.add(t.data_ptr())
    else:
        co.check_not_pruned(t, what)


def _oihw_strides(w: torch.Tensor) -> tuple:
    co_, ci, kh, kw = w.shape
    return (ci * kh * kw, kh * kw, 1)


def _check_image(img: torch.Tensor, n_blocks: int) -> None:
    if img.dim() != 4 or img.shape[1] != 3 or img.shape[2] != img.shape[3] or img.shape[2] != 4 << n_blocks:
        # same failure the reference produces at the torch.cat of get_logits (discriminator/model.py:54)
        raise RuntimeError(f"Sizes of tensors must match: image {tuple(img.shape)} does not reduce to a 4x4 "
                           f"embed grid through {n_blocks} down-blocks")


# ------------------------------------------------------------------------------------------------ trunk
class Stem:
    """Output of the stem convolution for a batch that arrives in parts (`stem_alloc` + `stem_run` per part): the
    trainer's D step runs the real half on a side stream next to the generator forward and the generated half when
    the images exist -- no [real ; fake] concatenation, and the HBM-bound stem of the real images leaves the chain."""

    def __init__(self, x: torch.Tensor, xcol: Optional[torch.Tensor]):
        self.x, self.xcol = x, xcol


def stem_alloc(pk: DPacked, n: int, h: int, w: int, device, need_xcol: bool = True) -> Stem:
    e = lambda *s: torch.empty(*s, device=device, dtype=BF16)  # noqa: E731
    return Stem(e(n, h, w, pk.n_c), e(n, h, w, 32) if need_xcol else None)


def stem_run(pk: DPacked, params, img: torch.Tensor, stem: Stem, first: int = 0) -> None:
    """discriminator/model.py:13 (img_forward.0) on samples [first, first + len(img)) of the stem buffers."""
    _check_image(img, len(pk.plan))
    n = img.shape[0]
    x = stem.x[first:first + n]
    xcol = stem.xcol[first:first + n] if stem.xcol is not None else None
    if tuple(x.shape[1:3]) != tuple(img.shape[2:]) or x.shape[0] != n:
        raise RuntimeError(f"stem_run: image part {tuple(img.shape)} does not fit rows [{first}, {first + n}) of "
                           f"{tuple(stem.x.shape)}")
    if FUSED_STEM:
        _use_operand(pk.w0_f, "stem_fwd")
        K.stem_fwd(img, pk.w0_f, x, bias=params["img_forward.0.bias"], xcol=xcol)
    else:
        tmp = xcol if xcol is not None else torch.empty(n, x.shape[1], x.shape[2], 32, device=img.device, dtype=BF16)
        K.im2col3(img, tmp)
        co.conv_run([tmp], co.taps_1x1(), pk.w0_f, pk.n_c, x, bias=params["img_forward.0.bias"])


def trunk_forward(pk: DPacked, params, img: Optional[torch.Tensor], ctx: Optional[dict] = None,
                  need_xcol: bool = True, stem: Optional[Stem] = None) -> torch.Tensor:
    """discriminator/model.py:36-41.  img: fp32 NCHW [N,3,H,H] -> embeds NHWC bf16 [N,4,4,16*n_c].
    ctx (optional dict) receives the activations the backward / tangent sweeps need; need_xcol=False skips the
    im2col tensor of the stem (only its weight gradient reads it: a backward with grads=None never does).
    stem: the stem outputs computed beforehand (`stem_alloc` / `stem_run`), img is then ignored."""
    if stem is None:
        _check_image(img, len(pk.plan))
        n, _, h, w = img.shape
        stem = stem_alloc(pk, n, h, w, img.device, need_xcol=ctx is not None and need_xcol)
        stem_run(pk, params, img, stem)
    x, xcol = stem.x, stem.xcol
    n, h, w = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
    saved = []
    for i, ((p, cin, cout, sc), b) in enumerate(zip(pk.plan, pk.blocks)):
        if i == pk.first_late:
            pk.sync_late()
        h, w = h // 2, w // 2
        pv = co.parity_views(x)
        s = e(n, h, w, cout)
        if sc:   # avg_pool2d(conv1x1(x)+b) == conv1x1(avg_pool2d(x))+b, pooling folded into 4 taps (App. C-2)
            co.conv_run(pv, TAPS_POOL, b["ws_f"], cout, s, bias=params[p + ".scale_conv.bias"])
        else:
            K.pool(x, s, shift=1, scale=0.25)
        c1 = e(n, h, w, cout)
        co.conv_run(pv, co.taps_4x4s2(cout), b["w4_f"], cout, c1, lrelu_pre=True)
        y, r = e(n, h, w, cout), e(n, h, w, cout)
        co.conv_run([c1], co.taps_3x3(cout), b["w3_f"], cout, y, lrelu_pre=True, aux=r,
                    scale_dev=params[p + ".gamma"], addend=s)
        saved.append((x, c1, r))
        x = y
    if ctx is not None:
        ctx.update(xcol=xcol, saved=saved, n=n, emb=x)
    return x


def trunk_backward(pk: DPacked, params, ctx: dict, g_emb: torch.Tensor, grads: Optional[dict],
                   need_gx: bool, cots: Optional[dict] = None, on_late=None) -> Optional[torch.Tensor]:
    """First-order backward of the trunk.  g_emb: NHWC bf16 cotangent of the embeds.  grads: name -> fp32
    accumulator (None = skip all parameter gradients, i.e. the G step / the first GP sweep).  cots (optional
    dict) receives the cotangent at every conv output for the MA-GP second sweep.  Returns d/d(image) as fp32
    NCHW when need_gx.  on_late(prefix) (optional) is called on the fold-back side stream once every parameter
    gradient of the blocks from `prefix` to the end of the trunk is final (data parallel: their exchange starts
    there, next to the rest of the sweep); it is not called when the early fold-back is off."""
    pk.sync_late()
    dev = g_emb.device
    n = ctx["n"]
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
    g_y = g_emb
    kept = []     # also keeps every tensor a side stream reads alive until the join below
    n_side = 0
    flushed: set = set()
    for (p, cin, cout, sc), b, (x, c1, r) in zip(reversed(pk.plan), reversed(pk.blocks), reversed(ctx["saved"])):
        h, w = c1.shape[1], c1.shape[2]
        g_c2 = e(n, h, w, cout)
        K.dres_bwd(g_y, r, g_c2, params[p + ".gamma"],
                   dgamma=grads[p + ".gamma"] if grads is not None else None,
                   colsum=grads[p + ".scale_conv.bias"] if (grads is not None and sc) else None)
        g_c1 = e(n, h, w, cout)
        co.conv_run([g_c2], co.taps_3x3_dgrad(cout), b["w3_d"], cout, g_c1, mask_src=c1)
        if sc:
            g_s = e(n, h, w, cin)
            co.conv_run([g_y], co.taps_1x1(), b["ws_d"], cin, g_s)
        else:
            g_s = g_y
        g_x = e(n, 2 * h, 2 * w, cin)
        if cin <= MERGE_MAX:
            # data gradient of the stride-2 conv: the four input phases in one launch (9 shifts, N = 4*cin)
            co.conv_run([g_c1], co.taps_merged(cin), b["w4_d"], 4 * cin, g_x, grid=(n, h, w), addend=g_s,
                        add_scale=0.25, phase_cout=cin)
        else:
            # the four input phases are independent launches: phases 1..3 on side streams (parallel graph branches)
            views = [g_x[:, a::2, c::2, :] for a in (0, 1) for c in (0, 1)]
            for ph in (1, 2, 3):
                with co.side(ph - 1, dev, co.POOL_PHASE):
                    co.conv_run([g_c1], co.taps_4x4s2_dgrad(cin, ph >> 1, ph & 1), b["w4_d"], cin, views[ph],
                                addend=g_s, add_scale=0.25)
            co.conv_run([g_c1], co.taps_4x4s2_dgrad(cin, 0, 0), b["w4_d"], cin, views[0], addend=g_s,
                        add_scale=0.25)
            co.join(dev, co.POOL_PHASE)
        if grads is not None:
            # weight gradients leave the data-gradient chain: side stream, joined at the end of the sweep
            pv = co.parity_views(x)
            n3, n4 = p + ".residual_conv.2.weight", p + ".residual_conv.0.weight"
            with co.side(n_side, dev):
                co.wgrad_run([c1], TAPS3_ID, g_c2, pk.ws.view(n3), pk.ws.strides(n3))
            with co.side(n_side + 1, dev):
                co.wgrad_run(pv, TAPS4_ID, g_c1, pk.ws.view(n4), pk.ws.strides(n4))
                if sc:
                    ws = params[p + ".scale_conv.weight"]
                    co.wgrad_run(pv, TAPS_POOL, g_y, grads[p + ".scale_conv.weight"], _oihw_strides(ws),
                                 scale_imm=0.25)
            n_side += 2
        kept.append((g_y, g_c2, g_c1, g_s))
        g_y = g_x
        if EARLY_FLUSH and grads is not None and len(kept) == len(pk.plan) - pk.first_late and pk.first_late > 0:
            # the low-resolution blocks hold > 90 % of the weights: fold their gradients back (tap-major scratch ->
            # OIHW) as soon as their weight-gradient launches are queued, on a side stream next to the rest of the sweep
            late_names = [nm for (p_, _, _, _) in pk.plan[pk.first_late:]
                          for nm in (p_ + ".residual_conv.0.weight", p_ + ".residual_conv.2.weight")]
            with co.side(1, dev, co.POOL_AUX, after_pool=co.POOL_WGRAD):
                pk.ws.flush(grads, late_names)
                if on_late is not None:
                    on_late(pk.plan[pk.first_late][0])
            flushed = set(late_names)
    if grads is not None:
        if ctx["xcol"] is None:
            raise RuntimeError("trunk_backward: parameter gradients need the stem im2col tensor "
                               "(trunk_forward(..., need_xcol=True))")
        with co.side(n_side, dev):
            co.wgrad_run([ctx["xcol"]], [(0, 0, 0, 0)], g_y, grads["img_forward.0.weight"], (27, 1, 0), ci_valid=27)
        K.pool(g_y, None, shift=0, colsum=grads["img_forward.0.bias"])
    if cots is not None:
        cots.update(blocks=[k[:3] for k in reversed(kept)], x0=g_y)
    gx = None
    if need_gx:
        hh = g_y.shape[1]
        gx = torch.empty(n, 3, hh, hh, device=dev, dtype=F32)
        co.conv_run([g_y], co.taps_3x3_dgrad(16), pk.w0_d, 16, None, out_nchw=gx, co_valid=3)
    co.join(dev)
    if grads is not None:
        co.join(dev, co.POOL_AUX)
        pk.ws.flush(grads, [n for n in pk.ws.shapes if n.startswith("img_forward.") and n not in flushed])
    return gx


def trunk_tangent(pk: DPacked, params, ctx: dict, cots: dict, t_img: torch.Tensor,
                  rowscale: Optional[torch.Tensor], grads: dict) -> torch.Tensor:
    """Tangent-forward sweep with frozen LeakyReLU masks: pushes the tangent image (scaled per sample by
    rowscale) through the linearised trunk and accumulates, per conv, wgrad(tangent input, first-order
    cotangent at its output) into grads -- exactly the parameter gradient of <tangent, d logits/d image>
    (aten::_convolution_double_backward at deep_fusion_gan/model.py:207 without the zero-gradient waste).
    Bias gradients stay zero (gotcha G3).  Returns the tangent of the embeds (NHWC bf16)."""
    pk.sync_late()
    n, _, h, w = t_img.shape
    dev = t_img.device
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
    tcol = e(n, h, w, 32)
    t_x = e(n, h, w, pk.n_c)
    if FUSED_STEM:
        _use_operand(pk.w0_f, "stem_fwd")
        K.stem_fwd(t_img, pk.w0_f, t_x, rowscale=rowscale, xcol=tcol)
    else:
        K.im2col3(t_img, tcol, rowscale=rowscale)
        co.conv_run([tcol], co.taps_1x1(), pk.w0_f, pk.n_c, t_x)
    # every weight gradient only consumes the tangent chain: side streams, joined at the end of the sweep
    with co.side(0, dev):
        co.wgrad_run([tcol], [(0, 0, 0, 0)], cots["x0"], grads["img_forward.0.weight"], (27, 1, 0), ci_valid=27)
    keep, n_side = [tcol], 1
    for (p, cin, cout, sc), b, (x, c1, r), (c_y, c_c2, c_c1) in zip(pk.plan, pk.blocks, ctx["saved"],
                                                                  cots["blocks"]):
        h, w = h // 2, w // 2
        pv = co.parity_views(t_x)
        t_s = e(n, h, w, cout)
        if sc:
            co.conv_run(pv, TAPS_POOL, b["ws_f"], cout, t_s)
        else:
            K.pool(t_x, t_s, shift=1, scale=0.25)
        with co.side(n_side, dev):
            if sc:
                co.wgrad_run(pv, TAPS_POOL, c_y, grads[p + ".scale_conv.weight"],
                             _oihw_strides(params[p + ".scale_conv.weight"]), scale_imm=0.25)
            n4 = p + ".residual_conv.0.weight"
            co.wgrad_run(pv, TAPS4_ID, c_c1, pk.ws.view(n4), pk.ws.strides(n4))
        t_c1 = e(n, h, w, cout)
        co.conv_run(pv, co.taps_4x4s2(cout), b["w4_f"], cout, t_c1, mask_src=c1)
        with co.side(n_side + 1, dev):
            n3 = p + ".residual_conv.2.weight"
            co.wgrad_run([t_c1], TAPS3_ID, c_c2, pk.ws.view(n3), pk.ws.strides(n3))
        t_y, t_r = e(n, h, w, cout), e(n, h, w, cout)
        co.conv_run([t_c1], co.taps_3x3(cout), b["w3_f"], cout, t_y, mask_src=r, aux=t_r,
                    scale_dev=params[p + ".gamma"], addend=t_s)
        K.dot(c_y, t_r, grads[p + ".gamma"])
        keep += [t_x, t_c1]
        n_side += 2
        t_x = t_y
    co.join(dev)
    pk.ws.flush(grads, [n for n in pk.ws.shapes if n.startswith("img_forward.")])
    return t_x


# ------------------------------------------------------------------------------------------------ head
def head_forward(pk: DPacked, params, embed: torch.Tensor, sent: torch.Tensor, ctx: Optional[dict] = None,
                 sscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """discriminator/model.py:43-59.  embed fp32 NCHW [b,16*n_c,4,4], sent fp32 [b,256] -> logits fp32 [b]."""
    pk.sync_late()
    if embed.dim() != 4 or tuple(embed.shape[2:]) != (4, 4) or sent.shape[-1] != SENT_DIM or \
            embed.shape[1] + SENT_DIM != pk.ccat or sent.shape[0] != embed.shape[0]:
        raise RuntimeError(f"Sizes of tensors must match: embed {tuple(embed.shape)} / sentence "
                           f"{tuple(sent.shape)} (discriminator/model.py:49-54)")
    b = embed.shape[0]
    dev = embed.device
    cat = torch.empty(b, 4, 4, pk.ccat, device=dev, dtype=BF16)
    K.head_cat(embed, sent, cat, sscale=sscale)
    hid = torch.empty(b, 4, 4, pk.ch, device=dev, dtype=BF16)
    co.conv_run([cat], co.taps_3x3(pk.ch), pk.wh_f, pk.ch, hid, lrelu_pre=True)
    logits = torch.empty(b, device=dev, dtype=F32)
    K.head_out(hid, params["img_sentence_forward.2.weight"], logits)
    if ctx is not None:
        ctx.update(cat=cat, hid=hid, b=b)
    return logits


def head_backward(pk: DPacked, params, ctx: dict, g_logit: Optional[torch.Tensor], grads: Optional[dict],
                  need_sent: bool, cots: Optional[dict] = None):
    """g_logit fp32 [b] (None = ones).  Returns (g_embed fp32 NCHW, g_sent fp32 [b,256] or None)."""
    pk.sync_late()
    b, cat, hid = ctx["b"], ctx["cat"], ctx["hid"]
    dev = cat.device
    w2 = params["img_sentence_forward.2.weight"]
    g_hid = torch.empty_like(hid)
    K.head_out_bwd(g_logit, hid, hid, w2, g_hid,
                   grads["img_sentence_forward.2.weight"] if grads is not None else None)
    g_cat = torch.empty_like(cat)
    co.conv_run([g_hid], co.taps_3x3_dgrad(pk.ccat), pk.wh_d, pk.ccat, g_cat)
    if grads is not None:
        nh = "img_sentence_forward.0.weight"
        co.wgrad_run([cat], TAPS3_ID, g_hid, pk.ws.view(nh), pk.ws.strides(nh))
        pk.ws.flush(grads, [nh])
    ce = pk.ccat - SENT_DIM
    g_embed = torch.empty(b, ce, 4, 4, device=dev, dtype=F32)
    g_sent = torch.empty(b, SENT_DIM, device=dev, dtype=F32) if need_sent else None
    K.head_cat_bwd(g_cat, g_embed, g_sent)
    if cots is not None:
        cots.update(hid=g_hid, logit=g_logit)
    return g_embed, g_sent


def head_tangent(pk: DPacked, params, ctx: dict, cots: dict, t_embed: torch.Tensor, t_sent: torch.Tensor,
                 sscale: Optional[torch.Tensor], grads: dict) -> torch.Tensor:
    """Tangent sweep of the head (second half of App. C-3).  t_embed fp32 NCHW, t_sent fp32 [b,256] (scaled per
    sample by sscale).  Accumulates the two head weight gradients, returns the tangent logits fp32 [b]."""
    pk.sync_late()
    b, hid = ctx["b"], ctx["hid"]
    dev = hid.device
    t_cat = torch.empty_like(ctx["cat"])
    K.head_cat(t_embed, t_sent, t_cat, sscale=sscale)
    t_hid = torch.empty_like(hid)
    co.conv_run([t_cat], co.taps_3x3(pk.ch), pk.wh_f, pk.ch, t_hid, mask_src=hid)
    nh = "img_sentence_forward.0.weight"
    co.wgrad_run([t_cat], TAPS3_ID, cots["hid"], pk.ws.view(nh), pk.ws.strides(nh))
    pk.ws.flush(grads, [nh])
    K.head_out_bwd(cots["logit"], None, t_hid, params["img_sentence_forward.2.weight"], None,
                   grads["img_sentence_forward.2.weight"])
    t_logit = torch.empty(b, device=dev, dtype=F32)
    K.head_out(t_hid, params["img_sentence_forward.2.weight"], t_logit)
    return t_logit


def embeds_to_nchw(emb: torch.Tensor) -> torch.Tensor:
    n, h, w, c = emb.shape
    out = torch.empty(n, c, h, w, device=emb.device, dtype=F32)
    K.nhwc_to_nchw(emb, out)
    return out


def embeds_to_nhwc(emb: torch.Tensor) -> torch.Tensor:
    n, c, h, w = emb.shape
    out = torch.empty(n, h, w, c, device=emb.device, dtype=BF16)
    K.nchw_to_nhwc(emb.contiguous(), out)
    return out
