"""Kernel programs of the generator: the 56 gamma/beta MLPs + linear_in as batched launches, the
ResidualBlockG stack with the nearest-2x upsample folded into sub-pixel convolutions, and its backward.

Reference arithmetic (Deep_Fusion_GAN/src/): generator/model.py:32-47, generator/residual_block.py:29-59,
generator/fusion_block.py:33-39.  Identities used: C-1 (the pointwise prefix and the 1x1
shortcut of a block commute with the preceding upsample, so they run at the low resolution) and C-4
(upsample -> conv3x3 == four 2x2 phase convolutions on the low-res input with row/column-folded weights).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import convops as co
from . import kernels as K
from .dnet import TAPS3_ID, _oihw_strides

BF16, F32 = torch.bfloat16, torch.float32
SENT_DIM, HID = 256, 256
MERGE_MAX = 64   # widest phase (channels) for which the merged 9-shift launch beats four 4-tap launches
MLPS = [(fb, m) for fb in (1, 2, 3, 4) for m in ("gamma_mlp", "beta_mlp")]
MLP_BWD_PER_BLOCK = __import__("os").environ.get("DFGAN_MLP_BWD_WHOLE", "0") != "1"    # A/B: one launch set at the end


def g_plan(params: Dict[str, torch.Tensor]) -> List[tuple]:
    """(prefix, C_in, C_out, has_scale_conv) per ResidualBlockG in execution order."""
    names = []
    i = 0
    while f"res_blocks.{i}.gamma" in params:
        names.append(f"res_blocks.{i}")
        i += 1
    names.append("res_block_out")
    plan = []
    for p in names:
        w = params[p + ".conv_1.weight"]
        plan.append((p, w.shape[1], w.shape[0], p + ".scale_conv.weight" in params))
    return plan


class GPacked:
    """bf16 GEMM operands of one generator parameter set."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.plan = g_plan(params)
        dev = params["linear_in.weight"].device
        e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
        self.c0 = params["linear_in.weight"].shape[0] // 16
        self.latent = params["linear_in.weight"].shape[1]
        self.blocks = []
        for i, (_, c1, c2, sc) in enumerate(self.plan):
            if c1 % 32 or c2 % 32:
                raise ValueError("channel counts must be multiples of 32 (n_channels % 32 == 0)")
            taps1 = 9 if i == 0 else 16
            # up-blocks: the four sub-pixel phases run as ONE launch over 9 shifts when 4*c2 <= 256 (an MMA costs
            # the same for any N <= 256, so the zero blocks are free); wider layers keep four 4-tap launches
            n1 = 9 if i == 0 else (36 if c2 <= MERGE_MAX else 16)
            self.blocks.append(dict(w1_f=e(n1 * c2, c1), w1_d=e(taps1 * c1, c2), w2_f=e(9 * c2, c2),
                                    w2_d=e(9 * c2, c2), ws_f=e(c2, c1) if sc else None,
                                    ws_d=e(c1, c2) if sc else None))
        self.c_last = self.plan[-1][2]
        self.wo_f, self.wo_d = e(9 * 16, self.c_last), e(9 * self.c_last, 32)
        self.bo16 = torch.zeros(16, device=dev, dtype=F32)
        self._ptrs = None
        # tap-major scratch for the 3x3 weight gradients that are accumulated straight into OIHW tensors (conv_2 of
        # every block, conv_1 of the first block, conv_out); see K.WgradScratch
        shapes = {p + ".conv_2.weight": tuple(params[p + ".conv_2.weight"].shape) for (p, _, _, _) in self.plan}
        shapes[self.plan[0][0] + ".conv_1.weight"] = tuple(params[self.plan[0][0] + ".conv_1.weight"].shape)
        shapes["conv_out.1.weight"] = tuple(params["conv_out.1.weight"].shape)
        self.ws = K.WgradScratch(shapes, dev)
        self.refresh(params)

    def _build(self, params: Dict[str, torch.Tensor]) -> None:
        pp = K.PackPlan(params["linear_in.weight"].device)
        slots = co.subpixel_slots()
        for i, ((p, c1, c2, sc), b) in enumerate(zip(self.plan, self.blocks)):
            w1, w2 = params[p + ".conv_1.weight"], params[p + ".conv_2.weight"]
            merged = i > 0 and c2 <= MERGE_MAX
            pp.add_with_image(w1, b["w1_f"], slots=None if i == 0 else (co.subpixel_merged_slots() if merged else slots),
                              group=4 if merged else 1)
            pp.add_with_image(w1, b["w1_d"], transpose=True, slots=None if i == 0 else slots)
            pp.add_with_image(w2, b["w2_f"])
            pp.add_with_image(w2, b["w2_d"], transpose=True)
            if sc:
                pp.add(params[p + ".scale_conv.weight"], b["ws_f"])
                pp.add(params[p + ".scale_conv.weight"], b["ws_d"], transpose=True)
        pp.add(params["conv_out.1.weight"], self.wo_f, rows_pad=16)
        pp.add(params["conv_out.1.weight"], self.wo_d, transpose=True, cols_pad=32)
        self._pp = pp

    def refresh(self, params: Dict[str, torch.Tensor]) -> None:
        """Re-pack every GEMM operand from the fp32 master weights: one kernel launch (+ the padded bias)."""
        ptrs = tuple(t.data_ptr() for t in params.values())
        if ptrs != self._ptrs:
            self._build(params)
            self._ptrs = ptrs
        self._pp.run()
        self.bo16[:3].copy_(params["conv_out.1.bias"])


class GMlp:
    """Buffers + device work tables of the batched linears for one batch size: layer 1 of all 56 MLPs and
    linear_in in one launch, layer 2 in a second one; three launches for the backward.  `grads` maps parameter
    names to fp32 accumulators whose addresses are baked into the backward tables."""

    def __init__(self, params: Dict[str, torch.Tensor], plan, batch: int, grads: Optional[Dict[str, torch.Tensor]]):
        dev = params["linear_in.weight"].device
        self.batch = batch
        self.offsets = {}
        off = 0
        names = []
        for (p, c1, c2, _) in plan:
            for fb, m in MLPS:
                width = c1 if fb <= 2 else c2
                self.offsets[(p, fb, m)] = (off, width)
                names.append((f"{p}.fusion_block_{fb}.{m}", off, width))
                off += width
        self.T = off
        self.n_mlp = len(names)
        z = lambda *s: torch.zeros(*s, device=dev, dtype=F32)  # noqa: E731
        self.sent, self.noise = z(batch, SENT_DIM), z(batch, params["linear_in.weight"].shape[1])
        self.hid, self.tab = z(batch, self.n_mlp * HID), z(batch, self.T)
        self.x0f = z(batch, params["linear_in.weight"].shape[0])
        hs = self.n_mlp * HID
        l1, l2 = [], []
        for j, (name, o, width) in enumerate(names):
            l1.append(dict(inp=self.sent.data_ptr(), in_stride=SENT_DIM, w=params[name + ".0.weight"].data_ptr(),
                           bias=params[name + ".0.bias"].data_ptr(), out=self.hid.data_ptr() + 4 * j * HID,
                           out_stride=hs, k=SENT_DIM, n_out=HID, relu=1))
            l2.append(dict(inp=self.hid.data_ptr() + 4 * j * HID, in_stride=hs,
                           w=params[name + ".2.weight"].data_ptr(), bias=params[name + ".2.bias"].data_ptr(),
                           out=self.tab.data_ptr() + 4 * o, out_stride=self.T, k=HID, n_out=width))
        lw = params["linear_in.weight"]
        l1.append(dict(inp=self.noise.data_ptr(), in_stride=lw.shape[1], w=lw.data_ptr(),
                       bias=params["linear_in.bias"].data_ptr(), out=self.x0f.data_ptr(),
                       out_stride=lw.shape[0], k=lw.shape[1], n_out=lw.shape[0]))
        self.t_l1, self.t_l2 = K.LinearTable(l1, 0, dev), K.LinearTable(l2, 0, dev)
        self.bwd = None
        if grads is not None:
            self.d_tab, self.d_hid, self.g_x0f = z(batch, self.T), z(batch, hs), z(batch, lw.shape[0])
            w2, dd, w1 = [], [], []
            for j, (name, o, width) in enumerate(names):
                w2.append(dict(inp=self.hid.data_ptr() + 4 * j * HID, in_stride=hs,
                               w=grads[name + ".2.weight"].data_ptr(), bias=grads[name + ".2.bias"].data_ptr(),
                               out=self.d_tab.data_ptr() + 4 * o, out_stride=self.T, k=HID, n_out=width))
                dd.append(dict(inp=self.d_hid.data_ptr() + 4 * j * HID, in_stride=hs,
                               w=params[name + ".2.weight"].data_ptr(), out=self.d_tab.data_ptr() + 4 * o,
                               out_stride=self.T, mask=self.hid.data_ptr() + 4 * j * HID, k=HID, n_out=width))
                w1.append(dict(inp=self.sent.data_ptr(), in_stride=SENT_DIM,
                               w=grads[name + ".0.weight"].data_ptr(), bias=grads[name + ".0.bias"].data_ptr(),
                               out=self.d_hid.data_ptr() + 4 * j * HID, out_stride=hs, k=SENT_DIM, n_out=HID))
            w1.append(dict(inp=self.noise.data_ptr(), in_stride=lw.shape[1], w=grads["linear_in.weight"].data_ptr(),
                           bias=grads["linear_in.bias"].data_ptr(), out=self.g_x0f.data_ptr(),
                           out_stride=lw.shape[0], k=lw.shape[1], n_out=lw.shape[0]))
            self.bwd = (K.LinearTable(w2, 2, dev), K.LinearTable(dd, 1, dev), K.LinearTable(w1, 2, dev))
        self._keep = (params, grads)

    def forward(self, noise: torch.Tensor, sent: torch.Tensor) -> None:
        self.noise.copy_(noise)
        self.sent.copy_(sent)
        self.t_l1.run(self.batch)
        self.t_l2.run(self.batch)

    def tables(self, prefix: str, first: int, src=None):
        """(ga, ba, gb, bb) views of the AffineBlock pair (first, first+1) of a block."""
        t = self.tab if src is None else src
        out = []
        for fb in (first, first + 1):
            for m in ("gamma_mlp", "beta_mlp"):
                o, wdt = self.offsets[(prefix, fb, m)]
                out.append(t[:, o:o + wdt])
        return tuple(out)

    def backward(self) -> None:
        """d_tab (accumulated by the affine backward kernels) and g_x0f -> MLP / linear_in parameter grads."""
        t_w2, t_dd, t_w1 = self.bwd
        dev = self.tab.device
        with co.side(2, dev, co.POOL_AUX):     # layer-2 weight gradients only read d_tab / hid: parallel branch
            t_w2.run(self.batch)
        t_dd.run(self.batch)
        t_w1.run(self.batch)
        co.join(dev, co.POOL_AUX)

    def backward_block(self, block: int) -> None:
        """The same for the 8 MLPs of generator block `block` (plan order) alone, on a side stream: called by the
        generator backward as soon as the block's two affine backward launches have filled its columns of d_tab, so
        only the last block's share (and linear_in) is left behind the data-gradient chain -- as one launch over all
        MLPs the three kernels ran alone at the end of the G step (143 us at batch 32)."""
        t_w2, t_dd, t_w1 = self.bwd
        per = len(MLPS)
        dev = self.tab.device
        # the three launches are latency-bound (32 CTAs each): consecutive blocks go to different side streams so
        # their chains overlap instead of queueing behind one another
        with co.side(block, dev, co.POOL_AUX):
            t_dd.run(self.batch, block * per, per)      # hidden-layer cotangent
            t_w1.run(self.batch, block * per, per)      # -> layer-1 weight gradients
            t_w2.run(self.batch, block * per, per)      # layer-2 weight gradients (read d_tab / hid only)

    def backward_finish(self) -> None:
        """linear_in's parameter gradients (last problem of the layer-1 table) + join of the per-block branches."""
        t_w1 = self.bwd[2]
        t_w1.run(self.batch, self.n_mlp, 1)
        co.join(self.tab.device, co.POOL_AUX)


def _phases(t: torch.Tensor):
    return [t[:, a::2, c::2, :] for a in (0, 1) for c in (0, 1)]


def generator_forward(pk: GPacked, params, mlp: GMlp, noise: torch.Tensor, sent: torch.Tensor,
                      ctx: Optional[dict] = None) -> torch.Tensor:
    """generator/model.py:32-47 -> fp32 NCHW image [B,3,S,S] (S = 4 * 2**n_upsamples)."""
    if noise.dim() != 2 or noise.shape[1] != pk.latent or sent.shape != (noise.shape[0], SENT_DIM):
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: noise {tuple(noise.shape)}, sentence "
                           f"{tuple(sent.shape)} (generator/model.py:34, generator/fusion_block.py:34)")
    n = noise.shape[0]
    dev = noise.device
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
    mlp.forward(noise, sent)
    z = e(n, 4, 4, pk.c0)
    K.nchw_to_nhwc(mlp.x0f.view(n, pk.c0, 4, 4), z)
    a12 = e(n, 4, 4, pk.c0)
    K.affine2_fwd(z, a12, mlp.tables(pk.plan[0][0], 1), mlp.T)
    h = 4
    saved = []
    for i, ((p, c1, c2, sc), b) in enumerate(zip(pk.plan, pk.blocks)):
        up = i > 0
        hh = 2 * h if up else h
        if sc:
            s = e(n, h, h, c2)
            co.conv_run([z], co.taps_1x1(), b["ws_f"], c2, s, bias=params[p + ".scale_conv.bias"])
        else:
            s = z
        h1, a34 = e(n, hh, hh, c2), e(n, hh, hh, c2)
        t34 = mlp.tables(p, 3)
        if up and c2 <= MERGE_MAX:   # all four sub-pixel phases in one launch (9 shifts, N = 4*c2)
            co.conv_run([a12], co.taps_merged(c2), b["w1_f"], 4 * c2, None, grid=(n, h, h),
                        bias=params[p + ".conv_1.bias"], aux=h1, out2=a34, post_mode=2, tables=t34,
                        tab_stride=mlp.T, phase_cout=c2)
        elif up:
            # four independent phase launches: phases 1..3 on side streams (parallel branches of the CUDA graph)
            def phase(ph):
                a, c = ph >> 1, ph & 1
                co.conv_run([a12], co.taps_subpixel_fwd(c2, a, c), b["w1_f"], c2, None, grid=(n, h, h),
                            bias=params[p + ".conv_1.bias"], aux=h1[:, a::2, c::2, :],
                            out2=a34[:, a::2, c::2, :], post_mode=2, tables=t34, tab_stride=mlp.T)
            for ph in (1, 2, 3):
                with co.side(ph - 1, dev, co.POOL_PHASE):
                    phase(ph)
            phase(0)
            co.join(dev, co.POOL_PHASE)
        else:
            co.conv_run([a12], co.taps_3x3(c2), b["w1_f"], c2, None, bias=params[p + ".conv_1.bias"], aux=h1,
                        out2=a34, post_mode=2, tables=t34, tab_stride=mlp.T)
        z_new, nxt = e(n, hh, hh, c2), e(n, hh, hh, c2)
        last = i == len(pk.plan) - 1
        co.conv_run([a34], co.taps_3x3(c2), b["w2_f"], c2, z_new, bias=params[p + ".conv_2.bias"],
                    scale_dev=params[p + ".gamma"], addend=s, add_shift=1 if up else 0, out2=nxt,
                    post_mode=1 if last else 2, tables=None if last else mlp.tables(pk.plan[i + 1][0], 1),
                    tab_stride=mlp.T)
        saved.append((z, a12, h1, a34))
        z, a12, h = z_new, nxt, hh
    img = torch.empty(n, 3, h, h, device=dev, dtype=F32)
    co.conv_run([a12], co.taps_3x3(16), pk.wo_f, 16, None, bias=pk.bo16, out_nchw=img, co_valid=3, tanh_out=True)
    if ctx is not None:
        ctx.update(saved=saved, z_last=z, act_last=a12, img=img, n=n)
    return img


def generator_backward(pk: GPacked, params, mlp: GMlp, ctx: dict, g_img: torch.Tensor,
                       grads: Dict[str, torch.Tensor]) -> None:
    """Backward of generator_forward: accumulates every parameter gradient into grads (fp32, pre-zeroed by the
    caller; the MLP / linear_in entries must be the tensors mlp was built with)."""
    n = ctx["n"]
    dev = g_img.device
    e = lambda *s: torch.empty(*s, device=dev, dtype=BF16)  # noqa: E731
    img, z, act = ctx["img"], ctx["z_last"], ctx["act_last"]
    s_out = img.shape[2]
    g_t = e(n, s_out, s_out, 32)
    dbo = torch.zeros(3, device=dev, dtype=F32)
    K.tanh_bwd_pack(g_img.contiguous(), img, g_t, dbias=dbo)
    grads["conv_out.1.bias"].add_(dbo)
    # weight gradients leave the data-gradient chain: side streams, joined before the MLP backward; `keep` holds
    # every tensor they read until then
    keep, n_side = [g_t], 1
    with co.side(0, dev):
        co.wgrad_run([act], TAPS3_ID, g_t, pk.ws.view("conv_out.1.weight"), pk.ws.strides("conv_out.1.weight"),
                     co_valid=3)
    g_z = e(*z.shape)
    co.conv_run([g_t], co.taps_3x3_dgrad(pk.c_last), pk.wo_d, pk.c_last, g_z, mask_src=z)
    mlp.d_tab.zero_()
    cs_all = torch.zeros(sum(c2_ for (_, _, c2_, _) in pk.plan), device=dev, dtype=F32)   # per-block channel sums
    cs_off = 0
    for i in range(len(pk.plan) - 1, -1, -1):
        (p, c1, c2, sc), b = pk.plan[i], pk.blocks[i]
        z_prev, a12, h1, a34 = ctx["saved"][i]
        up = i > 0
        hh, h = g_z.shape[1], z_prev.shape[1]
        gamma = params[p + ".gamma"]
        # conv_2: data gradient (unscaled; gamma is applied inside the affine backward) and weight gradient
        q = e(n, hh, hh, c2)
        co.conv_run([g_z], co.taps_3x3_dgrad(c2), b["w2_d"], c2, q)
        with co.side(n_side, dev):
            co.wgrad_run([a34], TAPS3_ID, g_z, pk.ws.view(p + ".conv_2.weight"), pk.ws.strides(p + ".conv_2.weight"),
                         scale_dev=gamma)
        # shortcut branch: 2x2 sum (backward of the nearest upsample) + per-channel sums of g_z
        cs = cs_all[cs_off:cs_off + c2]
        cs_off += c2
        pooled = e(n, h, h, c2) if up else None
        K.pool(g_z, pooled, shift=1 if up else 0, scale=1.0, colsum=cs)
        if not up:
            pooled = g_z
        # conv_2.bias, scale_conv.bias and the bias part of the gate gradient: one small launch off the main chain
        with co.side(n_side + 2, dev):
            K.gblock_small_grads(cs, gamma, params[p + ".conv_2.bias"], grads[p + ".conv_2.bias"],
                                 grads[p + ".gamma"], grads[p + ".scale_conv.bias"] if sc else None)
        # affine 3/4 backward -> gradient at conv_1's output, gamma-gate dot product, conv_1 bias gradient
        g_h1 = e(n, hh, hh, c2)
        K.affine2_bwd(q, h1, g_h1, mlp.tables(p, 3), mlp.tables(p, 3, mlp.d_tab), mlp.T, gscale=gamma,
                      dot=grads[p + ".gamma"], colsum=grads[p + ".conv_1.bias"])
        g_a12 = e(n, h, h, c1)
        w1 = params[p + ".conv_1.weight"]
        if up:
            folded = torch.zeros(16, c2, c1, device=dev, dtype=F32)
            co.conv_run(_phases(g_h1), co.taps_subpixel_dgrad(c1), b["w1_d"], c1, g_a12)
            with co.side(n_side + 1, dev):
                for a in (0, 1):
                    for c in (0, 1):
                        co.wgrad_run([a12], co.taps_subpixel_wgrad(a, c), g_h1[:, a::2, c::2, :], folded,
                                     (c1, 1, c2 * c1))
                K.unfold_subpixel_wgrad(folded, grads[p + ".conv_1.weight"])
            keep.append(folded)
        else:
            co.conv_run([g_h1], co.taps_3x3_dgrad(c1), b["w1_d"], c1, g_a12)
            with co.side(n_side + 1, dev):
                co.wgrad_run([a12], TAPS3_ID, g_h1, pk.ws.view(p + ".conv_1.weight"), pk.ws.strides(p + ".conv_1.weight"))
        if sc:
            t = e(n, h, h, c1)
            co.conv_run([pooled], co.taps_1x1(), b["ws_d"], c1, t)
            with co.side(n_side + 2, dev):
                co.wgrad_run([z_prev], [(0, 0, 0, 0)], pooled, grads[p + ".scale_conv.weight"], (c1, 1, 0))
        else:
            t = pooled
        g_prev = e(n, h, h, c1)
        K.affine2_bwd(g_a12, z_prev, g_prev, mlp.tables(p, 1), mlp.tables(p, 1, mlp.d_tab), mlp.T, g_add=t)
        if MLP_BWD_PER_BLOCK:
            mlp.backward_block(i)
        keep += [g_z, g_h1, pooled]
        n_side += 3
        g_z = g_prev
    co.join(dev)
    pk.ws.flush(grads)
    K.nhwc_to_nchw(g_z, mlp.g_x0f.view(n, pk.c0, 4, 4))
    if MLP_BWD_PER_BLOCK:
        mlp.backward_finish()
    else:
        mlp.backward()
