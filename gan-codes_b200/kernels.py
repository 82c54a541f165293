"""Thin torch-tensor wrappers over the pointwise / MLP / packing / optimiser entry points of the C ABI
(include/dfgan_b200.h).  Every function makes exactly one C call on the current CUDA stream; there is no
fallback path -- a missing library or a non-CUDA tensor raises."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib_decl import LinearItem

BF16, F32 = torch.bfloat16, torch.float32


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _nhwc(t: torch.Tensor):
    assert t.is_cuda and t.dtype == BF16 and t.dim() == 4 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    return t.shape


def _call(name: str, *args) -> None:
    _lib.check(getattr(_lib.lib(), name)(*args), name)


def affine2_fwd(x, out, tables, tab_stride):
    n, h, w, c = _nhwc(x)
    _nhwc(out)
    ga, ba, gb, bb = tables
    _call("dfgan_affine2_fwd", _p(x), _p(out), _p(ga), _p(ba), _p(gb), _p(bb), tab_stride, n, h * w, c, _s())


def affine2_bwd(g_a, x, gx, tables, dtables, tab_stride, *, g_add=None, gscale=None, dot=None, colsum=None):
    n, h, w, c = _nhwc(x)
    assert g_a.shape == x.shape and gx.shape == x.shape
    _nhwc(g_a), _nhwc(gx)
    if g_add is not None:
        assert g_add.shape == x.shape
        _nhwc(g_add)
    ga, ba, gb, bb = tables
    dga, dba, dgb, dbb = dtables
    _call("dfgan_affine2_bwd", _p(g_a), _p(x), _p(g_add), _p(gx), _p(ga), _p(ba), _p(gb), _p(bb), _p(dga),
          _p(dba), _p(dgb), _p(dbb), tab_stride, _p(gscale), _p(dot), _p(colsum), n, h * w, c, _s())


def pool(x, out=None, *, shift=1, scale=1.0, colsum=None, cs_dev=None, cs_imm=1.0):
    n, h, w, c = _nhwc(x)
    ho, wo = h >> shift, w >> shift
    if out is not None:
        assert tuple(_nhwc(out)) == (n, ho, wo, c)
    _call("dfgan_pool", _p(x), _p(out), _p(colsum), _p(cs_dev), cs_imm, n, ho, wo, c, shift, scale, _s())


def dres_bwd(g_y, r, g_pre, gamma, dgamma=None, colsum=None):
    n, h, w, c = _nhwc(g_y)
    assert r.shape == g_y.shape and g_pre.shape == g_y.shape
    _nhwc(r), _nhwc(g_pre)
    _call("dfgan_dres_bwd", _p(g_y), _p(r), _p(g_pre), _p(gamma), _p(dgamma), _p(colsum), n, h * w, c, _s())


def dot(a, b, out, scale=1.0):
    assert a.dtype == BF16 and b.dtype == BF16 and a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    _call("dfgan_dot", _p(a), _p(b), a.numel(), _p(out), scale, _s())


def im2col3(img, out, rowscale=None):
    assert img.is_cuda and img.dtype == F32 and img.is_contiguous() and img.shape[1] == 3
    n, _, h, w = img.shape
    assert tuple(_nhwc(out)) == (n, h, w, 32)
    _call("dfgan_im2col3", _p(img), _p(rowscale), _p(out), n, h, w, _s())


def stem_fwd(img, w0_f, out, bias=None, rowscale=None, xcol=None):
    """Fused discriminator stem (discriminator/model.py:13): fp32 NCHW image -> NHWC bf16 [N,H,W,32] in one kernel;
    xcol (optional, NHWC bf16 [N,H,W,32]) additionally receives the im2col tensor the stem's weight gradient reads."""
    assert img.is_cuda and img.dtype == F32 and img.is_contiguous() and img.shape[1] == 3
    n, _, h, w = img.shape
    assert tuple(_nhwc(out)) == (n, h, w, 32) and w0_f.dtype == BF16 and tuple(w0_f.shape) == (32, 32)
    assert w0_f.is_contiguous() and (bias is None or (bias.dtype == F32 and bias.numel() == 32))
    if xcol is not None:
        assert tuple(_nhwc(xcol)) == (n, h, w, 32)
    _call("dfgan_stem_fwd", _p(img), _p(rowscale), _p(w0_f), _p(bias), _p(out), _p(xcol), n, h, w, _s())


def tanh_bwd_pack(g_img, img, out, dbias=None):
    assert g_img.dtype == F32 and img.dtype == F32 and g_img.is_contiguous() and img.is_contiguous()
    n, _, h, w = img.shape
    assert tuple(_nhwc(out)) == (n, h, w, 32)
    _call("dfgan_tanh_bwd_pack", _p(g_img), _p(img), _p(out), _p(dbias), n, h * w, _s())


def nchw_to_nhwc(x, out):
    """fp32 [N,C,H,W] -> bf16 [N,H,W,C]"""
    assert x.dtype == F32 and x.is_contiguous() and x.is_cuda
    n, c, h, w = x.shape
    assert tuple(_nhwc(out)) == (n, h, w, c)
    _call("dfgan_nchw_to_nhwc", _p(x), _p(out), n, c, h * w, _s())


def nhwc_to_nchw(x, out, accumulate=False):
    n, h, w, c = _nhwc(x)
    assert out.dtype == F32 and out.is_contiguous() and tuple(out.shape) == (n, c, h, w)
    _call("dfgan_nhwc_to_nchw", _p(x), _p(out), n, c, h * w, int(accumulate), _s())


def rowsq(x, out):
    assert x.dtype == F32 and x.is_contiguous() and out.dtype == F32
    n = x.shape[0]
    _call("dfgan_rowsq", _p(x), n, x.numel() // n, _p(out), _s())


def head_cat(embed, sent, cat, sscale=None):
    b, ce = embed.shape[0], embed.shape[1]
    assert embed.dtype == F32 and embed.is_contiguous() and tuple(embed.shape[2:]) == (4, 4)
    assert sent.dtype == F32 and sent.is_contiguous() and sent.shape[0] == b
    assert tuple(_nhwc(cat)) == (b, 4, 4, ce + sent.shape[1])
    _call("dfgan_head_cat", _p(embed), _p(sent), _p(sscale), _p(cat), b, ce, sent.shape[1], _s())


def head_cat_bwd(g_cat, g_embed, g_sent):
    b = g_cat.shape[0]
    ce = g_embed.shape[1]
    cs = g_cat.shape[3] - ce
    assert g_embed.dtype == F32 and g_embed.is_contiguous()
    if g_sent is not None:
        assert g_sent.dtype == F32 and g_sent.is_contiguous() and tuple(g_sent.shape) == (b, cs)
    _call("dfgan_head_cat_bwd", _p(g_cat), _p(g_embed), _p(g_sent), b, ce, cs, _s())


def head_out(hid, w2, logits):
    b, _, _, ch = _nhwc(hid)
    assert w2.dtype == F32 and w2.is_contiguous() and w2.numel() == 16 * ch and logits.numel() == b
    _call("dfgan_head_out", _p(hid), _p(w2), _p(logits), b, ch, _s())


def head_out_bwd(g_logit, hid, act, w2, g_hid, dw2):
    ref = hid if hid is not None else act
    b, _, _, ch = _nhwc(ref)
    _call("dfgan_head_out_bwd", _p(g_logit), _p(hid), _p(act), _p(w2), _p(g_hid), _p(dw2), b, ch, _s())


# ------------------------------------------------------------------------------------ batched linears
class LinearTable:
    """Device-resident table of 64-output work items for dfgan_linear_batched.  Built once from the
    pointers of the tensors involved; the tensors must stay alive and in place while the table is used."""

    def __init__(self, problems: Sequence[dict], mode: int, device):
        items = []
        self.k_max, self.n_out_max = 1, 1
        self.starts = []        # first work item of every problem (run(first_problem, n_problems) launches a slice)
        for pr in problems:
            k, n_out = pr["k"], pr["n_out"]
            self.k_max, self.n_out_max = max(self.k_max, k), max(self.n_out_max, n_out)
            self.starts.append(len(items))
            span = k if mode == 1 else n_out
            for o0 in range(0, span, 64):
                items.append(LinearItem(pr["inp"], pr["in_stride"], pr["w"], pr.get("bias"), pr["out"],
                                        pr["out_stride"], pr.get("mask"), k, n_out, o0, int(pr.get("relu", 0))))
        self.n = len(items)
        self.starts.append(self.n)
        arr = (LinearItem * self.n)(*items)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.dev = raw.to(device)
        self.mode = mode

    def run(self, batch: int, first_problem: int = 0, n_problems: Optional[int] = None) -> None:
        """All work items, or those of problems [first_problem, first_problem + n_problems)."""
        i0 = self.starts[first_problem]
        i1 = self.n if n_problems is None else self.starts[first_problem + n_problems]
        if i1 <= i0:
            return
        _call("dfgan_linear_batched", self.dev.data_ptr() + i0 * C.sizeof(LinearItem), i1 - i0, batch, self.k_max,
              self.n_out_max, self.mode, _s())


# ------------------------------------------------------------------------------------ packing / optimiser
def pack_weights(w, out, *, transpose=False, rows_pad=0, cols_pad=0, slots=None, scale=1.0, as_matrix=False,
                 image=False):
    """fp32 OIHW -> bf16 [n_slots*rows_pad, cols_pad].  slots: list of lists of source tap indices (default:
    one slot per tap).  as_matrix: treat [Co, Ci, kh, kw] as [Co, Ci*kh*kw, 1, 1] (stem im2col weights)."""
    assert w.dtype == F32 and w.is_contiguous() and w.is_cuda and out.dtype == BF16 and out.is_contiguous()
    co, ci, kh, kw = w.shape
    if as_matrix:
        ci, kh, kw = ci * kh * kw, 1, 1
    kk = kh * kw
    if slots is None:
        slots = [[t] for t in range(kk)]
    rows = ci if transpose else co
    cols = co if transpose else ci
    rows_pad, cols_pad = rows_pad or rows, cols_pad or cols
    assert out.numel() == len(slots) * rows_pad * cols_pad, (out.shape, len(slots), rows_pad, cols_pad)
    from . import convops
    convops.IMAGE_OF.pop(out.data_ptr(), None)   # a re-used address must not inherit a stale image
    flat = (C.c_int8 * (len(slots) * 4))(*[(s[j] if j < len(s) else -1) for s in slots for j in range(4)])
    _call("dfgan_pack_weights", _p(w), _p(out), co, ci, kk, int(transpose), rows_pad, cols_pad, len(slots),
          C.cast(flat, C.c_void_p), scale, int(image), _s())


class PackPlan:
    """All weight re-packs of a network as ONE launch: `add` mirrors pack_weights and records a table entry,
    `run` replays the table (built on first use; tensors must stay in place)."""

    def __init__(self, device):
        self.device, self.entries, self.blocks, self.dev = device, [], 0, None
        self.images, self.outs, self.nblocks = [], [], []
        self.bs_off = _lib.lib().dfgan_pack_entry_bs_offset()
        self.esize = _lib.lib().dfgan_pack_entry_size()

    def add(self, w, out, *, transpose=False, rows_pad=0, cols_pad=0, slots=None, scale=1.0, as_matrix=False,
            image=False):
        assert w.dtype == F32 and w.is_contiguous() and out.dtype == BF16 and out.is_contiguous()
        from . import convops
        convops.PRUNED.discard(out.data_ptr())   # a re-used address starts a new life
        convops.IMAGE_OF.pop(out.data_ptr(), None)   # a re-used address must not inherit a stale image
        co, ci, kh, kw = w.shape
        if as_matrix:
            ci, kh, kw = ci * kh * kw, 1, 1
        if slots is None:
            slots = [[t] for t in range(kh * kw)]
        rows, cols = (ci, co) if transpose else (co, ci)
        rows_pad, cols_pad = rows_pad or rows, cols_pad or cols
        assert out.numel() == len(slots) * rows_pad * cols_pad
        flat = (C.c_int8 * (len(slots) * 4))(*[(s[j] if j < len(s) else -1) for s in slots for j in range(4)])
        buf = (C.c_uint8 * self.esize)()
        if image:
            assert cols_pad % 64 == 0
        nb = _lib.lib().dfgan_pack_entry_fill(C.cast(buf, C.c_void_p), _p(w), _p(out), co, ci, kh * kw,
                                              int(transpose), rows_pad, cols_pad, len(slots),
                                              C.cast(flat, C.c_void_p), scale, int(image), self.blocks)
        if nb < 0:
            _lib.check(nb, "dfgan_pack_entry_fill")
        self.entries.append(bytes(buf))
        self.outs.append(out.data_ptr())
        self.nblocks.append(nb)
        self.blocks += nb

    def prune(self, used: set) -> int:
        """Drop the entries whose output operand is not in `used` (data_ptrs); returns how many were dropped.
        block_start of the kept entries is rewritten in place (last 4 bytes of an entry)."""
        import struct
        from . import convops
        keep = [i for i, o in enumerate(self.outs) if o in used]
        dropped = len(self.outs) - len(keep)
        if dropped == 0 or not keep:
            return 0
        # a later launch that would consume a dropped (no longer refreshed) copy fails loudly in convops.conv_run
        convops.PRUNED.update(o for o in self.outs if o not in used)
        entries, start = [], 0
        for i in keep:
            e = bytearray(self.entries[i])
            e[self.bs_off:self.bs_off + 4] = struct.pack("<i", start)
            entries.append(bytes(e))
            start += self.nblocks[i]
        self.entries = entries
        self.outs = [self.outs[i] for i in keep]
        self.nblocks = [self.nblocks[i] for i in keep]
        self.blocks, self.dev = start, None
        return dropped

    def add_with_image(self, w, out, *, group=1, **kw):
        """`add` plus, for wide operands (K = columns a multiple of 64 and >= 128), a second copy in the swizzled
        shared-memory image layout registered for the wide-layer halo kernel (convops.IMAGE_OF)."""
        from . import convops
        self.add(w, out, **kw)
        cols = out.shape[1]
        if cols % 64 == 0 and cols >= 128:
            img = torch.empty_like(out)
            self.add(w, img, image=group, **kw)
            n_slots = len(kw["slots"]) if kw.get("slots") is not None else w.shape[2] * w.shape[3]
            convops.register_image(out, img, (out.shape[0] // n_slots) * group)
            self.images.append(img)

    def run(self):
        if self.dev is None:
            raw = torch.frombuffer(bytearray(b"".join(self.entries)), dtype=torch.uint8)
            self.dev = raw.to(self.device)
        _call("dfgan_pack_batched", self.dev.data_ptr(), len(self.entries), self.blocks, _s())


def unfold_subpixel_wgrad(folded, dw):
    co, ci = dw.shape[0], dw.shape[1]
    assert folded.dtype == F32 and folded.numel() == 16 * co * ci and dw.dtype == F32 and dw.is_contiguous()
    _call("dfgan_unfold_subpixel_wgrad", _p(folded), _p(dw), co, ci, _s())


class WgradScratch:
    """Tap-major fp32 scratch [KK][Co][Ci] for every k x k convolution weight of a network, the target of the
    weight-gradient GEMMs (their fp32 atomics are coalesced only when ci is the fastest index), and the batched
    fold-back into the reference's OIHW gradients: `flush(grads)` adds every scratch tensor into grads[name] and
    zeroes the scratch again (one launch)."""

    def __init__(self, shapes: dict, device):
        self.shapes, self.off, total = dict(shapes), {}, 0
        for name, (co, ci, kh, kw) in self.shapes.items():
            self.off[name] = total
            total += ((co * ci * kh * kw + 7) // 8) * 8
        self.buf = torch.zeros(max(total, 8), device=device, dtype=F32)
        self.device = device
        self._tables: dict = {}
        self.esize = _lib.lib().dfgan_unpack_entry_size()

    def __contains__(self, name: str) -> bool:
        return name in self.shapes

    def view(self, name: str):
        co, ci, kh, kw = self.shapes[name]
        n = co * ci * kh * kw
        return self.buf[self.off[name]:self.off[name] + n].view(kh * kw, co, ci)

    def strides(self, name: str) -> tuple:
        co, ci, _, _ = self.shapes[name]
        return (ci, 1, co * ci)          # (s_co, s_ci, s_tap) of the [tap][co][ci] scratch

    def flush(self, grads: dict, names=None) -> None:
        names = tuple(n for n in (names if names is not None else self.shapes) if n in self.shapes)
        if not names:
            return
        key = (names, tuple(grads[n].data_ptr() for n in names))
        tab = self._tables.get(key)
        if tab is None:
            entries, blocks = [], 0
            for n in names:
                co, ci, kh, kw = self.shapes[n]
                g = grads[n]
                assert g.dtype == F32 and g.is_contiguous() and tuple(g.shape) == (co, ci, kh, kw)
                buf = (C.c_uint8 * self.esize)()
                nb = _lib.lib().dfgan_unpack_entry_fill(C.cast(buf, C.c_void_p), self.view(n).data_ptr(),
                                                        g.data_ptr(), co, ci, kh * kw, blocks)
                if nb < 0:
                    _lib.check(nb, "dfgan_unpack_entry_fill")
                entries.append(bytes(buf))
                blocks += nb
            raw = torch.frombuffer(bytearray(b"".join(entries)), dtype=torch.uint8).to(self.device)
            tab = (raw, len(entries), blocks)
            if len(self._tables) > 8:
                self._tables.clear()
            self._tables[key] = tab
        _call("dfgan_wgrad_unpack", tab[0].data_ptr(), tab[1], tab[2], _s())


def gblock_small_grads(cs, gamma, bias2, g_bias2, g_gamma, g_sbias=None):
    """conv_2.bias / scale_conv.bias / gate gradients of a generator block from the per-channel cotangent sums."""
    for t in (cs, gamma, bias2, g_bias2, g_gamma):
        assert t.dtype == F32 and t.is_contiguous()
    assert g_sbias is None or (g_sbias.dtype == F32 and g_sbias.is_contiguous())
    _call("dfgan_gblock_small_grads", _p(cs), _p(gamma), _p(bias2), _p(g_bias2), _p(g_gamma), _p(g_sbias),
          cs.numel(), _s())


def d_hinge(logits, loss, g_logits, batch):
    """logits fp32 [3B-1] = [real; mismatched; fake] -> loss (1 element, overwritten) and d loss / d logits."""
    assert logits.dtype == F32 and logits.is_contiguous() and logits.numel() == 3 * batch - 1
    assert g_logits.dtype == F32 and g_logits.numel() == logits.numel() and loss.dtype == F32
    _call("dfgan_d_hinge", _p(logits), _p(loss), _p(g_logits), batch, _s())


def gp_coef(sq, loss, coef):
    """sq fp32 [B] -> MA-GP loss (1 element, overwritten) and the per-sample coefficient of its gradient."""
    assert sq.dtype == F32 and coef.dtype == F32 and sq.numel() == coef.numel() and loss.dtype == F32
    _call("dfgan_gp_coef", _p(sq), _p(loss), _p(coef), sq.numel(), _s())


def txtimg_loss(img, sent, loss, partial=None):
    """img fp32 NCHW [B,C,H,W], sent fp32 [B,E] -> loss (1 element, overwritten) = 1 - mean cosine(padded channel
    means, sentence embedding): the logged text-image loss of damsm/loss.py:4-26."""
    b, c = img.shape[0], img.shape[1]
    assert img.dtype == F32 and img.is_contiguous() and sent.dtype == F32 and sent.is_contiguous()
    assert sent.shape[0] == b and loss.dtype == F32
    if partial is None:
        partial = torch.empty(4 * b * c, device=img.device, dtype=F32)
    _call("dfgan_txtimg_loss", _p(img), b, c, img[0, 0].numel(), _p(sent), sent.shape[1], _p(partial),
          partial.numel(), _p(loss), _s())


def sqnorm_det(g, out, scratch):
    """out += sum(g^2) with a scheduling-independent (bit-reproducible) result; scratch: zero-initialised fp32 buffer
    of at least 4 * #SMs + 1 elements owned by the caller."""
    assert g.dtype == F32 and g.is_contiguous() and out.dtype == F32 and scratch.dtype == F32
    _call("dfgan_sqnorm_det", _p(g), g.numel(), _p(out), _p(scratch), scratch.numel(), _s())


def sqnorm(g, out):
    assert g.dtype == F32 and g.is_contiguous() and out.dtype == F32
    _call("dfgan_sqnorm", _p(g), g.numel(), _p(out), _s())


def adam_step(p, g, m, v, sumsq, step, *, lr, gscale=1.0, max_norm=5.0, b1=0.0, b2=0.9, eps=1e-8):
    for t in (p, g, m, v):
        assert t.dtype == F32 and t.is_contiguous() and t.numel() == p.numel()
    assert step.dtype == torch.int32
    _call("dfgan_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(sumsq), _p(step), gscale, max_norm, lr,
          b1, b2, eps, _s())
