"""Host-side builders for the tap-list GEMM descriptors (dfgan_conv_run / dfgan_wgrad_run).

Activations are NHWC bf16 torch tensors ``[N, H, W, C]``; packed weights are bf16 matrices
``[taps * rows, K]`` (see :mod:`pack`).  Every function here only fills a C struct and makes one C-ABI
call on the current CUDA stream; the arithmetic lives in csrc/conv_umma.cu.

Reference call sites covered (Deep_Fusion_GAN/src/): generator/residual_block.py:16,21,25 (3x3 and 1x1
convs), discriminator/residual_block.py:11,13,19 (4x4 stride-2, 3x3, 1x1), discriminator/model.py:31.
"""
from __future__ import annotations

import ctypes as C
import os as _os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, Tap, View4, WgradDesc

SLOPE = 0.2
SPLITK_MAX_PIXELS = 128 * 74   # more 128-pixel tiles than half the SMs never split (see plan_generic)
_SPLITK_WS: dict = {}           # (stream, device) -> zeroed fp32 scratch of the split-K launches on that stream

# packed operand (data_ptr) -> (the same operand in the swizzled shared-memory image layout, rows per image slot);
# registered by the packers for C_in % 64 == 0, C_in >= 128 and picked up by conv_run for the wide-layer halo kernel
IMAGE_OF: dict = {}
# data_ptrs of the operand copies conv_run actually consumed while TRACK_USAGE is on (DFGANTrainer turns it on for
# its first eager iteration and then drops the never-used copies from the per-step re-pack)
USED: set = set()
TRACK_USAGE = False
# data_ptrs of operand copies a PackPlan.prune() stopped refreshing: consuming one of them is an error (stale weights)
PRUNED: set = set()


def check_not_pruned(t: torch.Tensor, what: str) -> None:
    if PRUNED and t.data_ptr() in PRUNED:
        raise RuntimeError(f"{what}: this packed weight copy is no longer re-packed after optimiser steps (it was not "
                           "consumed by the trainer's first iteration and PackPlan.prune dropped it); a launch that "
                           "reads it would see stale weights")


# ------------------------------------------------------------------------------------ side streams
# Launches that are independent of the chain on the current stream (weight gradients next to the data-gradient
# chain, the four sub-pixel phases of one convolution) are issued on side streams: forked from / joined into the
# current stream with events, so under stream capture they become parallel branches of the CUDA graph and fill the
# SMs a small-grid kernel leaves idle.  SIDE_ENABLED = False serialises everything on the current stream (used by
# the per-kernel timing pass of bench.py and by DFGAN_NO_SIDE=1).
SIDE_ENABLED = _os.environ.get("DFGAN_NO_SIDE", "0") != "1"
N_SIDE = 3            # streams per pool
POOL_PHASE, POOL_WGRAD, POOL_AUX = 0, 1, 2   # AUX: operand re-packs / gradient fold-backs next to the main chain
N_POOLS = 3
_SIDE_STREAMS: dict = {}
_SIDE_DIRTY: dict = {}


def _dev_key(dev) -> int:
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _side_streams(dev) -> list:
    key = _dev_key(dev)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [[torch.cuda.Stream(device=dev) for _ in range(N_SIDE)] for _ in range(N_POOLS)]
        _SIDE_DIRTY[key] = [[False] * N_SIDE for _ in range(N_POOLS)]
    return _SIDE_STREAMS[key]


class side:
    """``with side(i, dev, pool):`` -- the body's launches go to side stream i of the pool, ordered after everything
    issued so far on the current stream.  The body must not allocate; every tensor it touches must stay alive
    until ``join`` of that pool."""

    def __init__(self, idx: int, dev, pool: int = POOL_WGRAD, after_pool=None):
        """after_pool: the body is ordered after everything issued so far on the side streams of THAT pool instead
        of after the current stream (a consumer of weight-gradient launches that must not stall the main chain)."""
        self.idx, self.dev, self.pool, self.ctx, self.after_pool = idx % N_SIDE, dev, pool, None, after_pool

    def __enter__(self):
        if not SIDE_ENABLED:
            return self
        s = _side_streams(self.dev)[self.pool][self.idx]
        if self.after_pool is None:
            s.wait_stream(torch.cuda.current_stream(self.dev))
        else:
            key = _dev_key(self.dev)
            waited = False
            for i, o in enumerate(_SIDE_STREAMS[key][self.after_pool]):
                if _SIDE_DIRTY[key][self.after_pool][i]:
                    s.wait_stream(o)
                    waited = True
            if not waited:
                s.wait_stream(torch.cuda.current_stream(self.dev))
        _SIDE_DIRTY[_dev_key(self.dev)][self.pool][self.idx] = True
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join(dev, pool: int = POOL_WGRAD) -> None:
    """The current stream waits for every side stream of the pool that received work since its last join."""
    if not SIDE_ENABLED:
        return
    key = _dev_key(dev)
    if key not in _SIDE_STREAMS:
        return
    cur = torch.cuda.current_stream(dev)
    for i, s in enumerate(_SIDE_STREAMS[key][pool]):
        if _SIDE_DIRTY[key][pool][i]:
            cur.wait_stream(s)
            _SIDE_DIRTY[key][pool][i] = False


def register_image(wpack: torch.Tensor, wimg: torch.Tensor, rows_per_slot: int) -> None:
    IMAGE_OF[wpack.data_ptr()] = (wimg, rows_per_slot)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def view_nhwc(t: torch.Tensor) -> View4:
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.stride(3) == 1, (t.dtype, t.shape, t.stride())
    n, h, w, c = t.shape
    return View4(t.data_ptr(), c, w, h, n, t.stride(2), t.stride(1), t.stride(0))


def parity_views(t: torch.Tensor) -> list:
    """Four stride-2 sub-lattices of an NHWC tensor, index = (row parity)*2 + (col parity)."""
    return [t[:, ph::2, pw::2, :] for ph in (0, 1) for pw in (0, 1)]


# ------------------------------------------------------------------------------------ tap tables
def taps_3x3(rows: int) -> list:
    return [(0, r - 1, s - 1, (r * 3 + s) * rows) for r in range(3) for s in range(3)]


def taps_3x3_dgrad(rows: int) -> list:
    return [(0, 1 - r, 1 - s, (r * 3 + s) * rows) for r in range(3) for s in range(3)]


def taps_1x1() -> list:
    return [(0, 0, 0, 0)]


def _s2_split(r: int):
    """kernel row r of a 4x4/stride-2/pad-1 conv -> (parity, offset) of the input row 2*o + r - 1."""
    v = r - 1
    return v & 1, (v - (v & 1)) // 2


def taps_4x4s2(rows: int) -> list:
    out = []
    for r in range(4):
        ph, dy = _s2_split(r)
        for s in range(4):
            pw, dx = _s2_split(s)
            out.append((ph * 2 + pw, dy, dx, (r * 4 + s) * rows))
    return out


def taps_4x4s2_dgrad(rows: int, a: int, c: int) -> list:
    """Taps of output phase (a, c) of the stride-2 data gradient: in-pixel (2i+a, 2j+c) receives
    g[i + dy, j + dx] * W[:, :, r, s] for the two r with r = a+1 (mod 2) (same for s)."""
    def sel(par):
        return [(1, 0), (3, -1)] if par == 0 else [(0, 1), (2, 0)]
    return [(0, dy, dx, (r * 4 + s) * rows) for r, dy in sel(a) for s, dx in sel(c)]


# ------------------------------------------------------------------------------------ conv
def conv_run(srcs: Sequence[torch.Tensor], taps: Sequence[tuple], wpack: torch.Tensor, cout: int,
             out: Optional[torch.Tensor] = None, *, grid: Optional[tuple] = None,
             bias: Optional[torch.Tensor] = None, scale_dev: Optional[torch.Tensor] = None,
             scale_imm: float = 1.0, lrelu_pre: bool = False, aux: Optional[torch.Tensor] = None,
             addend: Optional[torch.Tensor] = None, add_shift: int = 0, add_scale: float = 1.0,
             mask_src: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None, post_mode: int = 0,
             tables: Optional[tuple] = None, tab_stride: int = 0, out_nchw: Optional[torch.Tensor] = None,
             co_valid: int = 0, tanh_out: bool = False, n_tile: int = 0, max_ctas: int = 0,
             phase_cout: int = 0) -> None:
    """epilogue(sum_taps src[tap] (*) wpack[tap rows]) -- see dfgan_conv_run in include/dfgan_b200.h.

    ``out`` / ``aux`` / ``out2`` / ``mask_src`` are NHWC bf16 tensors (possibly strided phase views) that
    must share one set of strides.  ``grid`` = (N, H, W) of the output grid (default: shape of the first
    given output).  ``tables`` = (ga, ba, gb, bb) fp32 views [N, C] with row stride ``tab_stride``.
    """
    d = ConvDesc()
    assert 1 <= len(srcs) <= _lib.MAX_SRC and 1 <= len(taps) <= _lib.MAX_TAPS
    for i, s in enumerate(srcs):
        d.src[i] = view_nhwc(s)
    d.n_src, d.n_taps = len(srcs), len(taps)
    for i, (si, dy, dx, wrow) in enumerate(taps):
        d.taps[i] = Tap(si, dy, dx, wrow)
    assert wpack.dtype == torch.bfloat16 and wpack.is_contiguous() and wpack.dim() == 2
    d.wpack, d.w_rows, d.Cin, d.Cout = wpack.data_ptr(), wpack.shape[0], wpack.shape[1], cout
    ref = next((t for t in (out, aux, out2) if t is not None), None)
    if ref is not None:
        assert ref.dtype == torch.bfloat16 and ref.stride(3) == 1 and ref.shape[3] == (phase_cout or cout)
        d.out_sn, d.out_sh, d.out_sw = ref.stride(0), ref.stride(1), ref.stride(2)
        for t in (out, aux, out2, mask_src):
            if t is not None:
                assert t.dtype == torch.bfloat16 and t.stride() == ref.stride(), (t.stride(), ref.stride())
    else:
        assert out_nchw is not None and mask_src is None
    n, h, w = grid if grid is not None else (ref.shape[:3] if ref is not None else
                                              (out_nchw.shape[0], out_nchw.shape[2], out_nchw.shape[3]))
    d.N, d.H, d.W = n, h, w
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == (phase_cout or cout)
    d.out, d.aux, d.out2, d.mask_src = _ptr(out), _ptr(aux), _ptr(out2), _ptr(mask_src)
    d.bias, d.scale_dev, d.scale_imm = _ptr(bias), _ptr(scale_dev), scale_imm
    d.lrelu_pre, d.slope = int(lrelu_pre), SLOPE
    if addend is not None:
        assert addend.dtype == torch.bfloat16 and addend.stride(3) == 1 and addend.shape[3] == (phase_cout or cout)
        d.addend = addend.data_ptr()
        d.add_sn, d.add_sh, d.add_sw = addend.stride(0), addend.stride(1), addend.stride(2)
    d.add_shift, d.add_scale = add_shift, add_scale
    d.post_mode = post_mode
    if post_mode == 2:
        assert tables is not None and all(t.dtype == torch.float32 and t.stride(-1) == 1 for t in tables)
        d.tab_ga, d.tab_ba, d.tab_gb, d.tab_bb = (t.data_ptr() for t in tables)
        d.tab_stride = tab_stride
    if out_nchw is not None:
        assert out_nchw.dtype == torch.float32 and out_nchw.is_contiguous()
        assert tuple(out_nchw.shape) == (n, co_valid, h, w), (out_nchw.shape, (n, co_valid, h, w))
        d.out_nchw, d.co_valid, d.tanh_out = out_nchw.data_ptr(), co_valid, int(tanh_out)
    d.n_tile_hint, d.max_ctas = n_tile, max_ctas
    if phase_cout:
        assert grid is not None and cout == 4 * phase_cout
    d.phase_cout = phase_cout
    img = IMAGE_OF.get(wpack.data_ptr())
    if img is not None:
        d.wpack_img, d.img_rows_per_slot = img[0].data_ptr(), img[1]
    if TRACK_USAGE:
        uses_img = img is not None and _lib.lib().dfgan_conv_uses_image(C.byref(d)) == 1
        USED.add(img[0].data_ptr() if uses_img else wpack.data_ptr())
    elif PRUNED and (wpack.data_ptr() in PRUNED or (img is not None and img[0].data_ptr() in PRUNED)):
        uses_img = img is not None and _lib.lib().dfgan_conv_uses_image(C.byref(d)) == 1
        check_not_pruned(img[0] if uses_img else wpack, "conv_run")
    if n * h * w <= SPLITK_MAX_PIXELS:
        # low-resolution launch: offer a zeroed fp32 scratch so the C side can split the K loop over the idle SMs.
        # One persistent scratch per stream (launches on a stream are serialised; the second pass of a split launch
        # leaves the scratch zeroed again), allocated outside any stream capture by the eager warm-up iterations.
        nbytes = _lib.lib().dfgan_conv_splitk_bytes(C.byref(d))
        if nbytes > 0:
            key = (_stream(), wpack.device.index)
            ws = _SPLITK_WS.get(key)
            if ws is None or ws.numel() * 4 < nbytes:
                ws = torch.zeros(max(nbytes // 4, 1 << 20), device=wpack.device, dtype=torch.float32)
                _SPLITK_WS[key] = ws
            d.workspace, d.ws_bytes = ws.data_ptr(), nbytes
    _lib.check(_lib.lib().dfgan_conv_run(C.byref(d), _stream()), "dfgan_conv_run")


def wgrad_run(srcs: Sequence[torch.Tensor], taps: Sequence[tuple], g: torch.Tensor, dw: torch.Tensor,
              strides: tuple, *, scale_dev: Optional[torch.Tensor] = None, scale_imm: float = 1.0,
              k_splits: int = 0, ci_valid: int = 0, co_valid: int = 0) -> None:
    """dw[co, ci, tap] += scale * sum_pixels g[p, co] * src[tap][p + shift, ci]  (fp32 atomics).

    ``taps`` entries are (src, dy, dx, tap_id); ``strides`` = element strides (s_co, s_ci, s_tap) of dw.
    """
    d = WgradDesc()
    for i, s in enumerate(srcs):
        d.src[i] = view_nhwc(s)
    d.n_src, d.n_taps = len(srcs), len(taps)
    for i, (si, dy, dx, tid) in enumerate(taps):
        d.taps[i] = Tap(si, dy, dx, tid)
    d.g = view_nhwc(g)
    d.Cin, d.Cout = srcs[0].shape[3], g.shape[3]
    assert dw.dtype == torch.float32
    d.dw = dw.data_ptr()
    d.s_co, d.s_ci, d.s_tap = strides
    d.scale_dev, d.scale_imm, d.k_splits = _ptr(scale_dev), scale_imm, k_splits
    d.ci_valid, d.co_valid = ci_valid, co_valid
    _lib.check(_lib.lib().dfgan_wgrad_run(C.byref(d), _stream()), "dfgan_wgrad_run")


# ------------------------------------------------------------------------------------ sub-pixel taps
# nearest-2x upsample followed by conv3x3/pad1 == four phase-specific 2x2 convolutions on the low-res
# input.  Output row 2i+a reads low-res rows i+dy with the 3x3 rows folded as below.
SUBPIX_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}   # phase -> ((dy, rows summed), ..)


def subpixel_slots() -> list:
    """16 slots (phase a, phase c, dy index, dx index) -> list of source taps r*3+s summed into the slot."""
    slots = []
    for a in (0, 1):
        for c in (0, 1):
            for (_, rs) in SUBPIX_ROWS[a]:
                for (_, ss) in SUBPIX_ROWS[c]:
                    slots.append([r * 3 + s for r in rs for s in ss])
    return slots


def taps_subpixel_fwd(rows: int, a: int, c: int) -> list:
    base = (a * 2 + c) * 4
    return [(0, dy, dx, (base + yi * 2 + xi) * rows)
            for yi, (dy, _) in enumerate(SUBPIX_ROWS[a]) for xi, (dx, _) in enumerate(SUBPIX_ROWS[c])]


def taps_subpixel_dgrad(rows: int) -> list:
    """Data gradient wrt the low-res input: sources are the four phase views (index a*2+c) of the hi-res
    gradient, pixel (i,j) receives g_phase[i-dy, j-dx] * W_slot^T."""
    out = []
    for a in (0, 1):
        for c in (0, 1):
            base = (a * 2 + c) * 4
            for yi, (dy, _) in enumerate(SUBPIX_ROWS[a]):
                for xi, (dx, _) in enumerate(SUBPIX_ROWS[c]):
                    out.append((a * 2 + c, -dy, -dx, (base + yi * 2 + xi) * rows))
    return out


def taps_subpixel_wgrad(a: int, c: int) -> list:
    base = (a * 2 + c) * 4
    return [(0, dy, dx, base + yi * 2 + xi)
            for yi, (dy, _) in enumerate(SUBPIX_ROWS[a]) for xi, (dx, _) in enumerate(SUBPIX_ROWS[c])]


# ------------------------------------------------------------------------------------ merged phases
# One launch for the four phases of a 2x-upsampling layer: 9 input shifts (dy,dx in -1..1), weight rows
# [shift][phase a*2+c][channel]; a (shift, phase) block is zero when the phase does not read that shift.
def taps_merged(rows_per_phase: int) -> list:
    return [(0, dy, dx, ((dy + 1) * 3 + (dx + 1)) * 4 * rows_per_phase) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def subpixel_merged_slots() -> list:
    """36 slots (shift-major, then phase) for upsample -> conv3x3 (App. C-4)."""
    slots = []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            for a in (0, 1):
                for c in (0, 1):
                    rs = dict(SUBPIX_ROWS[a]).get(dy, ())
                    ss = dict(SUBPIX_ROWS[c]).get(dx, ())
                    slots.append([r * 3 + s for r in rs for s in ss])
    return slots


def s2dgrad_merged_slots() -> list:
    """36 slots for the data gradient of a 4x4 / stride-2 / pad-1 conv: phase (a,c) of the input grid reads
    the output gradient at shift (dy,dx) through kernel tap (r,s)."""
    sel = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}
    slots = []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            for a in (0, 1):
                for c in (0, 1):
                    r, s_ = sel[a].get(dy), sel[c].get(dx)
                    slots.append([] if r is None or s_ is None else [r * 4 + s_])
    return slots
