"""Host-side builders for the tap-list GEMM descriptors (dfgan_conv_run / dfgan_wgrad_run).

Activations are NHWC bf16 torch tensors ``[N, H, W, C]``; packed weights are bf16 matrices
``[taps * rows, K]`` (see :mod:`pack`).  Every function here only fills a C struct and makes one C-ABI
call on the current CUDA stream; the arithmetic lives in csrc/conv_umma.cu.

Reference call sites covered (Deep_Fusion_GAN/src/): generator/residual_block.py:16,21,25 (3x3 and 1x1
convs), discriminator/residual_block.py:11,13,19 (4x4 stride-2, 3x3, 1x1), discriminator/model.py:31.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, Tap, View4, WgradDesc

SLOPE = 0.2


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
             out: torch.Tensor, *, grid: Optional[tuple] = None, bias: Optional[torch.Tensor] = None,
             scale_dev: Optional[torch.Tensor] = None, scale_imm: float = 1.0, lrelu_pre: bool = False,
             aux: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None,
             add_shift: int = 0, add_scale: float = 1.0, mask_src: Optional[torch.Tensor] = None,
             n_tile: int = 0, max_ctas: int = 0) -> None:
    """out = epilogue(sum_taps src[tap] (*) wpack[tap rows]) -- see dfgan_conv_run in the header.

    ``out`` may be a strided NHWC view (phase-interleaved writes); ``aux`` / ``mask_src`` must have
    the same strides as ``out``.  ``grid`` = (N, H, W) of the output grid (default: out.shape[:3]).
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
    n, h, w = grid if grid is not None else out.shape[:3]
    d.N, d.H, d.W = n, h, w
    assert out.dtype == torch.bfloat16 and out.stride(3) == 1 and out.shape[3] == cout
    d.out, d.out_sn, d.out_sh, d.out_sw = out.data_ptr(), out.stride(0), out.stride(1), out.stride(2)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout
    d.bias, d.scale_dev, d.scale_imm = _ptr(bias), _ptr(scale_dev), scale_imm
    d.lrelu_pre, d.slope = int(lrelu_pre), SLOPE
    for t in (aux, mask_src):
        if t is not None:
            assert t.dtype == torch.bfloat16 and t.stride() == out.stride(), (t.stride(), out.stride())
    d.aux, d.mask_src = _ptr(aux), _ptr(mask_src)
    if addend is not None:
        assert addend.dtype == torch.bfloat16 and addend.stride(3) == 1 and addend.shape[3] == cout
        d.addend = addend.data_ptr()
        d.add_sn, d.add_sh, d.add_sw = addend.stride(0), addend.stride(1), addend.stride(2)
    d.add_shift, d.add_scale = add_shift, add_scale
    d.n_tile_hint, d.max_ctas = n_tile, max_ctas
    _lib.check(_lib.lib().dfgan_conv_run(C.byref(d), _stream()), "dfgan_conv_run")


def wgrad_run(srcs: Sequence[torch.Tensor], taps: Sequence[tuple], g: torch.Tensor, dw: torch.Tensor,
              strides: tuple, *, scale_dev: Optional[torch.Tensor] = None, scale_imm: float = 1.0,
              k_splits: int = 0) -> None:
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
    _lib.check(_lib.lib().dfgan_wgrad_run(C.byref(d), _stream()), "dfgan_wgrad_run")


# ------------------------------------------------------------------------------------ packing (torch)
def pack_fwd(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> bf16 [kh*kw*Cout, Cin] (row = tap*Cout + co): B operand of the forward GEMM."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw * co, ci).to(torch.bfloat16).contiguous()


def pack_dgrad(w: torch.Tensor) -> torch.Tensor:
    """OIHW fp32 -> bf16 [kh*kw*Cin, Cout] (row = tap*Cin + ci): B operand of the data-gradient GEMM."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 1, 0).reshape(kh * kw * ci, co).to(torch.bfloat16).contiguous()
