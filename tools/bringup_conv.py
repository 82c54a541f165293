"""GPU bring-up of the tcgen05 GEMM engines against torch (cuDNN fp32, TF32 off) on bf16-rounded data.
Usage: python tools/bringup_conv.py <case>|all   (each case is run in its own process by `all`)."""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    return torch


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _pack(w, transpose=False, **kw):
    import torch
    from gan_codes_b200 import kernels as K
    co_, ci, kh, kw_ = w.shape
    rows = ci if transpose else co_
    cols = co_ if transpose else ci
    slots = kw.get("slots")
    n_slots = len(slots) if slots else kh * kw_
    out = torch.empty(n_slots * kw.get("rows_pad", rows), kw.get("cols_pad", cols), device=w.device, dtype=torch.bfloat16)
    K.pack_weights(w.float().contiguous(), out, transpose=transpose, **kw)
    return out


def _run_fprop3(N, C, K, H, W, **kw):
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    x = torch.randn(N, C, H, W, device="cuda").bfloat16()
    w = (torch.randn(K, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
    b = torch.randn(K, device="cuda")
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    out = torch.empty(N, H, W, K, device="cuda", dtype=torch.bfloat16)
    co.conv_run([_nhwc(x)], co.taps_3x3(K), _pack(w), K, out, bias=b, **kw)
    torch.cuda.synchronize()
    err = _rel(out.permute(0, 3, 1, 2), ref)
    print(f"fprop3 N{N} C{C} K{K} {H}x{W} {kw}: rel {err:.3e}")
    return err < 8e-3


@case
def fprop3_a():
    return _run_fprop3(2, 64, 64, 16, 16)


@case
def fprop3_b():
    ok = _run_fprop3(4, 256, 256, 32, 32)
    ok &= _run_fprop3(3, 128, 64, 64, 64)
    ok &= _run_fprop3(5, 256, 256, 4, 4)
    ok &= _run_fprop3(2, 512, 512, 8, 8)
    ok &= _run_fprop3(2, 32, 32, 64, 64)
    ok &= _run_fprop3(2, 64, 32, 128, 128)
    ok &= _run_fprop3(2, 256, 128, 16, 16, n_tile=64)
    return ok


@case
def fprop4s2():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, K, H) in [(2, 32, 64, 64), (2, 64, 128, 32), (3, 512, 512, 8), (2, 256, 512, 32)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(K, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        ref = F.leaky_relu(F.conv2d(x.float(), w.float(), None, stride=2, padding=1), 0.2)
        out = torch.empty(N, H // 2, H // 2, K, device="cuda", dtype=torch.bfloat16)
        co.conv_run(co.parity_views(_nhwc(x)), co.taps_4x4s2(K), _pack(w), K, out, lrelu_pre=True)
        torch.cuda.synchronize()
        err = _rel(out.permute(0, 3, 1, 2), ref)
        print(f"fprop4s2 N{N} C{C} K{K} {H}: rel {err:.3e}")
        ok &= err < 8e-3
    return ok


@case
def dgrad():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, K, H) in [(2, 64, 128, 32), (2, 256, 256, 8), (2, 32, 64, 64)]:
        w = (torch.randn(K, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
        g = torch.randn(N, K, H, H, device="cuda").bfloat16()
        ref = F.conv_transpose2d(g.float(), w.float(), padding=1)
        out = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        co.conv_run([_nhwc(g)], co.taps_3x3_dgrad(C), _pack(w, True), C, out)
        torch.cuda.synchronize()
        err = _rel(out.permute(0, 3, 1, 2), ref)
        print(f"dgrad3 N{N} C{C} K{K} {H}: rel {err:.3e}")
        ok &= err < 8e-3
        # stride-2 4x4: g on the H/2 grid, dx on the H grid, written phase by phase
        w4 = (torch.randn(K, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        g2 = torch.randn(N, K, H // 2, H // 2, device="cuda").bfloat16()
        ref = F.conv_transpose2d(g2.float(), w4.float(), stride=2, padding=1)
        out = torch.zeros(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        wp = _pack(w4, True)
        g2n = _nhwc(g2)
        for a in (0, 1):
            for c in (0, 1):
                co.conv_run([g2n], co.taps_4x4s2_dgrad(C, a, c), wp, C, out[:, a::2, c::2, :])
        torch.cuda.synchronize()
        err = _rel(out.permute(0, 3, 1, 2), ref)
        print(f"dgrad4s2 N{N} C{C} K{K} {H}: rel {err:.3e}")
        ok &= err < 8e-3
    return ok


@case
def epilogue():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    N, C, K, H = 2, 64, 64, 32
    x = torch.randn(N, C, H, H, device="cuda").bfloat16()
    w = (torch.randn(K, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
    b = torch.randn(K, device="cuda")
    gamma = torch.tensor([0.37], device="cuda")
    sc = torch.randn(N, K, H // 2, H // 2, device="cuda").bfloat16()
    msk = torch.randn(N, K, H, H, device="cuda").bfloat16()
    pre = F.leaky_relu(F.conv2d(x.float(), w.float(), b, padding=1), 0.2)
    pre = pre * torch.where(msk.float() > 0, 1.0, 0.2)
    ref = pre * gamma * 2.0 + 0.5 * F.interpolate(sc.float(), scale_factor=2)
    out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty_like(out)
    co.conv_run([_nhwc(x)], co.taps_3x3(K), _pack(w), K, out, bias=b, lrelu_pre=True,
                scale_dev=gamma, scale_imm=2.0, aux=aux, addend=_nhwc(sc), add_shift=1, add_scale=0.5,
                mask_src=_nhwc(msk))
    torch.cuda.synchronize()
    e1, e2 = _rel(out.permute(0, 3, 1, 2), ref), _rel(aux.permute(0, 3, 1, 2), pre)
    print(f"epilogue: out rel {e1:.3e} aux rel {e2:.3e}")
    return e1 < 8e-3 and e2 < 8e-3


@case
def subpixel():
    """upsample -> conv3x3 as four phase convs (fwd), its data gradient and its folded weight gradient."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co, kernels as K
    ok = True
    for (N, C, Kc, H) in [(2, 64, 32, 16), (3, 256, 128, 8), (2, 128, 64, 32), (2, 64, 32, 64)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5))
        b = torch.randn(Kc, device="cuda")
        xr = x.float().requires_grad_()
        wr = w.clone().requires_grad_()
        ref = F.conv2d(F.interpolate(xr, scale_factor=2), wr, b, padding=1)
        g = torch.randn_like(ref).bfloat16()
        gx_ref, gw_ref = torch.autograd.grad(ref, (xr, wr), g.float())
        slots = co.subpixel_slots()
        wf, wd = _pack(w, slots=slots), _pack(w, True, slots=slots)
        out = torch.empty(N, 2 * H, 2 * H, Kc, device="cuda", dtype=torch.bfloat16)
        xn = _nhwc(x)
        for a in (0, 1):
            for c in (0, 1):
                co.conv_run([xn], co.taps_subpixel_fwd(Kc, a, c), wf, Kc, out[:, a::2, c::2, :], grid=(N, H, H), bias=b)
        gn = _nhwc(g)
        gx = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        co.conv_run([gn[:, a::2, c::2, :] for a in (0, 1) for c in (0, 1)], co.taps_subpixel_dgrad(C), wd, C, gx)
        folded = torch.zeros(16, Kc, C, device="cuda")
        for a in (0, 1):
            for c in (0, 1):
                co.wgrad_run([xn], co.taps_subpixel_wgrad(a, c), gn[:, a::2, c::2, :], folded, (C, 1, Kc * C))
        dw = torch.zeros(Kc, C, 3, 3, device="cuda")
        K.unfold_subpixel_wgrad(folded, dw)
        torch.cuda.synchronize()
        e = (_rel(out.permute(0, 3, 1, 2), ref), _rel(gx.permute(0, 3, 1, 2), gx_ref), _rel(dw, gw_ref))
        print(f"subpixel N{N} C{C} K{Kc} {H}: fwd {e[0]:.3e} dgrad {e[1]:.3e} wgrad {e[2]:.3e}")
        ok &= max(e) < 8e-3
    return ok


def _pack_img(w, transpose=False, group=1, **kw):
    """normal pack + the swizzled-image copy, registered for the wide-layer halo kernel"""
    import torch
    from gan_codes_b200 import convops as co, kernels as K
    out = _pack(w, transpose, **kw)
    img = torch.empty_like(out)
    K.pack_weights(w.float().contiguous(), img, transpose=transpose, image=group, **kw)
    slots = kw.get("slots")
    n_slots = len(slots) if slots else w.shape[2] * w.shape[3]
    co.register_image(out, img, out.shape[0] // n_slots * group)
    return out, img


@case
def halo2():
    """Wide layers (C_in >= 128): halo tile + bulk-copied pre-swizzled weight tiles vs torch and the generic kernel."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, Kc, H) in [(2, 128, 128, 64), (2, 256, 256, 32), (3, 512, 256, 16), (2, 128, 512, 32), (1, 256, 64, 64)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
        b = torch.randn(Kc, device="cuda")
        ref = F.conv2d(x.float(), w.float(), b, padding=1)
        wp, keep = _pack_img(w)
        o1 = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        xn = _nhwc(x)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o1, bias=b)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o2, bias=b, n_tile=min(Kc, 256))
        # 4x4 stride 2 over four parity views (2x2 window per view)
        w4 = (torch.randn(Kc, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        r4 = F.leaky_relu(F.conv2d(x.float(), w4.float(), None, stride=2, padding=1), 0.2)
        w4p, keep4 = _pack_img(w4)
        o4 = torch.empty(N, H // 2, H // 2, Kc, device="cuda", dtype=torch.bfloat16)
        if H // 2 >= 16:
            co.conv_run(co.parity_views(xn), co.taps_4x4s2(Kc), w4p, Kc, o4, lrelu_pre=True)
        torch.cuda.synchronize()
        e = [_rel(o1.permute(0, 3, 1, 2), ref), _rel(o1, o2)]
        if H // 2 >= 16:
            e.append(_rel(o4.permute(0, 3, 1, 2), r4))
        print(f"halo2 N{N} C{C} K{Kc} {H}: " + " ".join(f"{v:.3e}" for v in e))
        ok &= e[0] < 8e-3 and e[1] < 1e-3 and (len(e) < 3 or e[2] < 8e-3)
    # merged sub-pixel phases with a wide input (C_in 128 -> 4 x 64)
    N, C, Kc, H = 2, 128, 64, 32
    x = torch.randn(N, C, H, H, device="cuda").bfloat16()
    w = torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5)
    b = torch.randn(Kc, device="cuda")
    ref = F.conv2d(F.interpolate(x.float(), scale_factor=2), w, b, padding=1)
    wm, keepm = _pack_img(w, slots=co.subpixel_merged_slots(), group=4)
    out = torch.empty(N, 2 * H, 2 * H, Kc, device="cuda", dtype=torch.bfloat16)
    co.conv_run([_nhwc(x)], co.taps_merged(Kc), wm, 4 * Kc, out, grid=(N, H, H), bias=b, phase_cout=Kc)
    torch.cuda.synchronize()
    e = _rel(out.permute(0, 3, 1, 2), ref)
    print(f"halo2 merged sub-pixel C128->64 @32: {e:.3e}")
    return ok and e < 8e-3


@case
def lowres():
    """Low-resolution grids (4x4, 8x8) through the generic kernel's two round-2 paths: weight tiles bulk-copied from the
    swizzled image, and on-chip split-K over a thread-block cluster with the partial tiles reduced through distributed
    shared memory -- plain and with every fused epilogue, against torch and against the single-CTA path (n_tile hint
    switches the cluster off)."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, Kc, H) in [(32, 512, 512, 4), (5, 256, 256, 4), (32, 256, 256, 8), (7, 512, 128, 8), (32, 768, 64, 4),
                          (3, 128, 256, 8)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
        b = torch.randn(Kc, device="cuda")
        ref = F.conv2d(x.float(), w.float(), b, padding=1)
        xn = _nhwc(x)
        wp, keep = _pack_img(w)
        o1, o2 = (torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16) for _ in range(2))
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o1, bias=b)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o2, bias=b, n_tile=min(Kc, 128))
        # fused epilogue on the cluster path: D conv3 (lrelu, aux, gate, addend) and the masked data gradient
        gamma = torch.tensor([0.37], device="cuda")
        add = torch.randn(N, H, H, Kc, device="cuda").bfloat16()
        msk = torch.randn(N, H, H, Kc, device="cuda").bfloat16()
        o3, aux = (torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16) for _ in range(2))
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o3, lrelu_pre=True, aux=aux, scale_dev=gamma, addend=add)
        r3 = F.leaky_relu(F.conv2d(x.float(), w.float(), None, padding=1), 0.2)
        ref3 = 0.37 * r3.bfloat16().float() + add.float().permute(0, 3, 1, 2)
        o4 = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o4, mask_src=msk)
        ref4 = F.conv2d(x.float(), w.float(), None, padding=1) * torch.where(msk.float() > 0, 1.0, 0.2).permute(0, 3, 1, 2)
        # 4x4 stride 2 onto this grid (16 taps over four parity views of a 2H x 2H input)
        x2 = torch.randn(N, C, 2 * H, 2 * H, device="cuda").bfloat16()
        w4 = (torch.randn(Kc, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        r5 = F.leaky_relu(F.conv2d(x2.float(), w4.float(), None, stride=2, padding=1), 0.2)
        w4p, keep4 = _pack_img(w4)
        o5 = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
        co.conv_run(co.parity_views(_nhwc(x2)), co.taps_4x4s2(Kc), w4p, Kc, o5, lrelu_pre=True)
        torch.cuda.synchronize()
        e = [_rel(o1.permute(0, 3, 1, 2), ref), _rel(o1, o2), _rel(o3.permute(0, 3, 1, 2), ref3),
             _rel(aux.permute(0, 3, 1, 2), r3), _rel(o4.permute(0, 3, 1, 2), ref4), _rel(o5.permute(0, 3, 1, 2), r5)]
        print(f"lowres N{N} C{C} K{Kc} {H}: 3x3 {e[0]:.3e} vs single-CTA {e[1]:.3e} | D conv3 {e[2]:.3e} aux {e[3]:.3e} "
              f"| masked {e[4]:.3e} | 4x4s2 {e[5]:.3e}")
        ok &= e[0] < 8e-3 and e[1] < 2e-3 and max(e[2:]) < 8e-3
    return ok


@case
def merged():
    """Merged-phase launches: upsample->conv3x3 forward and the 4x4/stride-2 data gradient as ONE conv over 9 shifts."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, Kc, H) in [(2, 64, 32, 64), (2, 256, 128, 16), (3, 128, 64, 32), (2, 256, 256, 8)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5))
        b = torch.randn(Kc, device="cuda")
        ref = F.conv2d(F.interpolate(x.float(), scale_factor=2), w, b, padding=1)
        wm = _pack(w, slots=co.subpixel_merged_slots())
        out = torch.empty(N, 2 * H, 2 * H, Kc, device="cuda", dtype=torch.bfloat16)
        co.conv_run([_nhwc(x)], co.taps_merged(Kc), wm, 4 * Kc, out, grid=(N, H, H), bias=b, phase_cout=Kc)
        # stride-2 data gradient: g on the HxH grid, dx on the 2Hx2H grid, plus the pooled-shortcut addend
        w4 = (torch.randn(C, Kc, 4, 4, device="cuda") / (4 * Kc ** 0.5)).bfloat16()     # conv Kc -> C, dgrad C -> Kc
        g = torch.randn(N, C, H, H, device="cuda").bfloat16()
        add = torch.randn(N, H, H, Kc, device="cuda").bfloat16()
        gref = F.conv_transpose2d(g.float(), w4.float(), stride=2, padding=1) + \
            0.25 * F.interpolate(add.float().permute(0, 3, 1, 2), scale_factor=2)
        wd = _pack(w4, True, slots=co.s2dgrad_merged_slots())
        gx = torch.empty(N, 2 * H, 2 * H, Kc, device="cuda", dtype=torch.bfloat16)
        co.conv_run([_nhwc(g)], co.taps_merged(Kc), wd, 4 * Kc, gx, grid=(N, H, H), addend=add, add_scale=0.25,
                    phase_cout=Kc)
        torch.cuda.synchronize()
        e = (_rel(out.permute(0, 3, 1, 2), ref), _rel(gx.permute(0, 3, 1, 2), gref))
        print(f"merged N{N} C{C} K{Kc} {H}: subpixel fwd {e[0]:.3e} s2 dgrad {e[1]:.3e}")
        ok &= max(e) < 8e-3
    return ok


@case
def edge():
    """image-boundary convs: stem via im2col (fwd / wgrad / dgrad to fp32 NCHW) and conv_out (+tanh)."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co, kernels as K
    N, H = 2, 64
    img = torch.randn(N, 3, H, H, device="cuda")
    w = torch.randn(32, 3, 3, 3, device="cuda") / 5
    b = torch.randn(32, device="cuda")
    imgr, wr = img.clone().requires_grad_(), w.clone().requires_grad_()
    ref = F.conv2d(imgr, wr, b, padding=1)
    g = torch.randn_like(ref).bfloat16()
    gi_ref, gw_ref = torch.autograd.grad(ref, (imgr, wr), g.float())
    xcol = torch.empty(N, H, H, 32, device="cuda", dtype=torch.bfloat16)
    K.im2col3(img, xcol)
    w0f = torch.empty(32, 32, device="cuda", dtype=torch.bfloat16)
    K.pack_weights(w, w0f, as_matrix=True, cols_pad=32)
    out = torch.empty(N, H, H, 32, device="cuda", dtype=torch.bfloat16)
    co.conv_run([xcol], co.taps_1x1(), w0f, 32, out, bias=b)
    dw = torch.zeros_like(w)
    gn = _nhwc(g)
    co.wgrad_run([xcol], [(0, 0, 0, 0)], gn, dw, (27, 1, 0), ci_valid=27)
    w0d = _pack(w, True, rows_pad=16)
    gi = torch.empty(N, 3, H, H, device="cuda")
    co.conv_run([gn], co.taps_3x3_dgrad(16), w0d, 16, None, out_nchw=gi, co_valid=3)
    torch.cuda.synchronize()
    e = [_rel(out.permute(0, 3, 1, 2), ref), _rel(dw, gw_ref), _rel(gi, gi_ref)]
    print(f"stem: fwd {e[0]:.3e} wgrad {e[1]:.3e} dgrad {e[2]:.3e}")
    # conv_out: lrelu'd 32-ch input -> 3 ch + tanh, and its backward
    x = torch.randn(N, 32, H, H, device="cuda").bfloat16()
    wo = torch.randn(3, 32, 3, 3, device="cuda") / 17
    bo = torch.randn(3, device="cuda") / 3
    xr, wor, bor = x.float().requires_grad_(), wo.clone().requires_grad_(), bo.clone().requires_grad_()
    ref = torch.tanh(F.conv2d(xr, wor, bor, padding=1))
    gimg = torch.randn_like(ref)
    gx_ref, gwo_ref, gbo_ref = torch.autograd.grad(ref, (xr, wor, bor), gimg)
    wof = _pack(wo, rows_pad=16)
    bo16 = torch.zeros(16, device="cuda")
    bo16[:3] = bo
    o = torch.empty(N, 3, H, H, device="cuda")
    xn = _nhwc(x)
    co.conv_run([xn], co.taps_3x3(16), wof, 16, None, bias=bo16, out_nchw=o, co_valid=3, tanh_out=True)
    g_t = torch.empty(N, H, H, 32, device="cuda", dtype=torch.bfloat16)
    dbo = torch.zeros(3, device="cuda")
    K.tanh_bwd_pack(gimg, o, g_t, dbias=dbo)
    dwo = torch.zeros_like(wo)
    co.wgrad_run([xn], [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)], g_t, dwo, (288, 9, 1), co_valid=3)
    wod = _pack(wo, True, cols_pad=32)
    gx = torch.empty(N, H, H, 32, device="cuda", dtype=torch.bfloat16)
    co.conv_run([g_t], co.taps_3x3_dgrad(32), wod, 32, gx)
    torch.cuda.synchronize()
    e2 = [_rel(o, ref), _rel(dwo, gwo_ref), _rel(dbo, gbo_ref), _rel(gx.permute(0, 3, 1, 2), gx_ref)]
    print(f"conv_out: fwd {e2[0]:.3e} wgrad {e2[1]:.3e} dbias {e2[2]:.3e} dgrad {e2[3]:.3e}")
    return max(e + e2) < 8e-3


@case
def stem():
    """fused stem kernel (csrc/stem.cu) vs torch fp32 conv2d on the bf16-rounded operands, vs the round-1 pair
    im2col3 + 1x1 GEMM (bit-identical im2col tensor, outputs equal to the last bf16 bit up to fp32 summation order),
    with / without bias, per-sample scale, im2col side output; sizes with a ragged last 128-pixel tile."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co, kernels as K
    ok = True
    for (N, H, W, use_bias, use_scale, want_col) in [(2, 64, 64, True, False, True), (3, 256, 256, True, False, False),
                                                     (2, 40, 24, False, True, True), (1, 8, 8, True, True, True),
                                                     (5, 12, 20, True, False, False)]:
        img = torch.randn(N, 3, H, W, device="cuda")
        w = torch.randn(32, 3, 3, 3, device="cuda") / 5
        b = torch.randn(32, device="cuda") if use_bias else None
        rs = (torch.rand(N, device="cuda") + 0.5) if use_scale else None
        w0f = torch.empty(32, 32, device="cuda", dtype=torch.bfloat16)
        K.pack_weights(w, w0f, as_matrix=True, cols_pad=32)
        out = torch.empty(N, H, W, 32, device="cuda", dtype=torch.bfloat16)
        col = torch.full((N, H, W, 32), 7.0, device="cuda", dtype=torch.bfloat16) if want_col else None
        K.stem_fwd(img, w0f, out, bias=b, rowscale=rs, xcol=col)
        # the pair it replaces
        col2 = torch.empty(N, H, W, 32, device="cuda", dtype=torch.bfloat16)
        K.im2col3(img, col2, rowscale=rs)
        out2 = torch.empty_like(out)
        co.conv_run([col2], co.taps_1x1(), w0f, 32, out2, bias=b)
        torch.cuda.synchronize()
        xs = img if rs is None else img * rs.view(N, 1, 1, 1)
        ref = F.conv2d(xs.bfloat16().float(), w.bfloat16().float(), b, padding=1)
        e_ref = _rel(out.permute(0, 3, 1, 2), ref)
        e_pair = float((out.float() - out2.float()).abs().max() / (out2.float().abs().max() + 1e-30))
        col_same = True if col is None else bool(torch.equal(col, col2))
        print(f"stem N{N} {H}x{W} bias {use_bias} scale {use_scale}: vs torch {e_ref:.3e}, vs im2col+GEMM pair "
              f"{e_pair:.3e}, im2col side output identical: {col_same}")
        ok &= e_ref < 4e-3 and e_pair < 8e-3 and col_same
    return ok


@case
def pointwise():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import kernels as K
    ok = True
    lr = lambda t: F.leaky_relu(t, 0.2)  # noqa: E731
    for (N, H, C) in [(3, 16, 64), (2, 64, 32), (4, 4, 256)]:
        T = 4 * C + 64
        tab = torch.randn(N, T, device="cuda")
        x = torch.randn(N, H, H, C, device="cuda").bfloat16()
        g = torch.randn(N, H, H, C, device="cuda").bfloat16()
        gadd = torch.randn(N, H, H, C, device="cuda").bfloat16()
        tabs = tuple(tab[:, 32 + i * C: 32 + (i + 1) * C] for i in range(4))
        xr = x.float().requires_grad_()
        tr = [t.clone().requires_grad_() for t in tabs]
        bc = lambda t: t[:, None, None, :]  # noqa: E731
        u1 = bc(tr[0]) * xr + bc(tr[1])
        a = lr(bc(tr[2]) * lr(u1) + bc(tr[3]))
        gsc = torch.tensor([0.7], device="cuda")
        grads = torch.autograd.grad(a, [xr] + tr, g.float() * gsc)
        out = torch.empty_like(x)
        K.affine2_fwd(x, out, tabs, T)
        dtab = torch.zeros_like(tab)
        dts = tuple(dtab[:, 32 + i * C: 32 + (i + 1) * C] for i in range(4))
        gx = torch.empty_like(x)
        dot = torch.zeros(1, device="cuda")
        cs = torch.zeros(C, device="cuda")
        K.affine2_bwd(g, x, gx, tabs, dts, T, g_add=gadd, gscale=gsc, dot=dot, colsum=cs)
        torch.cuda.synchronize()
        gx_ref = grads[0] + gadd.float()
        e = [_rel(out, a), _rel(gx, gx_ref), _rel(dot, (g.float() * a).sum().reshape(1)), _rel(cs, gx_ref.sum((0, 1, 2)))]
        e += [_rel(dts[i], grads[1 + i]) for i in range(4)]
        print(f"affine2 N{N} H{H} C{C}: " + " ".join(f"{v:.2e}" for v in e))
        ok &= max(e) < 8e-3
        # pooling / dres
        if H >= 2:
            po = torch.empty(N, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16)
            cs = torch.zeros(C, device="cuda")
            K.pool(x, po, shift=1, scale=0.25, colsum=cs, cs_dev=gsc, cs_imm=2.0)
            ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
            r = torch.randn(N, H, H, C, device="cuda").bfloat16()
            gp = torch.empty_like(x)
            dg = torch.zeros(1, device="cuda")
            cs2 = torch.zeros(C, device="cuda")
            K.dres_bwd(g, r, gp, gsc, dg, cs2)
            d2 = torch.zeros(1, device="cuda")
            K.dot(g, r, d2, 0.5)
            torch.cuda.synchronize()
            e = [_rel(po, ref), _rel(cs, 1.4 * x.float().sum((0, 1, 2))),
                 _rel(gp, g.float() * 0.7 * torch.where(r.float() > 0, 1.0, 0.2)),
                 _rel(dg, (g.float() * r.float()).sum().reshape(1)), _rel(cs2, g.float().sum((0, 1, 2))),
                 _rel(d2, 0.5 * (g.float() * r.float()).sum().reshape(1))]
            print(f"pool/dres N{N} H{H} C{C}: " + " ".join(f"{v:.2e}" for v in e))
            ok &= max(e) < 8e-3
    # layout + head + rowsq
    emb = torch.randn(3, 512, 4, 4, device="cuda")
    sent = torch.randn(3, 256, device="cuda")
    ssc = torch.rand(3, device="cuda")
    cat = torch.empty(3, 4, 4, 768, device="cuda", dtype=torch.bfloat16)
    K.head_cat(emb, sent, cat, sscale=ssc)
    ref = torch.cat((emb, (sent * ssc[:, None])[:, :, None, None].expand(-1, -1, 4, 4)), 1)
    ge, gs = torch.empty(3, 512, 4, 4, device="cuda"), torch.empty(3, 256, device="cuda")
    K.head_cat_bwd(cat, ge, gs)
    hid = torch.randn(3, 4, 4, 64, device="cuda").bfloat16()
    w2 = torch.randn(1, 64, 4, 4, device="cuda")
    lg = torch.empty(3, device="cuda")
    K.head_out(hid, w2, lg)
    gl = torch.randn(3, device="cuda")
    ghid = torch.empty_like(hid)
    dw2 = torch.zeros_like(w2)
    K.head_out_bwd(gl, hid, hid, w2, ghid, dw2)
    rs = torch.zeros(3, device="cuda")
    K.rowsq(emb, rs)
    back = torch.empty(3, 4, 4, 512, device="cuda", dtype=torch.bfloat16)
    K.nchw_to_nhwc(emb, back)
    rt = torch.empty_like(emb)
    K.nhwc_to_nchw(back, rt)
    torch.cuda.synchronize()
    hidn = hid.float().permute(0, 3, 1, 2)
    e = [_rel(cat.permute(0, 3, 1, 2), ref), _rel(ge, cat.float().permute(0, 3, 1, 2)[:, :512]),
         _rel(gs, cat.float()[..., 512:].sum((1, 2))), _rel(lg, (hidn * w2).sum((1, 2, 3))),
         _rel(ghid.permute(0, 3, 1, 2), gl[:, None, None, None] * w2 * torch.where(hidn > 0, 1.0, 0.2)),
         _rel(dw2, (gl[:, None, None, None] * hidn).sum(0, keepdim=True)), _rel(rs, (emb ** 2).sum((1, 2, 3))),
         _rel(rt, emb)]
    print("head/layout: " + " ".join(f"{v:.2e}" for v in e))
    ok &= max(e) < 8e-3
    return ok


@case
def optim():
    torch = _setup()
    from gan_codes_b200 import kernels as K
    n = 1_000_003
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.01
    pr = p.clone().requires_grad_()
    opt = torch.optim.Adam([pr], lr=4e-4, betas=(0.0, 0.9))
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    ok = True
    for it in range(3):
        gi = g * (it + 1)
        pr.grad = gi.clone()
        torch.nn.utils.clip_grad_norm_([pr], 5.0)
        opt.step()
        ss = torch.zeros(1, device="cuda")
        K.sqnorm(gi, ss)
        step += 1
        K.adam_step(p, gi, m, v, ss, step, lr=4e-4)
        torch.cuda.synchronize()
        e = _rel(p - pr.detach(), torch.zeros_like(p)) if False else ((p - pr.detach()).abs().max() / 4e-4).item()
        print(f"adam it{it}: max |dp| / lr = {e:.3e}, sumsq rel {abs(ss.item() - (gi ** 2).sum().item()) / (gi ** 2).sum().item():.2e}")
        ok &= e < 1e-2
    # scheduling-independent norm (data-parallel ranks must derive bit-identical clip coefficients): same bits on
    # every run, value within fp32 rounding of the fp64 sum
    big = torch.randn(19_645_272, device="cuda") * 0.05
    scratch = torch.zeros(4 * torch.cuda.get_device_properties(0).multi_processor_count + 8, device="cuda")
    vals = []
    for _ in range(6):
        ss = torch.zeros(1, device="cuda")
        K.sqnorm_det(big, ss, scratch)
        vals.append(ss.item())
    ref = (big.double() ** 2).sum().item()
    print(f"sqnorm_det: {len(set(vals))} distinct value(s) over 6 runs, rel err {abs(vals[0] - ref) / ref:.2e}")
    ok &= len(set(vals)) == 1 and abs(vals[0] - ref) / ref < 1e-5
    return ok


@case
def halo():
    """High-resolution narrow layers take the halo kernel (conv_halo_kernel); every tap-list flavour is checked
    against torch and against the generic kernel (forced with an explicit n_tile)."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, Kc, H) in [(2, 32, 32, 64), (2, 64, 64, 64), (1, 64, 16, 128), (2, 32, 64, 96), (1, 64, 128, 64)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w = (torch.randn(Kc, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
        b = torch.randn(Kc, device="cuda")
        ref = F.conv2d(x.float(), w.float(), b, padding=1)
        o1 = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
        o2 = torch.empty_like(o1)
        xn, wp = _nhwc(x), _pack(w)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o1, bias=b)
        co.conv_run([xn], co.taps_3x3(Kc), wp, Kc, o2, bias=b, n_tile=Kc)
        w1 = (torch.randn(Kc, C, 1, 1, device="cuda") / C ** 0.5).bfloat16()
        o3 = torch.empty_like(o1)
        co.conv_run([xn], co.taps_1x1(), _pack(w1), Kc, o3)
        torch.cuda.synchronize()
        e = (_rel(o1.permute(0, 3, 1, 2), ref), _rel(o1, o2), _rel(o3.permute(0, 3, 1, 2), F.conv2d(x.float(), w1.float())))
        print(f"halo 3x3 N{N} C{C} K{Kc} {H}: vs torch {e[0]:.3e} vs generic {e[1]:.3e} 1x1 {e[2]:.3e}")
        ok &= e[0] < 8e-3 and e[1] < 1e-3 and e[2] < 8e-3
    # 4x4 stride 2 over four parity views, its phase data gradient, the pooled 1x1 shortcut
    for (N, C, Kc, H) in [(2, 32, 64, 256), (1, 64, 64, 128)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        w4 = (torch.randn(Kc, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        ref = F.leaky_relu(F.conv2d(x.float(), w4.float(), None, stride=2, padding=1), 0.2)
        out = torch.empty(N, H // 2, H // 2, Kc, device="cuda", dtype=torch.bfloat16)
        pv = co.parity_views(_nhwc(x))
        co.conv_run(pv, co.taps_4x4s2(Kc), _pack(w4), Kc, out, lrelu_pre=True)
        ws = (torch.randn(Kc, C, 1, 1, device="cuda") / C ** 0.5).bfloat16()
        bs = torch.randn(Kc, device="cuda")
        sref = F.avg_pool2d(F.conv2d(x.float(), ws.float(), bs), 2)
        so = torch.empty_like(out)
        co.conv_run(pv, [(i, 0, 0, 0) for i in range(4)], _pack(ws, scale=0.25), Kc, so, bias=bs)
        e = [_rel(out.permute(0, 3, 1, 2), ref), _rel(so.permute(0, 3, 1, 2), sref)]
        if Kc <= 64:
            g2 = torch.randn(N, Kc, H // 2, H // 2, device="cuda").bfloat16()
            gref = F.conv_transpose2d(g2.float(), w4.float(), stride=2, padding=1)
            gx = torch.zeros(N, H, H, C, device="cuda", dtype=torch.bfloat16)
            wd, g2n = _pack(w4, True), _nhwc(g2)
            add = torch.randn(N, H // 2, H // 2, C, device="cuda").bfloat16()
            for a in (0, 1):
                for c in (0, 1):
                    co.conv_run([g2n], co.taps_4x4s2_dgrad(C, a, c), wd, C, gx[:, a::2, c::2, :], addend=add, add_scale=0.25)
            gref = gref + 0.25 * F.interpolate(add.float().permute(0, 3, 1, 2), scale_factor=2)
            e.append(_rel(gx.permute(0, 3, 1, 2), gref))
        torch.cuda.synchronize()
        print(f"halo 4s2 N{N} C{C} K{Kc} {H}: " + " ".join(f"{v:.3e}" for v in e))
        ok &= max(e) < 8e-3
    return ok


@case
def wgrad():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, K, H, k) in [(2, 64, 64, 32, 3), (2, 32, 32, 64, 3), (4, 256, 256, 8, 3), (2, 128, 64, 32, 3),
                            (2, 512, 512, 4, 3), (2, 32, 64, 64, 4), (2, 256, 512, 16, 4), (2, 64, 128, 32, 1),
                            # high-resolution narrow layers -> wgrad_halo_kernel
                            (2, 64, 64, 64, 3), (3, 32, 64, 96, 3), (2, 64, 32, 128, 3), (2, 32, 64, 256, 4),
                            (2, 64, 64, 128, 4), (2, 32, 32, 128, 1), (2, 64, 32, 64, 1)]:
        x = torch.randn(N, C, H, H, device="cuda").bfloat16()
        Ho = H // 2 if k == 4 else H
        g = torch.randn(N, K, Ho, Ho, device="cuda").bfloat16()
        xr = x.float().requires_grad_()
        wr = torch.zeros(K, C, k, k, device="cuda", requires_grad=True)
        y = F.conv2d(xr, wr, None, stride=2 if k == 4 else 1, padding={1: 0, 3: 1, 4: 1}[k])
        (ref,) = torch.autograd.grad(y, wr, g.float())
        dw = torch.zeros(K, C, k, k, device="cuda")
        if k == 4:
            srcs, taps = co.parity_views(_nhwc(x)), [(s, dy, dx, i) for i, (s, dy, dx, _) in enumerate(co.taps_4x4s2(0))]
        elif k == 3:
            srcs, taps = [_nhwc(x)], [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
        else:
            srcs, taps = [_nhwc(x)], [(0, 0, 0, 0)]
        co.wgrad_run(srcs, taps, _nhwc(g), dw, (C * k * k, k * k, 1), scale_imm=1.0)
        torch.cuda.synchronize()
        err = _rel(dw, ref)
        print(f"wgrad N{N} C{C} K{K} {H} k{k}: rel {err:.3e}")
        ok &= err < 8e-3
    return ok


@case
def unpack():
    """K.WgradScratch: tap-major scratch -> OIHW gradients (dfgan_wgrad_unpack), ragged shapes, accumulate + re-zero."""
    torch = _setup()
    from gan_codes_b200 import kernels as K
    shapes = {"a": (3, 32, 3, 3), "b": (64, 768, 3, 3), "c": (40, 24, 4, 4), "d": (512, 256, 3, 3), "e": (33, 65, 1, 1)}
    ws = K.WgradScratch(shapes, torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(3)
    ok = True
    for rnd in range(2):
        grads = {n: torch.randn(*sh, device="cuda", generator=g) for n, sh in shapes.items()}
        want = {}
        for n, sh in shapes.items():
            v = ws.view(n)
            assert float(v.abs().max()) == 0.0, "scratch must be zero before a sweep"
            v.copy_(torch.randn(v.shape, device="cuda", generator=g))
            want[n] = grads[n] + v.permute(1, 2, 0).reshape(sh)      # [KK][Co][Ci] -> [Co][Ci][KK]
        ws.flush(grads)
        torch.cuda.synchronize()
        for n in shapes:
            err = float((grads[n] - want[n]).abs().max())
            zero = float(ws.view(n).abs().max())
            print(f"unpack round {rnd} {n} {shapes[n]}: max err {err:.2e}, scratch after flush {zero:.1e}")
            ok &= err == 0.0 and zero == 0.0
    return ok


@case
def perf():
    torch = _setup()
    from gan_codes_b200 import convops as co
    for (N, C, K, H) in [(32, 256, 256, 32), (32, 256, 128, 64), (32, 128, 128, 64), (32, 64, 64, 128),
                         (32, 64, 32, 256), (32, 32, 32, 256), (32, 512, 512, 16)]:
        x = torch.randn(N, H, H, C, device="cuda").bfloat16()
        wp = torch.randn(9 * K, C, device="cuda").bfloat16()
        out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
        g = torch.randn(N, H, H, K, device="cuda").bfloat16()
        dw = torch.zeros(K, C, 3, 3, device="cuda")
        taps_w = [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
        for name, fn in (("fprop", lambda: co.conv_run([x], co.taps_3x3(K), wp, K, out)),
                         ("wgrad", lambda: co.wgrad_run([x], taps_w, g, dw, (C * 9, 9, 1)))):
            for _ in range(3):
                fn()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                fn()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 10
            fl = 2.0 * N * H * H * C * K * 9
            print(f"{name} N{N} C{C} K{K} {H}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
    return True


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        bad = []
        for name in CASES:
            r = subprocess.run(["timeout", "180", sys.executable, __file__, name])
            if r.returncode != 0:
                bad.append(name)
        print("FAILED:" if bad else "ALL OK", bad)
        sys.exit(1 if bad else 0)
    ok = CASES[which]()
    print(which, "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
