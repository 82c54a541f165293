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


def _run_fprop3(N, C, K, H, W, **kw):
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    x = torch.randn(N, C, H, W, device="cuda").bfloat16()
    w = (torch.randn(K, C, 3, 3, device="cuda") / (3 * C ** 0.5)).bfloat16()
    b = torch.randn(K, device="cuda")
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    out = torch.empty(N, H, W, K, device="cuda", dtype=torch.bfloat16)
    co.conv_run([_nhwc(x)], co.taps_3x3(K), co.pack_fwd(w.float()), K, out, bias=b, **kw)
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
        co.conv_run(co.parity_views(_nhwc(x)), co.taps_4x4s2(K), co.pack_fwd(w.float()), K, out, lrelu_pre=True)
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
        co.conv_run([_nhwc(g)], co.taps_3x3_dgrad(C), co.pack_dgrad(w.float()), C, out)
        torch.cuda.synchronize()
        err = _rel(out.permute(0, 3, 1, 2), ref)
        print(f"dgrad3 N{N} C{C} K{K} {H}: rel {err:.3e}")
        ok &= err < 8e-3
        # stride-2 4x4: g on the H/2 grid, dx on the H grid, written phase by phase
        w4 = (torch.randn(K, C, 4, 4, device="cuda") / (4 * C ** 0.5)).bfloat16()
        g2 = torch.randn(N, K, H // 2, H // 2, device="cuda").bfloat16()
        ref = F.conv_transpose2d(g2.float(), w4.float(), stride=2, padding=1)
        out = torch.zeros(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        wp = co.pack_dgrad(w4.float())
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
    ref = pre * gamma * 2.0 + 0.5 * F.interpolate(sc.float(), scale_factor=2)
    ref = ref * torch.where(msk.float() > 0, 1.0, 0.2)
    out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty_like(out)
    co.conv_run([_nhwc(x)], co.taps_3x3(K), co.pack_fwd(w.float()), K, out, bias=b, lrelu_pre=True,
                scale_dev=gamma, scale_imm=2.0, aux=aux, addend=_nhwc(sc), add_shift=1, add_scale=0.5,
                mask_src=_nhwc(msk))
    torch.cuda.synchronize()
    e1, e2 = _rel(out.permute(0, 3, 1, 2), ref), _rel(aux.permute(0, 3, 1, 2), pre)
    print(f"epilogue: out rel {e1:.3e} aux rel {e2:.3e}")
    return e1 < 8e-3 and e2 < 8e-3


@case
def wgrad():
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import convops as co
    ok = True
    for (N, C, K, H, k) in [(2, 64, 64, 32, 3), (2, 32, 32, 64, 3), (4, 256, 256, 8, 3), (2, 128, 64, 32, 3),
                            (2, 512, 512, 4, 3), (2, 32, 64, 64, 4), (2, 256, 512, 16, 4), (2, 64, 128, 32, 1)]:
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
