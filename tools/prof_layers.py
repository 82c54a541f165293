"""Runs a few representative layer shapes (for ncu captures and quick timing).
Usage: python tools/prof_layers.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
LAYERS = [(32, 32, 32, 256, 9), (32, 64, 64, 128, 9), (32, 32, 32, 256, 1), (32, 32, 32, 256, 9)]
for (N, C, K, H, taps) in LAYERS:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    wp = torch.randn(taps * K, C, device="cuda").bfloat16()
    out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
    g = torch.randn(N, H, H, K, device="cuda").bfloat16()
    dw = torch.zeros(K, C, 3, 3, device="cuda")
    tl = co.taps_3x3(K) if taps == 9 else co.taps_1x1()
    tw = [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)] if taps == 9 else [(0, 0, 0, 0)]
    for name, fn in (("conv", lambda: co.conv_run([x], tl, wp, K, out)),
                     ("wgrad", lambda: co.wgrad_run([x], tw, g, dw, (C * 9, 9, 1)))):
        for _ in range(2):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        fl = 2.0 * N * H * H * C * K * taps
        gb = 2.0 * N * H * H * (C + K) / 1e9
        print(f"{name:5s} N{N} C{C} K{K} {H}x{H} taps{taps}: {ms * 1e3:8.1f} us {fl / ms / 1e9:7.1f} TF/s  {gb / ms * 1e3:7.1f} GB/s(min-traffic)")

# epilogue variants of the halo conv (C64 -> C64 @128^2 and C32 -> C32 @256^2): what the fused epilogues cost
for (N, C, K, H) in [(32, 64, 64, 128), (32, 32, 32, 256)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    wp = torch.randn(9 * K, C, device="cuda").bfloat16()
    out, aux, out2 = (torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    msk = torch.randn(N, H, H, K, device="cuda").bfloat16()
    add = torch.randn(N, H // 2, H // 2, K, device="cuda").bfloat16()
    bias, gam = torch.randn(K, device="cuda"), torch.tensor([0.5], device="cuda")
    tab = torch.randn(N, 4 * K, device="cuda")
    tabs = tuple(tab[:, i * K:(i + 1) * K] for i in range(4))
    tl = co.taps_3x3(K)
    variants = {
        "plain out": dict(out=out),
        "dgrad: out+mask": dict(out=out, mask_src=msk),
        "G conv_1: bias,aux,out2(affine)": dict(out=None, bias=bias, aux=aux, out2=out2, post_mode=2, tables=tabs, tab_stride=4 * K),
        "G conv_2: bias,gate,addend>>1,out,out2(affine)": dict(out=out, bias=bias, scale_dev=gam, addend=add, add_shift=1,
                                                             out2=out2, post_mode=2, tables=tabs, tab_stride=4 * K),
        "D conv3: lrelu,aux,gate,addend,out": dict(out=out, lrelu_pre=True, aux=aux, scale_dev=gam, addend=msk),
    }
    for name, kw in variants.items():
        o = kw.pop("out")
        fn = lambda: co.conv_run([x], tl, wp, K, o, **kw)  # noqa: E731
        for _ in range(2):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        print(f"epilogue N{N} C{C} {H}x{H} {name:48s}: {t0.elapsed_time(t1) / reps * 1e3:7.1f} us")
