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
