"""Weight-gradient launches with the gradient in the reference's OIHW layout (tap fastest: the fp32 atomics of a warp
scatter over 36-64-byte strides) vs a tap-major [tap][co][ci] scratch (lanes = consecutive ci -> coalesced atomics).
Usage: python tools/prof_wgrad_layout.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402

TAPS3 = [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
for (N, C, K, H) in [(32, 256, 256, 32), (32, 512, 512, 16), (32, 128, 128, 64), (64, 256, 256, 32), (32, 512, 512, 8),
                     (32, 64, 64, 128), (32, 32, 32, 256)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    g = torch.randn(N, H, H, K, device="cuda").bfloat16()
    for name, dw, strides in (("OIHW     ", torch.zeros(K, C, 3, 3, device="cuda"), (C * 9, 9, 1)),
                              ("tap-major", torch.zeros(9, K, C, device="cuda"), (C, 1, K * C))):
        fn = lambda: co.wgrad_run([x], TAPS3, g, dw, strides)  # noqa: E731
        for _ in range(3):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"wgrad 3x3 N{N} C{C}->{K} {H}x{H} {name}: {ms * 1e3:7.1f} us {2.0 * N * H * H * C * K * 9 / ms / 1e9:7.1f} TF/s")
