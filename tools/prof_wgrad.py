"""wgrad timing sweep over k_splits for representative shapes. Usage: python tools/prof_wgrad.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
for (N, C, K, H) in [(32, 256, 256, 32), (32, 512, 512, 16), (32, 128, 128, 64), (32, 64, 64, 128), (32, 32, 32, 256)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    g = torch.randn(N, H, H, K, device="cuda").bfloat16()
    dw = torch.zeros(K, C, 3, 3, device="cuda")
    tw = [(0, r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
    for ks in (0, 1, 2, 4, 8, 16, 32):
        fn = lambda: co.wgrad_run([x], tw, g, dw, (C * 9, 9, 1), k_splits=ks)  # noqa: E731
        for _ in range(2):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        fl = 2.0 * N * H * H * C * K * 9
        print(f"wgrad N{N} C{C} K{K} {H}x{H} ks={ks:2d}: {ms * 1e3:8.1f} us {fl / ms / 1e9:7.1f} TF/s")
