"""Is the operand-fill bound per SM or aggregate (L2)?  Generic conv kernel, fixed work, varying CTA count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402

torch.manual_seed(0)
for (N, C, K, H) in [(32, 256, 256, 32), (32, 128, 128, 64)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    wp = torch.randn(9 * K, C, device="cuda").bfloat16()
    out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
    for ctas in (148, 111, 74, 37):
        fn = lambda: co.conv_run([x], co.taps_3x3(K), wp, K, out, max_ctas=ctas)  # noqa: E731
        for _ in range(2):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        tiles = N * H * H // 128 * (K // min(K, 256))
        print(f"C{C} K{K} {H}x{H} ctas={ctas:3d}: {ms * 1e3:7.1f} us  ({tiles} tiles, {tiles / ctas:.1f}/CTA, "
              f"{ms * 1e3 / (tiles / ctas):.2f} us per tile per CTA)")
