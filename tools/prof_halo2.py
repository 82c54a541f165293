"""halo2 (wide-layer halo kernel) vs the generic kernel on plain 3x3 layers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co, kernels as K  # noqa: E402

torch.manual_seed(0)
for (N, C, Kc, H) in [(32, 128, 128, 64), (32, 256, 256, 32), (32, 512, 512, 16), (32, 256, 128, 64), (64, 128, 128, 64)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    w = torch.randn(Kc, C, 3, 3, device="cuda") / 50
    wp = torch.empty(9 * Kc, C, device="cuda", dtype=torch.bfloat16)
    wi = torch.empty_like(wp)
    K.pack_weights(w, wp)
    K.pack_weights(w, wi, image=1)
    out = torch.empty(N, H, H, Kc, device="cuda", dtype=torch.bfloat16)
    for name, reg in (("generic", False), ("halo2", True)):
        co.IMAGE_OF.pop(wp.data_ptr(), None)
        if reg:
            co.register_image(wp, wi, Kc)
        fn = lambda: co.conv_run([x], co.taps_3x3(Kc), wp, Kc, out)  # noqa: E731
        for _ in range(2):
            fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            fn()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print(f"{name:8s} N{N} C{C} K{Kc} {H}x{H}: {ms * 1e3:7.1f} us {2.0 * N * H * H * C * Kc * 9 / ms / 1e9:7.1f} TF/s")
