"""Times the HBM-bound pointwise kernels at their largest shapes (also the ncu target for them)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import kernels as K  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)


def timeit(name, fn, nbytes):
    for _ in range(2):
        fn()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    print(f"{name:34s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s")


for (N, H, C) in [(32, 256, 32), (32, 128, 64)]:
    T = 4 * C
    tab, dtab = torch.randn(N, T, device="cuda"), torch.zeros(N, T, device="cuda")
    tabs = tuple(tab[:, i * C:(i + 1) * C] for i in range(4))
    dts = tuple(dtab[:, i * C:(i + 1) * C] for i in range(4))
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    g = torch.randn(N, H, H, C, device="cuda").bfloat16()
    o = torch.empty_like(x)
    nb = x.numel() * 2
    gam = torch.tensor([0.5], device="cuda")
    dot, cs, dg = torch.zeros(1, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(1, device="cuda")
    timeit(f"affine2_bwd {N}x{H}x{H}x{C}", lambda: K.affine2_bwd(g, x, o, tabs, dts, T, gscale=gam, dot=dot, colsum=cs), 3 * nb)
    timeit(f"affine2_fwd {N}x{H}x{H}x{C}", lambda: K.affine2_fwd(x, o, tabs, T), 2 * nb)
    timeit(f"dres_bwd    {N}x{H}x{H}x{C}", lambda: K.dres_bwd(g, x, o, gam, dg, cs), 3 * nb)
    po = torch.empty(N, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16)
    timeit(f"pool        {N}x{H}x{H}x{C}", lambda: K.pool(x, po, shift=1, scale=0.25, colsum=cs), nb * 1.25)
    timeit(f"torch copy  {N}x{H}x{H}x{C}", lambda: o.copy_(x), 2 * nb)
img = torch.randn(32, 3, 256, 256, device="cuda")
xc = torch.empty(32, 256, 256, 32, device="cuda", dtype=torch.bfloat16)
timeit("im2col3 32x3x256x256", lambda: K.im2col3(img, xc), img.numel() * 4 + xc.numel() * 2)
