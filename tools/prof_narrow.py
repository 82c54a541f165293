"""A few narrow-layer launches (C32 @256^2, C64 @128^2, batch 32) with the fused epilogues of the training iteration,
for ncu source-level captures and quick CUDA-event timing:  python tools/prof_narrow.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
for (N, C, K, H) in [(32, 32, 32, 256), (32, 64, 64, 128)]:
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
        "G conv_2: bias,gate,addend>>1,out,out2(affine)": dict(out=out, bias=bias, scale_dev=gam, addend=add, add_shift=1,
                                                             out2=out2, post_mode=2, tables=tabs, tab_stride=4 * K),
        "D conv3: lrelu,aux,gate,addend,out": dict(out=out, lrelu_pre=True, aux=aux, scale_dev=gam, addend=msk),
    }
    for name, kw in variants.items():
        o = kw.pop("out")
        fn = lambda: co.conv_run([x], tl, wp, K, o, **kw)  # noqa: E731
        fn()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        print(f"N{N} C{C} {H}x{H} {name:48s}: {t0.elapsed_time(t1) / reps * 1e3:7.1f} us")
