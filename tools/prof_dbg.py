"""Bottleneck isolation of conv_halo_kernel: runs two narrow layers under DFGAN_DBG masks (one subprocess per mask).
Usage: python tools/prof_dbg.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from gan_codes_b200 import convops as co
    for (N, C, K, H) in [(64, 64, 64, 128), (32, 32, 32, 256), (64, 32, 32, 256)]:
        x = torch.randn(N, H, H, C, device="cuda").bfloat16()
        wp = torch.randn(9 * K, C, device="cuda").bfloat16()
        out = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
        m = torch.randn(N, H, H, K, device="cuda").bfloat16()
        aux = torch.empty_like(out)
        gam = torch.ones(1, device="cuda")
        variants = {"plain": lambda: co.conv_run([x], co.taps_3x3(K), wp, K, out),
                    "mask": lambda: co.conv_run([x], co.taps_3x3(K), wp, K, out, mask_src=m),
                    "Dconv3": lambda: co.conv_run([x], co.taps_3x3(K), wp, K, out, lrelu_pre=True, aux=aux, scale_dev=gam,
                                                  addend=m)}
        res = []
        for name, fn in variants.items():
            for _ in range(3):
                fn()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                fn()
            t1.record()
            torch.cuda.synchronize()
            res.append(f"{name} {t0.elapsed_time(t1) / 20 * 1e3:7.1f} us")
        print(f"  DBG={os.environ.get('DFGAN_DBG', '0'):>2s} N{N} C{C}->{K} {H}x{H}: " + " | ".join(res), flush=True)
else:
    for mask in (0, 1, 8, 2, 4, 3, 5, 6, 7):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DFGAN_DBG=str(mask)))
