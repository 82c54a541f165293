"""Graph-timed microbenchmark of the low-resolution convolution launches of one training iteration (<= 16x16 grids,
the logit head): every signature is captured 16x back to back in a CUDA graph (no host launch latency in the number)
with its real weight packing (tap-major pack + swizzled image) and a plain epilogue.

    python tools/prof_lowres.py            # table: us per launch, TFLOP/s
    DFGAN_NO_CLUSTER=1 DFGAN_NO_BIMG=1 python tools/prof_lowres.py     # the round-1 path
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co  # noqa: E402
from tools.bringup_conv import _pack_img  # noqa: E402

REPS = 16
#        N   Cin  Cout  H   kind
SHAPES = [(32, 512, 512, 4, "3x3"), (32, 256, 256, 4, "3x3"), (32, 512, 512, 8, "3x3"), (32, 256, 256, 8, "3x3"),
          (32, 256, 256, 16, "3x3"), (32, 512, 512, 16, "3x3"), (64, 512, 512, 4, "3x3"), (64, 512, 512, 8, "3x3"),
          (32, 512, 512, 4, "dgrad4s2-phase"), (32, 512, 512, 8, "dgrad4s2-phase"), (32, 512, 256, 16, "dgrad4s2-phase"),
          (32, 256, 256, 4, "subpixel-phase"), (32, 256, 256, 8, "subpixel-phase"), (32, 256, 256, 16, "subpixel-phase"),
          (32, 512, 512, 4, "4x4s2"), (32, 512, 512, 8, "4x4s2"), (32, 256, 512, 16, "4x4s2"),
          (95, 768, 64, 4, "3x3"), (32, 768, 64, 4, "3x3"), (32, 64, 768, 4, "3x3"),
          (32, 256, 512, 16, "1x1"), (32, 512, 256, 16, "1x1")]


def build(N, C, Kc, H, kind):
    dev = "cuda"
    if kind == "3x3":
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        w = torch.randn(Kc, C, 3, 3, device=dev) / (3 * C ** 0.5)
        wp, keep = _pack_img(w)
        out = torch.empty(N, H, H, Kc, device=dev, dtype=torch.bfloat16)
        return (lambda: co.conv_run([x], co.taps_3x3(Kc), wp, Kc, out)), 2.0 * N * H * H * Kc * 9 * C, (x, w, wp, keep, out)
    if kind == "1x1":
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        w = torch.randn(Kc, C, 1, 1, device=dev) / C ** 0.5
        wp, keep = _pack_img(w)
        out = torch.empty(N, H, H, Kc, device=dev, dtype=torch.bfloat16)
        return (lambda: co.conv_run([x], co.taps_1x1(), wp, Kc, out)), 2.0 * N * H * H * Kc * C, (x, w, wp, keep, out)
    if kind == "dgrad4s2-phase":     # conv Kc -> C (4x4 s2), data gradient: g [N,H,H,C]... one of four phases, 4 taps
        g = torch.randn(N, H, H, C, device=dev).bfloat16()
        w = torch.randn(C, Kc, 4, 4, device=dev) / (4 * Kc ** 0.5)
        wp, keep = _pack_img(w, True)
        out = torch.empty(N, 2 * H, 2 * H, Kc, device=dev, dtype=torch.bfloat16)
        return (lambda: co.conv_run([g], co.taps_4x4s2_dgrad(Kc, 0, 0), wp, Kc, out[:, 0::2, 0::2, :])), \
            2.0 * N * H * H * Kc * 4 * C, (g, w, wp, keep, out)
    if kind == "subpixel-phase":
        x = torch.randn(N, H, H, C, device=dev).bfloat16()
        w = torch.randn(Kc, C, 3, 3, device=dev) / (3 * C ** 0.5)
        wp, keep = _pack_img(w, slots=co.subpixel_slots())
        out = torch.empty(N, 2 * H, 2 * H, Kc, device=dev, dtype=torch.bfloat16)
        return (lambda: co.conv_run([x], co.taps_subpixel_fwd(Kc, 0, 0), wp, Kc, None, grid=(N, H, H),
                                    aux=out[:, 0::2, 0::2, :])), 2.0 * N * H * H * Kc * 4 * C, (x, w, wp, keep, out)
    if kind == "4x4s2":              # output grid HxH from a 2Hx2H input, four parity views
        x = torch.randn(N, 2 * H, 2 * H, C, device=dev).bfloat16()
        w = torch.randn(Kc, C, 4, 4, device=dev) / (4 * C ** 0.5)
        wp, keep = _pack_img(w)
        out = torch.empty(N, H, H, Kc, device=dev, dtype=torch.bfloat16)
        return (lambda: co.conv_run(co.parity_views(x), co.taps_4x4s2(Kc), wp, Kc, out, lrelu_pre=True)), \
            2.0 * N * H * H * Kc * 16 * C, (x, w, wp, keep, out)
    raise ValueError(kind)


def main():
    torch.manual_seed(0)
    side = torch.cuda.Stream()
    total = 0.0
    print(f"{'signature':46s} {'us/launch':>10s} {'TFLOP/s':>9s}")
    only = [int(a) for a in sys.argv[1:]]
    for idx, (N, C, Kc, H, kind) in enumerate(SHAPES):
        if only and idx not in only:
            continue
        fn, flops, keep = build(N, C, Kc, H, kind)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(REPS):
                    fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (4 * REPS)
        total += us
        print(f"N{N:<3d} {H:>2d}x{H:<2d} Cin{C:<4d} Cout{Kc:<4d} {kind:18s} {us:10.1f} {flops / us / 1e6:9.1f}")
    print(f"sum {total:.1f} us")


if __name__ == "__main__":
    main()
