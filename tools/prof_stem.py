"""Graph-timed comparison of the fused stem kernel (csrc/stem.cu) with the round-1 pair im2col3 + 1x1 GEMM.
    python tools/prof_stem.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co, kernels as K  # noqa: E402

REPS = 16


def graph_time(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * REPS)


for N in (32, 64):
    img = torch.randn(N, 3, 256, 256, device="cuda")
    w = torch.randn(32, 3, 3, 3, device="cuda") / 5
    b = torch.randn(32, device="cuda")
    w0f = torch.empty(32, 32, device="cuda", dtype=torch.bfloat16)
    K.pack_weights(w, w0f, as_matrix=True, cols_pad=32)
    out = torch.empty(N, 256, 256, 32, device="cuda", dtype=torch.bfloat16)
    col = torch.empty(N, 256, 256, 32, device="cuda", dtype=torch.bfloat16)
    px = N * 256 * 256

    def pair():
        K.im2col3(img, col)
        co.conv_run([col], co.taps_1x1(), w0f, 32, out, bias=b)
    t_pair = graph_time(pair)
    t_i2c = graph_time(lambda: K.im2col3(img, col))
    t_f = graph_time(lambda: K.stem_fwd(img, w0f, out, bias=b))
    t_fc = graph_time(lambda: K.stem_fwd(img, w0f, out, bias=b, xcol=col))
    print(f"N{N}: im2col3 {t_i2c:.1f} us | im2col3 + 1x1 GEMM {t_pair:.1f} us | fused stem {t_f:.1f} us "
          f"({px * 76 / t_f / 1e6:.2f} TB/s on 12 + 64 B/pixel) | fused stem + im2col side output {t_fc:.1f} us "
          f"({px * 140 / t_fc / 1e6:.2f} TB/s on 12 + 128 B/pixel)")
