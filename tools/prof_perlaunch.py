"""Per-launch durations of the two GEMM engines inside one training iteration, by kernel family and shape.

Every dfgan_conv_run / dfgan_wgrad_run launch is bracketed by a CUDA event pair in C (ProfileScope) and labelled with
its kernel family, shape, epilogue mask, N tile and grid (DFGAN_PROFILE_DUMP).  The iteration runs eagerly, serially
(no side streams) BEHIND a device-side sleep, so the host has queued every launch before the first one starts and the
event pairs see back-to-back kernels rather than host launch gaps.

    python tools/prof_perlaunch.py [batch] [steps]   ->  table sorted by total time: n, avg us, TFLOP/s, algorithmic
                                                         GB/s, share, SM-time (duration x fraction of the 148 SMs used)
"""
import collections
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dump = os.path.join(tempfile.gettempdir(), f"dfgan_perlaunch_{os.getpid()}.tsv")
os.environ["DFGAN_PROFILE_DUMP"] = dump
import torch  # noqa: E402

from gan_codes_b200 import _lib, convops as co  # noqa: E402
from gan_codes_b200.engine import DFGANTrainer  # noqa: E402
from gan_codes_b200.modules import Discriminator, Generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
tr = DFGANTrainer(G, D, B, use_graph=False)
img = (torch.randn(B, 3, 256, 256, device="cuda") * 0.5).clamp(-1, 1)
sent, noise = torch.randn(B, 256, device="cuda") * 0.3, torch.randn(B, 100, device="cuda")
for _ in range(3):
    tr.step(img, sent, noise)
co.SIDE_ENABLED = False
torch.cuda.synchronize()
if os.path.exists(dump):
    os.remove(dump)
_lib.profile_enable(True)
for _ in range(STEPS):
    torch.cuda._sleep(60_000_000)          # ~30 ms: the host queues the whole iteration behind it
    tr.step(img, sent, noise)
    torch.cuda.synchronize()
_lib.profile_read()
_lib.profile_enable(False)

agg = collections.OrderedDict()
for line in open(dump):
    eng, label, ms, fl, by = line.rstrip("\n").split("\t")
    a = agg.setdefault(label, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(ms)
    a[2] += float(fl)
    a[3] += float(by)
os.remove(dump)
tot = sum(a[1] for a in agg.values()) / STEPS
print(f"batch {B}: GEMM-engine launches per iteration {sum(a[0] for a in agg.values()) // STEPS}, "
      f"sum of durations {tot:.3f} ms / iteration (serial, eager, {STEPS} iterations averaged)")
print(f"{'ms/iter':>8s} {'share':>6s} {'n':>3s} {'avg us':>8s} {'TFLOP/s':>8s} {'alg GB/s':>9s} {'SM-ms':>7s}  launch")
smt = 0.0
for label, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    grid = int(label.rsplit("grid", 1)[1])
    frac = min(grid, 148) / 148.0
    smt += ms / STEPS * frac
    print(f"{ms / STEPS:8.4f} {100 * ms / STEPS / tot:5.1f}% {n // STEPS:3d} {1e3 * ms / n:8.1f} {fl / ms / 1e9:8.1f} "
          f"{by / ms / 1e6:9.0f} {ms / STEPS * frac:7.4f}  {label}")
print(f"SM-time total {smt:.3f} ms")
