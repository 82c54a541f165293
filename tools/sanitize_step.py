"""Smallest complete run of the hot path for compute-sanitizer (one tool per gpurun call, B200_PROFILING.md):
one eager training iteration (D hinge + MA-GP + G, clip + Adam, side streams and programmatic dependent launches on)
at batch 2, 256x256, then one generator sampling call; exits 0 when the losses are finite.

    compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_step.py
    compute-sanitizer --tool synccheck python tools/sanitize_step.py
    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
"""
import os
import sys

os.environ.setdefault("DFGAN_NO_GRAPH", "1")     # eager launches: the tools instrument kernels, not graph replays
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200.engine import DFGANTrainer  # noqa: E402
from gan_codes_b200.modules import Discriminator, Generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(7)
G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
tr = DFGANTrainer(G, D, B)
img = (torch.randn(B, 3, 256, 256, device="cuda") * 0.5).clamp(-1, 1)
sent, noise = torch.randn(B, 256, device="cuda") * 0.3, torch.randn(B, 100, device="cuda")
out = tr.step(img, sent, noise)
torch.cuda.synchronize()
vals = {k: float(v) for k, v in out.items()}
x = tr.sample(noise, sent)
torch.cuda.synchronize()
print("losses", vals, "sample", tuple(x.shape), float(x.abs().max()))
ok = all(v == v and abs(v) != float("inf") for v in vals.values()) and bool(torch.isfinite(x).all())
print("SANITIZE RUN OK" if ok else "SANITIZE RUN FAILED")
sys.exit(0 if ok else 1)
