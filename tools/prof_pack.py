"""Timing of the weight re-pack per operand variant (512x512 4x4 and 3x3 weights, plain / transposed / image layouts)
and of the discriminator's whole batched re-pack.  Usage: python tools/prof_pack.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import kernels as K  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps * 1e3


for (co, ci, k) in [(512, 512, 4), (512, 512, 3), (64, 32, 4)]:
    w = torch.randn(co, ci, k, k, device="cuda")
    for tr in (False, True):
        for image in (0, 1):
            out = torch.empty(k * k * (ci if tr else co), (co if tr else ci), device="cuda", dtype=torch.bfloat16)
            if image and out.shape[1] % 64:
                continue
            us = timed(lambda: K.pack_weights(w, out, transpose=tr, image=image))
            mb = (w.numel() * 4 + out.numel() * 2) / 1e6
            print(f"{co}x{ci} k{k} transpose={int(tr)} image={image}: {us:7.1f} us  {mb / us * 1e-3:6.2f} TB/s ({mb:.1f} MB)")

from gan_codes_b200.modules import Discriminator  # noqa: E402
from gan_codes_b200 import dnet  # noqa: E402
D = Discriminator(32).cuda()
sd = {k: v.detach() for k, v in D.named_parameters()}
pk = dnet.DPacked(sd)
print(f"DPacked.refresh (all {len(pk._pp.entries)} entries, {pk._pp.blocks} blocks): {timed(lambda: pk.refresh(sd)):7.1f} us")

for (co, ci, k) in [(2048, 1024, 3), (2048, 1024, 4)]:
    w = torch.randn(co, ci, k, k, device="cuda")
    for tr in (False, True):
        out = torch.empty(k * k * (ci if tr else co), (co if tr else ci), device="cuda", dtype=torch.bfloat16)
        us = timed(lambda: K.pack_weights(w, out, transpose=tr, image=0))
        mb = (w.numel() * 4 + out.numel() * 2) / 1e6
        print(f"{co}x{ci} k{k} transpose={int(tr)}: {us:7.1f} us  {mb / us * 1e-6:6.2f} TB/s ({mb:.1f} MB)")
