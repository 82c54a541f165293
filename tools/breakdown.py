"""Per-shape time breakdown of one training iteration (CUDA events around every C-ABI call).
Usage: python tools/breakdown.py [batch] -> table sorted by total time."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gan_codes_b200 import convops as co, kernels as K  # noqa: E402
from gan_codes_b200.engine import DFGANTrainer  # noqa: E402
from gan_codes_b200.modules import Discriminator, Generator  # noqa: E402

REC = []


def wrap(mod, name, sig):
    fn = getattr(mod, name)

    def inner(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **kw)
        e1.record()
        REC.append((name, sig(*a, **kw), e0, e1))
        return r
    setattr(mod, name, inner)


def conv_sig(srcs, taps, wpack, cout, out=None, **kw):
    g = kw.get("grid")
    ref = next((t for t in (out, kw.get("aux"), kw.get("out2"), kw.get("out_nchw")) if t is not None))
    if g is None:
        g = tuple(ref.shape[:3]) if ref.dtype == torch.bfloat16 else (ref.shape[0], ref.shape[2], ref.shape[3])
    n, h, w = g
    cin = wpack.shape[1]
    fl = 2.0 * n * h * w * cout * len(taps) * cin
    return (f"N{n} {h}x{w} Cin{cin} Cout{cout} taps{len(taps)} src{len(srcs)}", fl)


def wgrad_sig(srcs, taps, g, dw, strides, **kw):
    n, h, w, cout = g.shape
    cin = srcs[0].shape[3]
    return (f"N{n} {h}x{w} Cin{cin} Cout{cout} taps{len(taps)} src{len(srcs)}", 2.0 * n * h * w * cout * len(taps) * cin)


def pw_sig(*a, **kw):
    t = next(x for x in a if isinstance(x, torch.Tensor))
    return (str(tuple(t.shape)), 0.0)


wrap(co, "conv_run", conv_sig)
wrap(co, "wgrad_run", wgrad_sig)
for nm in ("affine2_fwd", "affine2_bwd", "pool", "dres_bwd", "dot", "im2col3", "tanh_bwd_pack", "pack_weights",
           "adam_step", "sqnorm", "rowsq", "unfold_subpixel_wgrad"):
    wrap(K, nm, pw_sig)

_run = K.PackPlan.run


def _timed_run(self):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _run(self)
    e1.record()
    REC.append(("pack_plan", (f"{len(self.entries)} entries, {self.blocks} blocks", 0.0), e0, e1))


K.PackPlan.run = _timed_run

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
tr = DFGANTrainer(G, D, B, use_graph=False)
img = (torch.randn(B, 3, 256, 256, device="cuda") * 0.5).clamp(-1, 1)
sent, noise = torch.randn(B, 256, device="cuda") * 0.3, torch.randn(B, 100, device="cuda")
for _ in range(3):
    tr.step(img, sent, noise)
torch.cuda.synchronize()
REC.clear()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
tr.step(img, sent, noise)
t1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for name, (sig, fl), e0, e1 in REC:
    a = agg[(name, sig)]
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"step {t0.elapsed_time(t1):.2f} ms (with event overhead), sum of timed calls {tot:.2f} ms, {len(REC)} calls")
for (name, sig), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    tf = fl / ms / 1e9 if ms > 0 and fl > 0 else 0.0
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {tf:7.1f} TF/s  {name:14s} {sig}")
byname = collections.defaultdict(float)
for (name, sig), (n, ms, fl) in agg.items():
    byname[name] += ms
print({k: round(v, 2) for k, v in sorted(byname.items(), key=lambda kv: -kv[1])})
