"""Parity at BASELINE.json's FULL sizes through size-independent properties (configs[3]: MA-GP discriminator step at
batch 64; configs[4]: generator sampling at batch 256), where the CPU oracle cannot run the whole batch in seconds.

    python tools/bringup_configs.py [case ...]

    sample_b256   every image of a 256-image `trainer.sample` batch equals the image the same (noise, sentence) row gets
                  in a 4-image batch (the forward has no cross-sample term; since round 2 the low-resolution layers of
                  the SMALL batch split their K loops over a thread-block cluster, so the fp32 partial sums are formed
                  in a different order and bf16-stored activations differ in the last bit: 1e-2 on the [-1, 1] image
                  instead of bit-exact), and the first 4 match the fp32 oracle to the bf16 tolerance
    gp_step_b64   MA-GP at batch 64: the loss is the mean of the per-sample penalties, so it equals the mean over 16
                  chunks of 4 of the batch-4 program's loss (four of which are checked against the ORACLE); the flat
                  gradient equals the mean of the 16 chunk gradients (linearity in the batch), and `gp_step` applies
                  exactly one clipped Adam update of it

First B200 run: round 2 (profiles/r02_configs_b64_b256.log, all rows green); part of `pytest -m gpu`
(tests/test_gpu_parity.py::test_full_size_properties).  tools/bringup_parity.py holds the direct comparisons with the
oracle at the same sizes.
"""
from __future__ import annotations

import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _report(name, err, tol):
    bad = not (err < tol)
    print(f"  {name:60s} {err:.3e} (tol {tol:.0e}){'   <-- FAIL' if bad else ''}")
    return not bad


def _modules(gated=True, ws=1.0):
    import torch
    from gan_codes_b200.modules import Discriminator, Generator
    from oracle import dfgan_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gr = (0.2, 1.0) if gated else None
    g_sd = O.init_generator_state(32, 100, gamma_range=gr)
    d_sd = O.init_discriminator_state(32, gamma_range=gr, weight_scale=ws)
    G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    return torch, O, g_sd, d_sd, G, D


@case
def sample_b256():
    torch, O, g_sd, d_sd, G, D = _modules()
    from gan_codes_b200.engine import DFGANTrainer
    _, sent, noise = O.synthetic_batch(256, 8, seed=77)            # images unused: size 8 keeps the draw cheap
    big = DFGANTrainer(G, D, 256, use_graph=False)
    out = big.sample(noise.cuda(), sent.cuda())
    ok = tuple(out.shape) == (256, 3, 256, 256) and bool(torch.isfinite(out).all())
    worst = 0.0
    for lo in (0, 124, 252):
        part = big.sample(noise[lo:lo + 4].cuda(), sent[lo:lo + 4].cuda())       # batch-4 program, same weights
        worst = max(worst, float((part - out[lo:lo + 4]).abs().max()))
    ok &= _report("rows of the 256-batch vs the same rows as a 4-batch (max abs)", worst, 1e-2)
    with torch.no_grad():
        ref = O.generator_forward(g_sd, noise[:4], sent[:4])
    err = float((out[:4].cpu() - ref).norm() / ref.norm())
    ok &= _report("first 4 images vs fp32 oracle (rel)", err, 2e-2)
    return bool(ok)


@case
def gp_step_b64():
    torch, O, g_sd, d_sd, G, D = _modules(ws=2.2)                   # O(1e-2 .. 1) penalties, see make_shell_golden.py
    from gan_codes_b200.engine import DFGANTrainer
    images, sent, _ = O.synthetic_batch(64, 256, seed=78)
    img_d, sent_d = images.cuda(), sent.cuda()
    big = DFGANTrainer(G, D, 64, use_graph=False)
    G4, D4 = copy.deepcopy(G), copy.deepcopy(D)
    small = DFGANTrainer(G4, D4, 4, use_graph=False)
    # (1) loss and flat gradient of the batch-64 program (no update yet)
    with torch.no_grad():
        loss64 = float(big._gp_step(img_d, sent_d))
        grad64 = big.ds.grad.clone()
        # (2) the same from 16 chunks of the batch-4 program, and the oracle's chunk losses
        acc, losses4, ref_losses = torch.zeros_like(grad64), [], []
        for c in range(16):
            sl = slice(4 * c, 4 * c + 4)
            losses4.append(float(small._gp_step(img_d[sl], sent_d[sl])))
            acc += small.ds.grad
    for c in range(0, 16, 5):                                       # 4 of the 16 chunks on the CPU oracle (~1 s each)
        sl = slice(4 * c, 4 * c + 4)
        ref_losses.append((c, float(O.magp_loss(O._leaf(d_sd), images[sl], sent[sl]).detach())))
    mean4 = sum(losses4) / 16
    # (different K-loop splits at batch 4 and batch 64: bf16 last-bit differences, amplified by n^6 / n^4)
    ok = _report("loss(64) vs mean of 16 batch-4 losses (rel)", abs(loss64 - mean4) / abs(mean4), 1e-2)
    for c, r in ref_losses:
        ok &= _report(f"chunk {c}: batch-4 loss vs fp32 oracle (rel)", abs(losses4[c] - r) / abs(r), 1e-1)
    ok &= _report("grad(64) vs mean of 16 chunk gradients (rel L2)",
                  float((grad64 - acc / 16).norm() / (acc / 16).norm()), 2e-2)
    # (3) gp_step = exactly one clipped Adam update with that gradient: |delta| <= lr / sqrt(1 - beta2) element-wise,
    #     sign(delta) = -sign(grad) wherever the gradient is not rounding noise, step counter 1
    before = big.ds.flat.clone()
    big.gp_step(img_d, sent_d)
    torch.cuda.synchronize()
    delta = big.ds.flat - before
    ok &= int(big.ds.step_count.item()) == 1
    ok &= _report("max |update| / lr", float(delta.abs().max()) / 4e-4, 3.17)
    strong = grad64.abs() > 1e-3 * grad64.abs().max()
    agree = float((torch.sign(delta[strong]) == -torch.sign(grad64[strong])).float().mean())
    ok &= _report("1 - fraction of strong-gradient elements moving against the gradient", 1.0 - agree, 1e-3)
    return bool(ok)


def main(argv):
    names = argv or list(CASES)
    bad = []
    for n in names:
        print(f"[{n}]")
        if not CASES[n]():
            bad.append(n)
    print("FAILED: " + ", ".join(bad) if bad else "all config cases passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
