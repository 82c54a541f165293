"""Throughput of the trainer shell's training loop against the bare fused iteration (round-2 starting point).

    python tools/prof_shell.py [--batch 32] [--iters 40] [--decoded 0|1]

Prints one JSON line: images/s of (a) `DFGANTrainer.step` on device-resident inputs (bench.py's `value`), (b) the loop
body of `DeepFusionGAN.fit` -- batch preparation from HOST tensors, text encoder, noise draw, step, loss accumulation,
one read at the end -- on an in-memory loader, and (c) with --decoded 1 the same loop fed by
`dataset.DFGANDataset.collate` on synthetic decoded uint8 images (500x375, CUB-like boxes), i.e. including the batched
GPU image transform.  (b)/(a) is what the shell costs; (c) adds the input pipeline.  Written after round 1's GPU budget
was spent: NOT YET RUN ON A B200.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--decoded", type=int, default=1)
    ap.add_argument("--overlap", type=int, default=0, help="1: DFGAN_STAGE_MIXED (images staged by the trainer's copy stream)")
    a = ap.parse_args()
    if a.overlap:
        os.environ["DFGAN_STAGE_MIXED"] = "1"
    import torch
    warnings.filterwarnings("ignore", message="dropout option adds dropout")
    from gan_codes_b200.deep_fusion_gan import DeepFusionGAN
    from gan_codes_b200.text_encoder import RNNEncoder
    B, n_words = a.batch, 5450
    torch.manual_seed(123321)
    gen = torch.Generator().manual_seed(1)
    out = {"batch": B, "iters": a.iters, "overlap": a.overlap}
    with tempfile.TemporaryDirectory() as tmp:
        enc = os.path.join(tmp, "text_encoder.pth")
        torch.save(RNNEncoder(n_words, nhidden=256).state_dict(), enc)
        model = DeepFusionGAN(n_words, enc, tmp, tmp)
        lens = torch.randint(5, 19, (B,), generator=gen)
        caps = torch.randint(1, n_words, (B, 18, 1), generator=gen) * (torch.arange(18)[None, :, None] < lens[:, None, None])
        names = [f"img_{i}" for i in range(B)]
        images = (torch.randn(B, 3, 256, 256, generator=gen) * 0.5).clamp_(-1, 1)
        batch = (images, caps, lens, names)

        def loop(batches, n):
            """The training-loop body of DeepFusionGAN.fit (deep_fusion_gan.py), timed on the host around a final
            synchronising read, like an epoch."""
            sums = None
            t0 = time.perf_counter()
            for i in range(n):
                imgs, c, cl, _ = model._prepare(batches(i), overlap=model._overlap_h2d)
                tr = model._ensure_trainer(imgs.shape[0])
                sent = model.text_encoder(c, cl).detach()
                o = model._step(tr, imgs, sent, model._draw_noise(imgs.shape[0]))
                row = torch.stack([o[k] for k in ("g_loss", "d_loss", "gp_loss", "txtimg_loss")]).double()
                sums = row if sums is None else sums + row
            means = (sums / n).tolist()
            return time.perf_counter() - t0, means

        loop(lambda i: batch, 5)                         # eager warm-up + graph capture
        dt, means = loop(lambda i: batch, a.iters)
        out["shell_loop_images_per_s"] = round(B * a.iters / dt, 1)
        out["losses"] = [round(x, 5) for x in means]
        # bare iteration on resident inputs
        tr = model.trainer
        img_d, sent_d, noise_d = images.cuda(), (torch.randn(B, 256, generator=gen) * 0.3).cuda(), torch.randn(B, 100, generator=gen).cuda()
        for _ in range(3):
            tr.step(img_d, sent_d, noise_d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            tr.step(img_d, sent_d, noise_d)
        e1.record()
        torch.cuda.synchronize()
        out["bare_step_images_per_s"] = round(B * a.iters / (e0.elapsed_time(e1) / 1e3), 1)
        if a.decoded:
            from gan_codes_b200.augment import BatchImageTransform
            from gan_codes_b200.dataset import DFGANDataset
            ds = DFGANDataset.__new__(DFGANDataset)       # the collate only needs the transform
            ds.transform = BatchImageTransform(256)
            decoded = [torch.randint(0, 256, (375, 500, 3), generator=gen, dtype=torch.uint8) for _ in range(B)]
            samples = [(decoded[i], [60, 40, 380, 300], caps[i].numpy(), int(lens[i]), names[i]) for i in range(B)]
            loop(lambda i: ds.collate(samples), 3)
            dt, _ = loop(lambda i: ds.collate(samples), a.iters)
            out["shell_loop_with_gpu_transform_images_per_s"] = round(B * a.iters / dt, 1)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
