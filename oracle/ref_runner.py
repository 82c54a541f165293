"""TEST INFRASTRUCTURE -- not a product path.  Drives the UNMODIFIED reference trainer (`oracle/_ref`, vendored by
oracle/vendor_ref.py) through its own public API: `DeepFusionGAN(...).fit(loader, num_epochs=1)`
(Deep_Fusion_GAN/src/deep_fusion_gan/model.py:135-232), i.e. the reference's own loop body -- prepare_data, text
encoder, D hinge step, `_compute_gp` double backward, G step, clip_grad_norm_, torch.optim.Adam, the `.item()`
synchronisations -- on synthetic batches of BASELINE.json's shape.  Used by bench.py only: `--impl reference` /
`cpu_baseline` (host cores) and `library_baseline` (the same reference under stock PyTorch-CUDA: the bar the
hand-written kernels must beat).

How the iteration is timed without editing the reference: the *loader* is ours.  `_TimedLoader` stamps the clock
every time `fit`'s `for batch in train_loader` asks for the next batch (after a device synchronise on CUDA), so the
interval between request W and request W+K is exactly K iterations of the reference's loop; the second pass over the
loader (the epoch-end evaluation, :238-262 -- outside the path) raises `_EndOfMeasurement`, which ends `fit`.
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile
import time
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NTOKEN = 50


class _EndOfMeasurement(Exception):
    pass


class _TimedLoader:
    def __init__(self, batches: list, sync):
        self.batches, self.sync, self.stamps, self.passes = batches, sync, [], 0

    def _mark(self):
        self.sync()
        self.stamps.append(time.perf_counter())

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        self.passes += 1
        if self.passes > 1:
            raise _EndOfMeasurement()
        return self._gen()

    def _gen(self):
        for b in self.batches:
            self._mark()
            yield b
        self._mark()


def available() -> bool:
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import vendor_ref
    return vendor_ref.available()


def _import_reference():
    ref = os.path.join(HERE, "_ref")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import src.deep_fusion_gan.model as ref_model   # the reference's own module tree
    return ref_model


MODES = ("fp32", "tf32", "bf16_autocast_channels_last", "bf16_autocast_channels_last_cudnn_benchmark")


def run_fit(device: str, batch: int, steps: int, warmup: int, mode: str = "fp32", threads: int = 0) -> Dict:
    """K = `steps` timed iterations of the reference's `fit` loop after `warmup` untimed ones, on `device`
    ("cpu" | "cuda").  mode (CUDA): "fp32" = TF32 off everywhere (north_star's fp32 tier), "tf32" = TF32 on for
    cuDNN and cuBLAS, "bf16_autocast_channels_last" = the reference modules under torch.autocast(bf16) with
    channels_last parameters, "..._cudnn_benchmark" = the same with cuDNN's autotuner on (warm-up absorbs the
    search) -- all are switches of stock PyTorch, the reference's code is untouched."""
    import torch
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import dfgan_oracle as O
    from oracle import text_encoder_oracle as T
    assert mode in MODES and device in ("cpu", "cuda")
    ref_model = _import_reference()
    if threads:
        torch.set_num_threads(threads)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.cuda.is_available,
             torch.backends.cudnn.benchmark)
    tf32 = mode == "tf32"
    torch.backends.cudnn.benchmark = mode.endswith("cudnn_benchmark")
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        with tempfile.TemporaryDirectory() as tmp:
            enc = os.path.join(tmp, "text_encoder.pth")
            torch.save(T.init_state(NTOKEN), enc)
            img_dir, gen_dir = os.path.join(tmp, "gen_images"), os.path.join(tmp, "gen_weights")
            os.makedirs(img_dir), os.makedirs(gen_dir)
            if device == "cpu":
                # the reference picks its device in the constructor (:27); hide the GPU from that one decision
                torch.cuda.is_available = lambda: False
            try:
                torch.manual_seed(O.SEED)
                model = ref_model.DeepFusionGAN(NTOKEN, enc, img_dir, gen_dir)
            finally:
                torch.cuda.is_available = saved[2]
            assert model.device.type == device, (model.device, device)
            ctx = contextlib.nullcontext()
            if mode.startswith("bf16_autocast_channels_last"):
                model.generator.to(memory_format=torch.channels_last)
                model.discriminator.to(memory_format=torch.channels_last)
                ctx = torch.autocast(device_type=device, dtype=torch.bfloat16)
            images, _, _ = O.synthetic_batch(batch, 256)
            caps, lens = T.synthetic_captions(batch, NTOKEN)
            one = (images, caps.unsqueeze(-1), lens, tuple(f"img_{i}" for i in range(batch)))
            sync = (lambda: torch.cuda.synchronize()) if device == "cuda" else (lambda: None)
            loader = _TimedLoader([one] * (warmup + steps), sync)
            t0 = time.perf_counter()
            try:
                # the reference prints progress lines on stdout; bench.py's stdout carries ONE JSON line
                with ctx, contextlib.redirect_stdout(sys.stderr):
                    model.fit(loader, test_loader=None, num_epochs=1, auto_resume=False)
            except _EndOfMeasurement:
                pass
            wall = time.perf_counter() - t0
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved[0], saved[1]
        torch.backends.cudnn.benchmark = saved[3]
    st: List[float] = loader.stamps
    assert len(st) == warmup + steps + 1, (len(st), warmup, steps)
    total = st[-1] - st[warmup]
    per = [st[i + 1] - st[i] for i in range(warmup, warmup + steps)]
    return {"images_per_s": batch * steps / total, "seconds": total, "ms_per_step": 1e3 * total / steps,
            "best_ms": 1e3 * min(per), "batch": batch, "steps": steps, "warmup": warmup, "mode": mode,
            "device": device, "wall_s": wall,
            "what": "unmodified reference DeepFusionGAN.fit loop (oracle/_ref), synthetic loader"}


if __name__ == "__main__":
    import json
    dev = sys.argv[1] if len(sys.argv) > 1 else "cpu"
    print(json.dumps(run_fit(dev, int(sys.argv[2]) if len(sys.argv) > 2 else 2, 2, 1,
                             sys.argv[3] if len(sys.argv) > 3 else "fp32")))
