"""TEST INFRASTRUCTURE.  Generates tests/golden/sample_golden.json by running the reference's UNMODIFIED
sample.sample / save_image / generate_images (Deep_Fusion_GAN/src/sample.py) and test_generate.preprocess_caption
(imported read-only from /root/reference) on the inputs of `sample_inputs` below, with the stand-in generator / text
encoder callables defined here (both functions treat them as opaque callables).  Build container only:

    python oracle/make_sample_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/Deep_Fusion_GAN")
sys.path.insert(0, ROOT)

SEED = 31
PROMPTS = ["A small bird, with a red head.", "this Bird has wings that are blue and has a yellow belly and a very long "
           "pointed beak and sits on a thin branch of a tree", "", "zzz unknown words only"]
WORD2CODE = {w: i for i, w in enumerate(["<end>", "<unk>", "a", "small", "bird", "with", "red", "head", "this", "has",
                                         "wings", "that", "are", "blue", "and", "yellow", "belly", "very", "long"])}


def stand_in_generator(noise, sent):
    """[B,100] x [B,E] -> [B,3,8,8] in (-1,1): depends on both inputs, so noise order and embedding order matter."""
    b = noise.shape[0]
    base = torch.tanh(noise[:, :3, None, None] + sent[:, :3, None, None])
    ramp = torch.linspace(-0.5, 0.5, 64).view(1, 1, 8, 8)
    return (base * 0.8 + ramp * 0.4).expand(b, 3, 8, 8).clone()


def stand_in_encoder(captions, cap_lens):
    return (captions.float().cumsum(1)[:, -4:] / 50.0) * cap_lens[:, None].float() / 18.0


def sample_inputs():
    g = torch.Generator().manual_seed(5)
    images = torch.zeros(3, 3, 8, 8)
    captions = torch.randint(1, 19, (3, 18, 1), generator=g)
    lens = torch.tensor([7, 18, 11])
    return images, captions, lens, ("001.A/bird_1", "002.B/bird_2", "plain")


def png_digest(path):
    from PIL import Image
    a = np.array(Image.open(path))
    return {"shape": list(a.shape), "sha1": hashlib.sha1(a.tobytes()).hexdigest()}


def run(sample_mod, preprocess):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(SEED)
        sample_mod.sample(stand_in_generator, stand_in_encoder, sample_inputs(), os.path.join(tmp, "out"))
        out["files"] = {f: png_digest(os.path.join(tmp, "out", f)) for f in sorted(os.listdir(os.path.join(tmp, "out")))}
    out["prompts"] = []
    for p in PROMPTS:
        caps, lens = preprocess(p, WORD2CODE)
        out["prompts"].append({"captions": caps.flatten().tolist(), "shape": list(caps.shape), "len": lens.tolist()})
    return out


def main() -> int:
    import src.sample as ref_sample
    from src.test_generate import preprocess_caption
    gold = run(ref_sample, preprocess_caption)
    path = os.path.join(ROOT, "tests", "golden", "sample_golden.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1)
    print(json.dumps(gold)[:600])
    print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
