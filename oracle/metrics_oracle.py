"""TEST INFRASTRUCTURE -- not a product path.  Only tests/, tools/bringup_shell.py (the bodies of the GPU tests) and the
golden generators oracle/make_metrics_golden.py / make_shell_golden.py import this module.

CPU restatement (numpy + scipy, the libraries the reference uses) of the statistics of
Deep_Fusion_GAN/src/evaluation/metrics.py:

    inception_score(probs, splits)      :44-57   exp(mean_i sum_y p(y|x_i) (log p(y|x_i) - log p(y))), +1e-10 guards,
                                                 averaged over `splits` contiguous parts
    fid(act1, act2)                     :97-108  |mu1-mu2|^2 + Tr(S1 + S2 - 2 sqrtm(S1 S2)), np.cov(rowvar=False),
                                                 scipy.linalg.sqrtm, imaginary part discarded
    preprocess(images)                  :25-28   (x+1)/2, clamp to [0,1], bilinear 299x299 (align_corners=False)

and `StandInInception`, a small deterministic network with the two properties the reference relies on (an `fc`
layer that can be replaced by Identity, an `aux_logits` attribute): torchvision's pretrained Inception-v3 cannot be
downloaded here, so the parity pin swaps the factory `inception_v3` of the UNMODIFIED reference module for this
stand-in (oracle/make_metrics_golden.py) -- every line of the reference's own pre-processing, batching, softmax,
IS and FID arithmetic runs as shipped, only the feature extractor is substituted.  Its outputs are committed as
tests/golden/metrics_golden.npz; tests/test_evaluation_cpu.py checks this module and the product
(gan_codes_b200/evaluation.py) against that file.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

N_FEAT, N_CLASS = 24, 40


class StandInInception(nn.Module):
    """[n,3,299,299] in [0,1] -> logits [n, N_CLASS]; pooled N_FEAT-vector is the input of `fc`."""

    def __init__(self, weights: Dict[str, np.ndarray]):
        super().__init__()
        # layer constructors draw their default init from the global generator; keep the stand-in RNG-neutral so
        # that building it (the reference does so twice per epoch, evaluation/metrics.py:12, 67) does not shift the
        # noise stream of the training loop around it
        with torch.random.fork_rng(devices=[]):
            self.conv = nn.Conv2d(3, N_FEAT, 7, stride=4)
            self.fc = nn.Linear(N_FEAT, N_CLASS)
        self.aux_logits = True
        with torch.no_grad():
            self.conv.weight.copy_(torch.from_numpy(weights["conv_w"]))
            self.conv.bias.copy_(torch.from_numpy(weights["conv_b"]))
            self.fc.weight.copy_(torch.from_numpy(weights["fc_w"]))
            self.fc.bias.copy_(torch.from_numpy(weights["fc_b"]))

    def forward(self, x):
        h = torch.tanh(self.conv(x))
        # mean and a second moment of the feature maps, so the pooled vector is not a linear image statistic
        h = F.adaptive_avg_pool2d(h, 1).flatten(1) + 0.5 * F.adaptive_avg_pool2d(h * h, 1).flatten(1)
        return self.fc(h)


def stand_in_weights(seed: int = 11) -> Dict[str, np.ndarray]:
    r = np.random.RandomState(seed)
    return {"conv_w": (r.standard_normal((N_FEAT, 3, 7, 7)) * 0.6).astype(np.float32),
            "conv_b": (r.standard_normal(N_FEAT) * 0.3).astype(np.float32),
            "fc_w": (r.standard_normal((N_CLASS, N_FEAT)) * 4.0).astype(np.float32),
            "fc_b": (r.standard_normal(N_CLASS) * 0.5).astype(np.float32)}


def synthetic_images(n: int, size: int, seed: int, shift: float = 0.0) -> np.ndarray:
    """int8 grid values v; the image is v / 64 + shift -- exactly representable, a few per cent outside [-1, 1]
    so the clamp of :27 is exercised.  Smooth blobs + noise, so that the stand-in features vary across images."""
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size] / size
    imgs = np.empty((n, 3, size, size), np.float64)
    for i in range(n):
        for c in range(3):
            fx, fy, ph = r.uniform(0.5, 3.0), r.uniform(0.5, 3.0), r.uniform(0, 6.28)
            imgs[i, c] = r.uniform(0.3, 1.1) * np.sin(6.28 * (fx * xx + fy * yy) + ph) + r.uniform(-0.3, 0.3)
    imgs += r.standard_normal(imgs.shape) * 0.1 + shift
    return np.clip(np.round(imgs * 64), -127, 127).astype(np.int8)


def to_images(grid: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(grid.astype(np.float32) / 64.0)


def preprocess(images: torch.Tensor) -> torch.Tensor:
    batch = torch.clamp((images + 1) / 2, min=0, max=1)
    return F.interpolate(batch, size=(299, 299), mode="bilinear", align_corners=False)


def inception_score(probs: np.ndarray, splits: int = 1) -> float:
    n = probs.shape[0]
    out = []
    for k in range(splits):
        part = probs[k * (n // splits):(k + 1) * (n // splits), :].astype(np.float64)
        py = part.mean(axis=0) + 1e-10
        kl = [np.sum((p + 1e-10) * (np.log(p + 1e-10) - np.log(py))) for p in part]
        out.append(np.exp(np.mean(kl)))
    return float(np.mean(out))


def fid(act1: np.ndarray, act2: np.ndarray) -> float:
    from scipy import linalg
    a1, a2 = act1.astype(np.float64), act2.astype(np.float64)
    mu1, s1 = a1.mean(axis=0), np.cov(a1, rowvar=False)
    mu2, s2 = a2.mean(axis=0), np.cov(a2, rowvar=False)
    covmean = linalg.sqrtm(s1.dot(s2))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    d = mu1 - mu2
    return float(d.dot(d) + np.trace(s1 + s2 - 2 * covmean))
