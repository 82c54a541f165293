"""Per-epoch Inception Score / FID (evaluation/metrics.py:7-118), resident and batched.

Same two entry points and signatures as the reference -- `compute_inception_score(images, cuda, batch_size,
splits)` and `compute_fid(real_images, fake_images, cuda, batch_size)` -- with the same pre-processing
([-1,1] -> [0,1], clamp, bilinear 299x299, logits clamped to +-50 before the softmax), the same NaN handling and
the same failure values (1.0 / inf).  What changes is where the work happens:

  * the reference builds a fresh Inception-v3 inside each function (twice per epoch, :12 and :67) and moves every
    batch of activations to numpy; here ONE network per device stays resident (`inception()`), and the pooled
    2048-vector is taken from the input of its `fc` layer, so a single pass over the generated images yields both
    the class probabilities (IS) and the features (FID): `compute_is_fid`;
  * the statistics stay on the device in fp64: IS as one fused KL reduction over the probability matrix, FID mean
    / covariance as one GEMM, and Tr sqrt(S1 S2) from two symmetric eigen-decompositions
    (Tr (S1 S2)^(1/2) = sum_i sqrt(lambda_i(S1^(1/2) S2 S1^(1/2)))) instead of scipy's Schur-based `sqrtm` of the
    non-symmetric product on the host (:105).  For full-rank covariances both agree to rounding (tests: 1e-6); when
    there are fewer images than feature dimensions (the trainer's 768 images against 2048 features, deep_fusion_gan/
    model.py:246) the product is singular, scipy's result is then ill-defined (it warns and returns a matrix with a
    discarded imaginary part) while the symmetric form stays exact -- documented difference, DESIGN section 6.

Inception-v3 itself is torchvision library code (`torchvision.models.inception_v3`, weights fetched by torchvision:
not available offline -- the functions then fail like the reference's, returning 1.0 / inf); any module with an
`fc` layer can be injected with `set_inception_factory` (the tests pin the arithmetic against the reference's own
functions that way).  Like the reference these functions run wherever the images can go (`cuda=True` falls back to
the CPU when there is no GPU, :11, :64): they are torch library calls end to end, none of the repo's kernels.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_FACTORY: Optional[Callable[[], nn.Module]] = None
_RESIDENT: Dict[Tuple[str, int], nn.Module] = {}


def set_inception_factory(factory: Optional[Callable[[], nn.Module]]) -> None:
    """Use `factory()` instead of torchvision's pretrained Inception-v3 (None restores the default).  The module
    must map [n,3,299,299] images in [0,1] to logits and own an `fc` layer whose input is the pooled feature."""
    global _FACTORY
    _FACTORY = factory
    _RESIDENT.clear()


def _default_factory() -> nn.Module:
    from torchvision.models import inception_v3
    return inception_v3(pretrained=True, transform_input=False)     # evaluation/metrics.py:12, 67


class _Resident(nn.Module):
    """Inception-v3 in eval mode + a hook on `fc` that keeps the pooled features of the last forward."""

    def __init__(self, net: nn.Module):
        super().__init__()
        if not hasattr(net, "fc"):
            raise AttributeError("the inception network needs an `fc` layer (evaluation/metrics.py:70 replaces it)")
        self.net = net.eval()
        self.pooled: Optional[torch.Tensor] = None
        net.fc.register_forward_pre_hook(self._keep)

    def _keep(self, _module, args):
        self.pooled = args[0]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        out = self.net(x)
        logits = out[0] if isinstance(out, tuple) else getattr(out, "logits", out)
        pooled, self.pooled = self.pooled, None
        return logits, pooled.reshape(pooled.shape[0], -1)


def inception(device: torch.device) -> _Resident:
    """The resident network of `device` (built on first use)."""
    key = (device.type, device.index or 0)
    if key not in _RESIDENT:
        _RESIDENT[key] = _Resident((_FACTORY or _default_factory)().to(device))
    return _RESIDENT[key]


def _device(cuda: bool) -> torch.device:
    return torch.device("cuda" if cuda and torch.cuda.is_available() else "cpu")


def _sanitise(images: torch.Tensor, what: str) -> torch.Tensor:
    # :19-22, :78-80 -- one fused finiteness test instead of two host-synchronising ones
    if not bool(torch.isfinite(images).all()):
        print(f"Warning: Input images contain NaN or Inf values{what}")
        images = torch.nan_to_num(images, nan=0.0, posinf=1.0, neginf=-1.0)
    return images


@torch.no_grad()
def inception_outputs(images: torch.Tensor, device: torch.device, batch_size: int = 8,
                      want_probs: bool = True, want_pooled: bool = True):
    """(class probabilities [N,classes] fp32 or None, pooled features [N,D] fp32 or None), both on `device`."""
    model = inception(device)
    probs, pooled = [], []
    for i in range(0, images.size(0), batch_size):
        batch = images[i:i + batch_size].to(device, torch.float32, non_blocking=True)
        batch = torch.clamp((batch + 1) / 2, min=0, max=1)                              # :25-27, :84-86
        batch = F.interpolate(batch, size=(299, 299), mode="bilinear", align_corners=False)
        logits, feat = model(batch)
        if want_probs:
            probs.append(F.softmax(torch.clamp(logits, min=-50, max=50), dim=1))         # :33-35
        if want_pooled:
            pooled.append(feat)
    return (torch.cat(probs) if want_probs else None), (torch.cat(pooled) if want_pooled else None)


def inception_score_from_probs(probs: torch.Tensor, splits: int = 1) -> float:
    """exp(mean_i KL(p(y|x_i) || p(y))) averaged over `splits` contiguous parts (:46-57), fp64 on the device."""
    n = probs.shape[0]
    p = probs.double()
    bad = ~torch.isfinite(p)
    if bool(bad.any()):                                                                  # :38-40
        print("Warning: NaN or Inf values detected in IS probabilities")
        p = torch.where(bad, torch.full_like(p, 1.0 / 1000), p)
    scores = []
    for k in range(splits):
        part = p[k * (n // splits):(k + 1) * (n // splits)]
        py = part.mean(dim=0, keepdim=True) + 1e-10
        pyx = part + 1e-10
        kl = (pyx * (pyx.log() - py.log())).sum(dim=1)
        scores.append(kl.mean().exp())
    return float(torch.stack(scores).mean())


def _mean_cov(act: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    a = act.double()
    mu = a.mean(dim=0)
    c = a - mu
    return mu, (c.t() @ c) / (a.shape[0] - 1)                                            # np.cov(rowvar=False), :100


def _sqrt_psd(s: torch.Tensor) -> torch.Tensor:
    w, u = torch.linalg.eigh(s)
    return (u * w.clamp_min(0).sqrt()) @ u.t()


def fid_from_activations(act1: torch.Tensor, act2: torch.Tensor) -> float:
    """|mu1 - mu2|^2 + Tr(S1 + S2 - 2 (S1 S2)^(1/2)) (:97-108) in fp64 on the device of the activations."""
    if not bool(torch.isfinite(act1).all() & torch.isfinite(act2).all()):                # :94-96
        print("Warning: Activations contain NaN values")
        return float("inf")
    mu1, s1 = _mean_cov(act1)
    mu2, s2 = _mean_cov(act2)
    r = _sqrt_psd(s1)
    m = r @ s2 @ r
    tr_covmean = torch.linalg.eigvalsh((m + m.t()) * 0.5).clamp_min(0).sqrt().sum()
    diff = mu1 - mu2
    return float(diff.dot(diff) + torch.trace(s1) + torch.trace(s2) - 2 * tr_covmean)


@torch.no_grad()
def compute_inception_score(images, cuda=True, batch_size=8, splits=1):
    """evaluation/metrics.py:7-61.  images: (N, 3, H, W) in [-1, 1]."""
    try:
        device = _device(cuda)
        images = _sanitise(images, " for IS calculation")
        probs, _ = inception_outputs(images, device, batch_size, want_pooled=False)
        return inception_score_from_probs(probs, splits)
    except Exception as e:                                                               # :59-61
        print(f"Error in IS calculation: {e}")
        return 1.0


@torch.no_grad()
def compute_fid(real_images, fake_images, cuda=True, batch_size=8):
    """evaluation/metrics.py:63-118."""
    try:
        device = _device(cuda)
        _, act1 = inception_outputs(_sanitise(real_images, ""), device, batch_size, want_probs=False)
        _, act2 = inception_outputs(_sanitise(fake_images, ""), device, batch_size, want_probs=False)
        try:
            return fid_from_activations(act1, act2)
        except Exception as e:                                                           # :110-112
            print(f"Error in FID calculation: {e}")
            return float("inf")
    except Exception as e:                                                               # :113-115
        print(f"Exception in FID calculation: {e}")
        return float("inf")


@torch.no_grad()
def compute_is_fid(fake_images, real_images, cuda=True, batch_size=64, splits=1) -> Tuple[float, float]:
    """Both metrics of DeepFusionGAN.compute_is_fid (deep_fusion_gan/model.py:91-97) from ONE pass over the
    generated images and one over the real ones (the reference makes three passes with two freshly built
    networks).  Failure values as in the two functions above."""
    is_score, fid = 1.0, float("inf")
    try:
        device = _device(cuda)
        probs, act_fake = inception_outputs(_sanitise(fake_images.float(), ""), device, batch_size)
        is_score = inception_score_from_probs(probs, splits)
        _, act_real = inception_outputs(_sanitise(real_images.float(), ""), device, batch_size, want_probs=False)
        fid = fid_from_activations(act_real, act_fake)
    except Exception as e:
        print(f"Error computing metrics: {e}")
    return is_score, fid
