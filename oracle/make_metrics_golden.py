"""TEST INFRASTRUCTURE.  Generates tests/golden/metrics_golden.npz by running the UNMODIFIED reference functions
compute_inception_score / compute_fid (Deep_Fusion_GAN/src/evaluation/metrics.py, imported read-only from
/root/reference) with the module's `inception_v3` factory swapped for oracle/metrics_oracle.py:StandInInception
(the pretrained torchvision weights cannot be downloaded in this container).  Build container only:

    python oracle/make_metrics_golden.py

The file holds the inputs (int8 image grids), the stand-in's weights and the reference's results: IS for 1 and 3
splits, FID(real, fake), FID(real, real) and the values for inputs containing NaN / Inf.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/Deep_Fusion_GAN")
sys.path.insert(0, ROOT)

from oracle import metrics_oracle as M  # noqa: E402

N_IMAGES, SIZE = 96, 24


class _LinalgWithDisp:
    """The reference calls `linalg.sqrtm(A, disp=False)` and unpacks `(sqrtm, error_estimate)` (:105), the SciPy
    < 1.16 interface; the SciPy of this image (see the printed version) dropped the keyword, so the UNSHIMMED
    reference returns inf from its `except` branch for every input.  This adapter restores the old calling
    convention around the same scipy routine; nothing else of the reference is touched."""

    def sqrtm(self, a, disp=True):
        from scipy import linalg
        r = linalg.sqrtm(a)
        return r if disp else (r, None)


def main() -> int:
    import scipy
    print("scipy", scipy.__version__, "torch", torch.__version__)
    import src.evaluation.metrics as ref
    w = M.stand_in_weights()
    ref.inception_v3 = lambda pretrained=True, transform_input=False: M.StandInInception(w)
    ref.linalg = _LinalgWithDisp()
    real_g = M.synthetic_images(N_IMAGES, SIZE, seed=1)
    fake_g = M.synthetic_images(N_IMAGES, SIZE, seed=2, shift=0.15)
    real, fake = M.to_images(real_g), M.to_images(fake_g)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {
        "is_fake_1": ref.compute_inception_score(fake, cuda=False, batch_size=8, splits=1),
        "is_fake_3": ref.compute_inception_score(fake, cuda=False, batch_size=8, splits=3),
        "is_real_1": ref.compute_inception_score(real, cuda=False, batch_size=8, splits=1),
        "fid_real_fake": ref.compute_fid(real, fake, cuda=False, batch_size=8),
        "fid_real_real": ref.compute_fid(real, real, cuda=False, batch_size=8),
    }
    bad = fake.clone()
    bad[0, 0, 0, 0], bad[1, 1, 2, 3], bad[2, 2, 5, 5] = float("nan"), float("inf"), float("-inf")
    out["is_bad_1"] = ref.compute_inception_score(bad, cuda=False, batch_size=8, splits=1)
    out["fid_real_bad"] = ref.compute_fid(real, bad, cuda=False, batch_size=8)
    for k, v in out.items():
        print(f"{k:16s} {v!r}")
        assert np.isfinite(v), k
    # the restatement against the reference while it is at hand
    net = M.StandInInception(w).eval()
    with torch.no_grad():
        logits = net(M.preprocess(fake))
        probs = torch.softmax(torch.clamp(logits, -50, 50), dim=1).numpy()
        net.fc = torch.nn.Identity()
        a_real, a_fake = net(M.preprocess(real)).numpy(), net(M.preprocess(fake)).numpy()
    for name, mine in (("is_fake_1", M.inception_score(probs, 1)), ("is_fake_3", M.inception_score(probs, 3)),
                       ("fid_real_fake", M.fid(a_real, a_fake))):
        rel = abs(mine - out[name]) / abs(out[name])
        print(f"oracle vs reference {name}: rel {rel:.2e}")
        assert rel < 2e-5, (name, mine, out[name])
    path = os.path.join(ROOT, "tests", "golden", "metrics_golden.npz")
    np.savez_compressed(path, real=real_g, fake=fake_g, **w, **{k: np.float64(v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
