"""Image side of the input pipeline (objects/dataset.py:112-129, 28-31, utils.py:15-19), CPU side:
the integer-exact oracle against (a) the committed outputs of the reference's own `_get_image` and (b) the stock
torchvision / Pillow pipeline run on this box over random geometry, the host-side geometry and random-draw order of the
product module, and the C entry points' argument validation (no compute without a GPU)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
import torch

from gan_codes_b200 import _lib, augment as P
from gan_codes_b200._lib_decl import AugItem
from oracle import augment_oracle as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "augment_golden.npz"))
LUT = ((torch.arange(256, dtype=torch.uint8).float().div(255)) - 0.5) / 0.5


def golden_case(i):
    w, h = (int(v) for v in G[f"size{i}"])
    img = A.synthetic_image(w, h, i)
    assert hashlib.sha1(img.tobytes()).hexdigest() == str(G[f"img_sha1_{i}"]), "synthetic image differs on this box"
    top, left, flip = (int(v) for v in G[f"params{i}"])
    want = LUT[torch.from_numpy(G[f"levels{i}"]).long()].numpy()
    return img, [int(v) for v in G[f"bbox{i}"]], int(G[f"imsize{i}"]), int(G[f"seed{i}"]), (top, left, bool(flip)), want


@pytest.mark.parametrize("i", range(int(G["n"])))
def test_oracle_reproduces_the_reference_outputs(i):
    img, bbox, imsize, seed, (top, left, flip), want = golden_case(i)
    got = A.transform_image(img, bbox, imsize, top, left, flip)
    assert got.dtype == np.float32 and got.shape == (3, imsize, imsize)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("i", range(int(G["n"])))
def test_product_geometry_and_random_draw_order(i):
    """BatchImageTransform.plan's host arithmetic: same rectangle / resized size as the oracle, and with the seed of
    the golden run the torch generator yields the crop and flip the reference's transform drew."""
    img, bbox, imsize, seed, params, _ = golden_case(i)
    h, w = img.shape[:2]
    x1, y1, x2, y2 = P.bbox_crop_rect(w, h, bbox)
    assert (x1, y1, x2, y2) == A.bbox_crop_rect(w, h, bbox)
    rw, rh = P.resized_size(x2 - x1, y2 - y1, imsize)
    assert (rw, rh) == A.resized_size(x2 - x1, y2 - y1, imsize)
    torch.manual_seed(seed)
    assert P.draw_crop_flip(rw, rh, imsize) == params


def test_oracle_matches_stock_pillow_over_random_geometry():
    import torchvision.transforms as transforms
    from PIL import Image
    rng = np.random.default_rng(5)
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    for it in range(25):
        imsize = [64, 32, 96][it % 3]
        h, w = int(rng.integers(40, 330)), int(rng.integers(40, 330))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if it % 2 else A.synthetic_image(w, h, 100 + it)
        bw, bh = int(rng.integers(20, w + 1)), int(rng.integers(20, h + 1))
        bbox = [int(rng.integers(0, w - bw + 1)), int(rng.integers(0, h - bh + 1)), bw, bh]
        x1, y1, x2, y2 = A.bbox_crop_rect(w, h, bbox)
        tf = transforms.Compose([transforms.Resize(int(imsize * 76 / 64)), transforms.RandomCrop(imsize),
                                 transforms.RandomHorizontalFlip()])
        torch.manual_seed(it)
        want = norm(tf(Image.fromarray(img).crop((x1, y1, x2, y2)))).numpy()
        torch.manual_seed(it)
        rw, rh = P.resized_size(x2 - x1, y2 - y1, imsize)
        top, left, flip = P.draw_crop_flip(rw, rh, imsize)
        assert np.array_equal(A.transform_image(img, bbox, imsize, top, left, flip), want), (it, w, h, bbox)


def test_resample_coefficients_are_normalised_fixed_point():
    for in_size, out_size in ((390, 329), (58, 91), (400, 50), (76, 76), (1, 7)):
        lo, cnt, kk = A.resample_coeffs(in_size, out_size)
        assert (lo >= 0).all() and (lo + cnt <= in_size).all() and (cnt >= 1).all()
        assert np.abs(kk.sum(1) - (1 << A.PRECISION_BITS)).max() <= kk.shape[1]        # rounding of each tap
        assert kk.shape[1] == int(np.ceil(max(in_size / out_size, 1.0))) * 2 + 1


def test_c_entry_points_validate_arguments():
    l = _lib.lib()
    assert l.dfgan_augment_workspace_bytes(32, 256) == 4 * 32 * 2 * 256 * 26
    ok = AugItem(src=4096, pitch=3 * 500, src_w=500, src_h=375, x1=110, y1=60, cw=260, ch=210, rw=376, rh=304, top=3,
                 left=7, flip=1)
    assert l.dfgan_augment_check(C.byref(ok), 256) == 0

    def bad(**kw):
        it = AugItem.from_buffer_copy(ok)
        for k, v in kw.items():
            setattr(it, k, v)
        return l.dfgan_augment_check(C.byref(it), 256)
    assert bad(src=None) == -1 and bad(pitch=100) == -1
    assert bad(x1=300) == -1 and bad(ch=400) == -1 and bad(cw=0) == -1          # rectangle outside the image
    assert bad(rh=255) == -2                                                     # RandomCrop larger than the image
    assert bad(top=49) == -1 and bad(left=121) == -1 and bad(top=-1) == -1      # crop window outside the resized image
    assert bad(src_w=9000, pitch=27000, cw=8000, rw=640) == -2                   # down-scaling beyond the tap budget
    assert l.dfgan_augment(None, None, None, None, None, 0, 4, 256, None) == -1
    assert l.dfgan_augment(None, 4096, 4096, 4096, 4096, 16, 4, 256, None) == -1            # workspace too small


def test_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.BatchImageTransform(256)
    with pytest.raises(ValueError, match="larger than input image size"):
        P.draw_crop_flip(200, 304, 256)
