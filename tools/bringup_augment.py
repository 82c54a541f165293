"""GPU bring-up / parity cases of the batched image transform (gan_codes_b200.augment.BatchImageTransform over
csrc/augment.cu): BIT-EXACT against the committed outputs of the reference's `_get_image`
(tests/golden/augment_golden.npz), the integer-exact oracle and the stock torchvision / Pillow pipeline run on the box.
Bodies of tests/test_gpu_parity.py::test_augment; also runnable as  python tools/bringup_augment.py [case]
"""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _tables(tf, b, s):
    """Device coefficient tables of the last call as numpy: lo, cnt [B,2,S], k [B,2,KMAX,S]."""
    ws = tf._ws.view(dtype=__import__("torch").int32).cpu().numpy()
    kmax = ws.size // (b * 2 * s) - 2
    lo, cnt = ws[:b * 2 * s].reshape(b, 2, s), ws[b * 2 * s:b * 4 * s].reshape(b, 2, s)
    return lo, cnt, ws[b * 4 * s:b * (4 + 2 * kmax) * s].reshape(b, 2, kmax, s)


def _check_tables(tf, geo, s):
    """Diagnostic: the device tables against oracle.resample_coeffs for the kept window (reports the first miss)."""
    import numpy as np
    from oracle import augment_oracle as A
    lo, cnt, k = _tables(tf, len(geo), s)
    ok = True
    for b, g in enumerate(geo):
        for axis, (n_in, n_out, off) in enumerate(((g["cw"], g["rw"], g["left"]), (g["ch"], g["rh"], g["top"]))):
            xmin, xn, kk = A.resample_coeffs(n_in, n_out)
            idx = off + np.arange(s)
            if axis == 0 and g["flip"]:
                idx = off + s - 1 - np.arange(s)
            want_k = np.zeros((k.shape[2], s), np.int32)
            want_k[:kk.shape[1]] = kk[idx].T
            good = np.array_equal(lo[b, axis], xmin[idx]) and np.array_equal(cnt[b, axis], xn[idx]) and \
                np.array_equal(k[b, axis], want_k)
            if not good:
                o = int(np.argmax((lo[b, axis] != xmin[idx]) | (cnt[b, axis] != xn[idx]) | (k[b, axis] != want_k).any(0)))
                print(f"  table mismatch image {b} axis {axis} out {o}: lo {lo[b, axis, o]} vs {xmin[idx][o]}, "
                      f"n {cnt[b, axis, o]} vs {xn[idx][o]}, k {k[b, axis, :6, o]} vs {want_k[:6, o]}")
            ok &= good
    return ok


def _compare(name, got, want):
    import numpy as np
    same = np.array_equal(got, want)
    extra = "" if same else f"   <-- FAIL ({int((got != want).sum())} of {got.size} values differ, max {np.abs(got - want).max():.3e})"
    print(f"  {name:66s} {'bit-exact' if same else 'DIFFERENT'}{extra}")
    return same


@case
def golden():
    """The committed outputs of the reference's unmodified DFGANDataset._get_image, batched per image size: once with
    the recorded crop / flip decisions, once drawing them from the torch generator seeded like the golden run."""
    import numpy as np
    import torch
    from gan_codes_b200.augment import BatchImageTransform
    from tests.test_augment_cpu import golden_case, G
    by_size = {}
    for i in range(int(G["n"])):
        c = golden_case(i)
        by_size.setdefault(c[2], []).append(c)
    ok = True
    for imsize, cs in sorted(by_size.items()):
        tf = BatchImageTransform(imsize)
        imgs = [torch.from_numpy(c[0]) for c in cs]
        out = tf(imgs, [c[1] for c in cs], params=[c[4] for c in cs]).cpu().numpy()
        geo = tf.plan([(c[0].shape[1], c[0].shape[0]) for c in cs], [c[1] for c in cs], [c[4] for c in cs])
        ok &= _check_tables(tf, geo, imsize)
        ok &= _compare(f"imsize {imsize}: {len(cs)} images, recorded decisions vs reference _get_image",
                       out, np.stack([c[5] for c in cs]))
        drawn = []
        for c in cs:                      # one image per call, seeded like the reference call was
            torch.manual_seed(c[3])
            drawn.append(tf([torch.from_numpy(c[0]).cuda()], [c[1]])[0].cpu().numpy())
        ok &= _compare(f"imsize {imsize}: decisions drawn from the seeded torch generator", np.stack(drawn),
                       np.stack([c[5] for c in cs]))
    return ok


def _cub_like_batch(n, seed):
    import numpy as np
    from oracle import augment_oracle as A
    rng = np.random.default_rng(seed)
    imgs, boxes = [], []
    for i in range(n):
        w, h = [(500, 375), (500, 333), (375, 500), (500, 500), (320, 240), (428, 500)][i % 6]
        img = A.synthetic_image(w, h, seed * 100 + i) if i % 2 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bw, bh = int(rng.integers(w // 4, w)), int(rng.integers(h // 4, h))
        boxes.append([int(rng.integers(0, w - bw + 1)), int(rng.integers(0, h - bh + 1)), bw, bh])
        imgs.append(img)
    return imgs, boxes


@case
def batch32():
    """A benchmark-size batch of CUB-like photographs (mixed sizes and boxes, random crops and flips) against the
    oracle AND against the stock torchvision / Pillow pipeline executed on the box with the same generator seed."""
    import numpy as np
    import torch
    import torchvision.transforms as transforms
    from PIL import Image
    from gan_codes_b200.augment import BatchImageTransform
    from oracle import augment_oracle as A
    imgs, boxes = _cub_like_batch(32, 1)
    tf = BatchImageTransform(256)
    torch.manual_seed(77)
    out = tf([torch.from_numpy(im) for im in imgs], boxes).cpu().numpy()
    torch.manual_seed(77)
    geo = tf.plan([(im.shape[1], im.shape[0]) for im in imgs], boxes)          # same draws again
    ok = _check_tables(tf, geo, 256)
    want = np.stack([A.transform_image(im, bx, 256, g["top"], g["left"], bool(g["flip"]))
                     for im, bx, g in zip(imgs, boxes, geo)])
    ok &= _compare("32 images 256x256 vs oracle", out, want)
    t = transforms.Compose([transforms.Resize(int(256 * 76 / 64)), transforms.RandomCrop(256),
                            transforms.RandomHorizontalFlip()])
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    torch.manual_seed(77)
    stock = []
    for im, bx in zip(imgs, boxes):
        x1, y1, x2, y2 = A.bbox_crop_rect(im.shape[1], im.shape[0], bx)
        stock.append(norm(t(Image.fromarray(im).crop((x1, y1, x2, y2)))).numpy())
    ok &= _compare("32 images 256x256 vs stock torchvision + Pillow on this box", out, np.stack(stock))
    ok &= _compare("device-resident inputs, same decisions", tf(
        [torch.from_numpy(im).cuda() for im in imgs], boxes,
        params=[(g["top"], g["left"], g["flip"]) for g in geo]).cpu().numpy(), want)
    return ok


@case
def errors():
    import torch
    from gan_codes_b200.augment import BatchImageTransform
    tf = BatchImageTransform(256)
    img = torch.zeros(300, 400, 3, dtype=torch.uint8)

    def raises(fn, exc, text):
        try:
            fn()
        except exc as e:
            return text in str(e)
        return False
    ok = raises(lambda: tf([img.float()], [[0, 0, 100, 100]]), RuntimeError, "uint8 [H, W, 3]")
    ok &= raises(lambda: tf([img], []), RuntimeError, "one bounding box per image")
    ok &= raises(lambda: tf([img], [[0, 0, 100, 100]], params=[(60, 0, False)]), RuntimeError, "dfgan_augment_check")
    big = torch.zeros(3600, 3600, 3, dtype=torch.uint8)       # 3600 -> 304: factor 11.8, 25 taps > the budget of 24
    ok &= raises(lambda: tf([big], [[0, 0, 3600, 3600]], params=[(0, 0, False)]), RuntimeError, "unsupported shape")
    print("  error behaviour:", "ok" if ok else "MISMATCH")
    return ok


@case
def timing():
    """Informational: device time for a 32-image batch (inputs resident) next to the stock pipeline on one host thread."""
    import torch
    import torchvision.transforms as transforms
    from PIL import Image
    from gan_codes_b200.augment import BatchImageTransform
    from oracle import augment_oracle as A
    imgs, boxes = _cub_like_batch(32, 2)
    tf = BatchImageTransform(256)
    dimgs = [torch.from_numpy(im).cuda() for im in imgs]
    for _ in range(3):
        tf(dimgs, boxes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        tf(dimgs, boxes)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20
    print(f"  ours: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us device / {wall * 1e6:.0f} us wall per 32-image batch "
          f"({32 / wall:.0f} images/s incl. host geometry)")
    t = transforms.Compose([transforms.Resize(int(256 * 76 / 64)), transforms.RandomCrop(256),
                            transforms.RandomHorizontalFlip()])
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    pil = [Image.fromarray(im) for im in imgs]
    t0 = time.perf_counter()
    for im, p, bx in zip(imgs, pil, boxes):
        norm(t(p.crop(A.bbox_crop_rect(im.shape[1], im.shape[0], bx))))
    dt = time.perf_counter() - t0
    print(f"  stock torchvision + Pillow, one host thread (decode excluded): {dt * 1e3:.1f} ms per 32-image batch "
          f"({32 / dt:.0f} images/s)")
    return True


if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    bad = []
    for name in which:
        print("==", name, flush=True)
        if not CASES[name]():
            bad.append(name)
    print("FAILED: %s" % bad if bad else "ALL OK")
    sys.exit(1 if bad else 0)
