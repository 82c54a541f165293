"""TEST INFRASTRUCTURE.  Builds the miniature CUB-format data set tests/golden/mini_cub/ and generates
tests/golden/dataset_golden.npz by running the reference's UNMODIFIED `DFGANDataset` (objects/dataset.py, imported
read-only from /root/reference) under a `DataLoader` built exactly like `create_loader` (utils.py:13-24; that module
itself cannot be imported here because it imports matplotlib).  Build container only:

    python oracle/make_dataset_golden.py

mini_cub: 8 synthetic photographs (augment_oracle.synthetic_image, stored losslessly as PNG data in `.jpg`-named files --
PIL picks the decoder from the content, so every box decodes the same pixels), `images.txt`, `bounding_boxes.txt`,
`captions.pickle` (10 captions per image, lengths 3..25 so that the >18-word shuffle path of :104-110 is taken),
`train/filenames.pickle` (5 files), `test/filenames.pickle` (3 files).

Golden: for the seeds below, two epochs of the train loader (batch 2, imsize 64, drop_last, shuffle) and one of the test
loader -- file names, captions, caption lengths and the image tensors (as uint8 levels, lossless) of every batch, plus
the dataset attributes a trainer reads (n_words, num_examples, bbox table).
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/Deep_Fusion_GAN")
sys.path.insert(0, ROOT)

from oracle import augment_oracle as A  # noqa: E402

MINI = os.path.join(ROOT, "tests", "golden", "mini_cub")
IMSIZE, BATCH, SEED = 64, 2, 2024
SIZES = [(120, 90), (97, 131), (150, 100), (88, 88), (140, 105), (100, 140), (128, 96), (90, 120)]
BOXES = [(10.0, 5.0, 100.0, 80.0), (0.0, 20.0, 97.0, 100.0), (30.0, 10.0, 90.0, 85.0), (4.0, 4.0, 80.0, 80.0),
         (20.0, 15.0, 100.0, 70.0), (10.0, 30.0, 85.0, 95.0), (0.0, 0.0, 128.0, 96.0), (15.0, 25.0, 60.0, 80.0)]
N_WORDS = 40


def build_mini() -> None:
    from PIL import Image
    r = np.random.RandomState(7)
    names = [f"{k // 3 + 1:03d}.Species_{k // 3}/Bird_{k:04d}" for k in range(len(SIZES))]
    for k, ((w, h), name) in enumerate(zip(SIZES, names)):
        path = os.path.join(MINI, "birds", name + ".jpg")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            Image.fromarray(A.synthetic_image(w, h, 100 + k)).save(f, format="PNG", optimize=True)
    with open(os.path.join(MINI, "images.txt"), "w") as f:
        f.writelines(f"{k + 1} {n}.jpg\n" for k, n in enumerate(names))
    with open(os.path.join(MINI, "bounding_boxes.txt"), "w") as f:
        f.writelines(f"{k + 1} {b[0]} {b[1]} {b[2]} {b[3]}\n" for k, b in enumerate(BOXES))
    caps = lambda n: [[int(x) for x in r.randint(1, N_WORDS, size=r.randint(3, 26))] for _ in range(10 * n)]  # noqa: E731
    code2word = {i: ("<end>" if i == 0 else f"w{i}") for i in range(N_WORDS)}
    word2code = {w: i for i, w in code2word.items()}
    with open(os.path.join(MINI, "captions.pickle"), "wb") as f:
        pickle.dump([caps(5), caps(3), code2word, word2code], f, protocol=2)
    for split, sel in (("train", names[:5]), ("test", names[5:])):
        os.makedirs(os.path.join(MINI, split), exist_ok=True)
        with open(os.path.join(MINI, split, "filenames.pickle"), "wb") as f:
            pickle.dump(list(sel), f, protocol=2)


def reference_loader(split: str):
    """create_loader (utils.py:13-24) spelled out with the reference's dataset class."""
    from torch.utils.data import DataLoader
    from torchvision.transforms import transforms
    from src.objects.dataset import DFGANDataset
    image_transform = transforms.Compose([transforms.Resize(int(IMSIZE * 76 / 64)), transforms.RandomCrop(IMSIZE),
                                          transforms.RandomHorizontalFlip()])
    dataset = DFGANDataset(MINI, split, image_transform)
    return DataLoader(dataset, batch_size=BATCH, drop_last=True, shuffle=True)


def to_levels(images: torch.Tensor) -> np.ndarray:
    lv = torch.round((images * 0.5 + 0.5) * 255).to(torch.uint8)
    back = (lv.float().div(255) - 0.5) / 0.5
    assert torch.equal(back, images), "image tensor is not on the 256-level grid"
    return lv.numpy()


def main() -> int:
    build_mini()
    out = {}
    for split, epochs in (("train", 2), ("test", 1)):
        loader = reference_loader(split)
        ds = loader.dataset
        out[f"{split}_n_words"] = np.array(ds.n_words)
        out[f"{split}_num_examples"] = np.array(ds.num_examples)
        torch.manual_seed(SEED)
        np.random.seed(SEED)
        k = 0
        for _ in range(epochs):
            for images, captions, lens, names in loader:
                out[f"{split}_{k}_images"] = to_levels(images)
                out[f"{split}_{k}_captions"] = captions.numpy()
                out[f"{split}_{k}_lens"] = lens.numpy()
                out[f"{split}_{k}_names"] = np.array(list(names))
                assert captions.shape == (BATCH, 18, 1) and captions.dtype == torch.int64 and lens.dtype == torch.int64
                k += 1
        out[f"{split}_batches"] = np.array(k)
        print(split, "batches", k, "lens", [out[f"{split}_{i}_lens"].tolist() for i in range(k)])
    bbox = ds.bbox
    out["bbox_names"] = np.array(sorted(bbox))
    out["bbox_values"] = np.array([bbox[n] for n in sorted(bbox)])
    path = os.path.join(ROOT, "tests", "golden", "dataset_golden.npz")
    np.savez_compressed(path, **out)
    size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(MINI) for f in fs)
    print("wrote", path, os.path.getsize(path), "bytes; mini_cub", size, "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
