"""Drop-in DFGANDataset / loader on CPU: metadata, captions, sample order, random draws and the
collated batches of gan_codes_b200/dataset.py against tests/golden/dataset_golden.npz -- the batches the reference's
own unmodified DFGANDataset produced on tests/golden/mini_cub under a DataLoader built like its create_loader
(oracle/make_dataset_golden.py).  The pixels of a batch are the GPU transform's (tools/bringup_shell.py `loader` case
checks those bit-exactly on the device); here the CPU oracle of that transform stands in behind the same interface, so
the whole host side -- decode, box table, caption choice, crop / flip draws in sample order, collation -- is compared bit
for bit."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from gan_codes_b200 import augment as P
from gan_codes_b200.dataset import DFGANDataset, create_loader, make_loader
from oracle import augment_oracle as A

HERE = os.path.dirname(__file__)
MINI = os.path.join(HERE, "golden", "mini_cub")
IMSIZE, BATCH, SEED = 64, 2, 2024


class OracleTransform:
    """BatchImageTransform's interface on the CPU oracle (crop / flip drawn by the product's own draw_crop_flip)."""

    def __init__(self, imsize):
        self.imsize = imsize

    def plan(self, *a, **k):
        raise NotImplementedError

    def __call__(self, images, bboxes):
        out = []
        for im, bb in zip(images, bboxes):
            h, w = im.shape[:2]
            x1, y1, x2, y2 = P.bbox_crop_rect(w, h, bb)
            rw, rh = P.resized_size(x2 - x1, y2 - y1, self.imsize)
            top, left, flip = P.draw_crop_flip(rw, rh, self.imsize)
            out.append(torch.from_numpy(A.transform_image(im.numpy(), bb, self.imsize, top, left, flip)))
        return torch.stack(out)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "dataset_golden.npz"))


def loader(split):
    ds = DFGANDataset(MINI, split, OracleTransform(IMSIZE))
    return DataLoader(ds, batch_size=BATCH, drop_last=True, shuffle=True, collate_fn=ds.collate)


def levels(images):
    return torch.round((images * 0.5 + 0.5) * 255).to(torch.uint8).numpy()


@pytest.mark.parametrize("split,epochs", [("train", 2), ("test", 1)])
def test_seeded_epochs_reproduce_the_reference_batches(gold, split, epochs, capsys):
    ld = loader(split)
    out = capsys.readouterr().out
    assert "Total filenames: 8" in out and "Dictionary size: 40" in out and "Embeddings number: 10" in out
    ds = ld.dataset
    assert ds.n_words == int(gold[f"{split}_n_words"]) and len(ds) == ds.num_examples == int(gold[f"{split}_num_examples"])
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    k = 0
    for _ in range(epochs):
        for images, captions, lens, names in ld:
            assert images.dtype == torch.float32 and images.shape == (BATCH, 3, IMSIZE, IMSIZE)
            assert captions.dtype == torch.int64 and captions.shape == (BATCH, 18, 1) and lens.dtype == torch.int64
            assert list(names) == gold[f"{split}_{k}_names"].tolist()
            assert np.array_equal(captions.numpy(), gold[f"{split}_{k}_captions"])
            assert np.array_equal(lens.numpy(), gold[f"{split}_{k}_lens"])
            assert np.array_equal(levels(images), gold[f"{split}_{k}_images"])        # bit-exact pixels
            back = torch.from_numpy(gold[f"{split}_{k}_images"]).float().div(255).sub(0.5).div(0.5)
            assert torch.equal(back, images)
            k += 1
    assert k == int(gold[f"{split}_batches"])


def test_bbox_table_and_attributes(gold):
    ds = DFGANDataset(MINI, "train")
    assert sorted(ds.bbox) == gold["bbox_names"].tolist()
    assert np.array_equal(np.array([ds.bbox[n] for n in sorted(ds.bbox)]), gold["bbox_values"])
    assert all(isinstance(v, int) for box in ds.bbox.values() for v in box)
    assert ds.embeddings_num == 10 and ds.split == "train" and ds.images_dir.endswith("birds")
    assert ds.code2word[0] == "<end>" and ds.word2code["w7"] == 7 and len(ds.captions) == 10 * len(ds.file_names)
    image, box, cap, n, name = ds[0]
    assert image.dtype == torch.uint8 and image.shape == (90, 120, 3) and box == [10, 5, 100, 80]
    assert cap.shape == (18, 1) and cap.dtype == np.int64 and 1 <= n <= 18 and name == ds.file_names[0]


def test_long_captions_keep_18_words_in_order():
    ds = DFGANDataset(MINI, "train")
    long_idx = next(i for i, c in enumerate(ds.captions) if len(c) > 18)
    np.random.seed(1)
    cap, n = ds._get_caption(long_idx)
    assert n == 18 and (cap[:, 0] != 0).all()
    src = list(ds.captions[long_idx])
    pos = 0
    for w in cap[:, 0]:                       # a subsequence of the caption
        pos = src.index(w, pos) + 1
    short_idx = next(i for i, c in enumerate(ds.captions) if len(c) < 18)
    cap, n = ds._get_caption(short_idx)
    assert n == len(ds.captions[short_idx]) and (cap[n:, 0] == 0).all()


def test_errors():
    with pytest.raises(ValueError, match="filenames.pickle does not exist"):
        DFGANDataset(MINI, "valid")
    import torchvision.transforms as T
    with pytest.raises(TypeError, match="BatchImageTransform"):
        DFGANDataset(MINI, "train", T.Compose([T.Resize(76)]))
    with pytest.raises(RuntimeError, match="needs the BatchImageTransform"):
        ds = DFGANDataset(MINI, "train")
        ds.collate([ds[0], ds[1]])
    with pytest.raises(AssertionError, match="Wrong split type"):
        create_loader(64, 2, MINI, "valid")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA only"):
            create_loader(64, 2, MINI, "train")


def test_worker_processes_decode_and_the_main_process_transforms(gold):
    """num_workers > 0: same sample order, crops, flips and pixels as the seeded single-process run (all drawn from this
    process's torch generator); only the caption choice moves to the workers' numpy generators."""
    ds = DFGANDataset(MINI, "train", OracleTransform(IMSIZE))
    ld = make_loader(ds, BATCH, num_workers=2)
    assert len(ld) == 2 and ld.dataset is ds and ld.batch_size == BATCH
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    k = 0
    for _ in range(2):
        for images, captions, lens, names in ld:
            assert list(names) == gold[f"train_{k}_names"].tolist()
            assert np.array_equal(levels(images), gold[f"train_{k}_images"])
            assert captions.shape == (BATCH, 18, 1) and captions.dtype == torch.int64 and lens.dtype == torch.int64
            for b, name in enumerate(names):                      # one of the image's ten captions, correctly padded
                idx = ds.file_names.index(name)
                n = int(lens[b])
                cands = [c for c in ds.captions[idx * 10:(idx + 1) * 10] if min(len(c), 18) == n]
                row = captions[b, :n, 0].tolist()
                assert any(row == list(c) if len(c) <= 18 else set(row) <= set(c) for c in cands)
                assert (captions[b, n:, 0] == 0).all()
            k += 1
    assert k == 4


def test_data_parallel_loaders_are_disjoint():
    """create_loader(rank=, world=): the ranks of a data-parallel run read disjoint shards of one seeded shuffle, and
    set_epoch changes the shuffle (without a sampler every rank saw the same batches)."""
    from gan_codes_b200.dataset import make_loader
    ds = DFGANDataset(MINI, "train", None)
    per_rank = []
    for rank in (0, 1):
        ld = make_loader(ds, 1, rank=rank, world=2)
        ld.sampler.set_epoch(0)
        first = list(ld.sampler)
        ld.sampler.set_epoch(1)
        second = list(ld.sampler)
        per_rank.append((first, second))
    (a0, a1), (b0, b1) = per_rank
    assert set(a0).isdisjoint(b0) and set(a1).isdisjoint(b1)
    assert len(a0) == len(b0) == len(ds) // 2
    assert set(a0 + b0) <= set(range(len(ds))) and (a0, b0) != (a1, b1)      # another shuffle in the next epoch
