"""Drop-in `DFGANDataset` / `create_loader` (objects/dataset.py:15-141, utils.py:13-24) feeding the
batched GPU image transform.

The reference prepares every image inside `__getitem__` on the loader's single Python thread: PIL decode, bounding-box
crop, antialiased resize, random crop, flip, ToTensor, Normalize (objects/dataset.py:112-129) -- ~2.2 ms per image,
457 images/s, against the >2 700 images/s one B200 trains at.  Here `__getitem__` only decodes (PIL, the reference's
decoder) and picks the caption; the loader's `collate_fn` hands the whole batch of decoded images to
`augment.BatchImageTransform` (two CUDA launches, bit-identical pixels) and returns the reference's 4-tuple

    (images [B,3,S,S] fp32 -- already on the GPU, captions [B,18,1] int64, caption lengths [B] int64, file names)

which `prepare_data` / `DeepFusionGAN.fit` consume unchanged.  Same constructor arguments, attributes (`n_words`,
`code2word`, `word2code`, `file_names`, `captions`, `bbox`, `embeddings_num`, `num_examples`, ...), file layout
(`captions.pickle`, `{split}/filenames.pickle`, `bounding_boxes.txt`, `images.txt`, `birds/`), printed summary and random
draws: captions come from numpy's global generator with the reference's calls (`randint(0, 10)`, `shuffle` for captions
longer than 18 words), crop and flip from torch's global generator in sample order (augment.draw_crop_flip), and the
DataLoader is the same class with the same arguments (`drop_last=True, shuffle=True`), so a seeded epoch yields the
reference's batches bit for bit (tests/golden/dataset_golden.npz).

`transform` is `None` (decoded samples only) or a `BatchImageTransform`; torchvision's per-image Compose is the
reference's CPU path and is rejected -- this package has no CPU path for the pixels.
"""
from __future__ import annotations

import os
import pickle
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import numpy.random as random
import torch
from torch.utils.data import DataLoader, Dataset

N_STEPS = 18     # objects/dataset.py:100


class DFGANDataset(Dataset):
    def __init__(self, data_dir: str, split: str = "train", transform: Optional[Callable] = None):
        if transform is not None and not hasattr(transform, "plan"):
            raise TypeError("gan_codes_b200 DFGANDataset takes a BatchImageTransform (applied to whole batches on the "
                            "GPU by `collate`), not a per-image torchvision transform: create_loader builds it")
        self.split = split
        self.data_dir = data_dir
        self.split_dir = os.path.join(data_dir, split)
        self.captions_path = os.path.join(self.data_dir, "captions.pickle")
        self.filenames_path = os.path.join(self.split_dir, "filenames.pickle")
        self.transform = transform
        self.embeddings_num = 10
        self.images_dir = os.path.join(self.data_dir, "birds")
        self.bbox_path = self._locate("bounding_boxes.txt")
        self.images_path = self._locate("images.txt")
        self.bbox = self._load_bbox()
        self.file_names, self.captions, self.code2word, self.word2code = self._load_text_data()
        self.n_words = len(self.code2word)
        self.num_examples = len(self.file_names)
        self._print_info()

    def _locate(self, name: str) -> str:
        """data_dir, then ../data, then ./data (objects/dataset.py:35-47)."""
        path = os.path.join(self.data_dir, name)
        if not os.path.isfile(path):
            path = os.path.join("..", "data", name)
            if not os.path.isfile(path):
                path = os.path.join(os.getcwd(), "data", name)
        return path

    def _print_info(self):
        for line in (f"Total filenames: {len(self.bbox)}", f"Load captions from: {self.captions_path}",
                     f"Load file names from: {self.filenames_path} ({self.num_examples})",
                     f"Dictionary size: {self.n_words}", f"Embeddings number: {self.embeddings_num}"):
            print(line)

    @staticmethod
    def _rows(path: str) -> List[List[str]]:
        with open(path) as f:
            return [line.split() for line in f if line.strip()]

    def _load_bbox(self) -> Dict[str, List[int]]:
        """{file name without '.jpg': [x, y, width, height] as ints} (:63-75; the i-th row of bounding_boxes.txt
        belongs to the i-th row of images.txt, values truncated like DataFrame.astype(int))."""
        boxes = [[int(float(v)) for v in row[1:5]] for row in self._rows(self.bbox_path)]
        names = [row[1] for row in self._rows(self.images_path)]
        return {name[:-4]: boxes[i] for i, name in enumerate(names)}

    def _load_text_data(self) -> Tuple[List[str], List[List[int]], Dict[int, str], Dict[str, int]]:
        """captions.pickle = [train captions, test captions, code2word, word2code]; `{split}/filenames.pickle` = the
        split's file names, 10 consecutive captions each (:77-96)."""
        if not os.path.isfile(self.filenames_path):
            raise ValueError(f"File {self.filenames_path} does not exist")
        with open(self.captions_path, "rb") as f:
            by_split = pickle.load(f)
        with open(self.filenames_path, "rb") as f:
            names = pickle.load(f)
        return names, by_split[0 if self.split == "train" else 1], by_split[2], by_split[3]

    def _get_caption(self, caption_idx: int) -> Tuple[np.ndarray, int]:
        """:98-110: zero-padded [18,1] int64 column; longer captions keep 18 randomly chosen words in order."""
        words = np.asarray(self.captions[caption_idx])
        n = len(words)
        column = np.zeros((N_STEPS, 1), dtype="int64")
        if n > N_STEPS:
            order = list(np.arange(n))
            random.shuffle(order)                          # the reference's draw: a list shuffled by numpy's generator
            words, n = words[np.sort(order[:N_STEPS])], N_STEPS
        column[:n, 0] = words
        return column, n

    def _get_random_caption(self, idx: int) -> Tuple[np.ndarray, int]:
        caption_shift = random.randint(0, self.embeddings_num)
        return self._get_caption(idx * self.embeddings_num + caption_shift)

    def _decode(self, image_path: str) -> torch.Tensor:
        """Image.open(...).convert('RGB') (:113) as a uint8 [H, W, 3] tensor; everything after the decode is the
        batched transform's."""
        from PIL import Image
        with Image.open(image_path) as im:
            return torch.from_numpy(np.array(im.convert("RGB"), dtype=np.uint8))

    def __getitem__(self, idx: int):
        """(decoded uint8 [H,W,3], bounding box, caption [18,1], caption length, file name)."""
        file_name = self.file_names[idx]
        image = self._decode(f"{self.images_dir}/{file_name}.jpg")
        encoded_caption, caption_len = self._get_random_caption(idx)
        return image, self.bbox[file_name], encoded_caption, caption_len, file_name

    def __len__(self) -> int:
        return self.num_examples

    def collate(self, samples):
        """Batch of decoded samples -> the reference's (images, captions, captions_len, file_names) with the image
        transform applied to the whole batch on the GPU."""
        if self.transform is None:
            raise RuntimeError("DFGANDataset.collate needs the BatchImageTransform (create_loader sets it)")
        images, boxes, captions, lens, names = zip(*samples)
        batch = self.transform(list(images), list(boxes))
        return (batch, torch.from_numpy(np.stack(captions)), torch.tensor([int(n) for n in lens], dtype=torch.int64),
                list(names))


def _as_list(samples):
    return samples


class _MainProcessCollate:
    """A DataLoader whose worker processes only decode (`__getitem__` touches no CUDA state) while the batched GPU
    transform runs in the consuming process: iterating yields `dataset.collate(samples)` for every list of decoded
    samples the workers deliver.  Quacks like the DataLoader for what the trainer uses (`dataset`, `len`, iteration)."""

    def __init__(self, loader: DataLoader, dataset: "DFGANDataset"):
        self.loader = loader
        self.dataset = dataset
        self.batch_size = loader.batch_size
        self.sampler = loader.sampler       # DistributedSampler.set_epoch for the data-parallel trainer

    def __len__(self) -> int:
        return len(self.loader)

    def __iter__(self):
        if getattr(self.loader, "_iterator", None) is not None:
            # A DataLoader draws one number from the global generator whenever it builds an iterator (its worker base
            # seed); with persistent workers the iterator is built once and re-armed afterwards, so later epochs would
            # see a generator stream shifted by one draw against the single-process loader.  Spend that draw here: the
            # shuffling order and the crops of every epoch stay those of the seeded `num_workers=0` run (checked by
            # tests/test_dataset_cpu.py).
            torch.empty((), dtype=torch.int64).random_()
        for samples in self.loader:
            yield self.dataset.collate(samples)


def create_loader(imsize: int, batch_size: int, data_dir: str, split: str, num_workers: int = 0, rank: int = 0,
                  world: int = 1):
    """utils.py:13-24.  `num_workers=0` is the reference's loader (one thread decodes; a seeded epoch reproduces the
    reference's batches exactly).  PIL decodes ~250 CUB photographs per second per core, a tenth of what one B200
    trains at, so production runs pass `num_workers > 0`: the JPEG decode and the caption choice move to worker
    processes, the crop / flip draws and the GPU transform stay in this process.  Sample order, crops and flips (torch
    generator of this process) are then still those of the seeded reference run; the caption picked for an image is
    not (numpy's generator lives in the workers).  Data parallel: `rank` / `world` shard the samples with a
    DistributedSampler (same shuffle seed on every rank, disjoint indices, `loader.sampler.set_epoch(e)` per epoch --
    DeepFusionGAN.fit calls it)."""
    assert split in ["train", "test"], "Wrong split type, expected train or test"
    from .augment import BatchImageTransform
    dataset = DFGANDataset(data_dir, split, BatchImageTransform(imsize))
    return make_loader(dataset, batch_size, num_workers, rank, world)


def make_loader(dataset: DFGANDataset, batch_size: int, num_workers: int = 0, rank: int = 0, world: int = 1):
    order = {"shuffle": True}
    if world > 1:
        from torch.utils.data.distributed import DistributedSampler
        order = {"sampler": DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, drop_last=True)}
    if num_workers <= 0:
        return DataLoader(dataset, batch_size=batch_size, drop_last=True, collate_fn=dataset.collate, **order)
    import copy
    decode_only = copy.copy(dataset)        # what the workers receive: no transform, hence no CUDA object to fork / pickle
    decode_only.transform = None
    return _MainProcessCollate(DataLoader(decode_only, batch_size=batch_size, drop_last=True,
                                          collate_fn=_as_list, num_workers=num_workers, persistent_workers=True,
                                          prefetch_factor=4, **order), dataset)
