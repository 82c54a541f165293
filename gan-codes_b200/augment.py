"""Image side of the reference's input pipeline for a whole batch of decoded images.

The reference prepares every image on one Python thread (DataLoader num_workers=0, utils.py:24):
`DFGANDataset._get_image` (objects/dataset.py:112-129) crops the bird's bounding box, `create_loader`'s transform
(utils.py:15-19) resizes the smaller edge to int(imsize*76/64) with Pillow's antialiased bilinear filter, takes a
random imsize x imsize crop and flips it with probability 1/2, and `DFGANDataset.normalize` (:28-31) maps it to fp32
CHW in [-1, 1].  `BatchImageTransform` produces the same tensors -- bit for bit -- for a list of decoded uint8 images
in two CUDA launches (csrc/augment.cu):

    tf = BatchImageTransform(256)
    images = tf(decoded, bboxes)            # [B,3,256,256] fp32 on the GPU, ready for DFGANTrainer.step

The random decisions are drawn on the host from torch's global generator with the calls and in the order torchvision
makes them (RandomCrop.get_params: randint for the row, randint for the column -- none when the resized image already
has the crop size; RandomHorizontalFlip: rand(1) < p), image by image, so a seeded run crops and flips exactly like the
reference's transform applied to the same images in the same order.  JPEG decoding is outside this module (the
caller's decoder; pixel values are taken as given).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib_decl import AugItem

RESIZE_NUM, RESIZE_DEN = 76, 64        # utils.py:16


def bbox_crop_rect(width: int, height: int, bbox: Sequence[float]) -> Tuple[int, int, int, int]:
    """(x1, y1, x2, y2) of objects/dataset.py:116-124 for an image of the given size."""
    r = int(max(bbox[2], bbox[3]) * 0.75)
    cx = int((2 * bbox[0] + bbox[2]) / 2)
    cy = int((2 * bbox[1] + bbox[3]) / 2)
    return max(0, cx - r), max(0, cy - r), min(width, cx + r), min(height, cy + r)


def resized_size(w: int, h: int, imsize: int) -> Tuple[int, int]:
    """(width, height) transforms.Resize(int(imsize*76/64)) gives a w x h image (torchvision: smaller edge -> size,
    longer edge -> int(size * long / short))."""
    size = int(imsize * RESIZE_NUM / RESIZE_DEN)
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_short, new_long) if w <= h else (new_long, new_short)


def draw_crop_flip(rw: int, rh: int, imsize: int, p_flip: float = 0.5) -> Tuple[int, int, bool]:
    """(top, left, flip) from torch's global generator, consumed like RandomCrop.get_params followed by
    RandomHorizontalFlip.forward."""
    if rw < imsize or rh < imsize:
        raise ValueError(f"Required crop size {(imsize, imsize)} is larger than input image size {(rh, rw)}")
    if rw == imsize and rh == imsize:
        top = left = 0
    else:
        top = int(torch.randint(0, rh - imsize + 1, size=(1,)).item())
        left = int(torch.randint(0, rw - imsize + 1, size=(1,)).item())
    return top, left, bool(torch.rand(1) < p_flip)


class BatchImageTransform:
    def __init__(self, imsize: int = 256, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("gan_codes_b200.augment: CUDA only (no CPU path)")
        self.imsize = imsize
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        # the value ToTensor + Normalize(0.5, 0.5) give each uint8 level, from the reference's own fp32 expression
        levels = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
        self.lut = levels.sub(0.5).div(0.5).to(self.device)
        self._ws = None
        self._pin = None
        self._items_dev = None
        self._copied = torch.cuda.Event()

    def plan(self, sizes: Sequence[Tuple[int, int]], bboxes: Sequence[Sequence[float]],
             params: Optional[Sequence[Tuple[int, int, bool]]] = None) -> List[dict]:
        """Host-side geometry of every image: crop rectangle, resized size, (top, left, flip) -- drawn here unless
        given.  `sizes` are (width, height) of the decoded images."""
        out = []
        for i, ((w, h), bbox) in enumerate(zip(sizes, bboxes)):
            x1, y1, x2, y2 = bbox_crop_rect(w, h, bbox)
            rw, rh = resized_size(x2 - x1, y2 - y1, self.imsize)
            top, left, flip = draw_crop_flip(rw, rh, self.imsize) if params is None else params[i]
            out.append(dict(x1=int(x1), y1=int(y1), cw=int(x2 - x1), ch=int(y2 - y1), rw=rw, rh=rh, top=int(top),
                            left=int(left), flip=int(bool(flip))))
        return out

    @torch.no_grad()
    def __call__(self, images: Sequence[torch.Tensor], bboxes: Sequence[Sequence[float]],
                 params: Optional[Sequence[Tuple[int, int, bool]]] = None) -> torch.Tensor:
        """images: decoded RGB uint8 tensors [H, W, 3] (host or device), one per sample; bboxes: (x, y, width, height)
        per sample as in bounding_boxes.txt; params: optional (top, left, flip) per sample instead of random draws."""
        b, s, dev = len(images), self.imsize, self.device
        if b == 0 or len(bboxes) != b:
            raise RuntimeError(f"need one bounding box per image, got {b} images and {len(bboxes)} boxes")
        dimgs = []
        for im in images:
            if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
                raise RuntimeError(f"images must be uint8 [H, W, 3], got {im.dtype} {tuple(im.shape)}")
            im = im.to(dev, non_blocking=True)
            if im.stride(2) != 1 or im.stride(1) != 3:
                im = im.contiguous()
            dimgs.append(im)
        geo = self.plan([(im.shape[1], im.shape[0]) for im in dimgs], bboxes, params)
        items = (AugItem * b)()
        l = _lib.lib()
        for it, im, g in zip(items, dimgs, geo):
            it.src, it.pitch, it.src_w, it.src_h = im.data_ptr(), im.stride(0), im.shape[1], im.shape[0]
            for k, v in g.items():
                setattr(it, k, v)
            _lib.check(l.dfgan_augment_check(C.byref(it), s), "dfgan_augment_check")
        nbytes = C.sizeof(items)
        if self._pin is None or self._pin.numel() < nbytes:
            self._pin = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self._items_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        else:
            self._copied.synchronize()          # the previous call's table has left the staging buffer
        C.memmove(self._pin.data_ptr(), items, nbytes)
        self._items_dev[:nbytes].copy_(self._pin[:nbytes], non_blocking=True)
        self._copied.record()
        need = l.dfgan_augment_workspace_bytes(b, s)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty(b, 3, s, s, device=dev, dtype=torch.float32)
        _lib.check(l.dfgan_augment(items, self._items_dev.data_ptr(), self.lut.data_ptr(), out.data_ptr(),
                                   self._ws.data_ptr(), self._ws.numel(), b, s,
                                   torch.cuda.current_stream().cuda_stream), "dfgan_augment")
        return out
