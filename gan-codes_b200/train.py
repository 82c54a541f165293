"""`train()` of the reference's train.py:8-66 on the drop-in classes: seed, loaders, trainer, resume from
`checkpoint.pt`, `fit`, per-epoch summary.  The reference hard-codes its author's paths; here they are arguments
(defaults: the reference's relative output directories).  One deliberate difference: a restart continues with the epoch
after the checkpoint's (the reference script repeats the checkpoint's epoch, see the comment in `train`).

    python -m gan_codes_b200.train --data /data/birds --encoder /weights/text_encoder.pth [--epochs 600] [--batch 24]
                                   [--workers 12]
"""
from __future__ import annotations

import argparse
import os
import random
from typing import List, Tuple

import numpy as np
import torch


def fix_seed(seed: int = 123321):
    """utils.py:27-35."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    print(f"Seed {seed} fixed")


def train(data_path: str, encoder_weights_path: str, image_save_path: str = "../gen_images",
          gen_path_save: str = "../gen_weights", num_epochs: int = 600, batch_size: int = 24,
          imsize: int = 256, num_workers: int = 0) -> Tuple[List[float], ...]:
    from .dataset import create_loader
    from .deep_fusion_gan import DeepFusionGAN
    fix_seed()
    os.makedirs(image_save_path, exist_ok=True)
    os.makedirs(gen_path_save, exist_ok=True)
    extra = {"num_workers": num_workers} if num_workers else {}          # 0 = the reference's single-thread loader
    train_loader = create_loader(imsize, batch_size, data_path, "train", **extra)
    test_loader = create_loader(imsize, batch_size, data_path, "test", **extra)
    print(f"Test set size: {len(test_loader.dataset)} images")
    model = DeepFusionGAN(n_words=train_loader.dataset.n_words, encoder_weights_path=encoder_weights_path,
                          image_save_path=image_save_path, gen_path_save=gen_path_save)
    model.ixtoword = train_loader.dataset.code2word
    checkpoint_path = os.path.join(gen_path_save, "checkpoint.pt")
    histories = [[] for _ in range(6)]
    start_epoch = 0
    if os.path.exists(checkpoint_path):
        print(f"Loading checkpoint from latest: {checkpoint_path}")
        last_epoch, *histories = model.load_from_checkpoint(checkpoint_path)
        # The reference passes the checkpoint's own epoch on (train.py:35, 46), so it trains that epoch a second time
        # and its histories grow one entry longer than the epoch count on every restart; `fit`'s own resume path
        # (deep_fusion_gan/model.py:150) continues with the NEXT epoch, and so does this script.
        start_epoch = last_epoch + 1
    else:
        print(f"Warning: Checkpoint {checkpoint_path} not found. Starting from scratch.")
    out = model.fit(train_loader, test_loader=test_loader, num_epochs=num_epochs, start_epoch=start_epoch,
                    g_losses_epoch=histories[0], d_losses_epoch=histories[1], d_gp_losses_epoch=histories[2],
                    is_scores_epoch=histories[3], fid_scores_epoch=histories[4], txtimg_losses_epoch=histories[5],
                    auto_resume=False)
    g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses = out
    for epoch in range(len(g_losses)):
        print(f"Epoch {epoch + 1}: G Loss: {g_losses[epoch]:.4f}, D Loss: {d_losses[epoch]:.4f}, "
              f"D GP Loss: {d_gp_losses[epoch]:.4f}, Text-Image Loss: {txtimg_losses[epoch]:.4f}, "
              f"IS: {is_scores[epoch]:.4f}, FID: {fid_scores[epoch]:.4f}")
    return out


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--data", required=True)
    ap.add_argument("--encoder", required=True)
    ap.add_argument("--images-out", default="../gen_images")
    ap.add_argument("--weights-out", default="../gen_weights")
    ap.add_argument("--epochs", type=int, default=600)
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--imsize", type=int, default=256)
    ap.add_argument("--workers", type=int, default=0, help="decode worker processes per loader (0 = reference behaviour)")
    a = ap.parse_args()
    train(a.data, a.encoder, a.images_out, a.weights_out, a.epochs, a.batch, a.imsize, a.workers)


if __name__ == "__main__":
    main()
