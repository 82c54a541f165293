"""Drop-in `damsm_loss` (damsm/loss.py:4-26): the text-image score the reference's trainer logs each iteration.

    loss = 1 - mean_b cos( pad_E(mean over H,W of fake_images[b]), text_encoder(captions, cap_lens)[b] )

Same signature and value; one encoder pass (text_encoder.RNNEncoder, two launches) and `dfgan_txtimg_loss` (two
launches, bit-reproducible) instead of mean / eye / linear / normalize x2 / mul / sum / mean.  Inference-style like every
use in the reference (the result is logged, never differentiated: deep_fusion_gan/model.py:220-223, 231): the returned
scalar carries no grad_fn.  Inside `DeepFusionGAN.fit` the same kernel runs within the captured iteration on the
sentence embeddings it already holds.
"""
from __future__ import annotations

import torch

from . import kernels as K


@torch.no_grad()
def text_image_loss(fake_images: torch.Tensor, sent_emb: torch.Tensor) -> torch.Tensor:
    """The loss from given sentence embeddings [B, E] (fp32, CUDA)."""
    if not sent_emb.is_cuda:
        raise RuntimeError("gan_codes_b200 damsm: CUDA tensors only (no CPU path)")
    if fake_images.dim() != 4 or fake_images.shape[0] != sent_emb.shape[0]:
        raise RuntimeError(f"fake_images must be [B,C,H,W] with the batch of the embeddings, got "
                           f"{tuple(fake_images.shape)} and {tuple(sent_emb.shape)}")
    loss = torch.empty(1, device=sent_emb.device, dtype=torch.float32)
    K.txtimg_loss(fake_images.detach().to(sent_emb.device, torch.float32).contiguous(),
                  sent_emb.detach().to(torch.float32).contiguous(), loss)
    return loss[0]


def damsm_loss(fake_images, captions, cap_lens, text_encoder):
    """fake_images (B, 3, H, W); captions (B, max_len); cap_lens (B,); text_encoder: RNNEncoder."""
    return text_image_loss(fake_images, text_encoder(captions, cap_lens))
