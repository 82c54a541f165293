"""Drop-in sampling helpers (sample.py:13-45, test_generate.py:12-20) on the drop-in modules.

    generate_images(generator, sentence_embeds, device)   no_grad generator forward on fresh host-drawn noise (:13-18)
    save_image(image, save_dir, file_name)                 [-1,1] CHW array -> uint8 PNG (:21-28)
    sample(generator, text_encoder, batch, save_dir)       one loader batch -> one PNG per caption (:31-45)
    preprocess_caption(caption, word2code, max_len)        prompt string -> (captions [1,18,1], cap_lens [1])
    encode_prompt(text_encoder, caption, word2code)        the sentence embedding of one prompt

`generate_images` draws its noise like the reference, `torch.randn(batch_size, 100)` from the host generator, so a seeded
call sees the same latent vectors.  The generator forward is the bf16 tensor-core program of modules.Generator (BASELINE
config 5: 256-image batches); everything else here is file and string handling.

test_generate.py feeds `preprocess_caption`'s [1,18,1] tensor straight into the text encoder (:66), which the
reference's `pack_padded_sequence` rejects for lack of a squeeze (objects/utils.py:15 does it for loader batches);
`encode_prompt` does the squeeze the script forgot.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import Tensor

from .deep_fusion_gan import prepare_data


@torch.no_grad()
def generate_images(generator, sentence_embeds: Tensor, device: torch.device) -> Tensor:
    batch_size = sentence_embeds.shape[0]
    noise = torch.randn(batch_size, 100).to(device)
    return generator(noise, sentence_embeds)


def save_image(image: np.ndarray, save_dir: str, file_name: str):
    """[-1, 1] --> [0, 255] by truncation (:23-24), CHW -> HWC, PNG named after the sample ('/' -> '_')."""
    from PIL import Image
    image = ((image + 1.0) * 127.5).astype(np.uint8)
    image = Image.fromarray(np.transpose(image, (1, 2, 0)))
    image.save(os.path.join(save_dir, f"{file_name.replace('/', '_')}.png"))


def sample(generator, text_encoder, batch, save_dir: str):
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    os.makedirs(save_dir, exist_ok=True)
    images, captions, captions_len, file_names = prepare_data(batch, device)
    sent_emb = text_encoder(captions, captions_len).detach()
    fake_images = generate_images(generator, sent_emb, device).data.cpu().numpy()    # ONE device->host copy
    for i in range(images.shape[0]):
        save_image(fake_images[i], save_dir, file_names[i])


def preprocess_caption(caption, word2code, max_len=18):
    """test_generate.py:12-20: lower-case, drop '.' and ',', split on blanks, unknown words -> '<unk>' (or 0)."""
    tokens = caption.lower().replace(".", "").replace(",", "").split()
    encoded = [word2code.get(w, word2code.get("<unk>", 0)) for w in tokens]
    cap_len = min(len(encoded), max_len)
    arr = np.zeros((max_len, 1), dtype="int64")
    arr[:cap_len, 0] = encoded[:max_len]
    return torch.LongTensor(arr).unsqueeze(0), torch.LongTensor([cap_len])


@torch.no_grad()
def encode_prompt(text_encoder, caption: str, word2code, device=None) -> Tensor:
    """Sentence embedding [1, 256] of one prompt (test_generate.py:60-66 with the missing squeeze)."""
    captions, cap_lens = preprocess_caption(caption, word2code)
    if int(cap_lens[0]) == 0:
        raise ValueError("the prompt has no words")
    device = next(text_encoder.parameters()).device if device is None else device
    return text_encoder(captions[:, :, 0].to(device), cap_lens.to(device))
