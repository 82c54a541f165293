"""Drop-in `DeepFusionGAN` trainer shell (deep_fusion_gan/model.py:20-387) over the fused
iteration of engine.DFGANTrainer, the frozen text encoder (text_encoder.RNNEncoder) and the resident IS/FID
(evaluation.py).

Same constructor `DeepFusionGAN(n_words, encoder_weights_path, image_save_path, gen_path_save)`, the same public
attributes (`generator`, `discriminator`, `text_encoder`, `ixtoword`, `g_optim`, `d_optim`, `grad_clip_g/d`,
`device`) and methods (`fit`, `load_from_checkpoint`, `compute_is_fid`, `compute_text_image_loss`, `_compute_gp`,
`_check_nan`, `_save_checkpoint`, `_save_gen_weights`, `_save_fake_image[_with_prompt]`) with the reference's
argument meaning, return values, printed lines and files.  `train.py` of the reference runs unchanged against it.

What differs from the reference's loop, and why (each one is invisible in the returned histories):

  * one iteration is `DFGANTrainer.step` (one CUDA-graph replay), not ~3000 eager launches with eight host
    synchronisations (`.item()` x3, `_check_nan` x4, the pageable `noise.to(device)`; :182-231).  The losses are
    accumulated on the device in fp64 and read ONCE per epoch; the epoch means are the same numbers the reference
    gets from `np.mean` of its per-iteration `.item()` lists.
  * `_check_nan` (:99-104) swaps a non-finite loss for a fresh random leaf tensor, whose `.backward()` then raises
    (it has no grad_fn): the reference aborts on the first non-finite loss.  Here the check runs on the epoch sums
    (the one host read): a non-finite loss raises RuntimeError at the end of that epoch, before any checkpoint is
    written.  `_check_nan` itself is kept, with the reference's behaviour, for callers that use it.
  * noise is drawn exactly like the reference's `torch.randn(batch_size, 100)` -- on the host, from the global
    generator, in the same order (one draw per training batch, one per evaluation batch), so a seeded run sees the
    same noise -- but into pinned memory and copied asynchronously.
  * the text-image loss (`compute_text_image_loss` -> damsm/loss.py) is computed inside the iteration from the
    sentence embeddings it already holds (one 2-launch kernel, kernels.txtimg_loss); the reference runs its frozen
    encoder a second time on the same captions for the same embeddings.
  * the evaluation pass keeps real and generated images on the device (the reference moves 2 x 768 images to the
    host and back) and uses the resident Inception pass of evaluation.compute_is_fid.
  * batches must have the constant size of the first one (the reference's loaders are built with `drop_last=True`,
    utils.py:24); a different size raises.

Data parallel (the reference has none): `DeepFusionGAN(..., world=N)` inside an initialised
torch.distributed process group.  Every rank feeds its own shard of the data (e.g. a DistributedSampler in its loader;
all ranks must see the same number of batches); the three gradient all-reduces per iteration happen inside
DFGANTrainer; the epoch loss sums are averaged over the ranks with one more tiny all-reduce, so every rank holds the
same histories and a non-finite loss stops all ranks together; IS/FID are computed by every rank on its own evaluation
shard and rank 0's values are the ones saved; only rank 0 prints and writes files.
"""
from __future__ import annotations

import os
from typing import Callable, List

import torch
import torch.nn as nn
from torch import Tensor

from . import checkpoint as ckpt
from . import evaluation
from .modules import Discriminator, Generator
from .text_encoder import RNNEncoder

LOSS_KEYS = ("g_loss", "d_loss", "gp_loss", "txtimg_loss")


def prepare_data(batch, device: torch.device):
    """objects/utils.py:7-20: sort the batch by decreasing caption length, move it to the device."""
    images, captions, captions_len, file_names = batch
    sorted_cap_lens, sorted_cap_indices = torch.sort(captions_len, 0, True)
    sorted_cap_lens = sorted_cap_lens.to(device)
    sorted_images = images[sorted_cap_indices].to(device, non_blocking=True)
    sorted_captions = captions[sorted_cap_indices].squeeze().to(device)
    sorted_file_names = [file_names[i] for i in sorted_cap_indices.numpy()]
    return sorted_images, sorted_captions, sorted_cap_lens, sorted_file_names


class _PinnedStage:
    """Two alternating pinned host buffers of one shape: `fill(buf)` writes the next buffer on the host, the copy to
    the device is asynchronous, and a buffer is rewritten only after the copy issued from it has completed."""

    def __init__(self, shape, dtype, device: torch.device):
        self.device = device
        self.slots = [[torch.empty(shape, dtype=dtype).pin_memory(), torch.cuda.Event()] for _ in range(2)]
        self.i = 0

    def push(self, fill: Callable[[Tensor], None]) -> Tensor:
        buf, ev = self.slots[self.i]
        self.i ^= 1
        ev.synchronize()
        fill(buf)
        dev = buf.to(self.device, non_blocking=True)
        ev.record()
        return dev

    def fill_only(self, fill: Callable[[Tensor], None]):
        """(pinned buffer, its event): the consumer copies from the buffer itself and records the event on the stream
        it copied on (DFGAN_STAGE_MIXED: DFGANTrainer.step stages host tensors on its own copy stream)."""
        buf, ev = self.slots[self.i]
        self.i ^= 1
        ev.synchronize()
        fill(buf)
        return buf, ev


class _PendingAdam:
    """`g_optim` / `d_optim` before the first batch fixes the batch size (the flat optimiser lives in
    DFGANTrainer): keeps a loaded `torch.optim.Adam.state_dict()` until the trainer exists."""

    def __init__(self, lr: float, betas=(0.0, 0.9)):
        self.lr, self.betas, self.pending = lr, betas, None

    def load_state_dict(self, sd: dict) -> None:
        self.pending = sd

    def state_dict(self) -> dict:
        if self.pending is not None:
            return self.pending
        raise RuntimeError("the optimiser state exists once the first batch has been seen (fit) or after "
                           "load_from_checkpoint")


class DeepFusionGAN:
    def __init__(self, n_words, encoder_weights_path: str, image_save_path: str, gen_path_save: str, *,
                 world: int = 1):
        self.image_save_path = image_save_path
        self.gen_path_save = gen_path_save
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")           # :26
        self.generator = Generator(n_channels=32, latent_dim=100).to(self.device)
        self.discriminator = Discriminator(n_c=32).to(self.device)
        self.text_encoder = RNNEncoder.load(encoder_weights_path, n_words)
        self.text_encoder.to(self.device)
        self.ixtoword = None                                                                    # set from outside (:35)
        for p in self.text_encoder.parameters():
            p.requires_grad = False
        self.text_encoder.eval()
        self.g_optim = _PendingAdam(0.0001)                                                     # :42-43
        self.d_optim = _PendingAdam(0.0004)
        self.relu = nn.ReLU()
        self.grad_clip_g = 5.0
        self.grad_clip_d = 5.0
        self.world = world
        self.rank = torch.distributed.get_rank() if world > 1 else 0
        self._noise_gen = None
        self.trainer = None
        self._pin = {}
        # opt-in, not yet measured: leave the sorted image batch in pinned host memory and let DFGANTrainer.step copy
        # it on its side stream, overlapping the previous iteration
        self._overlap_h2d = os.environ.get("DFGAN_STAGE_MIXED", "0") == "1"
        self._h2d_done = None

    # ------------------------------------------------------------------------------------ engine plumbing
    def _make_trainer(self, batch_size: int):
        """The fused iteration for this batch size (CUDA only: DFGANTrainer raises on CPU modules)."""
        from .engine import DFGANTrainer
        if self.grad_clip_g != self.grad_clip_d or self.grad_clip_g <= 0:
            raise RuntimeError("grad_clip_g and grad_clip_d must be one positive threshold (the reference's 5.0 / "
                               "5.0, deep_fusion_gan/model.py:51-52): the clip is fused into the Adam kernel")
        return DFGANTrainer(self.generator, self.discriminator, batch_size, lr_g=self.g_optim.lr,
                            lr_d=self.d_optim.lr, betas=self.g_optim.betas, clip=self.grad_clip_g,
                            world=self.world, text_image_loss=hasattr(self, "compute_text_image_loss"))

    def _ensure_trainer(self, batch_size: int):
        if self.trainer is None:
            pend_g, pend_d = self.g_optim.pending, self.d_optim.pending
            self.trainer = self._make_trainer(batch_size)
            self.g_optim, self.d_optim = self.trainer.g_optim, self.trainer.d_optim
            if pend_g is not None:
                self.g_optim.load_state_dict(pend_g)
            if pend_d is not None:
                self.d_optim.load_state_dict(pend_d)
        elif self.trainer.batch != batch_size:
            raise RuntimeError(f"batch size changed from {self.trainer.batch} to {batch_size}: the fused iteration is "
                               "built for one batch size (the reference's loaders use drop_last=True, utils.py:24)")
        return self.trainer

    def _draw_noise(self, batch_size: int) -> Tensor:
        """torch.randn(batch_size, 100).to(device) of :182 / :257 -- the same draw from the global host generator,
        staged through two alternating pinned buffers so the copy never blocks the host.  Data parallel: every rank
        draws from its own generator (seed of the global one + rank) -- with the documented fix_seed() on every rank
        the global generators are identical and N ranks would otherwise train on the same latents."""
        gen = None
        if self.world > 1:
            if self._noise_gen is None:
                self._noise_gen = torch.Generator().manual_seed(torch.initial_seed() + self.rank)
            gen = self._noise_gen
        if self.device.type != "cuda":
            return torch.randn(batch_size, 100, generator=gen).to(self.device)
        return self._stage("noise", (batch_size, 100), torch.float32).push(
            lambda buf: torch.randn(batch_size, 100, generator=gen, out=buf))

    def _check_sharded(self, images: Tensor) -> None:
        """Data parallel, first batch of a run: the ranks must not all hold the same images (a loader built without
        a DistributedSampler / rank argument hands every rank the same shuffled batches, and N GPUs then compute one
        gradient N times without anything reporting it)."""
        probe = images.detach().double().flatten()[:: max(1, images.numel() // 4096)].sum().reshape(1).to(self.device)
        seen = [torch.zeros_like(probe) for _ in range(self.world)]
        torch.distributed.all_gather(seen, probe)
        if all(torch.equal(seen[0], t) for t in seen[1:]):
            raise RuntimeError("data parallel: every rank received the same first batch -- shard the loader "
                               "(dataset.create_loader(..., rank=, world=) or a DistributedSampler)")

    def _stage(self, what: str, shape, dtype) -> _PinnedStage:
        key = (what, tuple(shape), dtype)
        if key not in self._pin:
            self._pin[key] = _PinnedStage(shape, dtype, self.device)
        return self._pin[key]

    def _prepare(self, batch, overlap: bool = False):
        """prepare_data (objects/utils.py:7-20) with the image gather written straight into pinned memory and an
        asynchronous copy (25 MB per batch of 32): the host never waits for the transfer."""
        images, captions, captions_len, file_names = batch
        if self.device.type != "cuda" or images.is_cuda or images.dtype != torch.float32:
            return prepare_data(batch, self.device)
        sorted_cap_lens, sorted_cap_indices = torch.sort(captions_len, 0, True)
        stage = self._stage("images", images.shape, images.dtype)
        gather = lambda buf: torch.index_select(images, 0, sorted_cap_indices, out=buf)  # noqa: E731
        if overlap:
            sorted_images, self._h2d_done = stage.fill_only(gather)      # stays on the host; see `_step`
        else:
            sorted_images = stage.push(gather)
        sorted_captions = captions[sorted_cap_indices].squeeze().to(self.device)
        sorted_file_names = [file_names[i] for i in sorted_cap_indices.numpy()]
        return sorted_images, sorted_captions, sorted_cap_lens.to(self.device), sorted_file_names

    # ------------------------------------------------------------------------------------ reference methods
    def _zero_grad(self):
        """:54-56.  Gradients live in the trainer's flat buffers and are cleared inside every phase."""
        if self.trainer is not None:
            self.trainer.gs.zero_grad()
            self.trainer.ds.zero_grad()

    def _compute_gp(self, images: Tensor, sentence_embeds: Tensor) -> Tensor:
        """:58-85 through the drop-in modules' autograd path (create_graph=True).  `fit` does not call it -- the
        MA-GP phase is part of the fused iteration -- it is kept for callers of the reference's method."""
        batch_size = images.shape[0]
        images_interpolated = images.data.requires_grad_()
        sentences_interpolated = sentence_embeds.data.requires_grad_()
        embeds = self.discriminator.build_embeds(images_interpolated)
        logits = self.discriminator.get_logits(embeds, sentences_interpolated)
        grads = torch.autograd.grad(outputs=logits, inputs=(images_interpolated, sentences_interpolated),
                                    grad_outputs=torch.ones_like(logits), retain_graph=True, create_graph=True)
        grad = torch.cat((grads[0].reshape(batch_size, -1), grads[1].reshape(batch_size, -1)), dim=1)
        grad_norm = torch.sqrt(torch.sum(grad ** 2, dim=1) + 1e-8)
        return torch.clamp(grad_norm, 0, 10)

    def compute_text_image_loss(self, fake_images, captions, cap_lens):
        """:87-89 -> damsm/loss.py:4-26 for arbitrary images / captions (one encoder pass + one fused kernel).
        Inside `fit` the same quantity comes out of the iteration itself."""
        from .damsm import damsm_loss
        return damsm_loss(fake_images, captions, cap_lens, self.text_encoder)

    def compute_is_fid(self, fake_imgs, real_imgs):
        """:91-97, both metrics from the resident Inception network (evaluation.compute_is_fid)."""
        return evaluation.compute_is_fid(fake_imgs.float(), real_imgs.float(), cuda=True, batch_size=64, splits=1)

    def _check_nan(self, tensor_value, name=""):
        """:99-104, unchanged behaviour (synchronises; not used inside `fit`, see the module docstring)."""
        if torch.isnan(tensor_value).any() or torch.isinf(tensor_value).any():
            print(f"Warning: {name} contains NaN or Inf values, resetting to small random values")
            return torch.randn_like(tensor_value) * 0.01
        return tensor_value

    def _restore(self, checkpoint: dict) -> None:
        self.generator.load_state_dict(checkpoint["generator"])
        self.discriminator.load_state_dict(checkpoint["discriminator"])
        self.g_optim.load_state_dict(checkpoint["g_optimizer"])
        self.d_optim.load_state_dict(checkpoint["d_optimizer"])
        if self.trainer is not None:
            self.trainer.weights_changed()

    def load_from_checkpoint(self, checkpoint_path: str):
        """:106-133: restores networks and optimisers from a checkpoint of this class or of the reference's;
        returns (epoch, g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses)."""
        checkpoint = ckpt.read_checkpoint(checkpoint_path, map_location=self.device)
        self._restore(checkpoint)
        g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses = ckpt.histories_of(checkpoint)
        print(f"Loaded checkpoint from epoch {checkpoint['epoch']+1}.")
        print(f"Metrics at checkpoint - FID: {fid_scores[-1]:.4f}, IS: {is_scores[-1]:.4f}")
        return checkpoint["epoch"], g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses

    def fit(self, train_loader, test_loader=None, num_epochs: int = 600, start_epoch: int = 0,
            g_losses_epoch=None, d_losses_epoch=None, d_gp_losses_epoch=None,
            is_scores_epoch=None, fid_scores_epoch=None, txtimg_losses_epoch=None,
            auto_resume: bool = True):
        """:135-296."""
        from tqdm.auto import trange
        g_losses_epoch = g_losses_epoch if g_losses_epoch is not None else []
        d_losses_epoch = d_losses_epoch if d_losses_epoch is not None else []
        d_gp_losses_epoch = d_gp_losses_epoch if d_gp_losses_epoch is not None else []
        is_scores_epoch = is_scores_epoch if is_scores_epoch is not None else []
        fid_scores_epoch = fid_scores_epoch if fid_scores_epoch is not None else []
        txtimg_losses_epoch = txtimg_losses_epoch if txtimg_losses_epoch is not None else []

        latest = os.path.join(self.gen_path_save, "checkpoint.pt")
        if auto_resume and start_epoch == 0 and os.path.exists(latest):                        # :148-162
            checkpoint = ckpt.read_checkpoint(latest, map_location=self.device)
            start_epoch = checkpoint["epoch"] + 1
            self._restore(checkpoint)
            g_losses_epoch = checkpoint["g_losses"]
            d_losses_epoch = checkpoint["d_losses"]
            d_gp_losses_epoch = checkpoint["d_gp_losses"]
            is_scores_epoch = checkpoint["is_scores"]
            fid_scores_epoch = checkpoint["fid_scores"]
            txtimg_losses_epoch = checkpoint["txtimg_losses"]
            print(f"Resuming from epoch {start_epoch}")

        say = print if self.rank == 0 else (lambda *a, **k: None)
        for epoch in trange(start_epoch, num_epochs, desc="Train Deep Fusion GAN", disable=self.rank != 0):
            sums, n_iter = None, 0
            sampler = getattr(train_loader, "sampler", None)
            if self.world > 1 and hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)                 # DistributedSampler: a different shuffle every epoch
            for batch in train_loader:
                images, captions, captions_len, _ = self._prepare(batch, overlap=self._overlap_h2d)
                batch_size = images.shape[0]
                if self.world > 1 and epoch == start_epoch and n_iter == 0:
                    self._check_sharded(images)
                trainer = self._ensure_trainer(batch_size)
                sentence_embeds = self.text_encoder(captions, captions_len).detach()            # :171
                noise = self._draw_noise(batch_size)                                            # :182
                out = self._step(trainer, images, sentence_embeds, noise)                       # :173-228
                row = torch.stack([out[k] if k in out else out["g_loss"] * 0 for k in LOSS_KEYS]).double()
                sums = row if sums is None else sums + row
                n_iter += 1
            if n_iter == 0:
                raise RuntimeError("train_loader produced no batches")
            if self.world > 1:
                torch.distributed.all_reduce(sums)
                sums = sums / self.world
            means = (sums / n_iter).tolist()                     # the ONE device->host read of the epoch
            for name, v in zip(("generator loss", "discriminator loss", "gradient penalty loss", "text-image loss"),
                               means):
                if v != v or v in (float("inf"), float("-inf")):
                    raise RuntimeError(f"{name} contains NaN or Inf values in epoch {epoch} (the reference aborts in "
                                       "the backward of the iteration where it happens, deep_fusion_gan/model.py:99-104)")
            g_losses_epoch.append(means[0])
            d_losses_epoch.append(means[1])
            d_gp_losses_epoch.append(means[2])
            txtimg_losses_epoch.append(means[3])

            with torch.no_grad():                                                               # :237-274
                eval_loader = test_loader if test_loader is not None else train_loader
                say(f"Computing metrics on {'test' if test_loader is not None else 'training'} set")
                real_imgs_metric, fake_imgs_metric = [], []
                for i, batch in enumerate(eval_loader):
                    if i >= 32:
                        break
                    images, captions, captions_len, _ = self._prepare(batch)
                    sentence_embeds = self.text_encoder(captions, captions_len).detach()
                    noise = self._draw_noise(images.size(0))
                    fake_images = self._generate(noise, sentence_embeds)
                    real_imgs_metric.append(images)
                    fake_imgs_metric.append(fake_images)
                real_imgs_metric = torch.cat(real_imgs_metric, dim=0)
                fake_imgs_metric = torch.cat(fake_imgs_metric, dim=0)
                is_score, fid_score = 0.0, 0.0
                if hasattr(self, "compute_is_fid"):
                    try:
                        is_score, fid_score = self.compute_is_fid(fake_imgs_metric, real_imgs_metric)
                        say(f"Computed IS: {is_score:.4f}, FID: {fid_score:.4f}")
                    except Exception as e:
                        say(f"Error computing metrics: {e}")
                is_scores_epoch.append(is_score)
                fid_scores_epoch.append(fid_score)

            sync = getattr(self.trainer, "sync_optimizer_state", None)
            if self.world > 1 and sync is not None:
                sync()      # collective; a no-op unless the optimiser step is sharded over the ranks
            if self.rank == 0:
                self._save_fake_image_with_prompt(fake_images, captions, captions_len, epoch)
                self._save_checkpoint(epoch, g_losses_epoch, d_losses_epoch, d_gp_losses_epoch,
                                      is_scores_epoch, fid_scores_epoch, txtimg_losses_epoch)
            if self.world > 1:
                # rank 0 has just written files: the ranks enter the next epoch together (a rank spinning in the first
                # gradient exchange for the length of a checkpoint write would run into the exchange kernels' wait bound)
                torch.distributed.barrier()
            say(f"Epoch {epoch+1}: "
                  f"G Loss: {g_losses_epoch[-1]:.4f}, "
                  f"D Loss: {d_losses_epoch[-1]:.4f}, "
                  f"D GP Loss: {d_gp_losses_epoch[-1]:.4f}, "
                  f"Text-Image Loss: {txtimg_losses_epoch[-1]:.4f}, "
                  f"IS: {is_scores_epoch[-1]:.4f}, "
                  f"FID: {fid_scores_epoch[-1]:.4f}")

        return g_losses_epoch, d_losses_epoch, d_gp_losses_epoch, is_scores_epoch, fid_scores_epoch, txtimg_losses_epoch

    def _step(self, trainer, images: Tensor, sentence_embeds: Tensor, noise: Tensor):
        out = trainer.step(images, sentence_embeds, noise)
        if self._h2d_done is not None:          # images were handed over as a pinned host buffer (DFGAN_STAGE_MIXED)
            self._h2d_done.record(trainer._copy_stream)
            self._h2d_done = None
        return out

    def _generate(self, noise: Tensor, sentence_embeds: Tensor) -> Tensor:
        """self.generator(noise, sentence_embeds) under no_grad (:258) on the trainer's packed weights."""
        return self.trainer.sample(noise, sentence_embeds)

    def _save_checkpoint(self, epoch, g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses):
        """:298-322: checkpoint.pt, checkpoint_epoch_{e}.pt every tenth epoch, gen_{e}.pth."""
        if self.trainer is None:
            raise RuntimeError("nothing to save before the first iteration")
        return self.trainer.save_checkpoint(self.gen_path_save, epoch, g_losses=g_losses, d_losses=d_losses,
                                            d_gp_losses=d_gp_losses, is_scores=is_scores, fid_scores=fid_scores,
                                            txtimg_losses=txtimg_losses)

    def _save_gen_weights(self, epoch: int):
        """:324-327."""
        gen_path = os.path.join(self.gen_path_save, f"gen_{epoch}.pth")
        torch.save({k: v.detach().clone() for k, v in self.generator.state_dict().items()}, gen_path)

    def _caption_text(self, captions, cap_lens, num_images: int) -> List[str]:
        """:346-356, with ONE device->host copy of the captions instead of one `.item()` per word."""
        caps = captions[:num_images].detach().cpu().tolist()
        lens = cap_lens[:num_images].detach().cpu().tolist()
        text = []
        for i in range(num_images):
            sent = ""
            for j in range(lens[i]):
                word_idx = caps[i][j]
                if self.ixtoword is not None and word_idx in self.ixtoword:
                    sent += self.ixtoword[word_idx] + " "
                else:
                    sent += f"[{word_idx}] "
            text.append(sent.strip())
        return text

    def _save_fake_image_with_prompt(self, fake_images: Tensor, captions, cap_lens, epoch: int):
        """:329-381: samples_with_text_epoch_{e}.jpg (needs matplotlib) and fake_sample_epoch_{e}.png; without
        matplotlib only the grid, like the reference's except branch."""
        try:
            import matplotlib.pyplot as plt
            from matplotlib.backends.backend_agg import FigureCanvasAgg
            from matplotlib.figure import Figure
            import numpy as np
            num_images = min(4, fake_images.size(0))
            fig = Figure(figsize=(12, 3 * num_images))
            FigureCanvasAgg(fig)
            original_text = self._caption_text(captions, cap_lens, num_images)
            imgs = ((fake_images[:num_images].detach().float() + 1) / 2).cpu().numpy()
            for i in range(num_images):
                ax = fig.add_subplot(num_images, 1, i + 1)
                ax.imshow(np.clip(imgs[i].transpose(1, 2, 0), 0, 1))
                ax.set_title(f"Prompt: {original_text[i]}", fontsize=10)
                ax.axis("off")
            plt.tight_layout()
            fig.savefig(os.path.join(self.image_save_path, f"samples_with_text_epoch_{epoch}.jpg"))
            self._save_fake_image(fake_images, epoch)
        except Exception as e:
            print(f"Error saving images with text: {e}")
            self._save_fake_image(fake_images, epoch)

    def _save_fake_image(self, fake_images: Tensor, epoch: int):
        """:383-387."""
        import torchvision.utils as vutils
        img_path = os.path.join(self.image_save_path, f"fake_sample_epoch_{epoch}.png")
        vutils.save_image(fake_images.data, img_path, normalize=True)
