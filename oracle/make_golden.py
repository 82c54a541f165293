"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference modules
(/root/reference/Deep_Fusion_GAN/src, imported read-only) on state_dicts and inputs produced by
oracle/dfgan_oracle.py.  Run in the build container only:  python oracle/make_golden.py
The GPU box has no /root/reference; tests there read the committed .npz files.

Each golden file holds, per named tensor, a digest = [l2 norm, mean, 32 strided samples] so the files stay
small while still pinning every activation / gradient the parity tests look at.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/Deep_Fusion_GAN"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import dfgan_oracle as O  # noqa: E402

N_SAMPLES = 32


def digest(t: torch.Tensor) -> np.ndarray:
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, N_SAMPLES).long()
    return np.concatenate([[f.norm().item(), f.mean().item()], f[idx].numpy()])


def ref_modules(g_sd, d_sd, depth):
    """Reference modules carrying the oracle-initialised weights.  depth=6: the reference's own Generator /
    Discriminator classes.  depth<6 (64x64 config): the reference's *block classes* assembled at reduced
    depth -- the reference cannot express that size itself."""
    from src.discriminator.model import Discriminator
    from src.discriminator.residual_block import ResidualBlockD
    from src.generator.model import Generator
    from src.generator.residual_block import ResidualBlockG
    n_c = d_sd["img_forward.0.weight"].shape[0]
    n_ch = g_sd["conv_out.1.weight"].shape[1] if depth == 6 else g_sd["linear_in.weight"].shape[0] // 16 // 8
    latent = g_sd["linear_in.weight"].shape[1]
    G = Generator(n_ch, latent)
    D = Discriminator(n_c)
    if depth != 6:
        plan = O.generator_plan(n_ch, depth)
        G.res_blocks = nn.ModuleList([ResidualBlockG(a, b) for (_, a, b) in plan[:-1]])
        G.res_block_out = ResidualBlockG(plan[-1][1], plan[-1][2])
        G.conv_out[1] = nn.Conv2d(plan[-1][2], 3, 3, 1, 1)
        dplan = O.discriminator_plan(n_c, depth)
        D.img_forward = nn.Sequential(D.img_forward[0], *[ResidualBlockD(a, b) for (_, a, b) in dplan])
    dt = g_sd["linear_in.weight"].dtype
    G.to(dt)
    D.to(dt)
    G.load_state_dict(g_sd, strict=True)
    D.load_state_dict(d_sd, strict=True)
    assert list(G.state_dict().keys()) == list(g_sd.keys()), "generator key order differs from App. A"
    assert list(D.state_dict().keys()) == list(d_sd.keys()), "discriminator key order differs"
    return G, D


def ref_compute_gp(D, images, sent):
    """Verbatim call sequence of deep_fusion_gan/model.py:59-85 on the reference module."""
    b = images.shape[0]
    x = images.data.requires_grad_()
    s = sent.data.requires_grad_()
    logits = D.get_logits(D.build_embeds(x), s)
    grads = torch.autograd.grad(outputs=logits, inputs=(x, s), grad_outputs=torch.ones_like(logits),
                                retain_graph=True, create_graph=True)
    g = torch.cat((grads[0].reshape(b, -1), grads[1].reshape(b, -1)), dim=1)
    return torch.clamp(torch.sqrt(torch.sum(g ** 2, dim=1) + 1e-8), 0, 10), grads


def ref_step(G, D, g_opt, d_opt, images, sent, noise, out, tag):
    """deep_fusion_gan/model.py:173-228 on the reference modules with injected inputs."""
    b = images.shape[0]
    real_embeds = D.build_embeds(images)
    d_loss_real = F.relu(1.0 - D.get_logits(real_embeds, sent)).mean()
    d_loss_mis = F.relu(1.0 + D.get_logits(real_embeds[:b - 1], sent[1:b])).mean()
    fake = G(noise, sent)
    d_loss_fake = F.relu(1.0 + D.get_logits(D.build_embeds(fake.detach()), sent)).mean()
    d_loss = d_loss_real + (d_loss_fake + d_loss_mis) / 2.0
    d_opt.zero_grad(); g_opt.zero_grad()
    d_loss.backward()
    for k, p in D.named_parameters():
        out[f"{tag}/d_grad/{k}"] = digest(p.grad)
    out[f"{tag}/d_gnorm"] = digest(torch.nn.utils.clip_grad_norm_(D.parameters(), 5.0))
    d_opt.step()
    out[f"{tag}/d_loss"] = digest(d_loss)
    out[f"{tag}/fake"] = digest(fake)

    n, grads = ref_compute_gp(D, images, sent)
    gp = 2.0 * torch.mean(n ** 6)
    out[f"{tag}/gp_n"] = digest(n)
    out[f"{tag}/gp_gx"] = digest(grads[0])
    out[f"{tag}/gp_gs"] = digest(grads[1])
    d_opt.zero_grad(); g_opt.zero_grad()
    gp.backward()
    for k, p in D.named_parameters():
        assert p.grad is not None, k  # gotcha G3: zero tensors, never None
        out[f"{tag}/gp_grad/{k}"] = digest(p.grad)
    out[f"{tag}/gp_gnorm"] = digest(torch.nn.utils.clip_grad_norm_(D.parameters(), 5.0))
    d_opt.step()
    out[f"{tag}/gp_loss"] = digest(gp)

    g_loss = -D.get_logits(D.build_embeds(fake), sent).mean()
    d_opt.zero_grad(); g_opt.zero_grad()
    g_loss.backward()
    for k, p in G.named_parameters():
        out[f"{tag}/g_grad/{k}"] = digest(p.grad)
    out[f"{tag}/g_gnorm"] = digest(torch.nn.utils.clip_grad_norm_(G.parameters(), 5.0))
    g_opt.step()
    out[f"{tag}/g_loss"] = digest(g_loss)


CASES = {
    # name: (n_channels, depth, batch, gamma_range, d_weight_scale, steps, dtype)
    "native_b2_gated": (32, 6, 2, (0.2, 1.0), 3.0, 1, "float32"),
    "native_b2_default_init": (32, 6, 2, None, 1.0, 1, "float32"),
    "reduced64_b4_gated": (32, 4, 4, (0.2, 1.0), 3.0, 1, "float32"),
    # fp64 so that three consecutive optimiser steps can be pinned tightly (in fp32, Adam with beta1 = 0
    # amplifies rounding noise of near-zero gradients into +-lr updates and trajectories only track to ~5 %)
    "reduced64_b2_fp64_3steps": (32, 4, 2, (0.2, 1.0), 3.0, 3, "float64"),
}


def make(name):
    n_ch, depth, batch, gr, ws, steps, dt = CASES[name]
    dt = getattr(torch, dt)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    g_sd = O.init_generator_state(n_ch, 100, depth=depth, gamma_range=gr, dtype=dt)
    d_sd = O.init_discriminator_state(n_ch, depth=depth, gamma_range=gr, weight_scale=ws, dtype=dt)
    size = 4 * 2 ** depth
    images, sent, noise = O.synthetic_batch(batch, size, dtype=dt)
    G, D = ref_modules(g_sd, d_sd, depth)
    out = {}
    # forward activations through hooks on the reference modules
    hooks = []
    for net, pre in ((G, "G"), (D, "D")):
        for mname, m in net.named_modules():
            if isinstance(m, nn.Conv2d) or type(m).__name__.startswith("ResidualBlock"):
                hooks.append(m.register_forward_hook(
                    lambda mod, i, o, key=f"act/{pre}/{mname}": out.__setitem__(key, digest(o))))
    with torch.no_grad():
        img = G(noise, sent)
        emb = D.build_embeds(images)
        out["act/G/image"] = digest(img)
        out["act/D/embeds"] = digest(emb)
        out["act/D/logits"] = digest(D.get_logits(emb, sent))
    for h in hooks:
        h.remove()
    with torch.no_grad():
        out["act/D/logits_fake"] = digest(D.get_logits(D.build_embeds(img), sent))
    g_opt = torch.optim.Adam(G.parameters(), lr=0.0001, betas=(0.0, 0.9))
    d_opt = torch.optim.Adam(D.parameters(), lr=0.0004, betas=(0.0, 0.9))
    for s in range(steps):
        ref_step(G, D, g_opt, d_opt, images, sent, noise, out, f"step{s}")
    for k, v in G.state_dict().items():
        out[f"final/G/{k}"] = digest(v)
    for k, v in D.state_dict().items():
        out[f"final/D/{k}"] = digest(v)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, len(out), "tensors ->", path, os.path.getsize(path) // 1024, "KiB")


def make_phases(name="native_b2_fp64_phases", ws=2.0, batch=2):
    """fp64 arbiter for the per-phase gradients at the native size: the reference's modules in float64, every phase
    of deep_fusion_gan/model.py:173-228 evaluated FROM THE SAME WEIGHTS (no optimiser step in between), so a GPU path
    can be held to 1e-4 on each phase without the +-lr sign chaos of Adam(beta1 = 0) on noise-level gradient
    elements separating the two trajectories.  Same call sequence on the modules as ref_step."""
    dt = torch.float64
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    g_sd = O.init_generator_state(32, 100, gamma_range=(0.2, 1.0), dtype=dt)
    d_sd = O.init_discriminator_state(32, gamma_range=(0.2, 1.0), weight_scale=ws, dtype=dt)
    images, sent, noise = O.synthetic_batch(batch, 256, dtype=dt)
    G, D = ref_modules(g_sd, d_sd, 6)
    out = {"config": np.array([ws, batch], dtype=np.float64)}
    b = batch
    real_embeds = D.build_embeds(images)
    d_loss_real = F.relu(1.0 - D.get_logits(real_embeds, sent)).mean()
    d_loss_mis = F.relu(1.0 + D.get_logits(real_embeds[:b - 1], sent[1:b])).mean()
    fake = G(noise, sent)
    d_loss_fake = F.relu(1.0 + D.get_logits(D.build_embeds(fake.detach()), sent)).mean()
    d_loss = d_loss_real + (d_loss_fake + d_loss_mis) / 2.0
    D.zero_grad(); G.zero_grad()
    d_loss.backward()
    for k, p in D.named_parameters():
        out[f"phase/d_grad/{k}"] = digest(p.grad)
    out["phase/d_loss"], out["phase/fake"] = digest(d_loss), digest(fake)
    n, grads = ref_compute_gp(D, images, sent)
    gp = 2.0 * torch.mean(n ** 6)
    out["phase/gp_n"], out["phase/gp_gx"], out["phase/gp_gs"] = digest(n), digest(grads[0]), digest(grads[1])
    D.zero_grad(); G.zero_grad()
    gp.backward()
    for k, p in D.named_parameters():
        out[f"phase/gp_grad/{k}"] = digest(p.grad)
    out["phase/gp_loss"] = digest(gp)
    g_loss = -D.get_logits(D.build_embeds(fake), sent).mean()
    D.zero_grad(); G.zero_grad()
    g_loss.backward()
    for k, p in G.named_parameters():
        out[f"phase/g_grad/{k}"] = digest(p.grad)
    out["phase/g_loss"] = digest(g_loss)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, len(out), "tensors ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES) + ["native_b2_fp64_phases"]):
        make_phases() if n == "native_b2_fp64_phases" else make(n)
