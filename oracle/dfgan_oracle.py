"""TEST INFRASTRUCTURE -- not a product path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Functional restatement (plain torch ops on a state_dict, any float dtype, any device) of the DF-GAN hot
path of Nandan-Upadhyaya/GAN-codes.  Citations are relative to Deep_Fusion_GAN/src/ of the reference.

Parity pin: the reference ships no tests or golden vectors for this path, so the pin
is the reference code itself: oracle/make_golden.py imports the reference's own modules in the build
container, loads the state_dicts produced by `init_generator_state` / `init_discriminator_state` below
(which also proves the key/shape layout of App. A), runs forward, the three losses, all gradients and
the MA-GP double backward, and commits the results under tests/golden/.  tests/test_oracle_golden.py
checks this module against those files on every box.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]
SLOPE = 0.2          # generator/residual_block.py:37 ; discriminator/residual_block.py:12
SENT_DIM = 256       # generator/residual_block.py:10-11 ; discriminator/model.py:49
SEED = 123321        # utils.py:27


# ------------------------------------------------------------------------------------------------ layout
def generator_plan(n_channels: int, depth: int = 6) -> List[Tuple[str, int, int]]:
    """(prefix, C_in, C_out) of every ResidualBlockG in execution order (generator/model.py:15-24).
    depth=6 is the reference network (4 -> 256 px); smaller depths keep the first `depth` up-stages
    and derive the output block from the last stage's width (64x64 config: depth=4)."""
    c = n_channels
    widths = [(8 * c, 8 * c)] * 4 + [(8 * c, 4 * c), (4 * c, 2 * c)]
    plan = [(f"res_blocks.{i}", a, b) for i, (a, b) in enumerate(widths[:depth])]
    last = plan[-1][2]
    plan.append(("res_block_out", last, last // 2))
    return plan


def discriminator_plan(n_c: int, depth: int = 6) -> List[Tuple[str, int, int]]:
    """(prefix, C_in, C_out) of every ResidualBlockD (discriminator/model.py:15-25).  Reduced depths keep
    the LAST `depth` width transitions (the trunk must still end at 16*n_c channels on a 4x4 grid for the
    head, discriminator/model.py:29-31) and feed the first kept block from the n_c-wide stem."""
    mult = [1, 2, 4, 8, 16, 16, 16]
    pairs = [(n_c * mult[i], n_c * mult[i + 1]) for i in range(6)][6 - depth:]
    pairs[0] = (n_c, pairs[0][1])
    return [(f"img_forward.{k + 1}", a, b) for k, (a, b) in enumerate(pairs)]


def _uniform(gen, shape, bound, dtype):
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)


def _normal(gen, shape, std, dtype):
    return (torch.randn(shape, generator=gen, dtype=torch.float64) * std).to(dtype)


def _conv_init(sd, gen, name, cout, cin, k, bias, dtype):
    # torch.nn.Conv2d default: kaiming_uniform(a=sqrt(5)) -> U(+-1/sqrt(fan_in)) for weight and bias
    bound = 1.0 / math.sqrt(cin * k * k)
    sd[name + ".weight"] = _uniform(gen, (cout, cin, k, k), bound, dtype)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (cout,), bound, dtype)


def init_generator_state(n_channels: int = 32, latent_dim: int = 100, seed: int = SEED, depth: int = 6,
                         gamma_range: Optional[Tuple[float, float]] = None,
                         dtype=torch.float32) -> State:
    """Random state_dict with the reference's key order, shapes and init *distributions*
    (generator/model.py:13-30, generator/residual_block.py:14-27, generator/fusion_block.py:22-31).
    gamma_range=None keeps the reference's zero-initialised residual gates (gotcha G1)."""
    gen = torch.Generator().manual_seed(seed)
    sd: State = OrderedDict()
    b = 1.0 / math.sqrt(latent_dim)
    c0 = 8 * n_channels
    sd["linear_in.weight"] = _uniform(gen, (c0 * 16, latent_dim), b, dtype)
    sd["linear_in.bias"] = _uniform(gen, (c0 * 16,), b, dtype)
    for prefix, c1, c2 in generator_plan(n_channels, depth):
        g = 0.0 if gamma_range is None else float(
            gamma_range[0] + (gamma_range[1] - gamma_range[0]) * torch.rand((), generator=gen))
        sd[prefix + ".gamma"] = torch.full((1,), g, dtype=dtype)

        def affine(idx, width):
            for mlp in ("gamma_mlp", "beta_mlp"):
                p = f"{prefix}.fusion_block_{idx}.{mlp}"
                sd[p + ".0.weight"] = _normal(gen, (256, SENT_DIM), math.sqrt(2.0 / (256 + SENT_DIM)), dtype)
                sd[p + ".0.bias"] = torch.zeros(256, dtype=dtype)
                sd[p + ".2.weight"] = _normal(gen, (width, 256), math.sqrt(2.0 / (256 + width)), dtype)
                sd[p + ".2.bias"] = torch.zeros(width, dtype=dtype)

        affine(1, c1)
        affine(2, c1)
        _conv_init(sd, gen, prefix + ".conv_1", c2, c1, 3, True, dtype)
        affine(3, c2)
        affine(4, c2)
        _conv_init(sd, gen, prefix + ".conv_2", c2, c2, 3, True, dtype)
        if c1 != c2:
            _conv_init(sd, gen, prefix + ".scale_conv", c2, c1, 1, True, dtype)
    c_last = generator_plan(n_channels, depth)[-1][2]
    _conv_init(sd, gen, "conv_out.1", 3, c_last, 3, True, dtype)
    return sd


def init_discriminator_state(n_c: int = 32, seed: int = SEED + 1, depth: int = 6,
                             gamma_range: Optional[Tuple[float, float]] = None, weight_scale: float = 1.0,
                             dtype=torch.float32) -> State:
    """discriminator/model.py:11-34, discriminator/residual_block.py:9-21.  weight_scale multiplies every
    conv weight (x3 makes the MA-GP norms O(1))."""
    gen = torch.Generator().manual_seed(seed)
    sd: State = OrderedDict()
    _conv_init(sd, gen, "img_forward.0", n_c, 3, 3, True, dtype)
    for prefix, cin, cout in discriminator_plan(n_c, depth):
        g = 0.0 if gamma_range is None else float(
            gamma_range[0] + (gamma_range[1] - gamma_range[0]) * torch.rand((), generator=gen))
        sd[prefix + ".gamma"] = torch.full((1,), g, dtype=dtype)
        _conv_init(sd, gen, prefix + ".residual_conv.0", cout, cin, 4, False, dtype)
        _conv_init(sd, gen, prefix + ".residual_conv.2", cout, cout, 3, False, dtype)
        if cin != cout:
            _conv_init(sd, gen, prefix + ".scale_conv", cout, cin, 1, True, dtype)
    _conv_init(sd, gen, "img_sentence_forward.0", 2 * n_c, 16 * n_c + SENT_DIM, 3, False, dtype)
    _conv_init(sd, gen, "img_sentence_forward.2", 1, 2 * n_c, 4, False, dtype)
    if weight_scale != 1.0:
        for k in sd:
            if k.endswith(".weight"):
                sd[k] = sd[k] * weight_scale
    return sd


def synthetic_batch(batch: int, size: int = 256, latent_dim: int = 100, seed: int = SEED + 2,
                    dtype=torch.float32):
    """images ~ clamp(N(0,0.5),-1,1), sentence ~ 0.3*N(0,1), noise ~ N(0,1)."""
    gen = torch.Generator().manual_seed(seed)
    images = (torch.randn((batch, 3, size, size), generator=gen) * 0.5).clamp_(-1, 1).to(dtype)
    sent = (torch.randn((batch, SENT_DIM), generator=gen) * 0.3).to(dtype)
    noise = torch.randn((batch, latent_dim), generator=gen).to(dtype)
    return images, sent, noise


# ------------------------------------------------------------------------------------------------ generator
def _mlp(sd: State, p: str, sent):
    # generator/fusion_block.py:8-18 : Linear -> ReLU -> Linear
    h = torch.relu(F.linear(sent, sd[p + ".0.weight"], sd[p + ".0.bias"]))
    return F.linear(h, sd[p + ".2.weight"], sd[p + ".2.bias"])


def affine(sd: State, p: str, x, sent):
    """generator/fusion_block.py:33-39 : gamma(sent) * x + beta(sent), broadcast over H, W."""
    return _mlp(sd, p + ".gamma_mlp", sent)[:, :, None, None] * x + _mlp(sd, p + ".beta_mlp", sent)[:, :, None, None]


def g_block(sd: State, p: str, x, sent, tap: Optional[Callable] = None):
    """generator/residual_block.py:29-59."""
    s = F.conv2d(x, sd[p + ".scale_conv.weight"], sd[p + ".scale_conv.bias"]) if p + ".scale_conv.weight" in sd else x
    h = F.leaky_relu(affine(sd, p + ".fusion_block_1", x, sent), SLOPE)
    h = F.leaky_relu(affine(sd, p + ".fusion_block_2", h, sent), SLOPE)
    h1 = F.conv2d(h, sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"], padding=1)
    if tap:
        tap(p + ".conv_1", h1)
    h = F.leaky_relu(affine(sd, p + ".fusion_block_3", h1, sent), SLOPE)
    h = F.leaky_relu(affine(sd, p + ".fusion_block_4", h, sent), SLOPE)
    h2 = F.conv2d(h, sd[p + ".conv_2.weight"], sd[p + ".conv_2.bias"], padding=1)
    if tap:
        tap(p + ".conv_2", h2)
    y = s + sd[p + ".gamma"] * h2
    if tap:
        tap(p, y)
    return y


def generator_forward(sd: State, noise, sent, tap: Optional[Callable] = None):
    """generator/model.py:32-47.  The block list is read from the state_dict so reduced-depth
    compositions (64x64 config) run through the same code."""
    n_up = len({k.split(".")[1] for k in sd if k.startswith("res_blocks.")})
    c0 = sd["linear_in.weight"].shape[0] // 16
    x = F.linear(noise, sd["linear_in.weight"], sd["linear_in.bias"]).view(noise.shape[0], c0, 4, 4)
    for i in range(n_up):
        x = F.interpolate(g_block(sd, f"res_blocks.{i}", x, sent, tap), scale_factor=2)
    x = g_block(sd, "res_block_out", x, sent, tap)
    img = torch.tanh(F.conv2d(F.leaky_relu(x, SLOPE), sd["conv_out.1.weight"], sd["conv_out.1.bias"], padding=1))
    if tap:
        tap("conv_out", img)
    return img


# ------------------------------------------------------------------------------------------------ discriminator
def d_block(sd: State, p: str, x, tap: Optional[Callable] = None):
    """discriminator/residual_block.py:23-30."""
    s = F.conv2d(x, sd[p + ".scale_conv.weight"], sd[p + ".scale_conv.bias"]) if p + ".scale_conv.weight" in sd else x
    s = F.avg_pool2d(s, 2)
    r = F.conv2d(x, sd[p + ".residual_conv.0.weight"], None, stride=2, padding=1)
    if tap:
        tap(p + ".residual_conv.0", r)
    r = F.conv2d(F.leaky_relu(r, SLOPE), sd[p + ".residual_conv.2.weight"], None, padding=1)
    if tap:
        tap(p + ".residual_conv.2", r)
    r = F.leaky_relu(r, SLOPE)
    y = s + sd[p + ".gamma"] * r
    if tap:
        tap(p, y)
    return y


def discriminator_embeds(sd: State, image, tap: Optional[Callable] = None):
    """discriminator/model.py:36-41 (no activation after the stem conv, gotcha G8)."""
    x = F.conv2d(image, sd["img_forward.0.weight"], sd["img_forward.0.bias"], padding=1)
    if tap:
        tap("img_forward.0", x)
    k = 1
    while f"img_forward.{k}.gamma" in sd:
        x = d_block(sd, f"img_forward.{k}", x, tap)
        k += 1
    return x


def discriminator_logits(sd: State, embed, sent):
    """discriminator/model.py:43-59 : tile the sentence over the 4x4 grid, concat, conv3x3, LReLU, conv4x4."""
    tiled = sent.reshape(-1, SENT_DIM, 1, 1).expand(-1, -1, 4, 4)
    h = torch.cat((embed, tiled), dim=1)
    h = F.leaky_relu(F.conv2d(h, sd["img_sentence_forward.0.weight"], None, padding=1), SLOPE)
    return F.conv2d(h, sd["img_sentence_forward.2.weight"], None)


# ------------------------------------------------------------------------------------------------ step math
def d_hinge_loss(d_sd: State, images, fake_images, sent):
    """deep_fusion_gan/model.py:173-189 (fake_images must already be detached)."""
    b = images.shape[0]
    real_emb = discriminator_embeds(d_sd, images)
    loss_real = torch.relu(1.0 - discriminator_logits(d_sd, real_emb, sent)).mean()
    loss_mis = torch.relu(1.0 + discriminator_logits(d_sd, real_emb[:b - 1], sent[1:b])).mean()
    fake_emb = discriminator_embeds(d_sd, fake_images)
    loss_fake = torch.relu(1.0 + discriminator_logits(d_sd, fake_emb, sent)).mean()
    return loss_real + (loss_fake + loss_mis) / 2.0


def magp_norms(d_sd: State, images, sent, create_graph: bool = True):
    """deep_fusion_gan/model.py:59-85 : clamp(sqrt(|dlogit/dx|^2 + |dlogit/ds|^2 + 1e-8), 0, 10) per sample."""
    x = images.detach().clone().requires_grad_()
    s = sent.detach().clone().requires_grad_()
    logits = discriminator_logits(d_sd, discriminator_embeds(d_sd, x), s)
    gx, gs = torch.autograd.grad(logits, (x, s), torch.ones_like(logits), retain_graph=True,
                                 create_graph=create_graph)
    g = torch.cat((gx.reshape(x.shape[0], -1), gs.reshape(x.shape[0], -1)), dim=1)
    return torch.clamp(torch.sqrt((g ** 2).sum(dim=1) + 1e-8), 0, 10), (gx, gs)


def magp_loss(d_sd: State, images, sent, probe: Optional[dict] = None):
    """deep_fusion_gan/model.py:202-203 : 2 * mean(n^6)."""
    n, (gx, gs) = magp_norms(d_sd, images, sent)
    if probe is not None:
        probe.update(gp_n=n.detach(), gp_gx=gx.detach(), gp_gs=gs.detach())
    return 2.0 * torch.mean(n ** 6)


def g_hinge_loss(d_sd: State, fake_images, sent):
    """deep_fusion_gan/model.py:215-217."""
    return -discriminator_logits(d_sd, discriminator_embeds(d_sd, fake_images), sent).mean()


def text_image_loss(fake_images, sent):
    """damsm/loss.py:4-26 with the sentence embeddings given (the reference re-runs its frozen, deterministic text
    encoder on the same captions at :14, which returns the embeddings the iteration already holds): global average
    pool of the image (:12), zero-padding of the 3 channel means to the embedding width (F.linear with eye, :16-18),
    F.normalize of both (:20-21), loss = 1 - mean cosine (:23-25).  Logged by the trainer, never differentiated
    (deep_fusion_gan/model.py:220-223, 231).  Pinned by tests/golden/txtimg_loss.json (oracle/make_shell_golden.py)."""
    feat = fake_images.mean(dim=(2, 3))
    e = sent.shape[1]
    if feat.shape[1] != e:
        pad = torch.zeros(feat.shape[0], e, dtype=feat.dtype, device=feat.device)
        n = min(e, feat.shape[1])
        pad[:, :n] = feat[:, :n]
        feat = pad
    feat = feat / feat.norm(dim=1, keepdim=True).clamp_min(1e-12)
    s = sent / sent.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return 1 - (feat * s).sum(dim=1).mean()


def clip_coef(grads: List[torch.Tensor], max_norm: float = 5.0):
    """torch.nn.utils.clip_grad_norm_ as used at deep_fusion_gan/model.py:196,209,227 (L2, all tensors)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    return torch.clamp(max_norm / (total + 1e-6), max=1.0), total


class AdamState:
    """torch.optim.Adam(lr, betas=(0, 0.9), eps=1e-8) restated (deep_fusion_gan/model.py:42-43)."""

    def __init__(self, sd: State, lr: float, beta1: float = 0.0, beta2: float = 0.9, eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()}

    def step(self, sd: State, grads: Dict[str, torch.Tensor]) -> None:
        self.t += 1
        c1 = 1.0 - self.b1 ** self.t
        c2 = 1.0 - self.b2 ** self.t
        with torch.no_grad():
            for k, g in grads.items():
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                denom = (self.v[k].sqrt() / math.sqrt(c2)).add_(self.eps)
                sd[k].addcdiv_(self.m[k], denom, value=-self.lr / c1)


def _grads(loss, sd: State, zero_if_unused: bool) -> Dict[str, torch.Tensor]:
    keys = list(sd)
    gs = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    out = {}
    for k, g in zip(keys, gs):
        if g is None:
            if not zero_if_unused:
                continue
            g = torch.zeros_like(sd[k])
        out[k] = g
    return out


def _leaf(sd: State) -> State:
    return OrderedDict((k, v.detach().requires_grad_()) for k, v in sd.items())


def train_step(g_sd: State, d_sd: State, g_opt: AdamState, d_opt: AdamState, images, sent, noise,
               grad_hook: Optional[Callable] = None, probe: Optional[dict] = None):
    """One iteration of deep_fusion_gan/model.py:173-228 without the text encoder / damsm (gotcha G6):
    D hinge update, MA-GP update (biases get zero grads, gotcha G3), G update through the twice-updated D.
    g_sd / d_sd are updated in place; returns the three losses.  grad_hook(phase, grads) lets a
    data-parallel caller average the gradients before clipping."""
    losses = {}
    # --- D step
    gl, dl = _leaf(g_sd), _leaf(d_sd)
    fake = generator_forward(gl, noise, sent)
    d_loss = d_hinge_loss(dl, images, fake.detach(), sent)
    grads = _grads(d_loss, dl, zero_if_unused=False)
    if grad_hook:
        grads = grad_hook("d", grads)
    coef, _ = clip_coef(list(grads.values()))
    d_opt.step(d_sd, {k: g * coef for k, g in grads.items()})
    losses["d"] = d_loss.detach()
    if probe is not None:
        probe["fake"] = fake.detach()
    # --- MA-GP step
    dl = _leaf(d_sd)
    gp = magp_loss(dl, images, sent, probe)
    grads = _grads(gp, dl, zero_if_unused=True)
    if grad_hook:
        grads = grad_hook("gp", grads)
    coef, _ = clip_coef(list(grads.values()))
    d_opt.step(d_sd, {k: g * coef for k, g in grads.items()})
    losses["gp"] = gp.detach()
    # --- G step (graph of `fake` from the D step is reused, gotcha G4)
    g_loss = g_hinge_loss(d_sd, fake, sent)
    grads = _grads(g_loss, gl, zero_if_unused=False)
    if grad_hook:
        grads = grad_hook("g", grads)
    coef, _ = clip_coef(list(grads.values()))
    g_opt.step(g_sd, {k: g * coef for k, g in grads.items()})
    losses["g"] = g_loss.detach()
    return losses


def to(sd: State, **kw) -> State:
    return OrderedDict((k, v.to(**kw)) for k, v in sd.items())
