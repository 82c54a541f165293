"""Drop-in nn.Module API of the reference's generator / discriminator (same constructor signatures, attribute
names and state_dict keys -- SURVEY App. A) running on the sm_100a kernels.

Reference classes mirrored (Deep_Fusion_GAN/src/):
    generator/model.py:8-47          Generator(n_channels, latent_dim).forward(noise_vector, sentence_embed)
    generator/residual_block.py:9    ResidualBlockG(df_1_c_out, df_2_c_out, affine_h_d, df_1_c_in, df_2_c_in)
    generator/fusion_block.py:5      AffineBlock(input_dim, hidden_dim, output_dim)
    discriminator/model.py:8-59      Discriminator(n_c, sentence_embed_dim).build_embeds / get_logits
    discriminator/residual_block.py:7 ResidualBlockD(in_channels, out_channels)

Autograd contract: outputs carry grad_fn, `.backward()` fills every parameter's `.grad`, and
`torch.autograd.grad(logits, (image, sentence), create_graph=True)` returns tensors whose backward is the
MA-GP double backward (deep_fusion_gan/model.py:59-85, 202-210), implemented as a tangent-forward sweep
(SURVEY App. C-3).  Parameters are fp32 in the reference layout; compute is bf16 tensor-core with fp32
accumulation.  CUDA only: calling a module whose parameters are not on a CUDA device raises.

Precision tiers: the default "bf16" is the fast path above.  `module.set_precision("fp32")` runs the same module
graph on the CUDA-core fp32 convolution kernels (gan_codes_b200.fp32), which agree with the reference's fp32
(TF32-off) path to 1e-4 -- the verification tier of north_star.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn as nn

from . import dnet, fp32, gnet
from . import kernels as K

F32 = torch.float32
KEEP_DEBUG = False   # tests set this to keep the last forward's activation dicts on the module (`_debug`)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: gan_codes_b200 runs on sm_100a CUDA devices only (got {t.device}); "
                           "there is no CPU path")


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


class _PrecisionMixin:
    _precision = "bf16"

    def set_precision(self, precision: str):
        """"bf16" (tensor-core fast path, default) or "fp32" (CUDA-core verification tier, gan_codes_b200.fp32)."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self._precision = precision
        return self


class _FlatGrads:
    """One zero-filled fp32 buffer with a named view per parameter (a single memset instead of one per tensor)."""

    def __init__(self, names: List[str], params: List[torch.Tensor]):
        sizes = [((p.numel() + 3) // 4) * 4 for p in params]   # keep every view 16-byte aligned
        self.flat = torch.zeros(sum(sizes), device=params[0].device, dtype=F32)
        self.views: Dict[str, torch.Tensor] = {}
        o = 0
        for n, p, s in zip(names, params, sizes):
            self.views[n] = self.flat[o:o + p.numel()].view(p.shape)
            o += s


class _PackCache:
    def __init__(self, cls):
        self.cls, self.key, self.packed = cls, None, None

    def get(self, names, tensors):
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self.packed is None or key != self.key:
            params = dict(zip(names, tensors))
            with torch.no_grad():
                if self.packed is None or tuple(k[0] for k in key) != tuple(k[0] for k in (self.key or ())):
                    self.packed = self.cls(params)
                else:
                    self.packed.refresh(params)
            self.key = key
        return self.packed


# =============================================================================================== generator
class AffineBlock(nn.Module):
    """generator/fusion_block.py:5-39.  Holds the two MLPs' parameters; inside Generator the 56 MLPs run as two
    batched launches, the broadcast gamma*x+beta is fused into the neighbouring convolution kernels."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        self.gamma_mlp = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(inplace=True),
                                       nn.Linear(hidden_dim, output_dim))
        self.beta_mlp = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.ReLU(inplace=True),
                                      nn.Linear(hidden_dim, output_dim))
        for mlp in (self.gamma_mlp, self.beta_mlp):       # generator/fusion_block.py:22-31
            for m in mlp:
                if isinstance(m, nn.Linear):
                    nn.init.xavier_normal_(m.weight)
                    nn.init.constant_(m.bias, 0.0)


class ResidualBlockG(nn.Module):
    """generator/residual_block.py:9-59 (parameter container; executed by the fused generator program)."""

    def __init__(self, df_1_c_out: int, df_2_c_out: int, affine_h_d: int = 256, df_1_c_in: int = 256,
                 df_2_c_in: int = 256):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1))
        self.fusion_block_1 = AffineBlock(df_1_c_in, affine_h_d, df_1_c_out)
        self.fusion_block_2 = AffineBlock(df_1_c_in, affine_h_d, df_1_c_out)
        self.conv_1 = nn.Conv2d(df_1_c_out, df_2_c_out, 3, 1, 1)
        self.fusion_block_3 = AffineBlock(df_2_c_in, affine_h_d, df_2_c_out)
        self.fusion_block_4 = AffineBlock(df_2_c_in, affine_h_d, df_2_c_out)
        self.conv_2 = nn.Conv2d(df_2_c_out, df_2_c_out, 3, 1, 1)
        self.scale_conv = None
        if df_1_c_out != df_2_c_out:
            self.scale_conv = nn.Conv2d(df_1_c_out, df_2_c_out, 1, 1, 0)


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, noise, sent, *tensors):
        names = module._names
        params = dict(zip(names, tensors))
        pk = module._cache.get(names, tensors)
        need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
        ctx.need = any(ctx.needs_input_grad[3:])
        keep = ctx.need
        if keep:
            fg = _FlatGrads(names, list(tensors))
            mlp = gnet.GMlp(params, pk.plan, noise.shape[0], fg.views)
            save: dict = {}
            img = gnet.generator_forward(pk, params, mlp, _as_f32(noise), _as_f32(sent), save)
            ctx.state = (module, pk, params, mlp, save, fg)
        else:
            mlp = module._inference_mlp(params, pk, noise.shape[0])
            img = gnet.generator_forward(pk, params, mlp, _as_f32(noise), _as_f32(sent), None)
        del need_grad
        return img

    @staticmethod
    def backward(ctx, g_img):
        module, pk, params, mlp, save, fg = ctx.state
        fg.flat.zero_()
        gnet.generator_backward(pk, params, mlp, save, _as_f32(g_img), fg.views)
        # d/d(noise) and d/d(sentence) are not produced: the trainer feeds detached inputs
        # (deep_fusion_gan/model.py:171,182)
        return (None, None, None, *[fg.views[n] for n in module._names])


class Generator(_PrecisionMixin, nn.Module):
    """generator/model.py:8-47."""

    def __init__(self, n_channels: int, latent_dim: int):
        super().__init__()
        self.n_channels = n_channels
        c = n_channels
        self.linear_in = nn.Linear(latent_dim, c * 8 * 4 * 4)
        self.res_blocks = nn.ModuleList([
            ResidualBlockG(8 * c, 8 * c), ResidualBlockG(8 * c, 8 * c), ResidualBlockG(8 * c, 8 * c),
            ResidualBlockG(8 * c, 8 * c), ResidualBlockG(8 * c, 4 * c), ResidualBlockG(4 * c, 2 * c)])
        self.res_block_out = ResidualBlockG(2 * c, 1 * c)
        self.conv_out = nn.Sequential(nn.LeakyReLU(0.2, inplace=True), nn.Conv2d(c, 3, 3, 1, 1), nn.Tanh())
        self._cache = _PackCache(gnet.GPacked)
        self._inf = {}

    @property
    def _names(self) -> List[str]:
        return [n for n, _ in self.named_parameters()]

    def _inference_mlp(self, params, pk, batch):
        key = (batch, tuple(t.data_ptr() for t in params.values()))
        if key not in self._inf:
            self._inf = {key: gnet.GMlp(params, pk.plan, batch, None)}
        return self._inf[key]

    def forward(self, noise_vector: torch.Tensor, sentence_embed: torch.Tensor) -> torch.Tensor:
        tensors = [p for _, p in self.named_parameters()]
        _require_cuda(tensors[0], "Generator.forward")
        if self._precision == "fp32":
            return fp32.generator_forward(self, noise_vector, sentence_embed)
        if torch.is_grad_enabled() and (noise_vector.requires_grad or sentence_embed.requires_grad):
            # the tensor-core program produces parameter gradients only (the trainer feeds detached inputs,
            # deep_fusion_gan/model.py:171,182); returning None here would be a silent zero gradient
            raise RuntimeError("Generator.forward (bf16 tensor-core path): gradients with respect to noise_vector / "
                               "sentence_embed are not produced; detach the inputs, or use "
                               "set_precision('fp32'), whose autograd graph reaches them")
        return _GeneratorFn.apply(self, noise_vector, sentence_embed, *tensors)


# =============================================================================================== discriminator
class ResidualBlockD(nn.Module):
    """discriminator/residual_block.py:7-30 (parameter container; executed by the fused trunk program)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1))
        self.residual_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(out_channels, out_channels, 3, 1, 1, bias=False), nn.LeakyReLU(0.2, inplace=True))
        self.scale_conv = None
        if in_channels != out_channels:
            self.scale_conv = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


def _zeros_like_all(tensors):
    return [torch.zeros_like(t) for t in tensors]


class _EmbedsFn(torch.autograd.Function):
    """image -> embeds (fp32 NCHW at both ends)."""

    @staticmethod
    def forward(ctx, module, image, *tensors):
        names = module._trunk_names
        params = dict(zip(names, tensors))
        pk = module._cache.get(module._all_names, module._all_tensors())
        save: dict = {}
        img = _as_f32(image)
        emb = dnet.trunk_forward(pk, params, img, save)
        if KEEP_DEBUG:
            module._debug["trunk"] = save
        ctx.state = (module, pk, params, save, img)
        ctx.save_for_backward(*tensors)
        return dnet.embeds_to_nchw(emb)

    @staticmethod
    def backward(ctx, g_emb):
        module, pk, params, save, img = ctx.state
        need_gx = ctx.needs_input_grad[1]
        need_w = any(ctx.needs_input_grad[2:])
        # grad mode is on here exactly when the caller asked for create_graph=True (MA-GP first sweep)
        second = torch.is_grad_enabled()
        out = _EmbedsBwdFn.apply(module, pk, save, need_gx, need_w, second, g_emb, img, *ctx.saved_tensors)
        return (None, out[0], *out[1:])


class _EmbedsBwdFn(torch.autograd.Function):
    """First-order backward of the trunk as a differentiable op: its own backward is the tangent sweep."""

    @staticmethod
    def forward(ctx, module, pk, save, need_gx, need_w, second, g_emb, img, *tensors):
        names = module._trunk_names
        params = dict(zip(names, tensors))
        fg = _FlatGrads(names, list(tensors)) if need_w else None
        cots: dict = {} if second else None
        gx = dnet.trunk_backward(pk, params, save, dnet.embeds_to_nhwc(_as_f32(g_emb)),
                                 fg.views if fg else None, need_gx, cots)
        ctx.state = (module, pk, params, save, cots)
        ctx.n_t = len(tensors)
        outs = [gx if gx is not None else torch.zeros_like(img)]
        outs += [fg.views[n] for n in names] if fg else _zeros_like_all(tensors)
        ctx.mark_non_differentiable(*outs[1:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, gg_x, *_unused):
        module, pk, params, save, cots = ctx.state
        if cots is None:
            raise RuntimeError("double backward through build_embeds needs create_graph=True on the first "
                               "autograd.grad call (deep_fusion_gan/model.py:68-75)")
        names = module._trunk_names
        fg = _FlatGrads(names, [params[n] for n in names])   # biases stay exactly zero (gotcha G3)
        t_emb = dnet.trunk_tangent(pk, params, save, cots, _as_f32(gg_x), None, fg.views)
        return (None, None, None, None, None, None, dnet.embeds_to_nchw(t_emb), None,
                *[fg.views[n] for n in names])


class _LogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, embed, sent, w0, w2):
        params = {"img_sentence_forward.0.weight": w0, "img_sentence_forward.2.weight": w2}
        pk = module._cache.get(module._all_names, module._all_tensors())
        save: dict = {}
        logits = dnet.head_forward(pk, params, _as_f32(embed), _as_f32(sent), save)
        if KEEP_DEBUG:
            module._debug["head"] = save
        ctx.state = (module, pk, params, save)
        ctx.save_for_backward(w0, w2)
        return logits.view(-1, 1, 1, 1)

    @staticmethod
    def backward(ctx, g_logits):
        module, pk, params, save = ctx.state
        second = torch.is_grad_enabled()
        out = _LogitsBwdFn.apply(module, pk, save, second, g_logits, *ctx.saved_tensors)
        return (None, *out)


class _LogitsBwdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pk, save, second, g_logits, w0, w2):
        params = {"img_sentence_forward.0.weight": w0, "img_sentence_forward.2.weight": w2}
        names = list(params)
        fg = _FlatGrads(names, [w0, w2])
        cots: dict = {} if second else None
        g_embed, g_sent = dnet.head_backward(pk, params, save, _as_f32(g_logits).reshape(-1), fg.views, True, cots)
        ctx.state = (pk, params, save, cots)
        outs = (g_embed, g_sent, fg.views[names[0]], fg.views[names[1]])
        ctx.mark_non_differentiable(outs[2], outs[3])
        return outs

    @staticmethod
    def backward(ctx, gg_embed, gg_sent, *_unused):
        pk, params, save, cots = ctx.state
        if cots is None:
            raise RuntimeError("double backward through get_logits needs create_graph=True")
        names = list(params)
        fg = _FlatGrads(names, [params[n] for n in names])
        b = save["b"]
        dev = save["cat"].device
        if gg_embed is None:
            gg_embed = torch.zeros(b, pk.ccat - dnet.SENT_DIM, 4, 4, device=dev)
        if gg_sent is None:
            gg_sent = torch.zeros(b, dnet.SENT_DIM, device=dev)
        t_logit = dnet.head_tangent(pk, params, save, cots, _as_f32(gg_embed), _as_f32(gg_sent), None, fg.views)
        return (None, None, None, None, t_logit.view(-1, 1, 1, 1), fg.views[names[0]], fg.views[names[1]])


class Discriminator(_PrecisionMixin, nn.Module):
    """discriminator/model.py:8-59."""

    def __init__(self, n_c: int, sentence_embed_dim: int = 256):
        super().__init__()
        self.img_forward = nn.Sequential(
            nn.Conv2d(3, n_c, 3, 1, 1),
            ResidualBlockD(n_c * 1, n_c * 2), ResidualBlockD(n_c * 2, n_c * 4), ResidualBlockD(n_c * 4, n_c * 8),
            ResidualBlockD(n_c * 8, n_c * 16), ResidualBlockD(n_c * 16, n_c * 16),
            ResidualBlockD(n_c * 16, n_c * 16))
        self.img_sentence_forward = nn.Sequential(
            nn.Conv2d(16 * n_c + sentence_embed_dim, n_c * 2, 3, 1, 1, bias=False),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(n_c * 2, 1, 4, 1, 0, bias=False))
        self._cache = _PackCache(dnet.DPacked)
        self._debug = {}

    @property
    def _all_names(self):
        return [n for n, _ in self.named_parameters()]

    def _all_tensors(self):
        return [p.detach() for _, p in self.named_parameters()]

    @property
    def _trunk_names(self):
        return [n for n, _ in self.named_parameters() if n.startswith("img_forward.")]

    def build_embeds(self, image: torch.Tensor) -> torch.Tensor:
        tensors = [p for n, p in self.named_parameters() if n.startswith("img_forward.")]
        _require_cuda(tensors[0], "Discriminator.build_embeds")
        if self._precision == "fp32":
            return fp32.build_embeds(self, image)
        return _EmbedsFn.apply(self, image, *tensors)

    def get_logits(self, image_embed: torch.Tensor, sentence_embed: torch.Tensor) -> torch.Tensor:
        w0, w2 = self.img_sentence_forward[0].weight, self.img_sentence_forward[2].weight
        _require_cuda(w0, "Discriminator.get_logits")
        if self._precision == "fp32":
            return fp32.get_logits(self, image_embed, sentence_embed)
        return _LogitsFn.apply(self, image_embed, sentence_embed, w0, w2)

    def forward(self, image, sentence_embed):   # convenience; the reference trainer never calls it
        return self.get_logits(self.build_embeds(image), sentence_embed)


_ = (math, K)
