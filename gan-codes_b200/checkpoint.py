"""Checkpoint-format compatibility with the reference trainer.

The reference writes `checkpoint.pt` = {epoch, generator, discriminator, g_optimizer, d_optimizer, g_losses,
d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses} (deep_fusion_gan/model.py:298-319), a numbered copy
every 10 epochs (:318-319) and the bare generator state_dict as `gen_{epoch}.pth` (:324-327); it resumes from the
same dict (:106-133).  `generator` / `discriminator` are the modules' state_dicts (kept key for key by the
drop-in modules); `g_optimizer` / `d_optimizer` are `torch.optim.Adam.state_dict()`s, whose per-parameter
`exp_avg` / `exp_avg_sq` / `step` entries are indexed in `module.parameters()` order.

The engine keeps the Adam moments of a whole network in two flat fp32 buffers laid out like its flat parameter
buffer (engine.FlatState) and one device step counter.  The functions here convert between that layout and the
torch.optim.Adam dictionary, so a checkpoint written by either trainer resumes in the other.  They are pure
tensor bookkeeping (slices and copies, any device); no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch

# (parameter name, offset into the flat buffers, shape) in module.named_parameters() order
Layout = List[Tuple[str, int, torch.Size]]

HISTORY_KEYS = ("g_losses", "d_losses", "d_gp_losses", "is_scores", "fid_scores", "txtimg_losses")


def build_layout(named_shapes: Sequence[Tuple[str, torch.Size]], align: int = 4) -> Tuple[Layout, int]:
    """Offsets of every parameter in the flat buffers: each one starts on an `align`-element boundary (16 bytes,
    the vector width of the Adam / norm kernels).  Returns (layout, total elements)."""
    layout, o = [], 0
    for name, shape in named_shapes:
        n = 1
        for d in shape:
            n *= int(d)
        layout.append((name, o, torch.Size(shape)))
        o += ((n + align - 1) // align) * align
    return layout, o


def adam_state_dict(layout: Layout, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
                    betas: Tuple[float, float], eps: float) -> dict:
    """The flat moments as a `torch.optim.Adam.state_dict()` (the format of deep_fusion_gan/model.py:305-306).
    Like torch, an optimiser that has not stepped yet has an empty `state`."""
    state: Dict[int, dict] = {}
    if step > 0:
        for i, (_, o, shape) in enumerate(layout):
            n = shape.numel()
            state[i] = {"step": torch.tensor(float(step), dtype=torch.float32),
                        "exp_avg": m[o:o + n].view(shape).clone(),
                        "exp_avg_sq": v[o:o + n].view(shape).clone()}
    group = {"lr": lr, "betas": (float(betas[0]), float(betas[1])), "eps": eps, "weight_decay": 0,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
             "fused": None, "decoupled_weight_decay": False, "params": list(range(len(layout)))}
    return {"state": state, "param_groups": [group]}


def load_adam_state_dict(sd: dict, layout: Layout, m: torch.Tensor, v: torch.Tensor) -> dict:
    """Fill the flat moments from a `torch.optim.Adam.state_dict()`; returns {step, lr, betas, eps}.

    Error behaviour follows torch.optim.Optimizer.load_state_dict: ValueError on a different number of parameter
    groups or parameters.  Settings the fused Adam kernel does not implement (weight decay, amsgrad, maximize)
    are rejected instead of silently ignored; per-parameter step counts must agree (they always do for an
    optimiser that steps all of its parameters together, as the reference's do)."""
    groups = sd["param_groups"]
    if len(groups) != 1:
        raise ValueError("loaded state dict has a different number of parameter groups")
    g = groups[0]
    if len(g["params"]) != len(layout):
        raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                         "optimizer's group")
    if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
        raise ValueError("unsupported Adam settings in checkpoint (weight_decay / amsgrad / maximize); the "
                         "reference trainer uses none of them (deep_fusion_gan/model.py:42-43)")
    state = sd["state"]
    # validate everything first: a rejected checkpoint must leave the live moments untouched
    steps, plan = set(), []
    for i, (name, o, shape) in enumerate(layout):
        st = state.get(g["params"][i])
        if st is None:          # parameter never stepped: zero moments, like a fresh torch optimiser
            steps.add(0)
            continue
        for key in ("exp_avg", "exp_avg_sq"):
            if tuple(st[key].shape) != tuple(shape):
                raise ValueError(f"optimizer state of parameter {i} ({name}): {key} has shape "
                                 f"{tuple(st[key].shape)}, expected {tuple(shape)}")
        steps.add(int(float(st["step"])))
        plan.append((o, shape, st))
    if len(steps) > 1:
        raise ValueError(f"per-parameter Adam step counts differ ({sorted(steps)}); the flat optimiser keeps one")
    m.zero_()
    v.zero_()
    for o, shape, st in plan:
        n = shape.numel()
        m[o:o + n].view(shape).copy_(st["exp_avg"])
        v[o:o + n].view(shape).copy_(st["exp_avg_sq"])
    return {"step": steps.pop() if steps else 0, "lr": float(g["lr"]),
            "betas": (float(g["betas"][0]), float(g["betas"][1])), "eps": float(g["eps"])}


def make_checkpoint(epoch: int, generator_sd: dict, discriminator_sd: dict, g_optimizer_sd: dict,
                    d_optimizer_sd: dict, histories: Dict[str, list]) -> dict:
    """The dictionary of deep_fusion_gan/model.py:300-312, same keys in the same order."""
    unknown = set(histories) - set(HISTORY_KEYS)
    if unknown:
        raise KeyError(f"unknown metric histories {sorted(unknown)}; the checkpoint holds {HISTORY_KEYS}")
    ck = {"epoch": epoch, "generator": generator_sd, "discriminator": discriminator_sd,
          "g_optimizer": g_optimizer_sd, "d_optimizer": d_optimizer_sd}
    for k in HISTORY_KEYS:
        ck[k] = list(histories.get(k, []))
    return ck


def write_checkpoint(ck: dict, save_dir: str) -> List[str]:
    """`checkpoint.pt`, `checkpoint_epoch_{e}.pt` when (e + 1) % 10 == 0, and `gen_{e}.pth` (:315-327).
    Returns the paths written."""
    os.makedirs(save_dir, exist_ok=True)
    e = ck["epoch"]
    paths = [os.path.join(save_dir, "checkpoint.pt")]
    if (e + 1) % 10 == 0:
        paths.append(os.path.join(save_dir, f"checkpoint_epoch_{e}.pt"))
    for p in paths:
        torch.save(ck, p)
    gen = os.path.join(save_dir, f"gen_{e}.pth")
    torch.save(ck["generator"], gen)
    return paths + [gen]


def _safe_globals() -> list:
    """What a checkpoint of the reference's trainer holds besides tensors and builtins: numpy scalars in the loss /
    score histories (np.mean(...) results, deep_fusion_gan/model.py:154-162)."""
    import numpy as np
    out = [np.dtype, np.ndarray]
    try:
        from numpy._core.multiarray import _reconstruct, scalar
    except ImportError:  # numpy < 2
        from numpy.core.multiarray import _reconstruct, scalar
    out += [scalar, _reconstruct]
    out += [type(np.dtype(t)) for t in ("float64", "float32", "int64", "int32", "bool")]
    return out


def read_checkpoint(path: str, map_location=None, trusted: bool = False) -> dict:
    """FileNotFoundError with the reference's message when the file is missing (:108-109).  The file is read with
    torch's restricted unpickler (tensors, containers, numpy scalars); `trusted=True` (or DFGAN_TRUSTED_CHECKPOINTS=1)
    falls back to the unrestricted pickle load the reference uses -- only for files you wrote yourself."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Checkpoint file not found: {path}")
    if trusted or os.environ.get("DFGAN_TRUSTED_CHECKPOINTS", "0") == "1":
        ck = torch.load(path, map_location=map_location, weights_only=False)
    else:
        with torch.serialization.safe_globals(_safe_globals()):
            ck = torch.load(path, map_location=map_location, weights_only=True)
    for k in ("epoch", "generator", "discriminator", "g_optimizer", "d_optimizer"):
        if k not in ck:
            raise KeyError(k)
    return ck


def histories_of(ck: dict) -> Tuple[list, ...]:
    """(g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses), missing ones empty (:123-128)."""
    return tuple(ck.get(k, []) for k in HISTORY_KEYS)
