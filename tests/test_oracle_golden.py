"""Pins oracle/dfgan_oracle.py to the reference: every digest in tests/golden/*.npz was produced by the
UNMODIFIED reference modules (oracle/make_golden.py, run in the build container); the oracle must
reproduce all of them from the same seeds."""
import os

import numpy as np
import pytest
import torch

from oracle import dfgan_oracle as O
from oracle.make_golden import CASES, digest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(name, got, want, rtol=2e-4):
    got, want = np.asarray(got), np.asarray(want)
    scale = max(abs(want[0]), 1e-30)  # l2 norm of the tensor
    assert abs(got[0] - want[0]) <= rtol * scale + 1e-12, (name, got[0], want[0])
    tol = rtol * max(np.abs(want[2:]).max(), 1e-30) + 1e-12
    assert np.abs(got[2:] - want[2:]).max() <= 5 * tol, (name, np.abs(got[2:] - want[2:]).max(), tol)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_digests(name):
    n_ch, depth, batch, gr, ws, steps, dt = CASES[name]
    dt = getattr(torch, dt)
    tight = 1e-9 if dt == torch.float64 else None
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g_sd = O.init_generator_state(n_ch, 100, depth=depth, gamma_range=gr, dtype=dt)
    d_sd = O.init_discriminator_state(n_ch, depth=depth, gamma_range=gr, weight_scale=ws, dtype=dt)
    images, sent, noise = O.synthetic_batch(batch, 4 * 2 ** depth, dtype=dt)
    seen = {}
    with torch.no_grad():
        img = O.generator_forward(g_sd, noise, sent, tap=lambda k, v: seen.__setitem__("act/G/" + k, v))
        emb = O.discriminator_embeds(d_sd, images, tap=lambda k, v: seen.__setitem__("act/D/" + k, v))
        seen["act/G/image"] = img
        seen["act/D/embeds"] = emb
        seen["act/D/logits"] = O.discriminator_logits(d_sd, emb, sent)
        seen["act/D/logits_fake"] = O.discriminator_logits(d_sd, O.discriminator_embeds(d_sd, img), sent)
    checked = 0
    for k, v in seen.items():
        if k in gold.files:
            _close(k, digest(v), gold[k], rtol=tight or 2e-4)
            checked += 1
    assert checked > 25
    g_opt, d_opt = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    for s in range(steps):
        probe, rec = {}, {}

        def hook(phase, grads):
            rec[phase] = grads
            return grads
        losses = O.train_step(g_sd, d_sd, g_opt, d_opt, images, sent, noise, grad_hook=hook, probe=probe)
        tag = f"step{s}"
        rt = tight or 5e-4
        for ph, pre in (("d", "d_grad"), ("gp", "gp_grad"), ("g", "g_grad")):
            for k, g in rec[ph].items():
                _close(f"{tag}/{pre}/{k}", digest(g), gold[f"{tag}/{pre}/{k}"], rtol=rt)
                checked += 1
        for k in ("d", "gp", "g"):
            _close(f"{tag}/{k}_loss", digest(losses[k]), gold[f"{tag}/{k}_loss"], rtol=rt)
        for k in ("fake", "gp_n", "gp_gx", "gp_gs"):
            _close(f"{tag}/{k}", digest(probe[k]), gold[f"{tag}/{k}"], rtol=rt)
    if tight is None:
        return  # fp32: a +-lr Adam update of a near-zero gradient is rounding-chaotic; fp64 case pins Adam
    for k, v in g_sd.items():
        _close("final/G/" + k, digest(v), gold["final/G/" + k], rtol=tight or 2e-4)
    for k, v in d_sd.items():
        _close("final/D/" + k, digest(v), gold["final/D/" + k], rtol=tight or 2e-4)


def test_oracle_matches_reference_fp64_phase_digests():
    """tests/golden/native_b2_fp64_phases.npz: the reference's modules in float64 at the native size, the three
    phases' losses and every gradient from the SAME weights (oracle/make_golden.make_phases).  The oracle reproduces
    all 337 digests to 1e-9; the GPU tests hold the fp32 tier to 1e-4 and the bf16 path to its tier against them."""
    gold = np.load(os.path.join(GOLD, "native_b2_fp64_phases.npz"))
    ws, batch = float(gold["config"][0]), int(gold["config"][1])
    dt = torch.float64
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g_sd = O.init_generator_state(32, 100, gamma_range=(0.2, 1.0), dtype=dt)
    d_sd = O.init_discriminator_state(32, gamma_range=(0.2, 1.0), weight_scale=ws, dtype=dt)
    images, sent, noise = O.synthetic_batch(batch, 256, dtype=dt)
    gl, dl = O._leaf(g_sd), O._leaf(d_sd)
    fake = O.generator_forward(gl, noise, sent)
    d_loss = O.d_hinge_loss(dl, images, fake.detach(), sent)
    got = {"phase/d_loss": d_loss, "phase/fake": fake}
    got.update({"phase/d_grad/" + k: g for k, g in O._grads(d_loss, dl, zero_if_unused=False).items()})
    probe = {}
    dl2 = O._leaf(d_sd)
    gp = O.magp_loss(dl2, images, sent, probe)
    got.update({"phase/gp_loss": gp, "phase/gp_n": probe["gp_n"], "phase/gp_gx": probe["gp_gx"],
                "phase/gp_gs": probe["gp_gs"]})
    got.update({"phase/gp_grad/" + k: g for k, g in O._grads(gp, dl2, zero_if_unused=True).items()})
    g_loss = O.g_hinge_loss(d_sd, fake, sent)
    got["phase/g_loss"] = g_loss
    got.update({"phase/g_grad/" + k: g for k, g in O._grads(g_loss, gl, zero_if_unused=False).items()})
    assert set(got) == set(gold.files) - {"config"}
    for k, v in got.items():
        _close(k, digest(v), gold[k], rtol=1e-9)


def test_state_layout_counts():
    """269 generator tensors / 12 240 010 params, 30 discriminator tensors / 19 645 254."""
    g = O.init_generator_state()
    d = O.init_discriminator_state()
    assert len(g) == 269 and sum(v.numel() for v in g.values()) == 12_240_010
    assert len(d) == 30 and sum(v.numel() for v in d.values()) == 19_645_254


def test_default_init_known_answers():
    """gamma == 0 makes every block its shortcut; the GP step gives biases zero grads."""
    d = O.init_discriminator_state()
    images, sent, _ = O.synthetic_batch(1)
    dl = {k: v.clone().requires_grad_() for k, v in d.items()}
    loss = O.magp_loss(dl, images, sent)
    g_bias, = torch.autograd.grad(loss, [dl["img_forward.1.scale_conv.bias"]], allow_unused=True)
    assert g_bias is None or float(g_bias.abs().max()) == 0.0
