"""The storage-format emulation (oracle/dfgan_emul.py) with rounding disabled must be the fp32 oracle in a
different operator order (pool-before-1x1, sub-pixel folded conv) -- checked here on the 64x64 reduced-depth
composition so it runs in seconds on the CPU."""
import torch

from oracle import dfgan_emul as E
from oracle import dfgan_oracle as O


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_emulation_without_rounding_is_the_oracle():
    torch.manual_seed(0)
    g_sd = O.init_generator_state(32, 100, depth=4, gamma_range=(0.2, 1.0))
    d_sd = O.init_discriminator_state(32, depth=4, gamma_range=(0.2, 1.0), weight_scale=3.0)
    images, sent, noise = O.synthetic_batch(2, 64)
    q = E.Rounding(enabled=False)
    gl = {k: v.clone().requires_grad_() for k, v in g_sd.items()}
    dl = {k: v.clone().requires_grad_() for k, v in d_sd.items()}
    img_o, img_e = O.generator_forward(gl, noise, sent), E.generator_forward(gl, noise, sent, q)
    assert _rel(img_e, img_o) < 1e-5
    lo, le = O.g_hinge_loss(dl, img_o, sent), E.g_hinge_loss(dl, img_e, sent, q)
    go = torch.autograd.grad(lo, list(gl.values()))
    ge = torch.autograd.grad(le, list(gl.values()))
    for k, a, b in zip(gl, ge, go):
        assert _rel(a, b) < 2e-4, k
    fake = img_o.detach()
    do = torch.autograd.grad(O.d_hinge_loss(dl, images, fake, sent), list(dl.values()))
    de = torch.autograd.grad(E.d_hinge_loss(dl, images, fake, sent, q), list(dl.values()))
    for k, a, b in zip(dl, de, do):
        assert _rel(a, b) < 2e-4, k
    po = torch.autograd.grad(O.magp_loss(dl, images, sent), list(dl.values()), allow_unused=True)
    pe = torch.autograd.grad(E.magp_loss(dl, images, sent, q), list(dl.values()), allow_unused=True)
    for k, a, b in zip(dl, pe, po):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0, k
        else:
            assert _rel(a, b) < 5e-4, k


def test_rounding_changes_little():
    d_sd = O.init_discriminator_state(32, depth=4, gamma_range=(0.2, 1.0), weight_scale=3.0)
    images, sent, _ = O.synthetic_batch(2, 64)
    with torch.no_grad():
        a = E.discriminator_embeds(d_sd, images, E.Rounding(True))
        b = O.discriminator_embeds(d_sd, images)
    assert 1e-4 < _rel(a, b) < 2e-2
