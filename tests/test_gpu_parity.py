"""GPU parity tests (run with -m gpu on a B200).  Every case drives the CUDA path through the C ABI
(include/dfgan_b200.h) and compares with the oracle on the same seeded inputs:

  * kernel level  -- tools/bringup_conv.py cases: each tcgen05 GEMM variant and pointwise kernel vs torch fp32;
  * network level -- tools/bringup_nets.py cases: per-layer forward and all gradients vs the storage-format
    emulation of the oracle fed with the pipeline's own activations (rel 1e-2, the bf16 tier of north_star; see
    oracle/dfgan_emul.py for why teacher forcing is needed), the MA-GP double backward through the drop-in
    modules, whole training iterations vs the fp32 oracle, and oracle-free properties at the benchmark size.
"""
import os

import numpy as np
import pytest
import torch

from tools import bringup_augment, bringup_configs, bringup_conv, bringup_nets, bringup_parity, bringup_text

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in bringup_conv.CASES if n != "perf"])
def test_kernel(name):
    assert bringup_conv.CASES[name](), name


@pytest.mark.parametrize("name", list(bringup_nets.CASES))
def test_network(name):
    assert bringup_nets.CASES[name](), name


@pytest.mark.parametrize("name", list(bringup_parity.CASES))
def test_full_size_parity(name):
    """BASELINE configs[1] / [3] / [4] at their full batch sizes against the oracle run on the same device in fp32
    (TF32 off), the free-running bf16 gradient bound, and the committed digests of the unmodified reference's step."""
    assert bringup_parity.CASES[name](), name


@pytest.mark.parametrize("name", list(bringup_configs.CASES))
def test_full_size_properties(name):
    """Size-independent properties at batch 64 / 256 (rows of a big batch == the same rows of a small batch, mean of
    chunk gradients, one clipped Adam update)."""
    assert bringup_configs.CASES[name](), name


@pytest.mark.parametrize("name", list(bringup_augment.CASES))
def test_augment(name):
    """Batched image transform: bit-exact vs the committed outputs of the reference's _get_image,
    the integer-exact oracle and the stock torchvision / Pillow pipeline."""
    assert bringup_augment.CASES[name](), name


@pytest.mark.parametrize("name", list(bringup_text.CASES))
def test_text_encoder(name):
    """Frozen caption encoder vs the committed outputs of the reference's RNNEncoder, the fp64
    oracle and the stock torch path."""
    assert bringup_text.CASES[name](), name


@pytest.mark.parametrize("name", ["fprop3_b", "fprop4s2", "dgrad", "epilogue", "subpixel"])
def test_kernel_with_split_k(name):
    """The opt-in split-K path of the generic conv kernel (DFGAN_SPLITK=1 is read once per process, hence the
    subprocess): the low-resolution shapes of these cases split their K loop and run the second epilogue pass."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bringup_conv.py"), name],
                       env=dict(os.environ, DFGAN_SPLITK="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["dfwd", "gfwd", "engine"])
def test_network_with_split_k(name):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bringup_nets.py"), name],
                       env=dict(os.environ, DFGAN_SPLITK="1"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_extension_is_loaded_and_counts_launches():
    from gan_codes_b200 import _lib
    n0 = _lib.launch_count()
    assert bringup_conv.CASES["fprop3_a"]()
    assert _lib.launch_count() > n0
    maps = open("/proc/self/maps").read()
    assert "libdfgan_b200.so" in maps


def test_against_committed_reference_digests():
    """CUDA path vs the digests the UNMODIFIED reference produced (tests/golden/native_b2_gated.npz):
    generator image, discriminator embeds / logits from the same seeded weights and inputs."""
    from oracle import dfgan_oracle as O
    from oracle.make_golden import digest
    from gan_codes_b200.modules import Discriminator, Generator
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "native_b2_gated.npz"))
    g_sd = O.init_generator_state(32, 100, gamma_range=(0.2, 1.0))
    d_sd = O.init_discriminator_state(32, gamma_range=(0.2, 1.0), weight_scale=3.0)
    images, sent, noise = O.synthetic_batch(2, 256)
    G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    with torch.no_grad():
        img = G(noise.cuda(), sent.cuda())
        emb = D.build_embeds(images.cuda())
    for key, t in (("act/G/image", img), ("act/D/embeds", emb)):
        got, want = digest(t.cpu()), gold[key]
        assert abs(got[0] - want[0]) <= 1e-2 * want[0], (key, got[0], want[0])
        rms = want[0] / np.sqrt(t.numel())
        assert np.abs(got[2:] - want[2:]).max() <= 6e-2 * max(rms, np.abs(want[2:]).max()), key


def test_shape_errors_match_reference_behaviour():
    """Non-256 images / wrong sentence width raise RuntimeError like the reference."""
    from gan_codes_b200.modules import Discriminator, Generator
    G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
    with pytest.raises(RuntimeError):
        D.build_embeds(torch.randn(2, 3, 64, 64, device="cuda"))
    with pytest.raises(RuntimeError):
        D.get_logits(torch.randn(2, 512, 4, 4, device="cuda"), torch.randn(2, 128, device="cuda"))
    with pytest.raises(RuntimeError):
        G(torch.randn(2, 64, device="cuda"), torch.randn(2, 256, device="cuda"))


def test_generator_input_gradients_fail_loudly():
    """The tensor-core generator produces parameter gradients only; asking for d/d(noise) or d/d(sentence) raises
    instead of silently returning zeros, while detached inputs and no_grad calls keep working."""
    from gan_codes_b200.modules import Generator
    G = Generator(32, 100).cuda()
    noise, sent = torch.randn(2, 100, device="cuda"), torch.randn(2, 256, device="cuda")
    with pytest.raises(RuntimeError, match="not produced"):
        G(noise.clone().requires_grad_(), sent)
    with pytest.raises(RuntimeError, match="not produced"):
        G(noise, sent.clone().requires_grad_())
    with torch.no_grad():
        assert G(noise.clone().requires_grad_(), sent).shape == (2, 3, 256, 256)
    assert G(noise, sent).requires_grad
