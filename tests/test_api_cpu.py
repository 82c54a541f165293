"""CPU-side checks: drop-in API surface (constructor signatures, attributes, state_dict layout -- SURVEY
App. A / 8(b)), the C-ABI library loads and exports every symbol include/dfgan_b200.h declares, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import gan_codes_b200  # noqa: F401
from gan_codes_b200 import _lib, convops
from gan_codes_b200._lib_decl import EXPORTS
from gan_codes_b200.modules import AffineBlock, Discriminator, Generator, ResidualBlockD, ResidualBlockG
from oracle import dfgan_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_reference():
    G, D = Generator(32, 100), Discriminator(32)
    g_ref, d_ref = O.init_generator_state(), O.init_discriminator_state()
    assert list(G.state_dict().keys()) == list(g_ref.keys())
    assert list(D.state_dict().keys()) == list(d_ref.keys())
    for k, v in G.state_dict().items():
        assert v.shape == g_ref[k].shape and v.dtype == torch.float32, k
    for k, v in D.state_dict().items():
        assert v.shape == d_ref[k].shape and v.dtype == torch.float32, k
    G.load_state_dict(g_ref, strict=True)
    D.load_state_dict(d_ref, strict=True)


def test_default_init_semantics():
    """gamma gates start at zero, AffineBlock biases at zero (generator/residual_block.py:27, fusion_block.py:22-31)."""
    G, D = Generator(32, 100), Discriminator(32)
    for k, v in list(G.state_dict().items()) + list(D.state_dict().items()):
        if k.endswith(".gamma"):
            assert float(v.abs().max()) == 0.0
        if "_mlp." in k and k.endswith(".bias"):
            assert float(v.abs().max()) == 0.0


def test_attribute_surface():
    G, D = Generator(32, 100), Discriminator(32)
    assert isinstance(G.linear_in, torch.nn.Linear) and len(G.res_blocks) == 6
    assert isinstance(G.res_block_out, ResidualBlockG) and isinstance(G.conv_out[1], torch.nn.Conv2d)
    assert isinstance(G.res_blocks[4].fusion_block_1, AffineBlock) and G.res_blocks[4].scale_conv is not None
    assert G.res_blocks[0].scale_conv is None
    assert isinstance(D.img_forward[1], ResidualBlockD) and len(D.img_forward) == 7
    assert D.img_sentence_forward[0].weight.shape == (64, 768, 3, 3)
    assert callable(D.build_embeds) and callable(D.get_logits)
    b = ResidualBlockG(64, 32, affine_h_d=256, df_1_c_in=256, df_2_c_in=256)
    assert b.conv_1.weight.shape == (32, 64, 3, 3)
    a = AffineBlock(input_dim=256, hidden_dim=256, output_dim=64)
    assert a.gamma_mlp[2].weight.shape == (64, 256)


def test_no_cpu_fallback():
    G, D = Generator(32, 100), Discriminator(32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        G(torch.randn(1, 100), torch.randn(1, 256))
    with pytest.raises(RuntimeError, match="no CPU path"):
        D.build_embeds(torch.randn(1, 3, 256, 256))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "dfgan_b200.h")).read()
    declared = set(re.findall(r"\b(dfgan_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dfgan_arch() == b"sm_100a"
    assert ctypes.sizeof(_lib.ConvDesc) > 0


def test_tap_tables():
    """Host logic of the tap-list builders: stride-2 split, sub-pixel folding covers each 3x3 tap once per phase."""
    assert len(convops.taps_4x4s2(64)) == 16 and len({t[3] for t in convops.taps_4x4s2(64)}) == 16
    for a in (0, 1):
        for c in (0, 1):
            assert len(convops.taps_4x4s2_dgrad(32, a, c)) == 4
    slots = convops.subpixel_slots()
    assert len(slots) == 16
    for ph in range(4):
        covered = sorted(t for s in slots[ph * 4:(ph + 1) * 4] for t in s)
        assert covered == list(range(9))
    assert len(convops.taps_subpixel_dgrad(64)) == 16


def test_precision_tier_switch_and_no_cpu_path():
    """set_precision selects the bf16 fast path or the fp32 verification tier; neither has a CPU path, and the
    fp32 tier keeps the reference's state_dict layout (it runs on the same parameter containers)."""
    G, D = Generator(32, 100), Discriminator(32)
    keys_g, keys_d = list(G.state_dict()), list(D.state_dict())
    assert G.set_precision("fp32") is G and D.set_precision("fp32") is D
    assert list(G.state_dict()) == keys_g and list(D.state_dict()) == keys_d
    with pytest.raises(ValueError):
        G.set_precision("fp16")
    for prec in ("fp32", "bf16"):
        G.set_precision(prec), D.set_precision(prec)
        with pytest.raises(RuntimeError):
            G(torch.randn(2, 100), torch.randn(2, 256))
        with pytest.raises(RuntimeError):
            D.build_embeds(torch.randn(2, 3, 256, 256))
        with pytest.raises(RuntimeError):
            D.get_logits(torch.randn(2, 512, 4, 4), torch.randn(2, 256))


def test_side_stream_helpers_are_inert_without_cuda_work():
    """convops.side / join are pure scheduling helpers: disabled they must not touch any CUDA state."""
    was = convops.SIDE_ENABLED
    convops.SIDE_ENABLED = False
    try:
        with convops.side(0, torch.device("cpu")):
            pass
        convops.join(torch.device("cpu"))
    finally:
        convops.SIDE_ENABLED = was
