"""TEST INFRASTRUCTURE.  Pins the checkpoint format of the reference trainer and cross-checks
gan_codes_b200/checkpoint.py against it.  Run in the build container only (needs /root/reference):

    python oracle/make_checkpoint_golden.py      ->  tests/golden/checkpoint_format.json

DeepFusionGAN cannot be constructed here (its __init__ loads a text-encoder .pth, deep_fusion_gan/model.py:31),
so its UNMODIFIED `_save_checkpoint` / `_save_gen_weights` / `load_from_checkpoint` methods (:106-133, :298-327)
are called unbound on a stub that carries what they touch: the reference Generator / Discriminator, two
torch.optim.Adam built like :42-43 and `gen_path_save`.

What is checked here, with the reference present:
  1. a checkpoint written by the reference loads through checkpoint.load_adam_state_dict into flat moment
     buffers and exports again to a dictionary equal to the reference's, tensor for tensor;
  2. a checkpoint written by checkpoint.make_checkpoint / write_checkpoint is accepted by the reference's own
     load_from_checkpoint and restores the same weights and optimiser state.
What is committed: the STRUCTURE of the reference-written files (keys, order, shapes, dtypes, hyper-parameters,
file names) as a small JSON; tests/test_checkpoint_cpu.py compares our writer against it without the reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/Deep_Fusion_GAN"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


def shape_digest(sd: dict) -> str:
    txt = ";".join(f"{k}:{tuple(v.shape)}:{str(v.dtype)}" for k, v in sd.items())
    return hashlib.sha1(txt.encode()).hexdigest()


def stub(seed: int, save_dir: str, steps: int):
    """Reference networks + optimisers after `steps` iterations' worth of updates on random gradients (two D
    updates and one G update per iteration, each clipped like :196,209,227)."""
    from src.discriminator.model import Discriminator
    from src.generator.model import Generator
    torch.manual_seed(seed)
    s = types.SimpleNamespace()
    s.generator, s.discriminator = Generator(32, 100), Discriminator(32)
    s.g_optim = torch.optim.Adam(s.generator.parameters(), lr=0.0001, betas=(0.0, 0.9))
    s.d_optim = torch.optim.Adam(s.discriminator.parameters(), lr=0.0004, betas=(0.0, 0.9))
    s.gen_path_save = save_dir
    for _ in range(steps):
        for net, opt, n in ((s.discriminator, s.d_optim, 2), (s.generator, s.g_optim, 1)):
            for _ in range(n):
                for p in net.parameters():
                    p.grad = torch.randn_like(p) * 0.01
                torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
                opt.step()
    return s


def same_optim(a: dict, b: dict) -> bool:
    if a["param_groups"] != b["param_groups"] or a["state"].keys() != b["state"].keys():
        return False
    for i in a["state"]:
        for k in ("step", "exp_avg", "exp_avg_sq"):
            if not torch.equal(torch.as_tensor(a["state"][i][k]).float().cpu(),
                               torch.as_tensor(b["state"][i][k]).float().cpu()):
                return False
    return True


def main() -> int:
    from src.deep_fusion_gan.model import DeepFusionGAN as Ref
    from gan_codes_b200 import checkpoint as ck

    hist = dict(g_losses=[0.5, 0.4], d_losses=[1.9, 1.7], d_gp_losses=[0.1, 0.2], is_scores=[1.5, 1.6],
                fid_scores=[300.0, 280.0], txtimg_losses=[3.0, 2.5])
    with tempfile.TemporaryDirectory() as tmp:
        # ---- 1. reference writes (epoch 9 -> also the numbered copy), we read
        s = stub(1, tmp, steps=2)
        s._save_gen_weights = types.MethodType(Ref._save_gen_weights, s)
        Ref._save_checkpoint(s, 9, *[hist[k] for k in ck.HISTORY_KEYS])
        files = sorted(os.listdir(tmp))
        ref_ck = ck.read_checkpoint(os.path.join(tmp, "checkpoint.pt"))
        for net, key in ((s.generator, "g_optimizer"), (s.discriminator, "d_optimizer")):
            layout, total = ck.build_layout([(n, p.shape) for n, p in net.named_parameters()])
            m, v = torch.empty(total), torch.empty(total)
            h = ck.load_adam_state_dict(ref_ck[key], layout, m, v)
            back = ck.adam_state_dict(layout, m, v, h["step"], h["lr"], h["betas"], h["eps"])
            assert same_optim(back, ref_ck[key]), key
            assert h["step"] == (2 if key == "g_optimizer" else 4), h
        assert ck.histories_of(ref_ck) == tuple(hist[k] for k in ck.HISTORY_KEYS)
        gen_only = torch.load(os.path.join(tmp, "gen_9.pth"))
        assert list(gen_only) == list(ref_ck["generator"])

        golden = {
            "files_epoch_9": files,
            "top_level_keys": list(ref_ck),
            "generator": {"n": len(ref_ck["generator"]), "digest": shape_digest(ref_ck["generator"])},
            "discriminator": {"n": len(ref_ck["discriminator"]), "digest": shape_digest(ref_ck["discriminator"])},
            "optimizers": {},
            "torch": torch.__version__,
        }
        for key in ("g_optimizer", "d_optimizer"):
            o = ref_ck[key]
            g0 = dict(o["param_groups"][0])
            n_params = len(g0.pop("params"))
            g0["betas"] = list(g0["betas"])
            e0 = o["state"][0]
            golden["optimizers"][key] = {
                "n_param_groups": len(o["param_groups"]), "n_params": n_params, "group": g0,
                "state_indices_are_0_to_n_minus_1": sorted(o["state"]) == list(range(n_params)),
                "entry_keys": list(e0),
                "entry_dtypes": {k: str(torch.as_tensor(x).dtype) for k, x in e0.items()},
                "step_shape": list(torch.as_tensor(e0["step"]).shape),
                "moment_shapes_digest": hashlib.sha1(";".join(
                    str(tuple(o["state"][i]["exp_avg"].shape)) for i in range(n_params)).encode()).hexdigest(),
            }

        # ---- 2. we write, the reference reads
        with tempfile.TemporaryDirectory() as tmp2:
            mine = ck.make_checkpoint(9, ref_ck["generator"], ref_ck["discriminator"],
                                      *[ck.adam_state_dict(*_flat(net, ref_ck[key], ck))
                                        for net, key in ((s.generator, "g_optimizer"), (s.discriminator, "d_optimizer"))],
                                      hist)
            written = ck.write_checkpoint(mine, tmp2)
            assert sorted(os.path.basename(p) for p in written) == files, (written, files)
            assert list(mine) == list(ref_ck)
            fresh = stub(2, tmp2, steps=0)
            out = Ref.load_from_checkpoint(fresh, os.path.join(tmp2, "checkpoint.pt"))
            assert out[0] == 9 and tuple(out[1:]) == tuple(hist[k] for k in ck.HISTORY_KEYS)
            for a, b in ((fresh.generator, s.generator), (fresh.discriminator, s.discriminator)):
                for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
                    assert torch.equal(x, y), k
            assert same_optim(fresh.g_optim.state_dict(), s.g_optim.state_dict())
            assert same_optim(fresh.d_optim.state_dict(), s.d_optim.state_dict())
            # ... and both optimisers continue identically from there
            for net_a, opt_a, net_b, opt_b in ((fresh.generator, fresh.g_optim, s.generator, s.g_optim),):
                torch.manual_seed(7)
                for pa, pb in zip(net_a.parameters(), net_b.parameters()):
                    pa.grad = torch.randn_like(pa) * 0.01
                    pb.grad = pa.grad.clone()
                opt_a.step()
                opt_b.step()
                for pa, pb in zip(net_a.parameters(), net_b.parameters()):
                    assert torch.equal(pa, pb)

    out_path = os.path.join(ROOT, "tests", "golden", "checkpoint_format.json")
    with open(out_path, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
        f.write("\n")
    print("reference <-> checkpoint.py cross-loading OK; wrote", out_path)
    return 0


def _flat(net, opt_sd, ck):
    """(layout, m, v, step, lr, betas, eps) of a torch Adam dictionary, through the flat buffers."""
    layout, total = ck.build_layout([(n, p.shape) for n, p in net.named_parameters()])
    m, v = torch.empty(total), torch.empty(total)
    h = ck.load_adam_state_dict(opt_sd, layout, m, v)
    return layout, m, v, h["step"], h["lr"], h["betas"], h["eps"]


if __name__ == "__main__":
    sys.exit(main())
