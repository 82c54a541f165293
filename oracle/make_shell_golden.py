"""TEST INFRASTRUCTURE.  Generates tests/golden/trainer_shell.json by running the UNMODIFIED reference trainer
`DeepFusionGAN.fit` (Deep_Fusion_GAN/src/deep_fusion_gan/model.py:135-296, imported read-only from /root/reference)
on the CPU for 2 epochs of 2 iterations at batch 2, 256x256 -- its own loop, losses, MA-GP, clipping, torch.optim.Adam,
text encoder, damsm loss, evaluation pass, checkpoint writer.  Build container only:

    python oracle/make_shell_golden.py

Inputs are the committed generators of oracle/ (weights: dfgan_oracle.init_*_state with O(1) gates, text encoder:
text_encoder_oracle.init_state; batches: `shell_batches` below), so the tests rebuild them bit for bit; the file
holds the reference's six per-epoch histories, digests of the final parameters and optimiser moments, and the list
of files it wrote.  Two substitutions, both outside the path under test: torchvision's pretrained Inception-v3
(not downloadable here) is replaced by oracle/metrics_oracle.py:StandInInception inside the reference's metrics
module, and scipy's removed `sqrtm(disp=False)` calling convention is restored (see make_metrics_golden.py).
It also checks oracle.dfgan_oracle.text_image_loss against the reference's damsm_loss on the spot.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/Deep_Fusion_GAN")
sys.path.insert(0, ROOT)

from oracle import dfgan_oracle as O  # noqa: E402
from oracle import metrics_oracle as M  # noqa: E402
from oracle import text_encoder_oracle as T  # noqa: E402
from oracle.make_metrics_golden import _LinalgWithDisp  # noqa: E402

NTOKEN, BATCH, N_BATCHES, EPOCHS, SEED = 50, 2, 2, 2, 4242


def shell_states():
    g_sd = O.init_generator_state(32, 100, gamma_range=(0.2, 1.0))
    d_sd = O.init_discriminator_state(32, gamma_range=(0.2, 1.0), weight_scale=2.2)
    return g_sd, d_sd, T.init_state(NTOKEN)


def shell_batches():
    """The loader: a list of (images [B,3,256,256], captions [B,18,1], caption lengths [B], file names) with
    UNSORTED lengths, so that prepare_data's sort (objects/utils.py:10-16) matters."""
    out = []
    for k in range(N_BATCHES):
        images, _, _ = O.synthetic_batch(BATCH, 256, seed=900 + k)
        caps, lens = T.synthetic_captions(BATCH, NTOKEN, seed=50 + k)
        perm = torch.arange(BATCH - 1, -1, -1)               # increasing lengths
        out.append((images, caps[perm].unsqueeze(-1), lens[perm], tuple(f"img_{k}_{i}" for i in range(BATCH))))
    return out


def digest(sd) -> dict:
    return {k: float(v.double().norm()) for k, v in sd.items()}


def main() -> int:
    warnings.filterwarnings("ignore")
    import src.damsm.loss as ref_damsm
    import src.deep_fusion_gan.model as ref_model
    import src.evaluation.metrics as ref_metrics
    ref_metrics.inception_v3 = lambda pretrained=True, transform_input=False: M.StandInInception(M.stand_in_weights())
    ref_metrics.linalg = _LinalgWithDisp()
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, t_sd = shell_states()

    # the oracle's text-image loss against the reference's damsm_loss (stub encoder returning given embeddings)
    img, sent, _ = O.synthetic_batch(4, 16, seed=3)
    ref_val = ref_damsm.damsm_loss(img, None, None, lambda c, l: sent)
    err = abs(float(ref_val) - float(O.text_image_loss(img, sent)))
    print(f"text_image_loss: reference {float(ref_val):.8f}, |oracle - reference| = {err:.2e}")
    assert err < 1e-6

    with tempfile.TemporaryDirectory() as tmp:
        enc = os.path.join(tmp, "text_encoder.pth")
        torch.save(t_sd, enc)
        img_dir, gen_dir = os.path.join(tmp, "gen_images"), os.path.join(tmp, "gen_weights")
        os.makedirs(img_dir), os.makedirs(gen_dir)
        model = ref_model.DeepFusionGAN(NTOKEN, enc, img_dir, gen_dir)
        assert model.device.type == "cpu"
        model.generator.load_state_dict(g_sd, strict=True)
        model.discriminator.load_state_dict(d_sd, strict=True)
        torch.manual_seed(SEED)
        hist = model.fit(shell_batches(), test_loader=None, num_epochs=EPOCHS, auto_resume=False)
        files = sorted(os.listdir(gen_dir)) + sorted(os.listdir(img_dir))
        ck = torch.load(os.path.join(gen_dir, "checkpoint.pt"), weights_only=False)
    names = ("g_losses", "d_losses", "d_gp_losses", "is_scores", "fid_scores", "txtimg_losses")
    gold = {n: [float(x) for x in h] for n, h in zip(names, hist)}
    for n in names:
        print(n, gold[n])
    gold["files"] = files
    gold["epoch"] = ck["epoch"]
    gold["generator_norms"] = digest(ck["generator"])
    gold["discriminator_norms"] = digest(ck["discriminator"])
    gold["g_adam_step"] = float(ck["g_optimizer"]["state"][0]["step"])
    gold["d_adam_step"] = float(ck["d_optimizer"]["state"][0]["step"])
    gold["d_exp_avg_sq_norm_0"] = float(ck["d_optimizer"]["state"][0]["exp_avg_sq"].double().norm())
    gold["initial_norms"] = {"g": digest(g_sd)["linear_in.weight"], "d": digest(d_sd)["img_forward.0.weight"],
                             "t": digest(t_sd)["encoder.weight"]}
    gold["config"] = {"ntoken": NTOKEN, "batch": BATCH, "n_batches": N_BATCHES, "epochs": EPOCHS, "seed": SEED,
                      "torch": torch.__version__}
    path = os.path.join(ROOT, "tests", "golden", "trainer_shell.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes;", files)
    return 0


if __name__ == "__main__":
    sys.exit(main())
