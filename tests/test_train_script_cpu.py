"""`gan_codes_b200.train.train` (train.py:8-66) end to end on CPU through the package's seams: the real DFGANDataset /
DataLoader on tests/golden/mini_cub with the CPU oracle of the image transform, the real DeepFusionGAN shell with the
CPU oracle of the iteration and the text encoder.  Checks the orchestration: seeding, loaders, vocabulary hand-over,
checkpoint discovery and resume, the files of one epoch, the returned histories."""
import os

import pytest
import torch

from gan_codes_b200 import dataset as D
from gan_codes_b200 import deep_fusion_gan as S
from gan_codes_b200 import evaluation as E
from gan_codes_b200 import train as T
from oracle import metrics_oracle as M
from oracle import text_encoder_oracle as TE
from tests.test_dataset_cpu import MINI, OracleTransform
from tests.test_trainer_shell_cpu import ShellOnOracle


def _loader(imsize, batch_size, data_dir, split):
    ds = D.DFGANDataset(data_dir, split, OracleTransform(imsize))
    return D.DataLoader(ds, batch_size=batch_size, drop_last=True, shuffle=True, collate_fn=ds.collate)


def test_train_runs_an_epoch_and_resumes(tmp_path, monkeypatch, capsys):
    monkeypatch.setattr(D, "create_loader", _loader)
    monkeypatch.setattr(S, "DeepFusionGAN", ShellOnOracle)
    E.set_inception_factory(lambda: M.StandInInception(M.stand_in_weights()))
    enc = tmp_path / "text_encoder.pth"
    torch.save(TE.init_state(40), enc)                      # mini_cub's vocabulary has 40 words
    img_dir, gen_dir = str(tmp_path / "gen_images"), str(tmp_path / "gen_weights")
    try:
        out = T.train(MINI, str(enc), img_dir, gen_dir, num_epochs=1, batch_size=2)
        text = capsys.readouterr().out
        assert "Seed 123321 fixed" in text and "Test set size: 3 images" in text
        assert "Warning: Checkpoint" in text and "Starting from scratch" in text and "Epoch 1: G Loss:" in text
        assert [len(h) for h in out] == [1] * 6 and all(x == x for h in out for x in h)
        assert sorted(os.listdir(gen_dir)) == ["checkpoint.pt", "gen_0.pth"]
        assert sorted(os.listdir(img_dir)) == ["fake_sample_epoch_0.png"]
        # second call finds checkpoint.pt, restores and has nothing left to do for num_epochs=1
        out2 = T.train(MINI, str(enc), img_dir, gen_dir, num_epochs=1, batch_size=2)
        text = capsys.readouterr().out
        assert "Loading checkpoint from latest" in text and "Loaded checkpoint from epoch 1." in text
        assert [list(h) for h in out2] == [list(h) for h in out]
    finally:
        E.set_inception_factory(None)


def test_train_needs_cuda(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    with pytest.raises(RuntimeError, match="CUDA only"):
        T.train(MINI, str(tmp_path / "none.pth"), str(tmp_path / "a"), str(tmp_path / "b"), num_epochs=1, batch_size=2)
