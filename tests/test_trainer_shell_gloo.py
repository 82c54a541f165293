"""Data-parallel trainer shell on CPU with 2 gloo ranks: each rank runs DeepFusionGAN.fit on its own
batch with the oracle-backed iteration averaging gradients across ranks (the exchange DFGANTrainer does with NCCL).
Checks: identical weights and identical (rank-averaged) loss histories on both ranks, files written by rank 0 only."""
import os
import socket
import warnings

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _worker(rank, port, out_dir):
    warnings.filterwarnings("ignore")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // WORLD))
    from gan_codes_b200 import evaluation as E
    from oracle import dfgan_oracle as O
    from oracle import metrics_oracle as M
    from oracle.make_shell_golden import NTOKEN, shell_batches, shell_states
    from tests.test_trainer_shell_cpu import OracleTrainer, ShellOnOracle

    def hook(phase, grads):
        flat = torch.cat([g.reshape(-1) for g in grads.values()])
        dist.all_reduce(flat)
        flat /= WORLD
        out, o = {}, 0
        for k, g in grads.items():
            out[k] = flat[o:o + g.numel()].view_as(g)
            o += g.numel()
        return out

    class DPOracleTrainer(OracleTrainer):
        def step(self, images, sent, noise):
            probe = {}
            losses = O.train_step(self.g_sd, self.d_sd, self.g_optim.adam, self.d_optim.adam, images, sent, noise,
                                  grad_hook=hook, probe=probe)
            return {"d_loss": losses["d"], "gp_loss": losses["gp"], "g_loss": losses["g"],
                    "txtimg_loss": O.text_image_loss(probe["fake"], sent)}

    class Shell(ShellOnOracle):
        def _make_trainer(self, batch_size):
            return DPOracleTrainer(self.generator, self.discriminator, batch_size, True)

        def _save_checkpoint(self, *a, **k):
            open(os.path.join(out_dir, f"saved_by_rank{rank}"), "w").close()
            return super()._save_checkpoint(*a, **k)

    g_sd, d_sd, t_sd = shell_states()
    enc = os.path.join(out_dir, f"enc{rank}.pth")
    torch.save(t_sd, enc)
    img_dir, gen_dir = os.path.join(out_dir, "img"), os.path.join(out_dir, "gen")
    os.makedirs(img_dir, exist_ok=True), os.makedirs(gen_dir, exist_ok=True)
    E.set_inception_factory(lambda: M.StandInInception(M.stand_in_weights()))
    model = Shell(NTOKEN, enc, img_dir, gen_dir, world=WORLD)
    assert model.rank == rank
    model.generator.load_state_dict(g_sd)
    model.discriminator.load_state_dict(d_sd)
    torch.manual_seed(100 + rank)                      # every rank draws its own noise
    hist = model.fit([shell_batches()[rank]], num_epochs=1, auto_resume=False)     # its own shard of the data
    torch.save({"g": model.generator.state_dict(), "d": model.discriminator.state_dict(), "hist": hist},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shell(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"), weights_only=False)
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"), weights_only=False)
    for net in ("g", "d"):
        for k in r0[net]:
            assert torch.equal(r0[net][k], r1[net][k]), (net, k)
    for i in (0, 1, 2, 5):                              # g, d, gp, text-image: averaged over the ranks -> identical
        assert r0["hist"][i] == r1["hist"][i] and len(r0["hist"][i]) == 1
    # rank 0 alone wrote the epoch's files
    assert os.path.exists(os.path.join(tmp_path, "saved_by_rank0")) and not os.path.exists(os.path.join(tmp_path, "saved_by_rank1"))
    assert sorted(os.listdir(os.path.join(tmp_path, "gen"))) == ["checkpoint.pt", "gen_0.pth"]
    assert sorted(os.listdir(os.path.join(tmp_path, "img"))) == ["fake_sample_epoch_0.png"]
    ck = torch.load(os.path.join(tmp_path, "gen", "checkpoint.pt"), weights_only=False)
    assert ck["is_scores"] == r0["hist"][3] and ck["g_losses"] == r0["hist"][0]
    for k, v in ck["generator"].items():
        assert torch.equal(v, r0["g"][k])
