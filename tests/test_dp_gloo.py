"""Data-parallel semantics on the CPU with 2 gloo ranks: every rank runs the oracle iteration on
its own shard, gradients are summed with all_reduce and scaled by 1/N before the global-norm clip and Adam -- the
same exchange DFGANTrainer performs on the flat gradient buffer with NCCL.  Checks (1) both ranks end with
identical weights, (2) they equal a serial simulation of "N reference replicas on the N shards with averaged
gradients" (mismatch pairs formed inside each local shard)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dfgan_oracle as O

DEPTH, SIZE, B_LOCAL, WORLD = 4, 64, 2, 2


def _setup_states():
    g_sd = O.init_generator_state(32, 100, depth=DEPTH, gamma_range=(0.2, 1.0), dtype=torch.float64)
    d_sd = O.init_discriminator_state(32, depth=DEPTH, gamma_range=(0.2, 1.0), dtype=torch.float64)
    return g_sd, d_sd


def _shard(rank):
    return O.synthetic_batch(B_LOCAL, SIZE, seed=O.SEED + 100 + rank, dtype=torch.float64)


def _worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    g_sd, d_sd = _setup_states()
    go, do = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    images, sent, noise = _shard(rank)

    def hook(phase, grads):
        flat = torch.cat([g.reshape(-1) for g in grads.values()])   # one exchange per optimiser step
        dist.all_reduce(flat)
        flat /= WORLD
        out, o = {}, 0
        for k, g in grads.items():
            out[k] = flat[o:o + g.numel()].view_as(g)
            o += g.numel()
        return out
    for _ in range(2):
        O.train_step(g_sd, d_sd, go, do, images, sent, noise, grad_hook=hook)
    torch.save({"g": g_sd, "d": d_sd}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def _serial_reference():
    """N replicas stepped in lock-step with explicitly averaged gradients (no communication library)."""
    g_sd, d_sd = _setup_states()
    go, do = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    shards = [_shard(r) for r in range(WORLD)]
    for _ in range(2):
        # D step
        gls = [O._leaf(g_sd) for _ in range(WORLD)]
        fakes = [O.generator_forward(gl, sh[2], sh[1]) for gl, sh in zip(gls, shards)]
        per = []
        for sh, fk in zip(shards, fakes):
            dl = O._leaf(d_sd)
            per.append(O._grads(O.d_hinge_loss(dl, sh[0], fk.detach(), sh[1]), dl, False))
        avg = {k: sum(p[k] for p in per) / WORLD for k in per[0]}
        coef, _ = O.clip_coef(list(avg.values()))
        do.step(d_sd, {k: g * coef for k, g in avg.items()})
        per = []
        for sh in shards:
            dl = O._leaf(d_sd)
            per.append(O._grads(O.magp_loss(dl, sh[0], sh[1]), dl, True))
        avg = {k: sum(p[k] for p in per) / WORLD for k in per[0]}
        coef, _ = O.clip_coef(list(avg.values()))
        do.step(d_sd, {k: g * coef for k, g in avg.items()})
        per = [O._grads(O.g_hinge_loss(d_sd, fk, sh[1]), gl, False) for gl, fk, sh in zip(gls, fakes, shards)]
        avg = {k: sum(p[k] for p in per) / WORLD for k in per[0]}
        coef, _ = O.clip_coef(list(avg.values()))
        go.step(g_sd, {k: g * coef for k, g in avg.items()})
    return g_sd, d_sd


def test_two_rank_gloo_equals_averaged_replicas(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for net in ("g", "d"):
        for k in r0[net]:
            assert torch.equal(r0[net][k], r1[net][k]), (net, k)
    g_ref, d_ref = _serial_reference()
    for k, v in g_ref.items():
        assert float((r0["g"][k] - v).abs().max()) <= 1e-9 * (1 + float(v.abs().max())), k
    for k, v in d_ref.items():
        assert float((r0["d"][k] - v).abs().max()) <= 1e-9 * (1 + float(v.abs().max())), k


def test_owned_slices_partition_a_bucket():
    """Host side of the sharded optimiser step: the slices the ranks own (gan_codes_b200.peer.owned_slice, the
    partition csrc/peer_reduce.cu computes per kernel) tile every bucket exactly, in rank order, on 16-byte boundaries."""
    from gan_codes_b200.peer import owned_slice
    for lo, hi in ((0, 19_645_272), (1_529_728, 19_645_272), (0, 1_529_728), (8, 12), (0, 4), (16, 16 + 4 * 7)):
        for world in (2, 3, 4, 5, 8):
            at = lo
            for r in range(world):
                a, b = owned_slice(lo, hi, r, world)
                assert a == at and a <= b <= hi and (a - lo) % 4 == 0 and (b - lo) % 4 == 0
                at = b
            assert at == hi
            per = 4 * (((hi - lo) // 4 + world - 1) // world)
            assert all(owned_slice(lo, hi, r, world)[1] - owned_slice(lo, hi, r, world)[0] <= per for r in range(world))
