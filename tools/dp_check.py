"""Multi-GPU data-parallel check (run under torchrun on N GPUs): after a few DFGANTrainer iterations on different
shards every rank must hold bit-identical parameters (one all-reduce of the flat gradient per optimiser step, the
same clip coefficient and Adam update everywhere), and the update must differ from a world=1 run on the same shard
(i.e. the exchange really happened).  With the peer-memory exchange a third run repeats the FIRST D update with the
early (overlapped) exchange of the D step's low-resolution bucket switched off: with betas = (0, 0.9) the first Adam
update is lr * sign(gradient), so a gradient contribution that the early exchange missed (a kernel still writing into
the bucket) would flip the update of whole tensors, while the run-to-run noise of the fp32 atomics flips only elements
whose gradient is ~0 -- the share of D parameters whose first update differs must stay below 1e-3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gan_codes_b200.engine import DFGANTrainer  # noqa: E402
from gan_codes_b200.modules import Discriminator, Generator  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 4


def build(w):
    torch.manual_seed(1234)
    G, D = Generator(32, 100).to(dev), Discriminator(32).to(dev)
    with torch.no_grad():
        for n, p in list(G.named_parameters()) + list(D.named_parameters()):
            if n.endswith(".gamma"):
                p.fill_(0.5)
    return DFGANTrainer(G, D, B, world=w)


gen = torch.Generator().manual_seed(99 + rank)
img = (torch.randn((B, 3, 256, 256), generator=gen) * 0.5).clamp_(-1, 1).to(dev)
sent = (torch.randn((B, 256), generator=gen) * 0.3).to(dev)
noise = torch.randn((B, 100), generator=gen).to(dev)
tr = build(world)
for _ in range(4):          # 2 eager + capture + replay
    out = tr.step(img, sent, noise)
torch.cuda.synchronize()
ok = True
for name, st in (("G", tr.gs), ("D", tr.ds)):
    cs = torch.stack([st.flat.double().sum(), st.flat.double().abs().sum(), st.flat.double().pow(2).sum()])
    allcs = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    same = all(torch.equal(allcs[0], c) for c in allcs)
    ok &= same and bool(torch.isfinite(cs).all())
    if rank == 0:
        print(f"{name}: identical parameters on all {world} ranks: {same}; checksum {allcs[0].tolist()}")
solo = build(1)
for _ in range(4):
    solo.step(img, sent, noise)
torch.cuda.synchronize()
diff = float((solo.ds.flat - tr.ds.flat).abs().max())
if rank == 0:
    print(f"max |D params (world={world}) - D params (world=1, same shard)| = {diff:.3e} (must be > 0)")
    print("losses", {k: float(v) for k, v in out.items()})
ok &= diff > 0
from gan_codes_b200 import engine as E  # noqa: E402
if E.DP_SHARDED and tr.peer is not None:
    # sharded optimiser step: after the collective sync every rank holds the same, complete Adam moments, and they are
    # those of an unsharded run up to the rounding of the differently associated norm (first update: lr * sign(g))
    tr.sync_optimizer_state()
    for name, st in (("G", tr.gs), ("D", tr.ds)):
        cs = torch.stack([st.m.double().abs().sum(), st.v.double().sum()])
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        same = all(torch.equal(allcs[0], c) for c in allcs) and bool((cs > 0).all())
        ok &= same
        if rank == 0:
            print(f"{name}: sharded Adam moments identical and non-zero on all ranks after sync: {same}")
    # the FIRST D update, sharded against unsharded, from identical weights: the global squared norm (per-rank partial
    # sums added in rank order against one pass over the whole reduced gradient) and the update itself
    res = []
    for sh in (True, False):
        E.DP_SHARDED = sh
        t1 = build(world)
        start = t1.ds.flat.clone()
        with torch.no_grad():
            t1._seg_out = {}
            d_step, d_exchange = t1._segments(img, sent, noise)[0]
            d_step()
            d_exchange()
            t1.ds.apply(world)
        torch.cuda.synchronize()
        res.append((float(t1.ds.sumsq), (t1.ds.flat - start).clone()))
        del t1
    E.DP_SHARDED = True
    (s_sh, d_sh), (s_un, d_un) = res
    rel = abs(s_sh - s_un) / max(abs(s_un), 1e-30)
    moved = (d_sh != 0) | (d_un != 0)
    differ = float(((d_sh - d_un).abs() > 1e-5 * 4e-4).logical_and(moved).sum()) / max(float(moved.sum()), 1.0)
    if rank == 0:
        print(f"first D update, sharded vs unsharded step: |sum g^2| {s_sh:.6e} vs {s_un:.6e} (rel {rel:.1e}, must be "
              f"< 1e-4); share of parameters whose update differs {differ:.2e} (must be < 1e-3)")
    ok &= rel < 1e-4 and differ < 1e-3
if tr.peer is not None and E.DP_OVERLAP:
    firsts = []
    for overlap in (True, False):
        E.DP_OVERLAP = overlap
        t1 = build(world)
        start = t1.ds.flat.clone()
        with torch.no_grad():              # the D hinge step alone: backward, exchange(s), clip + Adam
            t1._seg_out = {}
            d_step, d_exchange = t1._segments(img, sent, noise)[0]
            d_step()
            d_exchange()
            t1.ds.apply(world)
        torch.cuda.synchronize()
        firsts.append((t1.ds.flat - start).clone())
        del t1
    E.DP_OVERLAP = True
    a, b = firsts
    moved = (a != 0) | (b != 0)
    differ = float(((a - b).abs() > 1e-5 * 4e-4).logical_and(moved).sum()) / max(float(moved.sum()), 1.0)
    if rank == 0:
        print(f"D parameters after the first D update, early exchange on vs off: share of moved parameters whose update "
              f"differs = {differ:.2e} (must be < 1e-3); moved {int(moved.sum())} of {moved.numel()}")
    ok &= differ < 1e-3
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}:", "DP OK" if ok else "DP MISMATCH")
sys.exit(0 if ok else 1)
