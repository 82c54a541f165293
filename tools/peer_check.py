"""Peer-memory all-reduce kernel (csrc/peer_reduce.cu, gan_codes_b200/peer.py) on N GPUs of one box, under torchrun:

  * correctness: random fp32 buffers per rank, in-place peer all-reduce of the whole buffer and of an unaligned-length
    tail range against the fp64 sum of all ranks' inputs (gathered with NCCL), bit-identical on every rank, and -- two
    ranks, where a + b has one rounding -- bit-identical to ncclAllReduce;
  * replays: the same launch captured in a CUDA graph and replayed 50 times (per-CTA call counters advance on device);
  * bandwidth: device time (CUDA events, max over ranks) of the D-gradient-sized (19.6 M), G-gradient-sized (12.3 M)
    and small-bucket (1.5 M elements) exchanges for a sweep of grid sizes, next to ncclAllReduce on the same buffer.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gan_codes_b200.peer import PeerGroup  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pg = PeerGroup(dev)
ok = True


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def all_same(t: torch.Tensor) -> bool:
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return all(torch.equal(parts[0], p) for p in parts)


# ------------------------------------------------------------------------------------ correctness
N = 19_600_000 + 4 * 37
buf = pg.empty(N)
gen = torch.Generator(device=dev).manual_seed(1000 + rank)
src = torch.randn(N, device=dev, generator=gen) * (1.0 + rank)
parts = [torch.empty_like(src) for _ in range(world)]
dist.all_gather(parts, src)
ref64 = torch.zeros(N, device=dev, dtype=torch.float64)
ref32 = torch.zeros(N, device=dev)
for p in parts:
    ref64 += p.double()
    ref32 += p                      # rank order 0..world-1, fp32: the kernel's order
del parts
for lo in (0, 4 * 1_234_567 + 4):
    buf.copy_(src)
    pg.all_reduce(buf[lo:], channel=0)
    torch.cuda.synchronize()
    err = float((buf[lo:].double() - ref64[lo:]).abs().max())
    # the unicast kernel adds in rank order (exactly ref32); the in-switch reduction's order is the hardware's
    exact = torch.equal(buf[lo:], ref32[lo:]) or (pg.multimem and err < 1e-4)
    untouched = torch.equal(buf[:lo], src[:lo])
    same = all_same(buf[lo:])
    say(f"range [{lo}, {N}): max |peer - fp64 sum| = {err:.3e}, == rank-order fp32 sum: {exact}, "
        f"prefix untouched: {untouched}, identical on all ranks: {same}")
    ok &= exact and untouched and same and err < 1e-4
say(f"exchange kernel: {'multimem (NVLS, in-switch reduction)' if pg.multimem else 'unicast pull / push'}")
if world == 2:
    nc = src.clone()
    dist.all_reduce(nc)
    buf.copy_(src)
    pg.all_reduce(buf)
    torch.cuda.synchronize()
    eq = torch.equal(nc, buf)
    say(f"ncclAllReduce == peer all-reduce bit for bit (2 ranks): {eq}")
    ok &= eq

# ------------------------------------------------------------------------------------ graph replays
buf.copy_(src)
torch.cuda.synchronize()
dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    pg.all_reduce(buf, channel=0)
    pg.all_reduce(buf[4 * 1000:4 * 5000], channel=1)
for _ in range(50):
    g.replay()
torch.cuda.synchronize()
same = all_same(buf)
say(f"50 graph replays of two captured exchanges: identical on all ranks: {same}")
ok &= same

# ------------------------------------------------------------------------------------ bandwidth
def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


buf.zero_()
for n in (19_600_000, 12_300_000, 1_500_000, 100_000):
    n = (n // 4) * 4
    v = buf[:n]
    t_nccl = timed(lambda: dist.all_reduce(v))
    row = [f"{n * 4 / 1e6:7.1f} MB  nccl {t_nccl * 1e3:7.1f} us"]
    for ctas in (8, 16, 32, 64, 128):
        t = timed(lambda: pg.all_reduce(v, channel=0, ctas=ctas))
        # bus bandwidth like nccl-tests: 2 (N-1)/N x bytes / time
        bw = 2 * (world - 1) / world * n * 4 / (t * 1e-3) / 1e9
        row.append(f"ctas {ctas:3d}: {t * 1e3:7.1f} us ({bw:5.0f} GB/s bus)")
    t = timed(lambda: pg.all_reduce(v, channel=0, ctas=128, threads=128))
    row.append(f"lite 128x128: {t * 1e3:7.1f} us")
    say(" | ".join(row))

if pg.multimem:
    pg.multimem = False
    say("unicast pull / push kernel on the same buffers:")
    buf.copy_(src)
    pg.all_reduce(buf)
    torch.cuda.synchronize()
    exact = torch.equal(buf, ref32)
    say(f"  == rank-order fp32 sum: {exact}, identical on all ranks: {all_same(buf)}")
    ok &= exact
    for n in (19_600_000, 12_300_000, 1_500_000):
        v = buf[:(n // 4) * 4]
        row = [f"{v.numel() * 4 / 1e6:7.1f} MB"]
        for ctas in (32, 64, 128):
            t = timed(lambda: pg.all_reduce(v, channel=0, ctas=ctas))
            row.append(f"ctas {ctas:3d}: {t * 1e3:7.1f} us")
        say(" | ".join(row))
pg.close()
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}:", "PEER OK" if ok else "PEER MISMATCH", flush=True)
sys.exit(0 if ok else 1)
