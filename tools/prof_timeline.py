"""Kernel timeline of CUDA-graph replays of the training iteration (torch.profiler / CUPTI activity records):
start, duration, stream of every kernel inside the replay -- which kernels overlap, where the device idles.

    python tools/prof_timeline.py [batch] > gpurun_out/timeline.csv     (then tools/timeline_report.py)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from gan_codes_b200.engine import DFGANTrainer  # noqa: E402
from gan_codes_b200.modules import Discriminator, Generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# under torchrun: data-parallel iteration, rank 0 prints its timeline
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
torch.manual_seed(0)
G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
tr = DFGANTrainer(G, D, B, world=world)
img = (torch.randn(B, 3, 256, 256, device="cuda") * 0.5).clamp(-1, 1)
sent, noise = torch.randn(B, 256, device="cuda") * 0.3, torch.randn(B, 100, device="cuda")
for _ in range(6):
    tr.step(img, sent, noise)
torch.cuda.synchronize()
assert tr._graph is not None
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(img, sent, noise)
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, ev.name))
rows.sort()
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    if rank != 0:
        sys.exit(0)
print("start_us,dur_us,name")
t0 = rows[0][0] if rows else 0
for s, d, n in rows:
    print(f"{s - t0:.3f},{d:.3f},\"{n[:120]}\"")
