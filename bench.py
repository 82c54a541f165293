#!/usr/bin/env python
"""Headline benchmark: training images/s of one full DF-GAN iteration (D hinge step + MA-GP step + G step, three
clip+Adam updates -- deep_fusion_gan/model.py:173-228) at 256x256, n_channels 32, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.  `value` = whole-job images/s
with inputs resident in HBM; `e2e` = the same through DFGANTrainer.step with pinned HOST inputs and a device->host
read of the three losses every step.  `roofline` is measured live with CUDA events around every launch of the
dominant kernel; `cpu_baseline` times the UNMODIFIED reference trainer's own `fit` loop (oracle/_ref, vendored by
oracle/vendor_ref.py; the oracle port only when that copy is absent) on the host cores (rank 0, N=1);
`library_baseline` times the same reference on the B200 under stock PyTorch-CUDA (fp32, TF32, autocast bf16) -- the
library bar.  `--impl reference` times the CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = 181.3   # algorithmic minimum of the full step, SURVEY 8(d) / BASELINE.md section 3


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def summary(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_step(batch: int, size: int, steps: int, warmup: int):
    """The reference's training iteration on the host cores.  Returns (img/s, cores, seconds, best step, kind, what):
    kind "reference" = the UNMODIFIED reference trainer's own `fit` loop from oracle/_ref (oracle/ref_runner.py);
    kind "port" = oracle/dfgan_oracle.py train_step, only when oracle/_ref is absent or the size is not 256."""
    import torch
    from oracle import dfgan_oracle as O
    from oracle import ref_runner
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if size == 256 and ref_runner.available():
        r = ref_runner.run_fit("cpu", batch, steps, warmup, "fp32", threads=cores)
        return (r["images_per_s"], cores, r["seconds"], r["best_ms"] / 1e3, "reference",
                "unmodified reference DeepFusionGAN.fit loop (oracle/_ref: text encoder + D hinge + MA-GP + G, "
                "clip, torch.optim.Adam, fp32, torch CPU)")
    depth = {256: 6, 128: 5, 64: 4}[size]
    g_sd = O.init_generator_state(32, 100, depth=depth)
    d_sd = O.init_discriminator_state(32, depth=depth)
    images, sent, noise = O.synthetic_batch(batch, size)
    go, do = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    for _ in range(warmup):
        O.train_step(g_sd, d_sd, go, do, images, sent, noise)
    best, total = None, 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        O.train_step(g_sd, d_sd, go, do, images, sent, noise)
        dt = time.perf_counter() - t0
        total += dt
        best = dt if best is None else min(best, dt)
    return batch * steps / total, cores, total, best, "port", "oracle/dfgan_oracle.py train_step (fp32, torch CPU)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = 4
    t0 = time.perf_counter()
    ips, cores, total, best, kind, what = cpu_reference_step(b, args.size, max(1, args.steps),
                                                             max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "train images/sec (G+D step, 256x256)", "value": round(ips, 4),
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * b / ips, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DF-GAN full training iteration {args.size}x{args.size}, n_channels 32, "
                               f"CPU sample batch {b}", "batch_per_gpu": b},
        "cpu_baseline": {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} iterations of batch {b} at {args.size}x{args.size} "
                                   f"({total:.1f} s), {what}"},
        "e2e": {"value": round(ips, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    from gan_codes_b200 import _lib
    from gan_codes_b200.engine import DFGANTrainer
    from gan_codes_b200.modules import Discriminator, Generator

    B, S = args.batch, args.size
    torch.manual_seed(123321)                        # utils.py:27 ; same init on every rank (data parallel)
    G, D = Generator(32, 100).to(dev), Discriminator(32).to(dev)
    trainer = DFGANTrainer(G, D, B, world=world)
    gen = torch.Generator().manual_seed(123321 + 17 * rank)   # every rank draws its own shard
    images_h = (torch.randn((B, 3, S, S), generator=gen) * 0.5).clamp_(-1, 1).pin_memory()
    sent_h = (torch.randn((B, 256), generator=gen) * 0.3).pin_memory()
    noise_h = torch.randn((B, 100), generator=gen).pin_memory()
    images_d, sent_d, noise_d = images_h.to(dev), sent_h.to(dev), noise_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_dev():
        trainer.step(images_d, sent_d, noise_d)

    # end-to-end step: pinned host inputs in, the three losses out to pinned host memory EVERY step.  The
    # device->host copy is asynchronous and the host waits for it one call later (double-buffered), so the host
    # keeps one iteration queued ahead and the next H2D overlaps the running iteration; the last step's losses
    # are awaited inside the timed region by `drain_e2e`.
    loss_h = [torch.zeros(3).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "read": 0}

    def step_e2e():
        i = e2e_state["i"]
        out = trainer.step(images_h, sent_h, noise_h)
        if i >= 1:
            loss_ev[(i - 1) & 1].synchronize()           # losses of the previous step are now on the host
            e2e_state["read"] += 1
        loss_h[i & 1].copy_(torch.stack([out["d_loss"], out["gp_loss"], out["g_loss"]]), non_blocking=True)
        loss_ev[i & 1].record()
        e2e_state["i"] = i + 1

    def drain_e2e():
        i = e2e_state["i"]
        if i >= 1:
            loss_ev[(i - 1) & 1].synchronize()
            e2e_state["read"] += 1

    # eager pass first (also the graph warm-up): counts OUR kernel launches per step and times every launch of
    # the two GEMM engines with CUDA events on the launching stream (roofline leg)
    for _ in range(2):
        step_dev()
    trainer.use_graph, graph_wanted = False, trainer.use_graph
    from gan_codes_b200 import convops as _co
    side_wanted, _co.SIDE_ENABLED = _co.SIDE_ENABLED, False   # serial launches: per-kernel durations stay clean
    _lib.profile_enable(True)
    l0 = _lib.launch_count()
    ms_eager = timed(step_dev, args.steps)
    launches = _lib.launch_count() - l0
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    _co.SIDE_ENABLED = side_wanted
    trainer.use_graph = graph_wanted
    for _ in range(max(3, args.warmup)):      # first call captures the CUDA graph, the rest replay it
        step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_dev, args.steps)
    clocks = sampler.summary() if sampler else None
    step_e2e()
    drain_e2e()
    e2e_state.update(i=0, read=0)

    def e2e_region():
        step_e2e()
        if e2e_state["i"] == args.steps:
            drain_e2e()
    ms_e2e = timed(e2e_region, args.steps)
    assert e2e_state["read"] == args.steps, e2e_state
    losses = {k: float(v) for k, v in trainer.last.items()}

    if rank == 0:
        hbm, tf_burst, tf_sust, how = _peaks()
        value = B * world * args.steps / (ms / 1e3)
        e2e = B * world * args.steps / (ms_e2e / 1e3)
        dom = max(("conv", "wgrad"), key=lambda k: prof[k]["ms"])
        gflop_exec = (prof["conv"]["flops"] + prof["wgrad"]["flops"]) / args.steps / B / 1e9
        p = prof[dom]
        ach = p["flops"] / (p["ms"] / 1e3) / 1e12 if p["ms"] > 0 else 0.0
        traffic = None
        pdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
        tname = next((n for n in ("r02_traffic.json", "r01_traffic.json") if os.path.exists(os.path.join(pdir, n))),
                     "r02_traffic.json")
        tpath = os.path.join(pdir, tname)
        if dom == "conv" and B == 32 and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = round(json.load(fh)["dram_bytes_per_launch"], 1)
        line = {
            "metric": "train images/sec (G+D step, 256x256)", "value": round(value, 2), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"DF-GAN full training iteration (D hinge + MA-GP + G, 3x clip+Adam) {S}x{S}, "
                                   f"n_channels 32, nz 100, cond 256, batch {B}/GPU (BASELINE configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations > 2 GB) far exceeds the 126 MB L2; no flush needed",
                       "pct_of_bf16_peak_algorithmic": round(100 * value / world * GFLOP_PER_IMAGE / 1e3 / tf_sust, 2),
                       "gflop_per_image_algorithmic": GFLOP_PER_IMAGE,
                       # executed = what the two GEMM engines actually multiplied (sub-pixel folding and the pooled
                       # shortcut save work; never credited), counted live by the launch counters of the eager pass
                       "gflop_per_image_executed": round(gflop_exec, 2),
                       "pct_of_bf16_peak": {
                           "executed_flops_vs_sustained": round(100 * value / world * gflop_exec / 1e3 / tf_sust, 2),
                           "executed_flops_vs_burst": round(100 * value / world * gflop_exec / 1e3 / tf_burst, 2),
                           "algorithmic_flops_vs_sustained": round(100 * value / world * GFLOP_PER_IMAGE / 1e3 / tf_sust, 2),
                           "algorithmic_flops_vs_burst": round(100 * value / world * GFLOP_PER_IMAGE / 1e3 / tf_burst, 2),
                           "peaks_tflops": {"sustained": tf_sust, "burst": tf_burst, "source": how}},
                       "losses": losses,
                       "cuda_graph": bool(trainer._graph is not None)},
            "e2e": {"value": round(e2e, 2), "unit": "images/s",
                    "h2d_bytes_per_step": int(images_h.numel() * 4 + sent_h.numel() * 4 + noise_h.numel() * 4),
                    "d2h_bytes_per_step": 12},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "conv_umma_kernel (dfgan_conv_run)" if dom == "conv" else
                         "wgrad_umma_kernel (dfgan_wgrad_run)",
                         "bound": "tensor", "achieved": round(ach, 2), "peak": tf_sust, "unit": "TFLOP/s",
                         "frac": round(ach / tf_sust, 4), "traffic": traffic,
                         "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu, "
                                         "cold cache), averaged over the conv-engine launches of one iteration; "
                                         f"see profiles/{tname}",
                         "algorithmic_bytes_per_launch": round(p["bytes"] / max(p["launches"], 1), 1),
                         "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step)",
                         "launches_timed": int(p["launches"]), "kernel_ms_per_step": round(p["ms"] / args.steps, 3),
                         "share_of_step": round(p["ms"] / ms, 4),
                         "timed_in": f"eager pass of the same {args.steps} steps ({round(ms_eager / args.steps, 3)} ms/step; "
                                     "the headline pass replays the identical launches as one CUDA graph)",
                         "executed_gflop_per_launch_avg": round(p["flops"] / max(p["launches"], 1) / 1e9, 3),
                         "other_engine": {k: round(v, 3) for k, v in prof["wgrad" if dom == "conv" else "conv"].items()}},
        }
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = 10     # bounded sample: ~10 s of CPU work on the box's host cores
            ips, cores, total, best, kind, what = cpu_reference_step(4, S, n_cpu, 1)
            line["cpu_baseline"] = {"value": round(ips, 4), "unit": "images/s", "cores": cores, "kind": kind,
                                    "sample": f"{n_cpu} iterations of batch 4 at {S}x{S} ({total:.1f} s) of {what}"}
        if world == 1 and not args.no_library_baseline:
            del trainer
            torch.cuda.empty_cache()
            line["library_baseline"] = library_baseline(B, S)
            best = line["library_baseline"].get("value")
            if best:
                line["vs_library"] = {"value_ratio": round(value / best, 2), "e2e_ratio": round(e2e / best, 2),
                                      "note": "this repo's images/s over the best stock PyTorch-CUDA mode of the "
                                              "unmodified reference on the same B200, same batch"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def library_baseline(batch: int, size: int, steps: int = 5, warmup: int = 3) -> dict:
    """SURVEY 2c / 8(d), BASELINE.md 5.2: the reference as it is, on this B200, under stock PyTorch-CUDA (cuDNN /
    cuBLAS) -- fp32 with TF32 off, TF32 on, and autocast(bf16) + channels_last.  The best of the three is the bar the
    hand-written sm_100a kernels have to beat.  oracle/_ref (the vendored unmodified reference) when present, else
    the oracle port's train_step on the device."""
    import torch
    from oracle import ref_runner
    modes, out = {}, {"unit": "images/s", "batch": batch, "steps": steps, "warmup": warmup}
    if size == 256 and ref_runner.available():
        out["kind"] = "reference"
        out["what"] = "unmodified reference DeepFusionGAN.fit loop (oracle/_ref) on cuda:0, stock PyTorch kernels"
        for mode in ref_runner.MODES:
            try:
                r = ref_runner.run_fit("cuda", batch, steps, warmup, mode)
                modes[mode] = {"images_per_s": round(r["images_per_s"], 1), "ms_per_step": round(r["ms_per_step"], 2)}
            except Exception as e:   # a mode stock PyTorch cannot run is reported, not hidden
                modes[mode] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            torch.cuda.empty_cache()
    else:
        from oracle import dfgan_oracle as O
        out["kind"] = "port"
        out["what"] = "oracle/dfgan_oracle.py train_step on cuda:0 (torch ops -> cuDNN / cuBLAS)"
        depth = {256: 6, 128: 5, 64: 4}[size]
        for mode in ref_runner.MODES:
            try:
                tf32 = mode == "tf32"
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = mode.endswith("cudnn_benchmark")
                g_sd = O.to(O.init_generator_state(32, 100, depth=depth), device="cuda")
                d_sd = O.to(O.init_discriminator_state(32, depth=depth), device="cuda")
                images, sent, noise = (t.cuda() for t in O.synthetic_batch(batch, size))
                go, do = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
                import contextlib
                ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode.startswith("bf16") else contextlib.nullcontext()
                with ctx:
                    for _ in range(warmup):
                        O.train_step(g_sd, d_sd, go, do, images, sent, noise)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(steps):
                        O.train_step(g_sd, d_sd, go, do, images, sent, noise)
                    torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                modes[mode] = {"images_per_s": round(batch * steps / dt, 1), "ms_per_step": round(1e3 * dt / steps, 2)}
            except Exception as e:
                modes[mode] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
            finally:
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
                torch.backends.cudnn.benchmark = False
            torch.cuda.empty_cache()
    out["modes"] = modes
    ok = {k: v["images_per_s"] for k, v in modes.items() if "images_per_s" in v}
    if ok:
        out["best_mode"] = max(ok, key=ok.get)
        out["value"] = ok[out["best_mode"]]
    return out


def run_side_workload(args):
    """BASELINE configs[3] (MA-GP step, 45.51 GFLOP/img) and configs[4] (generator sampling, 15.013 GFLOP/img):
    same harness, device-resident inputs; the MA-GP step replays the trainer's CUDA graph (gradient exchange included),
    sampling is eager launches; one JSON line."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gan_codes_b200.engine import DFGANTrainer
    from gan_codes_b200.modules import Discriminator, Generator
    torch.manual_seed(123321)
    G, D = Generator(32, 100).to(dev), Discriminator(32).to(dev)
    if args.workload == "gp":
        B = (args.batch if args.batch != 32 else 64) // world
        gflop = 45.51
    else:
        B = args.batch if args.batch != 32 else 256
        gflop = 15.013
    trainer = DFGANTrainer(G, D, B, world=world)
    gen = torch.Generator().manual_seed(5 + rank)
    img = (torch.randn((B, 3, 256, 256), generator=gen) * 0.5).clamp_(-1, 1).to(dev)
    sent = (torch.randn((B, 256), generator=gen) * 0.3).to(dev)
    noise = torch.randn((B, 100), generator=gen).to(dev)
    fn = (lambda: trainer.gp_step(img, sent)) if args.workload == "gp" else (lambda: trainer.sample(noise, sent))
    for _ in range(max(3, args.warmup)):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        _, _, tf_sust, how = _peaks()
        value = B * world * args.steps / (ms.item() / 1e3)
        print(json.dumps({
            "metric": "MA-GP discriminator step images/sec" if args.workload == "gp" else "generator sampling images/sec",
            "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms.item() / args.steps, 3), "higher_is_better": True,
            "scaling": "strong" if args.workload == "gp" else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{args.workload} 256x256 n_channels 32, batch {B}/GPU", "batch_per_gpu": B,
                       "pct_of_bf16_peak_algorithmic": round(100 * value / world * gflop / 1e3 / tf_sust, 2),
                       "gflop_per_image_algorithmic": gflop, "peak_source": how}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--workload", default="full", choices=["full", "gp", "sample"],
                    help="full = headline training iteration (BASELINE configs[1]/[2]); gp = MA-GP discriminator "
                         "update only (configs[3], default batch 64 split over the ranks); sample = no_grad generator "
                         "forward (configs[4], default batch 256)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "full":
        run_side_workload(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
