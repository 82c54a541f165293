"""Fused training iteration: the step body of DeepFusionGAN.fit (deep_fusion_gan/model.py:173-228) executed as
three kernel programs over flat parameter / gradient / Adam-moment buffers.

    D step   : G forward (graph kept), D trunk on [real ; fake] as one 2B batch, the three logit heads (real,
               mismatched, fake) as one (3B-1)-row head call, hinge loss, backward, clip + Adam        (:173-197)
    MA-GP    : D forward on the real images, first-order sweep for d logits / d(image, sentence), tangent
               sweep + wgrad against the saved cotangents (App. C-3), clip + Adam                      (:202-210)
    G step   : twice-updated D on the retained fake images, D dgrad only, G backward, clip + Adam       (:215-228)

Data parallelism: one process per GPU, batch sharded, ONE all-reduce per optimiser step on the
flat gradient buffer (three per iteration), global-norm clip and Adam on the reduced gradient.  The mismatch
pairs are formed inside the local shard, exactly like N reference replicas with averaged gradients.

No host synchronisation happens inside `step`; losses are returned as device scalars.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import checkpoint as ckpt
from . import convops as co
from . import dnet, gnet
from . import kernels as K

F32, BF16 = torch.float32, torch.bfloat16
STAGE_MIXED = os.environ.get("DFGAN_STAGE_MIXED", "0") == "1"
# world > 1: "1" captures the NCCL all-reduces inside the iteration's CUDA graph (one replay per iteration).  Measured on
# 2 B200s (round 2): the iteration is NOT faster (11.56 ms against 11.48 ms with segments) and the process then hangs in
# dist.barrier() / destroy_process_group() at teardown, so the default stays four graph segments with eager all-reduces
# in between.
DP_ONE_GRAPH = os.environ.get("DFGAN_DP_ONE_GRAPH", "0") == "1"
# world > 1: "1" (default) exchanges the gradients with the library's own peer-memory all-reduce kernel (peer.py,
# csrc/peer_reduce.cu) instead of NCCL: the kernel is captured, so N ranks replay ONE graph per iteration, and the D
# step's low-resolution bucket (> 90 % of the bytes) is reduced on a side stream while the rest of the backward sweep
# runs.  "0": NCCL all-reduces between graph segments (round 1).
DP_PEER = os.environ.get("DFGAN_DP_PEER", "1") == "1"
DP_OVERLAP = os.environ.get("DFGAN_DP_NO_OVERLAP", "0") != "1"
# "1": sharded optimiser step with the peer-memory exchange: reduce-scatter of the gradient, global norm from per-rank
# partial sums, clip + Adam on the owned slices only, all-gather of the parameters (same bytes on the links, (N-1)/N
# of the clip + Adam passes gone from every rank).  The Adam moments are then current only on their owner:
# DFGANTrainer.sync_optimizer_state() (collective) before a checkpoint.
DP_SHARDED = os.environ.get("DFGAN_DP_SHARDED", "0") == "1"
EARLY_REAL_STEM = os.environ.get("DFGAN_NO_EARLY_STEM", "0") != "1"    # A/B: real-image stem next to the G forward
PEER_OVERLAP_CTAS = int(os.environ.get("DFGAN_PEER_OVERLAP_CTAS", "128"))


class FlatState:
    """Parameters of a module re-homed into one flat fp32 buffer (module parameters become views, so
    state_dict()/load_state_dict() keep working), with same-layout gradient and Adam-moment buffers."""

    def __init__(self, module: nn.Module, lr: float, betas=(0.0, 0.9), eps: float = 1e-8, clip: float = 5.0,
                 peer=None, sharded: bool = False):
        named = list(module.named_parameters())
        dev = named[0][1].device
        self.layout, total = ckpt.build_layout([(n, p.shape) for n, p in named])
        # data parallel over peer memory: the gradient buffer is the exchange buffer (mapped by every rank).  Sharded
        # optimiser step (reduce-scatter -> clip + Adam on the owned slices -> all-gather of the parameters): the
        # parameter buffer is exchanged too, and the moments are peer-mapped so that a checkpoint can collect the
        # slices the other ranks own (sync_moments).
        self.peer = peer
        self.sharded = bool(sharded and peer is not None)
        shared = (lambda: peer.empty(total)) if self.sharded else (lambda: torch.zeros(total, device=dev, dtype=F32))
        self.flat = shared()
        self.grad = peer.empty(total) if peer is not None else torch.zeros(total, device=dev, dtype=F32)
        self.m, self.v = shared(), shared()
        self.total = total
        # the ranges the gradient is exchanged in (fixed for the life of the state: in the sharded step the owner of
        # an element -- the rank that keeps its moments current -- follows from them); the trainer splits the
        # discriminator's buffer into [low-resolution tail, head of the buffer] for the overlapped exchange
        self.buckets = [(0, total)]
        self._done: list = []          # buckets exchanged since the last optimiser step
        self.moments_current = True    # sharded: False between an optimiser step and sync_moments()
        self.partial = torch.zeros(1, device=dev, dtype=F32)
        self._scratches: list = []
        if self.sharded:
            peer.enable_scalar_exchange()
        self.params: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for (n, p), (_, o, _shape) in zip(named, self.layout):
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.params[n] = view
                self.grads[n] = self.grad[o:o + p.numel()].view(p.shape)
        self.n_real = sum(p.numel() for _, p in named)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.sumsq = torch.zeros(1, device=dev, dtype=F32)
        self.norm_scratch = torch.zeros(4 * torch.cuda.get_device_properties(dev).multi_processor_count + 8,
                                        device=dev, dtype=F32)
        self.lr, self.betas, self.eps, self.clip = lr, betas, eps, clip

    def zero_grad(self) -> None:
        self.grad.zero_()

    def offset_of(self, prefix: str) -> int:
        """Offset (elements, a multiple of 4) of the first parameter whose name starts with `prefix`."""
        for name, o, _ in self.layout:
            if name.startswith(prefix):
                return o
        raise KeyError(prefix)

    def reduce(self, world: int, lo: int = 0, hi: Optional[int] = None, channel: int = 0,
               overlapped: bool = False) -> None:
        """Data-parallel exchange of elements [lo, hi) of the flat gradient (default: all of it; sum over the ranks,
        the 1/world factor is applied inside the Adam kernel).  Peer-memory kernel on the current stream (capturable)
        when the buffer came from a PeerGroup -- an all-reduce, or in the sharded step a reduce-scatter -- else NCCL
        (issued eagerly).  overlapped: the exchange runs next to other kernels -- the small-footprint kernel shape."""
        if world <= 1:
            return
        hi = self.total if hi is None else hi
        g = self.grad if (lo == 0 and hi == self.total) else self.grad[lo:hi]
        if g.numel() == 0:
            return
        self._done.append((lo, hi))
        if self.peer is None:
            torch.distributed.all_reduce(g)
            return
        kw = dict(channel=channel, ctas=PEER_OVERLAP_CTAS, threads=128) if overlapped else dict(channel=channel)
        if self.sharded:
            if (lo, hi) not in self.buckets:
                raise RuntimeError(f"sharded optimiser step: [{lo}, {hi}) is not one of the buckets {self.buckets}")
            self.peer.reduce_scatter(g, **kw)
        else:
            self.peer.all_reduce(g, **kw)

    def reduce_rest(self, world: int) -> None:
        """Exchange whatever has not been exchanged since the last optimiser step: everything in ONE call when nothing
        has been (all-reduce path), else the remaining buckets."""
        if world <= 1:
            return
        if not self._done and not self.sharded:
            self.reduce(world)
            return
        for b in self.buckets:
            if b not in self._done:
                self.reduce(world, b[0], b[1])

    def apply(self, world: int = 1) -> None:
        """global L2 norm -> clip -> Adam on the flat buffers (two kernels).  Sharded: the norm of the owned slices,
        summed over the ranks in rank order, clip + Adam on the owned slices, all-gather of the parameters."""
        done, self._done = self._done, []
        if not (self.sharded and world > 1):
            self.sumsq.zero_()
            K.sqnorm_det(self.grad, self.sumsq, self.norm_scratch)   # bit-identical on every data-parallel rank
            self.step_count.add_(1)
            K.adam_step(self.flat, self.grad, self.m, self.v, self.sumsq, self.step_count, lr=self.lr,
                        gscale=1.0 / world, max_norm=self.clip, b1=self.betas[0], b2=self.betas[1], eps=self.eps)
            return
        if sorted(done) != sorted(self.buckets):
            raise RuntimeError(f"sharded optimiser step: exchanged {done}, buckets are {self.buckets}")
        own = [self.peer.owned(lo, hi) for lo, hi in self.buckets]
        self.partial.zero_()
        while len(self._scratches) < len(own):      # one scratch per range: its layout depends on the range's grid
            self._scratches.append(torch.zeros_like(self.norm_scratch))
        for (lo, hi), scratch in zip(own, self._scratches):
            if hi > lo:
                K.sqnorm_det(self.grad[lo:hi], self.partial, scratch)
        self.peer.sum_scalar(self.partial, self.sumsq)
        self.step_count.add_(1)
        for lo, hi in own:
            if hi > lo:
                K.adam_step(self.flat[lo:hi], self.grad[lo:hi], self.m[lo:hi], self.v[lo:hi], self.sumsq,
                            self.step_count, lr=self.lr, gscale=1.0 / world, max_norm=self.clip, b1=self.betas[0],
                            b2=self.betas[1], eps=self.eps)
        for lo, hi in self.buckets:
            self.peer.all_gather(self.flat[lo:hi])
        self.moments_current = False

    def sync_moments(self) -> None:
        """Sharded step, COLLECTIVE (every rank, outside an iteration): fetch the Adam moments of the slices the other
        ranks own, so that state_dict() is the full optimiser state on every rank."""
        if not self.sharded:
            return
        dist = torch.distributed
        torch.cuda.synchronize(self.flat.device)
        dist.barrier()
        for lo, hi in self.buckets:
            for r in range(self.peer.world):
                if r == self.peer.rank:
                    continue
                olo, ohi = self.peer.owned(lo, hi, r)
                if ohi > olo:
                    self.m[olo:ohi].copy_(self.peer.peer_view(self.m[olo:ohi], r))
                    self.v[olo:ohi].copy_(self.peer.peer_view(self.v[olo:ohi], r))
        torch.cuda.synchronize(self.flat.device)
        dist.barrier()
        self.moments_current = True

    def optimizer_step(self, world: int = 1) -> None:
        self.reduce(world)
        self.apply(world)

    # torch.optim.Adam-compatible state (deep_fusion_gan/model.py:118-119, 305-306); reads the device step
    # counter, i.e. synchronises -- checkpoint time only, never inside an iteration
    def state_dict(self) -> dict:
        if self.sharded and not self.moments_current:
            raise RuntimeError("sharded optimiser step: every rank keeps only its slices of the Adam moments current; "
                               "call DFGANTrainer.sync_optimizer_state() on ALL ranks before reading the state")
        return ckpt.adam_state_dict(self.layout, self.m, self.v, int(self.step_count.item()), self.lr, self.betas,
                                    self.eps)

    def load_state_dict(self, sd: dict) -> None:
        # every check runs before anything is written: a rejected checkpoint leaves the live optimiser as it was
        groups = sd.get("param_groups", [])
        if len(groups) == 1 and getattr(self, "frozen", False):
            g = groups[0]
            new = (float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"]))
            if new != (self.lr, (float(self.betas[0]), float(self.betas[1])), self.eps):
                # hyper-parameters are launch arguments baked into a captured graph: adopt them only before capture
                raise RuntimeError("checkpoint Adam hyper-parameters differ from the ones captured in the CUDA "
                                   f"graph (lr {new[0]} betas {new[1]} eps {new[2]})")
        with torch.no_grad():
            h = ckpt.load_adam_state_dict(sd, self.layout, self.m, self.v)
            self.step_count.fill_(h["step"])
        self.lr, self.betas, self.eps = h["lr"], h["betas"], h["eps"]
        self.moments_current = True


class DFGANTrainer:
    """step(images, sentence_embeds, noise) == one iteration of the reference's training loop."""

    def __init__(self, generator: nn.Module, discriminator: nn.Module, batch: int, lr_g: float = 1e-4,
                 lr_d: float = 4e-4, betas=(0.0, 0.9), clip: float = 5.0, world: int = 1,
                 use_graph: Optional[bool] = None, text_image_loss: bool = False):
        if not next(generator.parameters()).is_cuda:
            raise RuntimeError("DFGANTrainer: modules must live on a CUDA device (no CPU path)")
        self.G, self.D = generator, discriminator
        self.batch, self.world = batch, world
        self.peer = None
        if world > 1 and DP_PEER:
            self.peer = self._peer_group(next(generator.parameters()).device)
        self.gs = FlatState(generator, lr_g, betas, clip=clip, peer=self.peer, sharded=DP_SHARDED)
        self.ds = FlatState(discriminator, lr_d, betas, clip=clip, peer=self.peer, sharded=DP_SHARDED)
        self._d_reduced_from: Optional[int] = None
        self._late_bucket = False
        # the reference trainer's attribute names (deep_fusion_gan/model.py:28-29, 42-43)
        self.generator, self.discriminator = generator, discriminator
        self.g_optim, self.d_optim = self.gs, self.ds
        with torch.no_grad():
            self.gpk = gnet.GPacked(self.gs.params)
            self.dpk = dnet.DPacked(self.ds.params)
            self.mlp = gnet.GMlp(self.gs.params, self.gpk.plan, batch, self.gs.grads)
        if self.peer is not None and DP_OVERLAP and dnet.EARLY_FLUSH and self.dpk.first_late > 0:
            # the D gradient is exchanged in two ranges: the low-resolution tail early (next to the rest of the D
            # step's backward sweep), the head of the buffer after the sweep
            lo = self.ds.offset_of(self.dpk.plan[self.dpk.first_late][0] + ".")
            self.ds.buckets = [(lo, self.ds.total), (0, lo)]
            self._late_bucket = True
        dev = self.gs.flat.device
        co._side_streams(dev)      # created here, never inside a stream capture
        self.ones = torch.ones(batch, device=dev, dtype=F32)
        # the text-image loss the reference's trainer logs next to the generator loss (deep_fusion_gan/model.py:
        # 220-223 -> damsm/loss.py); opt-in, `step` then also returns "txtimg_loss"
        self.text_image_loss = text_image_loss
        self._txt_partial = torch.empty(4 * batch * 3, device=dev, dtype=F32) if text_image_loss else None
        self.last: Dict[str, torch.Tensor] = {}
        # The iteration is ~400 kernel launches; replaying it as ONE CUDA graph removes the host launch cost
        # (the eager path is host-bound).  Inputs are staged into static buffers, two eager iterations warm up
        # every lazily built table, the third call captures.
        if use_graph is None:
            use_graph = os.environ.get("DFGAN_NO_GRAPH", "0") != "1"
        self.use_graph = use_graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._eager_calls = 0
        self._in = None

    @staticmethod
    def _peer_group(dev):
        """The peer-memory exchange group, or None (-> NCCL all-reduces between graph segments) when ANY rank cannot
        map its peers' buffers (GPUs without P2P access to each other, IPC handles refused): every rank takes the same
        path, and the reason is printed once."""
        import torch.distributed as dist
        from .peer import PeerGroup
        pg, why = None, ""
        try:
            pg = PeerGroup(dev)
        except Exception as e:  # noqa: BLE001  (any failure here means "no peer memory", decided collectively below)
            why = f"{type(e).__name__}: {e}"
        okay = torch.tensor([1 if pg is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(okay, op=dist.ReduceOp.MIN)
        if int(okay.item()) == 1:
            if DP_SHARDED:
                pg.multimem = False      # the sharded step's kernels are unicast: plain CUDA-IPC buffers
            return pg
        if dist.get_rank() == 0 or why:
            print(f"[dfgan] peer-memory gradient exchange unavailable ({why or 'another rank failed'}): "
                  "falling back to ncclAllReduce between graph segments", file=sys.stderr, flush=True)
        return None

    # ------------------------------------------------------------------------------------ phases
    def _d_step(self, images, sent, noise):
        b = self.batch
        dp, dpk = self.ds.params, self.dpk
        self.g_ctx: dict = {}
        # the stem of the real half does not wait for the generator: side stream, next to the generator forward;
        # the generated half follows when the images exist (no [real ; fake] concatenation)
        dev = images.device
        stem = dnet.stem_alloc(dpk, 2 * b, images.shape[2], images.shape[3], dev)
        if EARLY_REAL_STEM:
            with co.side(2, dev, co.POOL_AUX):
                dnet.stem_run(dpk, dp, images, stem, 0)
        else:
            dnet.stem_run(dpk, dp, images, stem, 0)
        fake = gnet.generator_forward(self.gpk, self.gs.params, self.mlp, noise, sent, self.g_ctx)
        dnet.stem_run(dpk, dp, fake, stem, b)
        co.join(dev, co.POOL_AUX)
        tctx: dict = {}
        emb = dnet.embeds_to_nchw(dnet.trunk_forward(dpk, dp, None, tctx, stem=stem))         # [2B, C, 4, 4]
        emb3 = torch.cat((emb[:b], emb[:b - 1], emb[b:]), dim=0)
        sent3 = torch.cat((sent, sent[1:b], sent), dim=0)
        hctx: dict = {}
        lg = dnet.head_forward(dpk, dp, emb3, sent3, hctx)
        # hinge loss over [real ; mismatched ; fake] and its logit cotangents: one launch (:175-189)
        loss, g_lg = torch.empty(1, device=lg.device, dtype=F32), torch.empty_like(lg)
        K.d_hinge(lg, loss, g_lg, b)
        loss = loss[0]
        self.ds.zero_grad()
        ge, _ = dnet.head_backward(dpk, dp, hctx, g_lg, self.ds.grads, False)
        g_emb = torch.cat((ge[:b], ge[2 * b - 1:]), dim=0)
        g_emb[:b - 1] += ge[b:2 * b - 1]
        self._d_reduced_from = None
        dnet.trunk_backward(dpk, dp, tctx, dnet.embeds_to_nhwc(g_emb), self.ds.grads, False,
                            on_late=self._reduce_late if self._late_bucket else None)
        return loss

    def _reduce_late(self, prefix: str) -> None:
        """Called by the trunk backward on its fold-back side stream once every gradient of the low-resolution blocks
        (from `prefix` on) and of the head is final: that tail of the flat buffer -- > 90 % of its bytes -- is
        exchanged next to the remaining (high-resolution, long-running) part of the sweep."""
        lo = self.ds.offset_of(prefix + ".")
        assert (lo, self.ds.total) == self.ds.buckets[0], (lo, self.ds.buckets)
        self.ds.reduce(self.world, lo=lo, channel=1, overlapped=True)
        self._d_reduced_from = lo

    def _reduce_d_step(self) -> None:
        self.ds.reduce_rest(self.world)

    def _d_update(self):
        self.ds.apply(self.world)
        # the large (low-resolution) operands are re-packed next to the start of the following forward sweep
        self.dpk.refresh(self.ds.params, defer_late=True)

    def _g_update(self):
        self.gs.apply(self.world)
        self.gpk.refresh(self.gs.params)

    def _gp_step(self, images, sent):
        b = self.batch
        dp, dpk = self.ds.params, self.dpk
        tctx, hctx, cots = {}, {}, {}
        emb = dnet.embeds_to_nchw(dnet.trunk_forward(dpk, dp, images, tctx, need_xcol=False))
        dnet.head_forward(dpk, dp, emb, sent, hctx)
        g_embed, g_sent = dnet.head_backward(dpk, dp, hctx, None, None, True, cots)
        gx = dnet.trunk_backward(dpk, dp, tctx, dnet.embeds_to_nhwc(g_embed), None, True, cots)
        sq = torch.zeros(b, device=gx.device, dtype=F32)
        K.rowsq(gx, sq)
        K.rowsq(g_sent, sq)
        # n = sqrt(sq + 1e-8) (:82), clamp(n, 0, 10) (:85), loss = 2 mean(n^6) (:203) and d loss / d g_b = coef_b * g_b
        # with coef_b = (12 / B) n_b^4 [n_b <= 10] (G11): one launch
        loss, coef = torch.empty(1, device=sq.device, dtype=F32), torch.empty_like(sq)
        K.gp_coef(sq, loss, coef)
        loss = loss[0]
        self.ds.zero_grad()
        t_emb = dnet.trunk_tangent(dpk, dp, tctx, cots, gx, coef, self.ds.grads)
        dnet.head_tangent(dpk, dp, hctx, cots, dnet.embeds_to_nchw(t_emb), g_sent, coef, self.ds.grads)
        return loss

    def _g_step(self, sent):
        b = self.batch
        dp, dpk = self.ds.params, self.dpk
        fake = self.g_ctx["img"]
        tctx, hctx = {}, {}
        emb = dnet.embeds_to_nchw(dnet.trunk_forward(dpk, dp, fake, tctx, need_xcol=False))
        lg = dnet.head_forward(dpk, dp, emb, sent, hctx)
        loss = -lg.mean()
        if self.text_image_loss:
            # on the images of this iteration's D step (pre-update generator), like the reference (:183, :222)
            txt = torch.empty(1, device=fake.device, dtype=F32)
            K.txtimg_loss(fake, sent, txt, self._txt_partial)
            self._seg_out["txtimg_loss"] = txt[0]
        g_embed, _ = dnet.head_backward(dpk, dp, hctx, self.ones * (-1.0 / b), None, False)
        g_img = dnet.trunk_backward(dpk, dp, tctx, dnet.embeds_to_nhwc(g_embed), None, True)
        self.gs.zero_grad()
        gnet.generator_backward(self.gpk, self.gs.params, self.mlp, self.g_ctx, g_img, self.gs.grads)
        self.g_ctx = {}
        return loss

    # ------------------------------------------------------------------------------------ public
    # The iteration as four segments separated by the three data-parallel gradient exchanges.  With one GPU the
    # segments are captured as ONE CUDA graph; with several each segment is its own graph and the NCCL
    # all-reduces run eagerly between the replays (NCCL is never captured).
    def _segments(self, images, sent, noise):
        out = self._seg_out

        def s0():
            out["d_loss"] = self._d_step(images, sent, noise)

        def s1():
            self._d_update()
            out["gp_loss"] = self._gp_step(images, sent)

        def s2():
            self._d_update()
            out["g_loss"] = self._g_step(sent)

        def s3():
            self._g_update()
        return [(s0, self._reduce_d_step), (s1, lambda: self.ds.reduce_rest(self.world)),
                (s2, lambda: self.gs.reduce_rest(self.world)), (s3, None)]

    def _run_eager(self, images, sent, noise) -> Dict[str, torch.Tensor]:
        self._seg_out = {}
        for fn, reduce in self._segments(images, sent, noise):
            fn()
            if reduce is not None:
                reduce()
        return self._seg_out

    def _capture(self) -> None:
        self._seg_out = {}
        segs = self._segments(*self._in)
        torch.cuda.synchronize()
        if self.world == 1 or self.peer is not None or DP_ONE_GRAPH:
            # data parallel: the gradient exchanges are captured too -- the peer-memory kernel is an ordinary launch
            # (with DFGAN_DP_PEER=0 DFGAN_DP_ONE_GRAPH=1: NCCL's, which ProcessGroupNCCL supports once the
            # communicator exists) -- so N ranks replay ONE graph per iteration exactly like one rank does.
            # thread_local: the NCCL watchdog thread may touch the CUDA API while this thread captures.
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                for fn, reduce in segs:
                    fn()
                    if reduce is not None:
                        reduce()
            self._graphs = [(g, None)]
        else:
            self._graphs = []
            pool = None
            for fn, reduce in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    fn()
                pool = g.pool()
                self._graphs.append((g, reduce))
        self._graph = self._graphs[0][0]
        self._graph_out = self._seg_out
        self.gs.frozen = self.ds.frozen = True

    def _modules_stale(self) -> None:
        """The Adam kernel writes the parameters in place without touching their autograd version counters, which
        the drop-in modules' own operand caches are keyed on (modules._PackCache): drop those caches whenever the
        trainer has stepped, so `generator(noise, sent)` / `discriminator.build_embeds` called between iterations
        re-pack from the current weights."""
        for m in (self.G, self.D):
            cache = getattr(m, "_cache", None)
            if cache is not None:
                cache.key = None
                cache.packed = None

    def sync_optimizer_state(self) -> None:
        """Sharded optimiser step (DFGAN_DP_SHARDED=1, world > 1) only, COLLECTIVE: every rank collects the Adam moments
        of the slices the other ranks keep current; afterwards checkpoint() / g_optim.state_dict() hold the full state
        on every rank.  A no-op otherwise."""
        for st in (self.gs, self.ds):
            st.sync_moments()

    def _moments_stale(self) -> None:
        for st in (self.gs, self.ds):
            if st.sharded and self.world > 1:
                st.moments_current = False

    def weights_changed(self) -> None:
        """Call after writing the parameters from outside (load_state_dict, manual edits): re-packs the
        tensor-core operand copies and drops the retained generator graph."""
        with torch.no_grad():
            self.gpk.refresh(self.gs.params)
            self.dpk.refresh(self.ds.params)
        self.g_ctx = {}
        self._modules_stale()

    @torch.no_grad()
    def step(self, images: torch.Tensor, sentence_embeds: torch.Tensor, noise: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One training iteration.  Inputs may live on the host (pinned memory recommended) or on the device:
        images fp32 [B,3,S,S] in [-1,1], sentence_embeds fp32 [B,256], noise fp32 [B,latent].  Returns the
        three losses as device scalars (no synchronisation; with graphs the tensors are overwritten by the
        next call)."""
        dev = self.gs.flat.device
        if images.shape[0] != self.batch:
            raise RuntimeError(f"DFGANTrainer was built for batch {self.batch}, got {images.shape[0]}")
        self._modules_stale()
        self._moments_stale()
        if self._in is None:
            mk = lambda t: torch.empty(t.shape, device=dev, dtype=F32)  # noqa: E731
            self._in = (mk(images), mk(sentence_embeds), mk(noise))
            self._stage = (mk(images), mk(sentence_embeds), mk(noise))
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._ev_h2d, self._ev_free = torch.cuda.Event(), torch.cuda.Event()
            self._ev_free.record()
        srcs = (images, sentence_embeds, noise)
        if all(not t.is_cuda for t in srcs):
            # host inputs: the H2D copy goes to a staging buffer on a side stream, so it overlaps the previous
            # iteration still running on the main stream; a device-to-device copy then fills the static inputs
            main = torch.cuda.current_stream()
            self._copy_stream.wait_event(self._ev_free)          # previous staging -> static copy has finished
            with torch.cuda.stream(self._copy_stream):
                for dst, src in zip(self._stage, srcs):
                    dst.copy_(src, non_blocking=True)
                self._ev_h2d.record()
            main.wait_event(self._ev_h2d)
            for dst, src in zip(self._in, self._stage):
                dst.copy_(src, non_blocking=True)
            self._ev_free.record()
        elif STAGE_MIXED and any(not t.is_cuda for t in srcs):
            # opt-in (DFGAN_STAGE_MIXED=1, not yet measured): host tensors among device tensors take the same
            # overlapping route as the all-host case, device tensors are copied on the main stream as below
            host = [not t.is_cuda for t in srcs]
            main = torch.cuda.current_stream()
            self._copy_stream.wait_event(self._ev_free)
            with torch.cuda.stream(self._copy_stream):
                for dst, src, h in zip(self._stage, srcs, host):
                    if h:
                        dst.copy_(src, non_blocking=True)
                self._ev_h2d.record()
            main.wait_event(self._ev_h2d)
            for dst, st, src, h in zip(self._in, self._stage, srcs, host):
                dst.copy_(st if h else src, non_blocking=True)
            self._ev_free.record()
        else:
            for dst, src in zip(self._in, srcs):
                dst.copy_(src, non_blocking=True)
        if not self.use_graph or self._eager_calls < 2:
            first = self._eager_calls == 0
            self._eager_calls += 1
            if first:
                # learn which copy (tensor-map layout / shared-memory image) of every weight operand this shape
                # configuration consumes, then stop re-packing the other one after each optimiser step
                co.USED.clear()
                co.TRACK_USAGE = True
            self.last = self._run_eager(*self._in)
            if first:
                co.TRACK_USAGE = False
                self.dpk.prune(co.USED)
                self.gpk._pp.prune(co.USED)
            return self.last
        if self._graph is None:
            self._capture()
        for g, reduce in self._graphs:
            g.replay()
            if reduce is not None:
                reduce()
        self.last = self._graph_out
        return self.last

    @torch.no_grad()
    def gp_step(self, images: torch.Tensor, sentence_embeds: torch.Tensor) -> torch.Tensor:
        """MA-GP discriminator update only (BASELINE config 4): gradient penalty on the real images, backward
        through the double-backward sweep, gradient exchange, clip + Adam (deep_fusion_gan/model.py:202-210).  Like
        `step`: inputs are staged into static buffers, two eager calls warm up, the third captures ONE CUDA graph
        (exchange included when it is the peer-memory kernel) that later calls replay; the returned loss is a device
        scalar that the next call overwrites."""
        dev = self.gs.flat.device
        self._modules_stale()
        self._moments_stale()
        if getattr(self, "_gp_in", None) is None:
            self._gp_in = (torch.empty(images.shape, device=dev, dtype=F32),
                           torch.empty(sentence_embeds.shape, device=dev, dtype=F32))
            self._gp_graph, self._gp_calls, self._gp_out = None, 0, None
        for dst, src in zip(self._gp_in, (images, sentence_embeds)):
            if dst.shape != src.shape:
                raise RuntimeError(f"gp_step was first called with {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)

        def body():
            loss = self._gp_step(*self._gp_in)
            self.ds.reduce_rest(self.world)
            self._d_update()
            self.dpk.sync_late()
            return loss

        graphable = self.use_graph and (self.world == 1 or self.peer is not None)
        if not graphable or self._gp_calls < 2:
            self._gp_calls += 1
            return body()
        if self._gp_graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._gp_out = body()
            self._gp_graph = g
            self.ds.frozen = True
        self._gp_graph.replay()
        return self._gp_out

    @torch.no_grad()
    def sample(self, noise: torch.Tensor, sentence_embeds: torch.Tensor) -> torch.Tensor:
        """no_grad generator forward (sample.py:13-18, BASELINE config 5)."""
        dev = self.gs.flat.device
        n = noise.shape[0]
        mlp = self.mlp if n == self.batch else gnet.GMlp(self.gs.params, self.gpk.plan, n, None)
        return gnet.generator_forward(self.gpk, self.gs.params, mlp, noise.to(dev, F32).contiguous(),
                                      sentence_embeds.to(dev, F32).contiguous(), None)

    # ------------------------------------------------------------------------------------ checkpoints
    def checkpoint(self, epoch: int, **histories) -> dict:
        """The reference's checkpoint dictionary (deep_fusion_gan/model.py:300-312) for the current state;
        histories: g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses (lists)."""
        torch.cuda.synchronize(self.gs.flat.device)
        # parameters are views of the flat buffer: clone, so the file holds one compact storage per tensor (like the
        # reference's) and the dictionary does not change under the caller when training continues
        own = lambda sd: type(sd)((k, v.detach().clone()) for k, v in sd.items())  # noqa: E731
        return ckpt.make_checkpoint(epoch, own(self.G.state_dict()), own(self.D.state_dict()),
                                    self.gs.state_dict(), self.ds.state_dict(), histories)

    def save_checkpoint(self, save_dir: str, epoch: int, **histories):
        """_save_checkpoint + _save_gen_weights (:298-327): `checkpoint.pt`, every tenth epoch
        `checkpoint_epoch_{e}.pt`, and `gen_{e}.pth`.  Returns the paths written."""
        return ckpt.write_checkpoint(self.checkpoint(epoch, **histories), save_dir)

    def load_from_checkpoint(self, checkpoint_path: str):
        """load_from_checkpoint (:106-133): restores both networks and both optimisers from a checkpoint written
        by this trainer or by the reference's, re-packs the tensor-core operand copies of the weights, and returns
        (epoch, g_losses, d_losses, d_gp_losses, is_scores, fid_scores, txtimg_losses)."""
        dev = self.gs.flat.device
        ck = ckpt.read_checkpoint(checkpoint_path, map_location=dev)
        torch.cuda.synchronize(dev)
        # parameters are views of the flat buffers, so load_state_dict's in-place copies fill them
        self.G.load_state_dict(ck["generator"])
        self.D.load_state_dict(ck["discriminator"])
        self.gs.load_state_dict(ck["g_optimizer"])
        self.ds.load_state_dict(ck["d_optimizer"])
        self.weights_changed()
        return (ck["epoch"],) + ckpt.histories_of(ck)
