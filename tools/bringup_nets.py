"""GPU bring-up of the whole-network programs against the oracle (torch fp32 on the GPU, TF32 off).
Usage: python tools/bringup_nets.py <case>|all   -- prints a per-tensor relative-error table."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nchw(x):
    return x.permute(0, 3, 1, 2)


def _states(depth=6, batch=2, gated=True, ws=3.0):
    torch = _setup()
    from oracle import dfgan_oracle as O
    gr = (0.2, 1.0) if gated else None
    g_sd = O.to(O.init_generator_state(32, 100, depth=depth, gamma_range=gr), device="cuda")
    d_sd = O.to(O.init_discriminator_state(32, depth=depth, gamma_range=gr, weight_scale=ws), device="cuda")
    images, sent, noise = (t.cuda() for t in O.synthetic_batch(batch, 4 * 2 ** depth))
    return torch, O, g_sd, d_sd, images, sent, noise


def _report(rows, tol):
    bad = 0
    for name, err in rows:
        flag = "" if err < tol else "   <-- FAIL"
        bad += err >= tol or err != err
        print(f"  {name:55s} {err:.3e}{flag}")
    return bad == 0


def _bf(t):
    return t.bfloat16().float()


def _loss_rows(it, ours, ref):
    """bf16-vs-fp32 loss tolerances for a free-running comparison: the hinge D loss is a mean of O(1) terms
    (1e-2 relative); the MA-GP loss is n^6, i.e. six times the relative error of the gradient norm (1e-1); the G
    loss is a mean of raw logits that cancel to ~0, so it is compared on the logit scale (absolute)."""
    rows = [(f"it{it} d_loss (rel)", abs(ours["d"] - ref["d"].item()) / abs(ref["d"].item()))]
    # iterations after the first compare two free-running trajectories that have already taken different
    # +-lr Adam steps (beta1 = 0 makes the first update lr*sign(g)); they only need to stay close
    loose = 1.0 if it == 0 else 3.0
    ok = _report(rows, 1e-2 * loose)
    ok &= _report([(f"it{it} gp_loss (rel)", abs(ours["gp"] - ref["gp"].item()) / abs(ref["gp"].item()))], 1e-1 * loose)
    ok &= _report([(f"it{it} g_loss (abs)", abs(ours["g"] - ref["g"].item()))], 1e-1 * loose)
    return ok


def _ref_modules(torch, O, g_sd, d_sd):
    from gan_codes_b200.modules import Discriminator, Generator
    G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    return G, D


@case
def dfwd():
    """Forward of trunk + head: per-layer against the emulation fed with the pipeline's own inputs (tight), and
    end-to-end against the free-running emulation and the fp32 oracle (informational)."""
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import dnet
    from oracle import dfgan_emul as E
    pk = dnet.DPacked(d_sd)
    ctx, hctx = {}, {}
    emb = dnet.trunk_forward(pk, d_sd, images, ctx)
    logits = dnet.head_forward(pk, d_sd, dnet.embeds_to_nchw(emb), sent, hctx)
    torch.cuda.synchronize()
    q = E.Rounding(True)
    force = E.force_from_trunk_ctx(ctx, hctx)
    force["__computed__"] = comp = {}
    with torch.no_grad():
        emb_f = E.discriminator_embeds(d_sd, images, q, None, force)
        log_f = E.discriminator_logits(d_sd, emb_f, sent, q, force)
        emb_e = E.discriminator_embeds(d_sd, images, q)
        log_e = E.discriminator_logits(d_sd, emb_e, sent, q)
        emb_o = O.discriminator_embeds(d_sd, images)
        log_o = O.discriminator_logits(d_sd, emb_o, sent)
    rows = [(k, rel(_bf(force[k]), _bf(v))) for k, v in comp.items()]
    rows.append(("logits (forced inputs)", rel(logits, log_f.reshape(-1))))
    ok = _report(rows, 2e-3)
    print("  -- free-running (informational): embeds vs emul %.3e, vs fp32 %.3e; logits vs emul %.3e, vs fp32 %.3e"
          % (rel(nchw(emb), emb_e), rel(nchw(emb), emb_o), rel(logits, log_e.reshape(-1)), rel(logits, log_o.reshape(-1))))
    return ok


@case
def dbwd():
    """First-order backward of head + trunk for L = sum_b c_b * logit_b, against the teacher-forced emulation."""
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import dnet
    from gan_codes_b200.modules import _FlatGrads
    from oracle import dfgan_emul as E
    pk = dnet.DPacked(d_sd)
    ctx, hctx = {}, {}
    emb = dnet.trunk_forward(pk, d_sd, images, ctx)
    logits = dnet.head_forward(pk, d_sd, dnet.embeds_to_nchw(emb), sent, hctx)
    cvec = torch.tensor([0.7, -1.3], device="cuda")[:images.shape[0]]
    names = list(d_sd)
    fg = _FlatGrads(names, [d_sd[k] for k in names])
    g_embed, g_sent = dnet.head_backward(pk, d_sd, hctx, cvec, fg.views, True)
    cots = {}
    gx = dnet.trunk_backward(pk, d_sd, ctx, dnet.embeds_to_nhwc(g_embed), fg.views, True, cots)
    torch.cuda.synchronize()
    force = E.force_from_trunk_ctx(ctx, hctx)
    dl = {k: v.clone().requires_grad_() for k, v in d_sd.items()}
    x = images.clone().requires_grad_()
    s = sent.clone().requires_grad_()
    q = E.Rounding(True)
    lg = E.discriminator_logits(dl, E.discriminator_embeds(dl, x, q, None, force), s, q, force).reshape(-1)
    gref = torch.autograd.grad((lg * cvec).sum(), [x, s] + list(dl.values()))
    rows = [("logits", rel(logits, lg)), ("d/d image", rel(gx, gref[0])), ("d/d sentence", rel(g_sent, gref[1]))]
    ok = True
    for k, g in zip(names, gref[2:]):
        if k.endswith(".gamma"):
            # d gamma = <g_y, r> over ~1e6 elements cancels to a small scalar, so its error is measured against the
            # un-cancelled sum |g_y|.|r| (what bf16 rounding of g_y acts on), not against the scalar itself
            blk = int(k.split(".")[1]) - 1
            g_y, r = cots["blocks"][blk][0].float(), ctx["saved"][blk][2].float()
            scale = float((g_y.abs() * r.abs()).sum())
            ok &= _report([(f"grad {k} (vs sum|g_y||r|)", abs(float(fg.views[k]) - float(g)) / scale)], 2e-3)
        else:
            rows.append(("grad " + k, rel(fg.views[k], g)))
    return _report(rows, 1e-2) and ok


@case
def gp():
    """MA-GP through the drop-in modules (autograd.grad(create_graph) + backward) vs the forced emulation."""
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    import gan_codes_b200.modules as M
    from oracle import dfgan_emul as E
    M.KEEP_DEBUG = True
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    x = images.clone().requires_grad_()
    s = sent.clone().requires_grad_()
    logits = D.get_logits(D.build_embeds(x), s)
    gx, gs = torch.autograd.grad(logits, (x, s), torch.ones_like(logits), retain_graph=True, create_graph=True)
    g = torch.cat((gx.reshape(x.shape[0], -1), gs.reshape(x.shape[0], -1)), dim=1)
    n = torch.clamp(torch.sqrt((g ** 2).sum(dim=1) + 1e-8), 0, 10)
    loss = 2.0 * torch.mean(n ** 6)
    loss.backward()
    torch.cuda.synchronize()
    force = E.force_from_trunk_ctx(D._debug["trunk"], D._debug["head"])
    dl = {k: v.clone().requires_grad_() for k, v in d_sd.items()}
    probe = {}
    loss_ref = E.magp_loss(dl, images, sent, E.Rounding(True), probe, force)
    gref = torch.autograd.grad(loss_ref, list(dl.values()), allow_unused=True)
    p32 = {}
    loss32 = O.magp_loss(dl, images, sent, p32)
    print("  norms emul", probe["gp_n"].tolist(), "ours", n.tolist(), "fp32", p32["gp_n"].tolist())
    rows = [("gp_gx", rel(gx, probe["gp_gx"])), ("gp_gs", rel(gs, probe["gp_gs"])),
            ("gp_loss", abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()))]
    for (k, p), gr in zip(D.named_parameters(), gref):
        if gr is None:
            rows.append(("gp grad (zero) " + k, float(p.grad.abs().max()) if p.grad is not None else 1.0))
        else:
            rows.append(("gp grad " + k, rel(p.grad, gr)))
    return _report(rows, 1e-2)


@case
def gfwd():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import gnet
    from oracle import dfgan_emul as E
    pk = gnet.GPacked(g_sd)
    mlp = gnet.GMlp(g_sd, pk.plan, noise.shape[0], None)
    ctx = {}
    img = gnet.generator_forward(pk, g_sd, mlp, noise, sent, ctx)
    torch.cuda.synchronize()
    rows = []
    for (p, fb, m), (o, wdt) in list(mlp.offsets.items())[:4]:
        rows.append((f"mlp {p}.fusion_block_{fb}.{m}", rel(mlp.tab[:, o:o + wdt], O._mlp(g_sd, f"{p}.fusion_block_{fb}.{m}", sent))))
    names = [p[0] for p in pk.plan]
    force = E.force_from_g_ctx(names, ctx)
    force["__computed__"] = comp = {}
    q = E.Rounding(True)
    with torch.no_grad():
        E.generator_forward(g_sd, noise, sent, q, None, force)
        img_e = E.generator_forward(g_sd, noise, sent, q)
        img_o = O.generator_forward(g_sd, noise, sent)
    rows += [(k, rel(force[k] if k == "img" else _bf(force[k]), v if k == "img" else _bf(v))) for k, v in comp.items()]
    ok = _report(rows, 2e-3)
    print("  -- free-running (informational): image vs emul %.3e, vs fp32 %.3e" % (rel(img, img_e), rel(img, img_o)))
    return ok


@case
def gbwd():
    """Generator backward for L = <c, image> (random fixed c), against the teacher-forced emulation."""
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import gnet
    from gan_codes_b200.modules import _FlatGrads
    from oracle import dfgan_emul as E
    pk = gnet.GPacked(g_sd)
    names_p = list(g_sd)
    fg = _FlatGrads(names_p, [g_sd[k] for k in names_p])
    mlp = gnet.GMlp(g_sd, pk.plan, noise.shape[0], fg.views)
    ctx = {}
    img = gnet.generator_forward(pk, g_sd, mlp, noise, sent, ctx)
    cimg = torch.randn(img.shape, generator=torch.Generator().manual_seed(5)).cuda() * 0.01
    gnet.generator_backward(pk, g_sd, mlp, ctx, cimg, fg.views)
    torch.cuda.synchronize()
    force = E.force_from_g_ctx([p[0] for p in pk.plan], ctx)
    gl = {k: v.clone().requires_grad_() for k, v in g_sd.items()}
    img_e = E.generator_forward(gl, noise, sent, E.Rounding(True), None, force)
    gref = torch.autograd.grad((img_e * cimg).sum(), list(gl.values()))
    rows = [("grad " + k, rel(fg.views[k], g)) for k, g in zip(names_p, gref)]
    gates = [r for r in rows if r[0].endswith(".gamma")]      # cancellation-dominated scalars, see dbwd
    rows = [r for r in rows if not r[0].endswith(".gamma")]
    worst = sorted(rows, key=lambda r: -r[1])[:30]
    ok = _report(worst, 1e-2) and _report(gates, 6e-2)
    print(f"  ({len(rows)} tensors, worst 30 shown; median {sorted(r[1] for r in rows)[len(rows) // 2]:.3e})")
    return ok


@case
def step_modules():
    """Full training iteration through the drop-in modules + torch.optim.Adam vs the fp32 oracle step (losses)."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=4, ws=1.0)
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    g_opt = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.0, 0.9))
    d_opt = torch.optim.Adam(D.parameters(), lr=4e-4, betas=(0.0, 0.9))
    og, od = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    ok = True
    for it in range(2):
        ref = O.train_step(g_sd, d_sd, og, od, images, sent, noise)
        b = images.shape[0]
        real_emb = D.build_embeds(images)
        l_real = torch.relu(1.0 - D.get_logits(real_emb, sent)).mean()
        l_mis = torch.relu(1.0 + D.get_logits(real_emb[:b - 1], sent[1:b])).mean()
        fake = G(noise, sent)
        l_fake = torch.relu(1.0 + D.get_logits(D.build_embeds(fake.detach()), sent)).mean()
        d_loss = l_real + (l_fake + l_mis) / 2.0
        d_opt.zero_grad(); g_opt.zero_grad()
        d_loss.backward()
        torch.nn.utils.clip_grad_norm_(D.parameters(), 5.0)
        d_opt.step()
        x = images.data.requires_grad_()
        s = sent.data.requires_grad_()
        lg = D.get_logits(D.build_embeds(x), s)
        grads = torch.autograd.grad(lg, (x, s), torch.ones_like(lg), retain_graph=True, create_graph=True)
        gcat = torch.cat((grads[0].reshape(b, -1), grads[1].reshape(b, -1)), dim=1)
        gp = 2.0 * torch.mean(torch.clamp(torch.sqrt(torch.sum(gcat ** 2, dim=1) + 1e-8), 0, 10) ** 6)
        d_opt.zero_grad(); g_opt.zero_grad()
        gp.backward()
        torch.nn.utils.clip_grad_norm_(D.parameters(), 5.0)
        d_opt.step()
        g_loss = -D.get_logits(D.build_embeds(fake), sent).mean()
        d_opt.zero_grad(); g_opt.zero_grad()
        g_loss.backward()
        torch.nn.utils.clip_grad_norm_(G.parameters(), 5.0)
        g_opt.step()
        torch.cuda.synchronize()
        print("  ours d %.5f gp %.6g g %.5f | oracle d %.5f gp %.6g g %.5f" % (
            d_loss.item(), gp.item(), g_loss.item(), ref["d"].item(), ref["gp"].item(), ref["g"].item()))
        ok &= _loss_rows(it, {"d": d_loss.item(), "gp": gp.item(), "g": g_loss.item()}, ref)
    return ok


@case
def engine():
    """DFGANTrainer.step (fused engine) vs the fp32 oracle step: losses over 3 iterations, then parameters."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=4, ws=1.0)
    from gan_codes_b200.engine import DFGANTrainer
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    tr = DFGANTrainer(G, D, 4)
    og, od = O.AdamState(g_sd, 1e-4), O.AdamState(d_sd, 4e-4)
    ok = True
    for it in range(3):
        ref = O.train_step(g_sd, d_sd, og, od, images, sent, noise)
        out = tr.step(images, sent, noise)
        torch.cuda.synchronize()
        print("  ours d %.5f gp %.6g g %.5f | oracle d %.5f gp %.6g g %.5f" % (
            out["d_loss"].item(), out["gp_loss"].item(), out["g_loss"].item(), ref["d"].item(), ref["gp"].item(),
            ref["g"].item()))
        if it == 0:
            ok &= _loss_rows(it, {k: out[k + "_loss"].item() for k in ("d", "gp", "g")}, ref)
        else:
            # later iterations compare two INDEPENDENT trajectories: Adam(beta1 = 0) moves every weight by +-lr per
            # update, so noise-level gradient elements decide 2*lr of drift per step (batch 4: chaotic in g_loss).
            # Asserted instead: the same quantities at re-synchronised weights (tools/bringup_parity.py cfg1_step_b32),
            # the update rule against torch.optim.Adam, and the parameter bound below; these rows are informational.
            _loss_rows(it, {k: out[k + "_loss"].item() for k in ("d", "gp", "g")}, ref)
    # Adam invariant: |update| <= lr / sqrt(1 - beta2) per step, so the two parameter sets can differ by at most
    # 2 * 3.17 * lr * n_updates element-wise (6 D updates, 3 G updates) -- catches any scale / step-count error
    dsd, gsd = D.state_dict(), G.state_dict()
    d_max = max(float((dsd[k] - v).abs().max()) for k, v in d_sd.items())
    g_max = max(float((gsd[k] - v).abs().max()) for k, v in g_sd.items())
    print(f"  max |param - oracle param|: D {d_max:.3e} (bound {2 * 3.17 * 4e-4 * 6:.3e}), G {g_max:.3e} (bound {2 * 3.17 * 1e-4 * 3:.3e})")
    ok &= d_max <= 2 * 3.17 * 4e-4 * 6 and g_max <= 2 * 3.17 * 1e-4 * 3
    return ok


@case
def graph_side_consistency():
    """The CUDA-graph replay with side-stream branches and programmatic dependent launches must compute what
    the serial eager launches compute.  Three iterations bring the trainer to its captured state; the optimiser
    state is snapshotted, iteration 4 runs as a graph replay, the snapshot is restored and the same iteration runs
    again as serial eager launches: gradient buffers, losses and updated parameters agree up to atomic-order
    noise (Adam is sign-like at beta1 = 0, so runs from different states cannot be compared)."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=4, ws=1.0)
    from gan_codes_b200 import convops as co
    from gan_codes_b200.engine import DFGANTrainer
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    tr = DFGANTrainer(G, D, 4, use_graph=True)
    for _ in range(3):
        tr.step(images, sent, noise)
    torch.cuda.synchronize()
    assert tr._graph is not None
    bufs = [t for st in (tr.gs, tr.ds) for t in (st.flat, st.m, st.v, st.step_count)]
    snap = [t.clone() for t in bufs]

    def run():
        out = tr.step(images, sent, noise)
        torch.cuda.synchronize()
        return dict(dg=tr.ds.grad.clone(), gg=tr.gs.grad.clone(), dp=tr.ds.flat.clone(), gp=tr.gs.flat.clone(),
                    losses=[out[k].item() for k in ("d_loss", "gp_loss", "g_loss")])
    a = run()
    for t, s0 in zip(bufs, snap):
        t.copy_(s0)
    tr.dpk.refresh(tr.ds.params)
    tr.gpk.refresh(tr.gs.params)
    side_was, co.SIDE_ENABLED, tr.use_graph = co.SIDE_ENABLED, False, False
    try:
        b = run()
    finally:
        co.SIDE_ENABLED, tr.use_graph = side_was, True
    print("  losses graph+side", a["losses"], "| eager serial", b["losses"])
    # The MA-GP quantities scale with n^4 .. n^6 of a gradient norm that passes through six bf16-stored blocks:
    # ulp-level differences of the D update (atomic order) move it by ~0.3 %, i.e. 1-2 % on the GP loss / gradient;
    # a race would show up as an O(1) error in the parameters.  GP rows are therefore scaled by 1/4.
    rows = [("D gradient buffer (GP phase) / 4", rel(a["dg"], b["dg"]) / 4), ("G gradient buffer", rel(a["gg"], b["gg"])),
            ("D parameters", rel(a["dp"], b["dp"])), ("G parameters", rel(a["gp"], b["gp"]))]
    for i, (x, y) in enumerate(zip(a["losses"], b["losses"])):
        rows.append((f"loss {i}" + (" / 4" if i == 1 else ""), abs(x - y) / (abs(y) + 1e-6) / (4 if i == 1 else 1)))
    return _report(rows, 2e-2)


@case
def checkpoint():
    """Reference-format checkpoints (deep_fusion_gan/model.py:106-133, 298-327) through the trainer: trainer A runs
    four iterations (the last two as graph replays) and saves; trainer B -- fresh default-initialised networks that
    have already taken a step of their own -- loads the file.  B's parameters, Adam moments and step counters then
    equal A's bit for bit, its re-packed tensor-core operands produce bit-identical samples, the next iteration
    of both agrees up to atomic-order noise, A itself resumes from the file after its graph was captured, and
    torch.optim.Adam accepts the optimiser dictionaries."""
    import os
    import tempfile
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=4, ws=1.0)
    from gan_codes_b200.engine import DFGANTrainer
    from gan_codes_b200.modules import Discriminator, Generator
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    a = DFGANTrainer(G, D, 4, use_graph=True)
    for _ in range(4):
        a.step(images, sent, noise)
    assert a._graph is not None
    hist = dict(g_losses=[0.5], d_losses=[1.9], d_gp_losses=[0.1], is_scores=[1.5], fid_scores=[300.0],
                txtimg_losses=[3.0])
    G2, D2 = Generator(32, 100).cuda(), Discriminator(32).cuda()
    b = DFGANTrainer(G2, D2, 4, use_graph=False)
    b.step(images, sent, noise)
    ok = True
    with tempfile.TemporaryDirectory() as tmp:
        paths = a.save_checkpoint(tmp, 9, **hist)
        ok &= sorted(os.path.basename(p) for p in paths) == ["checkpoint.pt", "checkpoint_epoch_9.pt", "gen_9.pth"]
        ret = b.load_from_checkpoint(os.path.join(tmp, "checkpoint.pt"))
        ok &= ret[0] == 9 and list(ret[1:]) == [hist[k] for k in ("g_losses", "d_losses", "d_gp_losses", "is_scores",
                                                                 "fid_scores", "txtimg_losses")]
        exact = all(torch.equal(x, y) for sa, sb in ((a.gs, b.gs), (a.ds, b.ds))
                    for x, y in ((sa.flat, sb.flat), (sa.m, sb.m), (sa.v, sb.v), (sa.step_count, sb.step_count)))
        print("  restored parameters / moments / step counters bit-identical:", exact,
              "| steps G", int(b.gs.step_count), "D", int(b.ds.step_count))
        ok &= exact and int(b.gs.step_count) == 4 and int(b.ds.step_count) == 8
        rows = [("sample() after load vs source trainer", rel(b.sample(noise, sent), a.sample(noise, sent)))]
        ok &= _report(rows, 1e-6)

        # torch.optim.Adam takes the optimiser dictionaries (what the reference's load_from_checkpoint does, :118-119)
        ck = torch.load(os.path.join(tmp, "checkpoint.pt"), weights_only=False)
        for net, st, key in ((G2, b.gs, "g_optimizer"), (D2, b.ds, "d_optimizer")):
            opt = torch.optim.Adam(net.parameters(), lr=1.0)
            opt.load_state_dict(ck[key])
            same = all(torch.equal(opt.state[p]["exp_avg_sq"], st.v[o:o + p.numel()].view(p.shape))
                       for p, (_, o, _) in zip(net.parameters(), st.layout))
            print(f"  torch.optim.Adam.load_state_dict({key}): lr {opt.param_groups[0]['lr']}, moments equal: {same}")
            ok &= same and opt.param_groups[0]["lr"] == st.lr
        ok &= list(torch.load(os.path.join(tmp, "gen_9.pth"))) == list(G.state_dict())

        # the next iteration from the restored state: A as a graph replay, B as eager launches
        def run(tr):
            out = tr.step(images, sent, noise)
            torch.cuda.synchronize()
            return [out[k].item() for k in ("d_loss", "gp_loss", "g_loss")], tr.ds.flat.clone(), tr.gs.flat.clone()
        la, da, ga = run(a)
        lb, db, gb = run(b)
        # ... and A resumes from the file after capture: the replay reads the restored buffers
        a.load_from_checkpoint(os.path.join(tmp, "checkpoint.pt"))
        lc, dc, gc = run(a)
    print("  next-iteration losses: source", la, "| loaded", lb, "| source reloaded", lc)
    rows = [("D parameters after the next iteration (loaded)", rel(db, da)), ("G parameters (loaded)", rel(gb, ga)),
            ("D parameters (source reloaded)", rel(dc, da)), ("G parameters (source reloaded)", rel(gc, ga))]
    for name, l in (("loaded", lb), ("reloaded", lc)):
        for i, (x, y) in enumerate(zip(l, la)):
            rows.append((f"loss {i} ({name})" + (" / 4" if i == 1 else ""),
                         abs(x - y) / (abs(y) + 1e-6) / (4 if i == 1 else 1)))
    return ok & _report(rows, 2e-2)


@case
def fp32_kernels():
    """dfgan_conv2d_f32_{fprop,dgrad,wgrad} (through gan_codes_b200.fp32.conv2d and torch autograd, first and
    second order) vs torch fp64 for every convolution geometry of the two networks (+ ragged sizes)."""
    torch = _setup()
    import torch.nn.functional as F
    from gan_codes_b200 import fp32
    g = torch.Generator(device="cuda").manual_seed(5)
    rows = []
    #        N  Cin Cout H   W   k  st pad
    geos = [(2, 32, 32, 17, 19, 3, 1, 1), (2, 16, 24, 16, 16, 4, 2, 1), (2, 40, 24, 8, 8, 1, 1, 0),
            (3, 64, 1, 4, 4, 4, 1, 0), (4, 100, 72, 1, 1, 1, 1, 0), (2, 3, 32, 20, 20, 3, 1, 1),
            (1, 24, 3, 33, 18, 3, 1, 1), (2, 8, 40, 34, 22, 4, 2, 1)]
    for (n, ci, co, h, w, k, st, pad) in geos:
        x = torch.randn(n, ci, h, w, device="cuda", generator=g).requires_grad_()
        wt = (torch.randn(co, ci, k, k, device="cuda", generator=g) * 0.2).requires_grad_()
        b = torch.randn(co, device="cuda", generator=g).requires_grad_()
        xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, wt, b))
        y = fp32.conv2d(x, wt, b, st, pad)
        yd = F.conv2d(xd, wd, bd, st, pad)
        gy = torch.randn_like(y)
        gx, gw = torch.autograd.grad(y, (x, wt), gy, create_graph=True)
        gxd, gwd = torch.autograd.grad(yd, (xd, wd), gy.double(), create_graph=True)
        # second order: d/d(w, x) of |gx|^2 + <gw, gw> (exercises all six cross terms of the three operators)
        s2 = (gx ** 2).sum() + (gw ** 2).sum()
        s2d = (gxd ** 2).sum() + (gwd ** 2).sum()
        hx, hw = torch.autograd.grad(s2, (x, wt))
        hxd, hwd = torch.autograd.grad(s2d, (xd, wd))
        tag = f"N{n} {ci}->{co} {h}x{w} k{k} s{st} p{pad}"
        rows += [(tag + " fprop", rel(y, yd)), (tag + " dgrad", rel(gx, gxd)), (tag + " wgrad", rel(gw, gwd)),
                 (tag + " 2nd-order d/dx", rel(hx, hxd)), (tag + " 2nd-order d/dw", rel(hw, hwd))]
    lin = fp32.linear(torch.ones(3, 7, device="cuda"), torch.ones(5, 7, device="cuda"), torch.ones(5, device="cuda"))
    rows.append(("linear known answer (8)", float((lin - 8.0).abs().max())))
    return _report(rows, 2e-5)


@case
def fp32_tier():
    """north_star's fp32 / TF32-off tier: the drop-in modules in precision "fp32" (CUDA-core fp32 kernels) against
    the fp64 oracle on identical weights and inputs -- per-block activations, the three losses and every gradient
    of the three phases (D hinge, MA-GP double backward, G), all within rel 1e-4, free running (no teacher
    forcing)."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=2, ws=2.0)
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    G.set_precision("fp32"), D.set_precision("fp32")
    from gan_codes_b200 import fp32
    g64, d64 = O.to(g_sd, dtype=torch.float64), O.to(d_sd, dtype=torch.float64)
    im64, se64, no64 = images.double(), sent.double(), noise.double()
    b = images.shape[0]
    rows = []
    # ---- per-block activations
    ref_act = {}
    tap = lambda name, t: ref_act.__setitem__(name, t.detach())  # noqa: E731
    with torch.no_grad():
        img_o = O.generator_forward(g64, no64, se64, tap)
        emb_o = O.discriminator_embeds(d64, im64, tap)
        tg, td = {}, {}
        img = fp32.generator_forward(G, noise, sent, tg)
        emb = fp32.build_embeds(D, images, td)
    for name, t in list(tg.items()) + list(td.items()):
        rows.append(("act " + name, rel(t, ref_act[name])))
    rows += [("act G image", rel(img, img_o)), ("act D embeds", rel(emb, emb_o))]

    def grad_rows(tag, module, ref, ref32):
        """Every gradient tensor against the reference arithmetic in fp32 (torch, TF32 off -- north_star's target)
        and in fp64 (arbiter); the smaller deviation counts.  fp32 arithmetic itself is not 1e-4-exact on every
        tensor: the MA-GP coefficient n^4 quadruples the rounding error of the norm, a LeakyReLU whose
        pre-activation rounds to the other side of zero flips one mask element, scalar gate gradients are
        cancellation-dominated -- the two oracles differ from each other by 1e-4..3e-4 on such tensors (printed).
        Criterion: >= 85 % of the tensors within 1e-4 (one flipped mask element puts all ~26 tensors of its
        generator block above it), every tensor within 1e-3, exact zeros exactly zero."""
        errs, n_zero_bad = [], 0
        for k, p in module.named_parameters():
            if k not in ref:
                continue
            r = ref[k]
            if float(r.norm()) == 0.0:
                n_zero_bad += int(p.grad is not None and float(p.grad.abs().max()) != 0.0)
                continue
            # the target is the reference's fp32 path; fp64 is the arbiter where fp32 noise separates the two
            e32, e64 = rel(p.grad, ref32[k]), rel(p.grad, r)
            e = min(e32, e64)
            errs.append(e)
            if e > 1e-4:
                print(f"    {tag} {k}: vs fp32 oracle {e32:.3e}, vs fp64 oracle {e64:.3e} "
                      f"(fp32 oracle vs fp64 oracle: {rel(ref32[k], r):.3e})")
        within = sum(e <= 1e-4 for e in errs) / len(errs)
        print(f"  {tag}: {len(errs)} gradient tensors, {100 * within:.1f} % within 1e-4, worst {max(errs):.3e}")
        return [(f"{tag}: share of tensors above 1e-4 (scaled: 15 % = 1e-4)", (1.0 - within) * 1e-4 / 0.15),
                (f"{tag}: worst tensor / 10", max(errs) / 10.0),
                (f"{tag}: exact-zero tensors that are not zero", float(n_zero_bad))]

    g32, d32 = O.to(g_sd, dtype=torch.float32), O.to(d_sd, dtype=torch.float32)

    # ---- D step (deep_fusion_gan/model.py:173-193)
    gl, dl = O._leaf(g64), O._leaf(d64)
    fake_o = O.generator_forward(gl, no64, se64)
    d_loss_o = O.d_hinge_loss(dl, im64, fake_o.detach(), se64)
    ref = O._grads(d_loss_o, dl, zero_if_unused=False)
    gl32, dl32 = O._leaf(g32), O._leaf(d32)
    fake_32 = O.generator_forward(gl32, noise, sent)
    ref32 = O._grads(O.d_hinge_loss(dl32, images, fake_32.detach(), sent), dl32, zero_if_unused=False)
    fake = G(noise, sent)
    real_emb = D.build_embeds(images)
    d_loss = torch.relu(1.0 - D.get_logits(real_emb, sent)).mean() + 0.5 * (
        torch.relu(1.0 + D.get_logits(D.build_embeds(fake.detach()), sent)).mean() +
        torch.relu(1.0 + D.get_logits(real_emb[:b - 1], sent[1:b])).mean())
    D.zero_grad(), G.zero_grad()
    d_loss.backward()
    rows.append(("d_loss", abs(d_loss.item() - d_loss_o.item()) / abs(d_loss_o.item())))
    rows += grad_rows("D-step grad", D, ref, ref32)
    # ---- MA-GP step (:59-85, :202-207)
    dl = O._leaf(d64)
    gp_o = O.magp_loss(dl, im64, se64)
    ref = O._grads(gp_o, dl, zero_if_unused=True)
    dl32 = O._leaf(d32)
    ref32 = O._grads(O.magp_loss(dl32, images, sent), dl32, zero_if_unused=True)
    x, s = images.data.requires_grad_(), sent.data.requires_grad_()
    lg = D.get_logits(D.build_embeds(x), s)
    gx, gs = torch.autograd.grad(lg, (x, s), torch.ones_like(lg), retain_graph=True, create_graph=True)
    gcat = torch.cat((gx.reshape(b, -1), gs.reshape(b, -1)), dim=1)
    gp = 2.0 * torch.mean(torch.clamp(torch.sqrt(torch.sum(gcat ** 2, dim=1) + 1e-8), 0, 10) ** 6)
    D.zero_grad()
    gp.backward()
    rows.append(("gp_loss", abs(gp.item() - gp_o.item()) / abs(gp_o.item())))
    rows += grad_rows("MA-GP grad", D, ref, ref32)
    # ---- G step (:215-225)
    g_loss_o = O.g_hinge_loss(d64, fake_o, se64)
    ref = O._grads(g_loss_o, gl, zero_if_unused=False)
    ref32 = O._grads(O.g_hinge_loss(d32, fake_32, sent), gl32, zero_if_unused=False)
    g_loss = -D.get_logits(D.build_embeds(fake), sent).mean()
    G.zero_grad(), D.zero_grad()
    g_loss.backward()
    rows.append(("g_loss (abs, logit scale)", abs(g_loss.item() - g_loss_o.item())))
    rows += grad_rows("G-step grad", G, ref, ref32)
    torch.cuda.synchronize()
    return _report(rows, 1e-4)


@case
def gp_identity_full():
    """Size-independent property at the benchmark batch (B = 32, 256x256): the tangent sweep in direction
    v = coef * (gx, gs) must return tangent logits coef_b * (|gx_b|^2 + |gs_b|^2) -- ties the first-order dgrad
    chain to the tangent-forward chain without any oracle."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=32, ws=2.0)
    from gan_codes_b200 import dnet, kernels as K
    from gan_codes_b200.modules import _FlatGrads
    pk = dnet.DPacked(d_sd)
    tctx, hctx, cots = {}, {}, {}
    emb = dnet.embeds_to_nchw(dnet.trunk_forward(pk, d_sd, images, tctx))
    dnet.head_forward(pk, d_sd, emb, sent, hctx)
    g_embed, g_sent = dnet.head_backward(pk, d_sd, hctx, None, None, True, cots)
    gx = dnet.trunk_backward(pk, d_sd, tctx, dnet.embeds_to_nhwc(g_embed), None, True, cots)
    sq = torch.zeros(32, device="cuda")
    K.rowsq(gx, sq)
    K.rowsq(g_sent, sq)
    coef = torch.rand(32, device="cuda") + 0.5
    names = list(d_sd)
    fg = _FlatGrads(names, [d_sd[k] for k in names])
    t_emb = dnet.trunk_tangent(pk, d_sd, tctx, cots, gx, coef, fg.views)
    t_logit = dnet.head_tangent(pk, d_sd, hctx, cots, dnet.embeds_to_nchw(t_emb), g_sent, coef, fg.views)
    torch.cuda.synchronize()
    rows = [("tangent logits vs coef*|g|^2", rel(t_logit, coef * sq)),
            ("bias grads exactly zero", float(fg.views["img_forward.0.bias"].abs().max()))]
    return _report(rows, 1e-2)


@case
def default_init_known_answers():
    """at the reference's default init every gamma is 0, so the generator output equals the
    shortcut-only network, all conv-weight gradients inside residual branches are exactly 0 and gamma.grad != 0."""
    torch, O, g_sd, d_sd, images, sent, noise = _states(batch=2, gated=False, ws=1.0)
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    img = G(noise, sent)
    with torch.no_grad():
        img_ref = O.generator_forward(g_sd, noise, sent)
    loss = -D.get_logits(D.build_embeds(img), sent).mean()
    loss.backward()
    torch.cuda.synchronize()
    rows = [("image vs oracle", rel(img, img_ref))]
    zero_max, gam_min = 0.0, 1e30
    for k, p in G.named_parameters():
        if k.endswith(".gamma"):
            gam_min = min(gam_min, float(p.grad.abs()))
        elif ".conv_1." in k or ".conv_2." in k or "fusion_block" in k:
            zero_max = max(zero_max, float(p.grad.abs().max()))
    print(f"  max |grad| inside residual branches {zero_max:.3e}; min |gamma.grad| {gam_min:.3e}")
    return _report(rows, 1e-2) and zero_max == 0.0 and gam_min > 0.0



if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        bad = []
        for name in CASES:
            print(f"== {name}", flush=True)
            r = subprocess.run(["timeout", "300", sys.executable, __file__, name])
            if r.returncode != 0:
                bad.append(name)
        print("FAILED:" if bad else "ALL OK", bad)
        sys.exit(1 if bad else 0)
    ok = CASES[which]()
    print(which, "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
