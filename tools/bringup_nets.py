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


@case
def dfwd():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import dnet
    acts = {}
    with torch.no_grad():
        emb_ref = O.discriminator_embeds(d_sd, images, tap=lambda k, v: acts.__setitem__(k, v))
        log_ref = O.discriminator_logits(d_sd, emb_ref, sent)
    pk = dnet.DPacked(d_sd)
    ctx = {}
    emb = dnet.trunk_forward(pk, d_sd, images, ctx)
    logits = dnet.head_forward(pk, d_sd, dnet.embeds_to_nchw(emb), sent)
    torch.cuda.synchronize()
    rows = []
    for k, (x, c1, r) in enumerate(ctx["saved"]):
        p = f"img_forward.{k + 1}"
        if k == 0:
            rows.append(("img_forward.0", rel(nchw(x), acts["img_forward.0"])))
        rows.append((p + ".residual_conv.0 (post-lrelu)", rel(nchw(c1), torch.nn.functional.leaky_relu(acts[p + ".residual_conv.0"], 0.2))))
        rows.append((p + ".residual_conv.2 (post-lrelu)", rel(nchw(r), torch.nn.functional.leaky_relu(acts[p + ".residual_conv.2"], 0.2))))
    rows.append(("embeds", rel(nchw(emb), emb_ref)))
    rows.append(("logits", rel(logits, log_ref.reshape(-1))))
    return _report(rows, 1e-2)


def _ref_modules(torch, O, g_sd, d_sd):
    from gan_codes_b200.modules import Discriminator, Generator
    G, D = Generator(32, 100).cuda(), Discriminator(32).cuda()
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    return G, D


@case
def dbwd():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    fake = (torch.randn_like(images) * 0.5).clamp(-1, 1)
    dl = {k: v.clone().requires_grad_() for k, v in d_sd.items()}
    loss_ref = O.d_hinge_loss(dl, images, fake, sent)
    gref = torch.autograd.grad(loss_ref, list(dl.values()))
    b = images.shape[0]
    real_emb = D.build_embeds(images)
    l_real = torch.relu(1.0 - D.get_logits(real_emb, sent)).mean()
    l_mis = torch.relu(1.0 + D.get_logits(real_emb[:b - 1], sent[1:b])).mean()
    l_fake = torch.relu(1.0 + D.get_logits(D.build_embeds(fake), sent)).mean()
    loss = l_real + (l_fake + l_mis) / 2.0
    loss.backward()
    torch.cuda.synchronize()
    rows = [("d_loss", abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()))]
    for (k, p), g in zip(D.named_parameters(), gref):
        rows.append(("grad " + k, rel(p.grad, g)))
    return _report(rows, 2e-2)


@case
def gp():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    dl = {k: v.clone().requires_grad_() for k, v in d_sd.items()}
    probe = {}
    loss_ref = O.magp_loss(dl, images, sent, probe)
    gref = torch.autograd.grad(loss_ref, list(dl.values()), allow_unused=True)
    x = images.clone().requires_grad_()
    s = sent.clone().requires_grad_()
    logits = D.get_logits(D.build_embeds(x), s)
    gx, gs = torch.autograd.grad(logits, (x, s), torch.ones_like(logits), retain_graph=True, create_graph=True)
    g = torch.cat((gx.reshape(x.shape[0], -1), gs.reshape(x.shape[0], -1)), dim=1)
    n = torch.clamp(torch.sqrt((g ** 2).sum(dim=1) + 1e-8), 0, 10)
    loss = 2.0 * torch.mean(n ** 6)
    loss.backward()
    torch.cuda.synchronize()
    print("  norms ref", probe["gp_n"].tolist(), "ours", n.tolist())
    rows = [("gp_gx", rel(gx, probe["gp_gx"])), ("gp_gs", rel(gs, probe["gp_gs"])),
            ("gp_loss", abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()))]
    for (k, p), gr in zip(D.named_parameters(), gref):
        if gr is None:
            rows.append(("gp grad (zero) " + k, float(p.grad.abs().max()) if p.grad is not None else 1.0))
        else:
            rows.append(("gp grad " + k, rel(p.grad, gr)))
    return _report(rows, 3e-2)


@case
def gfwd():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    from gan_codes_b200 import gnet
    acts = {}
    with torch.no_grad():
        img_ref = O.generator_forward(g_sd, noise, sent, tap=lambda k, v: acts.__setitem__(k, v))
    pk = gnet.GPacked(g_sd)
    mlp = gnet.GMlp(g_sd, pk.plan, noise.shape[0], None)
    ctx = {}
    img = gnet.generator_forward(pk, g_sd, mlp, noise, sent, ctx)
    torch.cuda.synchronize()
    rows = []
    # gamma/beta tables against the oracle MLPs
    for (p, fb, m), (o, wdt) in list(mlp.offsets.items())[:8]:
        rows.append((f"mlp {p}.fusion_block_{fb}.{m}", rel(mlp.tab[:, o:o + wdt], O._mlp(g_sd, f"{p}.fusion_block_{fb}.{m}", sent))))
    for i, (z_prev, a12, h1, a34) in enumerate(ctx["saved"]):
        p = pk.plan[i][0]
        rows.append((p + ".conv_1", rel(nchw(h1), acts[p + ".conv_1"])))
        if i + 1 < len(ctx["saved"]):
            rows.append((p + " (block out)", rel(nchw(ctx["saved"][i + 1][0]), acts[p])))
    rows.append(("res_block_out (block out)", rel(nchw(ctx["z_last"]), acts["res_block_out"])))
    rows.append(("image", rel(img, img_ref)))
    return _report(rows, 1e-2)


@case
def gbwd():
    torch, O, g_sd, d_sd, images, sent, noise = _states()
    G, D = _ref_modules(torch, O, g_sd, d_sd)
    gl = {k: v.clone().requires_grad_() for k, v in g_sd.items()}
    loss_ref = O.g_hinge_loss(d_sd, O.generator_forward(gl, noise, sent), sent)
    gref = torch.autograd.grad(loss_ref, list(gl.values()))
    loss = -D.get_logits(D.build_embeds(G(noise, sent)), sent).mean()
    loss.backward()
    torch.cuda.synchronize()
    rows = [("g_loss", abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()))]
    for (k, p), g in zip(G.named_parameters(), gref):
        rows.append(("grad " + k, rel(p.grad, g)))
    worst = sorted(rows, key=lambda r: -r[1])[:25]
    ok = _report(worst, 3e-2)
    print(f"  ({len(rows)} tensors, worst 25 shown)")
    return ok


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
