"""GPU bring-up / parity cases of the frozen text encoder (gan_codes_b200.text_encoder.RNNEncoder over
csrc/text_encoder.cu) against oracle/text_encoder_oracle.py and the committed reference outputs
(tests/golden/text_encoder_b8.npz).  Bodies of tests/test_gpu_parity.py::test_text_encoder; also runnable as
    python tools/bringup_text.py [case]
Tolerance: fp32 arithmetic on both sides, 18 recurrent steps -> 1e-5 absolute on outputs bounded by 1.
"""
from __future__ import annotations

import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = {}
TOL = 1e-5


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _encoder(ntoken, sd=None, **kw):
    import torch
    from gan_codes_b200.text_encoder import RNNEncoder
    from oracle import text_encoder_oracle as T
    warnings.filterwarnings("ignore", message="dropout option adds dropout")
    enc = RNNEncoder(ntoken, nhidden=256, **kw)
    if sd is None:
        sd = T.init_state(ntoken) if kw.get("ninput", 300) == 300 else {k: v.detach().clone() for k, v in enc.state_dict().items()}
    enc.load_state_dict(sd)
    for p in enc.parameters():
        p.requires_grad = False
    return enc.cuda().eval(), {k: v.double() for k, v in sd.items()}, torch, T


def _row(name, got, want):
    err = float((got.double().cpu() - want.double().cpu()).abs().max())
    flag = "" if err < TOL else "   <-- FAIL"
    print(f"  {name:60s} {err:.3e}{flag}")
    return err < TOL


@case
def golden():
    """The committed outputs of the reference's own RNNEncoder (fp32 and fp64 runs) on the golden captions."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "text_encoder_b8.npz"))
    enc, sd64, torch, T = _encoder(int(g["ntoken"]))
    norms = np.array([float(v.norm()) for v in sd64.values()])
    assert np.allclose(norms, g["weight_norms"], rtol=1e-12), "init_state drew other weights than the golden run"
    caps, lens = torch.from_numpy(g["captions"]), torch.from_numpy(g["cap_lens"])
    out = enc(caps.cuda(), lens.cuda())
    ok = out.shape == (caps.shape[0], 256) and out.dtype == torch.float32 and not out.requires_grad
    ok &= _row("vs reference RNNEncoder, fp64 run (golden)", out, torch.from_numpy(g["sent_f64"]))
    ok &= _row("vs reference RNNEncoder, fp32 run (golden)", out, torch.from_numpy(g["sent_f32"]))
    ok &= _row("host cap_lens (the reference's checks applied) vs device", enc(caps.cuda(), lens), out)
    return ok


@case
def shapes():
    """Batch / length / width coverage against the fp64 oracle: the benchmark batch, unsorted device lengths, a
    batch that is not a multiple of the projection tile, one sentence, short padded width, another embedding width."""
    ok = True
    for (b, t, ninput, shuffle) in ((32, 18, 300, False), (32, 18, 300, True), (33, 18, 300, True), (1, 18, 300, False),
                                    (5, 7, 300, False), (64, 18, 300, True), (6, 18, 70, False)):
        enc, sd64, torch, T = _encoder(97, ninput=ninput) if ninput != 300 else _encoder(97)
        caps, lens = T.synthetic_captions(b, 97, seed=b + t, n_steps=t)
        if shuffle:
            perm = torch.randperm(b, generator=torch.Generator().manual_seed(b))
            caps, lens = caps[perm], lens[perm]
        want = T.sentence_embeds(sd64, caps, lens)
        ok &= _row(f"B={b} T={t} ninput={ninput} {'unsorted' if shuffle else 'sorted'} lens {lens[:6].tolist()}..",
                   enc(caps.cuda(), lens.cuda()), want)
    return ok


@case
def stock_torch():
    """Against the stock torch path the reference executes (nn.Embedding -> pack_padded_sequence -> nn.LSTM), run
    on the box's CPU in fp32 with the very same module parameters -- no oracle involved."""
    from torch.nn.utils.rnn import pack_padded_sequence
    enc, _, torch, T = _encoder(211)
    caps, lens = T.synthetic_captions(32, 211, seed=3)
    out = enc(caps.cuda(), lens.cuda())
    cpu = type(enc)(211, nhidden=256)
    cpu.load_state_dict(enc.state_dict())
    cpu.eval()
    with torch.no_grad():
        packed = pack_padded_sequence(cpu.encoder(caps), lens.tolist(), batch_first=True)
        hidden = cpu.rnn(packed)[1][0]
        want = hidden.transpose(0, 1).contiguous().view(-1, 256)
    return _row("vs nn.Embedding + pack_padded_sequence + nn.LSTM (CPU fp32)", out, want)


@case
def semantics():
    """Edge and error behaviour: zero / over-long device lengths are clamped, padding tokens never influence the
    result, out-of-range ids read as zero rows, host-side length checks and the training-mode guard raise."""
    enc, sd64, torch, T = _encoder(50)
    caps, lens = T.synthetic_captions(8, 50)
    base = enc(caps.cuda(), lens.cuda())
    ok = True
    noisy = caps.clone()
    noisy[torch.arange(18)[None, :] >= lens[:, None]] = 49                  # garbage in the padding
    ok &= _row("padding positions do not influence the embedding", enc(noisy.cuda(), lens.cuda()), base)
    l2 = lens.clone()
    l2[0], l2[-1] = 40, 0                                                   # clamped to 18 / empty sentence
    out = enc(caps.cuda(), l2.cuda())
    ok &= _row("length > n_steps clamps to n_steps", out[0], base[0])
    ok &= float(out[-1].abs().max()) == 0.0
    bad = caps.clone()
    bad[2, 1] = 50_000
    sd0 = {k: v.clone() for k, v in sd64.items()}
    ref_caps = caps.clone()
    ref_caps[2, 1] = 0
    sd0["encoder.weight"][0] = 0
    enc0, _, _, _ = _encoder(50, sd={k: v.float() for k, v in sd0.items()})
    ok &= _row("out-of-range token id == zero embedding row", enc(bad.cuda(), lens.cuda())[2],
               enc0(ref_caps.cuda(), lens.cuda())[2])

    def raises(fn, text):
        try:
            fn()
        except RuntimeError as e:
            return text in str(e)
        return False
    ok &= raises(lambda: enc(caps.cuda(), lens.flip(0)), "must be sorted in decreasing order")
    ok &= raises(lambda: enc(caps.cuda(), torch.zeros(8, dtype=torch.int64)), "has to be greater than 0")
    ok &= raises(lambda: enc(caps, lens), "no CPU path")
    enc.train()
    ok &= raises(lambda: enc(caps.cuda(), lens.cuda()), "call .eval() first")
    enc.eval()
    print("  error behaviour:", "ok" if ok else "MISMATCH")
    return ok


@case
def timing():
    """Informational: device time per call at the benchmark batch (CUDA events, 50 calls) next to the stock torch
    CUDA path of the reference (cuDNN LSTM + the cap_lens.tolist() synchronisation)."""
    from torch.nn.utils.rnn import pack_padded_sequence
    enc, _, torch, T = _encoder(5450)
    caps, lens = (x.cuda() for x in T.synthetic_captions(32, 5450, seed=5))

    def ours():
        return enc(caps, lens)

    def stock():
        with torch.no_grad():
            packed = pack_padded_sequence(enc.encoder(caps), lens.data.tolist(), batch_first=True)
            return enc.rnn(packed)[1][0].transpose(0, 1).contiguous().view(-1, 256)
    torch.backends.cudnn.allow_tf32 = False
    want = T.sentence_embeds({k: v.double().cpu() for k, v in enc.state_dict().items()}, caps.cpu(), lens.cpu())
    ok = _row("ours vs fp64 oracle", ours(), want)
    # informational only: cuDNN's RNN kernels choose their own math mode (TF32 by default: ~6e-4 from fp32)
    print("  stock torch CUDA path (cuDNN) vs fp64 oracle: %.3e" % float((stock().double().cpu() - want).abs().max()))
    for name, fn in (("ours (2 launches, no sync)", ours), ("stock torch CUDA", stock)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"  {name:40s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us / call")
    return ok


if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    bad = []
    for name in which:
        print("==", name, flush=True)
        if not CASES[name]():
            bad.append(name)
    print("FAILED: %s" % bad if bad else "ALL OK")
    sys.exit(1 if bad else 0)
