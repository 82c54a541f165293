"""TEST INFRASTRUCTURE -- not a product path.  Only tests/, tools/bringup_*.py (the bodies of the GPU tests) and
__graft_entry__.smoke() may import this module.

Restatement of the frozen caption encoder, RNNEncoder.forward of Deep_Fusion_GAN/src/text_encoder/model.py:33-56
in eval mode, as explicit per-sentence, per-time-step loops over a state_dict (any float dtype):

    emb      = encoder.weight[captions]                                            (:36, dropout is the identity)
    packed   : sentence b contributes time steps 0 .. cap_lens[b]-1 only          (:39-40)
    LSTM     : gates = W_ih x_t + b_ih + W_hh h + b_hh, split i | f | g | o,
               c' = sigmoid(f) c + sigmoid(i) tanh(g),  h' = sigmoid(o) tanh(c')  (:46, torch.nn.LSTM)
               forward direction t = 0 .. len-1, reverse direction t = len-1 .. 0, zero initial state
    sent_emb = [h_forward_final | h_reverse_final]                                 (:54-56)

Parity pin: the reference has no tests for this path; oracle/make_text_golden.py runs the reference's UNMODIFIED
RNNEncoder (imported from /root/reference in the build container) on `init_state` weights and `synthetic_captions`
inputs and commits its outputs as tests/golden/text_encoder_b8.npz; tests/test_text_encoder_cpu.py checks this
module against that file on every box.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

State = Dict[str, torch.Tensor]
N_STEPS = 18       # text_encoder/model.py:11
NINPUT = 300       # :8
NHIDDEN = 256      # RNNEncoder.load (:60): 128 per direction


def init_state(ntoken: int, seed: int = 7, dtype=torch.float32) -> State:
    """A state_dict with RNNEncoder's keys and shapes (encoder.weight ~ N(0,1) like nn.Embedding; LSTM tensors
    ~ U(-k, k), k = 1/sqrt(hidden), like nn.LSTM), drawn from a seeded CPU generator so every box gets the same
    numbers."""
    g = torch.Generator().manual_seed(seed)
    h = NHIDDEN // 2
    k = 1.0 / h ** 0.5
    sd: State = OrderedDict()
    sd["encoder.weight"] = torch.randn(ntoken, NINPUT, generator=g, dtype=torch.float64).to(dtype)
    for sfx in ("", "_reverse"):
        for name, shape in (("weight_ih", (4 * h, NINPUT)), ("weight_hh", (4 * h, h)), ("bias_ih", (4 * h,)),
                            ("bias_hh", (4 * h,))):
            sd[f"rnn.{name}_l0{sfx}"] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * k).to(dtype)
    return sd


def synthetic_captions(batch: int, ntoken: int, seed: int = 11, n_steps: int = N_STEPS) -> Tuple[torch.Tensor, torch.Tensor]:
    """(captions [B, n_steps] int64 zero-padded, cap_lens [B] int64 sorted decreasing, as objects/utils.py:10-16
    delivers them); lengths cover 1 .. n_steps including both extremes when the batch allows."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, n_steps + 1, (batch,), generator=g)
    if batch >= 2:
        lens[0], lens[-1] = n_steps, 1
    lens, _ = torch.sort(lens, 0, True)
    caps = torch.randint(1, ntoken, (batch, n_steps), generator=g)
    caps = caps * (torch.arange(n_steps)[None, :] < lens[:, None])
    return caps.to(torch.int64), lens.to(torch.int64)


def _direction(sd: State, sfx: str, x: torch.Tensor, reverse: bool) -> torch.Tensor:
    """x [len, ninput] -> final hidden state [h] of one direction of the layer."""
    w_ih, w_hh = sd[f"rnn.weight_ih_l0{sfx}"], sd[f"rnn.weight_hh_l0{sfx}"]
    b_ih, b_hh = sd[f"rnn.bias_ih_l0{sfx}"], sd[f"rnn.bias_hh_l0{sfx}"]
    hid = w_hh.shape[1]
    h = torch.zeros(hid, dtype=x.dtype, device=x.device)
    c = torch.zeros_like(h)
    order = range(x.shape[0] - 1, -1, -1) if reverse else range(x.shape[0])
    for t in order:
        gates = w_ih @ x[t] + b_ih + w_hh @ h + b_hh
        i, f, g, o = gates.split(hid)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


def sentence_embeds(sd: State, captions: torch.Tensor, cap_lens: torch.Tensor) -> torch.Tensor:
    """RNNEncoder.forward (text_encoder/model.py:33-56), eval mode.  Returns [B, 2 * hidden]."""
    out = []
    for b in range(captions.shape[0]):
        n = int(cap_lens[b])
        x = sd["encoder.weight"][captions[b, :n]]
        out.append(torch.cat((_direction(sd, "", x, False), _direction(sd, "_reverse", x, True))))
    return torch.stack(out)
