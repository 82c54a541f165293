"""Drop-in `RNNEncoder` (text_encoder/model.py:7-63): the frozen caption encoder the trainer calls once per
iteration (deep_fusion_gan/model.py:31-39, 171, 257) and once more inside damsm/loss.py:14.

Same constructor, attributes (`encoder` nn.Embedding, `drop`, `rnn` nn.LSTM -- hence the same state_dict keys,
so `RNNEncoder.load(weights_path, ntoken)` reads the reference's `.pth` files) and `forward(captions, cap_lens)
-> sent_emb [B, nhidden]`.  The arithmetic runs in two CUDA launches (csrc/text_encoder.cu) without the host
synchronisation of the reference's `cap_lens.data.tolist()`.

Differences a maintainer must know:
  * inference only, like every use in the reference (`requires_grad=False` + `.eval()`, deep_fusion_gan/model.py:
    37-39): the output carries no grad_fn, and calling it in training mode with drop_prob > 0 raises instead of
    silently skipping the embedding dropout;
  * the reference's only configuration: one layer, bidirectional, 128 hidden units per direction;
  * `cap_lens` need not be sorted.  pack_padded_sequence insists on decreasing lengths (RuntimeError) because the
    packed layout needs it; these kernels do not.  When `cap_lens` is a host tensor the reference's checks are
    applied anyway (free); a device tensor is not read back, its values are clamped to [0, n_steps] on the device.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

F32 = torch.float32
_SUFFIX = ("", "_reverse")


class RNNEncoder(nn.Module):
    def __init__(self, ntoken, ninput=300, drop_prob=0.5, nhidden=128, nlayers=1, bidirectional=True):
        super().__init__()
        self.n_steps = 18
        self.ntoken = ntoken
        self.ninput = ninput
        self.drop_prob = drop_prob
        self.nlayers = nlayers
        self.bidirectional = bidirectional
        self.num_directions = 2 if bidirectional else 1
        self.nhidden = nhidden // self.num_directions
        self.encoder = nn.Embedding(self.ntoken, self.ninput)
        self.drop = nn.Dropout(self.drop_prob)
        # nn.LSTM warns about dropout with a single layer exactly like the reference's constructor does
        self.rnn = nn.LSTM(self.ninput, self.nhidden, self.nlayers, batch_first=True, dropout=self.drop_prob,
                           bidirectional=self.bidirectional)
        self._ws = None

    def _weights(self):
        return [[getattr(self.rnn, f"{kind}_l0{sfx}") for sfx in _SUFFIX]
                for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]

    @torch.no_grad()
    def forward(self, captions: torch.Tensor, cap_lens: torch.Tensor) -> torch.Tensor:
        if self.nlayers != 1 or not self.bidirectional or self.nhidden != 128:
            raise RuntimeError("gan_codes_b200 RNNEncoder: only the reference's configuration is implemented "
                               "(1 layer, bidirectional, 128 hidden units per direction: RNNEncoder.load)")
        if self.training and self.drop_prob > 0:
            raise RuntimeError("gan_codes_b200 RNNEncoder is an inference encoder (the reference freezes it and "
                               "calls .eval(), deep_fusion_gan/model.py:37-39); call .eval() first")
        emb = self.encoder.weight
        if not emb.is_cuda or not captions.is_cuda:
            raise RuntimeError("gan_codes_b200 RNNEncoder: CUDA tensors only (no CPU path)")
        if captions.dim() != 2 or cap_lens.dim() != 1 or cap_lens.shape[0] != captions.shape[0]:
            raise RuntimeError(f"captions must be [batch, n_steps] and cap_lens [batch]; got {tuple(captions.shape)} "
                               f"and {tuple(cap_lens.shape)}")
        b, t = captions.shape
        if not cap_lens.is_cuda:
            # the checks pack_padded_sequence would make on the host list (text_encoder/model.py:39-40)
            lens = cap_lens.tolist()
            if any(x <= 0 for x in lens):
                raise RuntimeError("Length of all samples has to be greater than 0, but found an element in "
                                   "'lengths' that is <= 0")
            if any(x < y for x, y in zip(lens, lens[1:])):
                raise RuntimeError("`lengths` array must be sorted in decreasing order when `enforce_sorted` is "
                                   "True. You can pass `enforce_sorted=False` to pack_padded_sequence and/or "
                                   "pack_sequence to sidestep this requirement if you do not need ONNX exportability.")
            if lens and lens[0] > t:
                raise RuntimeError(f"Expected sequence length to be larger than or equal to the longest caption "
                                   f"({lens[0]}), got {t}")
        dev = emb.device
        tokens = captions.to(torch.int64).contiguous()
        lens_d = cap_lens.to(dev, torch.int64).contiguous()
        w_ih, w_hh, b_ih, b_hh = self._weights()
        for group in ([emb], w_ih, w_hh, b_ih, b_hh):
            for p in group:
                if p.dtype != F32 or not p.is_contiguous() or p.device != dev:
                    raise RuntimeError("gan_codes_b200 RNNEncoder: parameters must be contiguous fp32 on one CUDA device")
        l = _lib.lib()
        need = l.dfgan_text_encoder_workspace_bytes(b, t)
        if self._ws is None or self._ws.numel() * 4 < need or self._ws.device != dev:
            self._ws = torch.empty((need + 3) // 4, device=dev, dtype=F32)
        sent = torch.empty(b, 2 * self.nhidden, device=dev, dtype=F32)
        arr = lambda g: (C.c_void_p * 2)(g[0].data_ptr(), g[1].data_ptr())  # noqa: E731
        _lib.check(l.dfgan_text_encoder_fwd(tokens.data_ptr(), lens_d.data_ptr(), emb.data_ptr(), emb.shape[0],
                                            arr(w_ih), arr(w_hh), arr(b_ih), arr(b_hh), sent.data_ptr(),
                                            self._ws.data_ptr(), self._ws.numel() * 4, b, t, self.ninput,
                                            self.nhidden, torch.cuda.current_stream().cuda_stream),
                   "dfgan_text_encoder_fwd")
        return sent

    @staticmethod
    def load(weights_path: str, ntoken: int) -> "RNNEncoder":
        text_encoder = RNNEncoder(ntoken, nhidden=256)
        state_dict = torch.load(weights_path, map_location="cpu", weights_only=True)   # tensors only, no pickled code
        text_encoder.load_state_dict(state_dict)
        return text_encoder
