"""Frozen text encoder (text_encoder/model.py:7-63), CPU side: the oracle restatement against the
committed outputs of the reference's own RNNEncoder (tests/golden/text_encoder_b8.npz, made by
oracle/make_text_golden.py), the drop-in class's API surface / state_dict layout, the argument validation of the
C entry point (no compute without a GPU) and the refusal to run on the CPU."""
import ctypes as C
import os
import warnings

import numpy as np
import pytest
import torch

from gan_codes_b200 import _lib
from gan_codes_b200.text_encoder import RNNEncoder
from oracle import text_encoder_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "text_encoder_b8.npz"))
warnings.filterwarnings("ignore", message="dropout option adds dropout")


def test_oracle_reproduces_the_reference_outputs():
    ntoken = int(G["ntoken"])
    sd = T.init_state(ntoken)
    assert np.allclose([float(v.double().norm()) for v in sd.values()], G["weight_norms"], rtol=1e-12)
    caps, lens = T.synthetic_captions(8, ntoken)
    assert np.array_equal(caps.numpy(), G["captions"]) and np.array_equal(lens.numpy(), G["cap_lens"])
    assert lens.tolist() == sorted(lens.tolist(), reverse=True) and lens.max() == T.N_STEPS and lens.min() == 1
    out64 = T.sentence_embeds({k: v.double() for k, v in sd.items()}, caps, lens)
    assert float((out64 - torch.from_numpy(G["sent_f64"])).abs().max()) < 1e-13
    out32 = T.sentence_embeds(sd, caps, lens)
    assert float((out32 - torch.from_numpy(G["sent_f32"])).abs().max()) < 2e-6
    assert out32.shape == (8, 256)


def test_oracle_matches_stock_torch_lstm_on_unseen_inputs():
    """Independent of the golden file: nn.Embedding -> pack_padded_sequence -> nn.LSTM in fp64."""
    from torch.nn.utils.rnn import pack_padded_sequence
    enc = RNNEncoder(31, nhidden=256).double().eval()
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    caps, lens = T.synthetic_captions(5, 31, seed=99, n_steps=9)
    with torch.no_grad():
        hidden = enc.rnn(pack_padded_sequence(enc.encoder(caps), lens.tolist(), batch_first=True))[1][0]
    want = hidden.transpose(0, 1).contiguous().view(-1, 256)
    assert float((T.sentence_embeds(sd, caps, lens) - want).abs().max()) < 1e-13


def test_api_surface_and_state_dict_layout(tmp_path):
    enc = RNNEncoder(50, nhidden=256)
    assert (enc.n_steps, enc.ntoken, enc.ninput, enc.drop_prob, enc.nlayers, enc.bidirectional,
            enc.num_directions, enc.nhidden) == (18, 50, 300, 0.5, 1, True, 2, 128)
    assert isinstance(enc.encoder, torch.nn.Embedding) and isinstance(enc.rnn, torch.nn.LSTM)
    want = T.init_state(50)
    assert list(enc.state_dict()) == list(want)
    assert all(enc.state_dict()[k].shape == v.shape for k, v in want.items())
    path = str(tmp_path / "text_encoder.pth")
    torch.save(want, path)
    loaded = RNNEncoder.load(path, 50)
    assert all(torch.equal(loaded.state_dict()[k], v) for k, v in want.items())
    assert RNNEncoder(50).nhidden == 64           # default nhidden=128 over two directions, as in the reference


def test_no_cpu_path_and_guards():
    enc = RNNEncoder(50, nhidden=256).eval()
    caps, lens = T.synthetic_captions(4, 50)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(caps, lens)
    with pytest.raises(RuntimeError, match="call .eval"):
        enc.train()(caps, lens)
    with pytest.raises(RuntimeError, match="only the reference's configuration"):
        RNNEncoder(50, nhidden=128).eval()(caps, lens)


def test_c_entry_point_validates_arguments():
    l = _lib.lib()
    assert l.dfgan_text_encoder_workspace_bytes(32, 18) == 2 * 32 * 18 * 512 * 4
    assert l.dfgan_text_encoder_workspace_bytes(0, 18) == 0
    null2, some2 = (C.c_void_p * 2)(0, 0), (C.c_void_p * 2)(256, 256)
    call = l.dfgan_text_encoder_fwd
    assert call(None, 256, 256, 50, some2, some2, some2, some2, 256, 256, 1 << 30, 4, 18, 300, 128, None) == -1
    assert call(256, 256, 256, 50, null2, some2, some2, some2, 256, 256, 1 << 30, 4, 18, 300, 128, None) == -1
    assert call(256, 256, 256, 50, some2, some2, some2, some2, 256, 256, 1 << 30, 4, 18, 300, 64, None) == -2
    assert call(256, 256, 256, 50, some2, some2, some2, some2, 256, 256, 16, 4, 18, 300, 128, None) == -1   # workspace
    assert call(256, 256, 256, 50, some2, (C.c_void_p * 2)(260, 256), some2, some2, 256, 256, 1 << 30, 4, 18, 300, 128,
                None) == -1                                                                            # alignment
