"""gan_codes_b200/sample.py against the reference's own sample.sample / save_image / preprocess_caption
(tests/golden/sample_golden.json, oracle/make_sample_golden.py): file names, PNG pixels, noise draw, prompt encoding."""
import json
import os

import pytest
import torch

from gan_codes_b200 import sample as S
from oracle import make_sample_golden as G


def test_sample_and_prompt_encoding_match_the_reference():
    with open(os.path.join(os.path.dirname(__file__), "golden", "sample_golden.json")) as f:
        gold = json.load(f)
    assert G.run(S, S.preprocess_caption) == gold
    assert sorted(gold["files"]) == ["001.A_bird_1.png", "002.B_bird_2.png", "plain.png"]


def test_encode_prompt_squeezes_and_rejects_empty_prompts():
    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, captions, cap_lens):
            assert captions.shape == (1, 18) and cap_lens.tolist() == [7]
            return captions.float()
    out = S.encode_prompt(Enc(), "A small bird, with a red head.", G.WORD2CODE)
    assert out.shape == (1, 18) and out[0, :7].tolist() == [2, 3, 4, 5, 2, 6, 7]
    with pytest.raises(ValueError, match="no words"):
        S.encode_prompt(Enc(), " ., ", G.WORD2CODE)
