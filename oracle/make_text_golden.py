"""TEST INFRASTRUCTURE.  Generates tests/golden/text_encoder_b8.npz by running the UNMODIFIED reference
RNNEncoder (Deep_Fusion_GAN/src/text_encoder/model.py, imported read-only from /root/reference) on the weights of
oracle/text_encoder_oracle.py:init_state and the captions of synthetic_captions.  Build container only:

    python oracle/make_text_golden.py

The file holds the inputs (captions, lengths), a digest of the weights (so a box whose torch RNG produced other
numbers is noticed) and the reference's sentence embeddings in fp32 and fp64.  It also checks, while the reference
is at hand, that the RNNEncoder.load round trip through a .pth file works for the drop-in class's state_dict keys.
"""
from __future__ import annotations

import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference/Deep_Fusion_GAN")
sys.path.insert(0, ROOT)

from oracle import text_encoder_oracle as T  # noqa: E402

NTOKEN, BATCH = 50, 8


def main() -> int:
    warnings.filterwarnings("ignore", message="dropout option adds dropout")
    from src.text_encoder.model import RNNEncoder as Ref
    from gan_codes_b200.text_encoder import RNNEncoder as Mine
    sd = T.init_state(NTOKEN)
    caps, lens = T.synthetic_captions(BATCH, NTOKEN)
    out = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "text_encoder.pth")
            torch.save({k: v.to(dt) for k, v in sd.items()}, path)
            ref = Ref.load(path, NTOKEN) if dt == torch.float32 else Ref(NTOKEN, nhidden=256).double()
            if dt == torch.float64:
                ref.load_state_dict(torch.load(path))
            if dt == torch.float32:
                mine = Mine.load(path, NTOKEN)            # same keys, same shapes: strict load succeeds
                assert list(mine.state_dict()) == list(ref.state_dict())
        ref.eval()
        with torch.no_grad():
            out[tag] = ref(caps, lens)
        mine_oracle = T.sentence_embeds({k: v.to(dt) for k, v in sd.items()}, caps, lens)
        err = float((mine_oracle - out[tag]).abs().max())
        print(f"{tag}: reference vs oracle max |diff| = {err:.3e}")
        assert err < (2e-6 if dt == torch.float32 else 1e-13)
    digest = np.array([float(v.double().norm()) for v in sd.values()])
    path = os.path.join(ROOT, "tests", "golden", "text_encoder_b8.npz")
    np.savez_compressed(path, captions=caps.numpy(), cap_lens=lens.numpy(), weight_norms=digest,
                        sent_f32=out["f32"].numpy(), sent_f64=out["f64"].numpy(), ntoken=np.array(NTOKEN))
    print("wrote", path, os.path.getsize(path), "bytes; lens", lens.tolist())
    return 0


if __name__ == "__main__":
    sys.exit(main())
