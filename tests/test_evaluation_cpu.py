"""IS / FID: gan_codes_b200/evaluation.py and oracle/metrics_oracle.py against the results of the
reference's own compute_inception_score / compute_fid (tests/golden/metrics_golden.npz, oracle/make_metrics_golden.py).
These functions are torch library calls that, like the reference's, run on whatever device holds the images, so the
whole file runs on CPU."""
import os

import numpy as np
import pytest
import torch

from gan_codes_b200 import evaluation as E
from oracle import metrics_oracle as M

GOLD = os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    w = {k: z[k] for k in ("conv_w", "conv_b", "fc_w", "fc_b")}
    E.set_inception_factory(lambda: M.StandInInception(w))
    yield {"z": z, "real": M.to_images(z["real"]), "fake": M.to_images(z["fake"]), "w": w}
    E.set_inception_factory(None)


def rel(a, b):
    return abs(a - b) / abs(b)


def test_golden_inputs_are_the_committed_generator(gold):
    assert np.array_equal(gold["z"]["real"], M.synthetic_images(96, 24, seed=1))
    assert np.array_equal(gold["z"]["fake"], M.synthetic_images(96, 24, seed=2, shift=0.15))
    for k, v in M.stand_in_weights().items():
        assert np.array_equal(gold["w"][k], v)


def test_inception_score_matches_reference(gold):
    z = gold["z"]
    assert rel(E.compute_inception_score(gold["fake"], cuda=False, batch_size=8, splits=1), float(z["is_fake_1"])) < 1e-5
    assert rel(E.compute_inception_score(gold["fake"], cuda=False, batch_size=8, splits=3), float(z["is_fake_3"])) < 1e-5
    assert rel(E.compute_inception_score(gold["real"], cuda=False, batch_size=32, splits=1), float(z["is_real_1"])) < 1e-5


def test_fid_matches_reference(gold):
    z = gold["z"]
    assert rel(E.compute_fid(gold["real"], gold["fake"], cuda=False, batch_size=8), float(z["fid_real_fake"])) < 1e-5
    assert abs(E.compute_fid(gold["real"], gold["real"], cuda=False, batch_size=8)) < 1e-9
    assert abs(float(z["fid_real_real"])) < 1e-9


def test_nan_inf_inputs_follow_the_reference(gold, capsys):
    z = gold["z"]
    bad = gold["fake"].clone()
    bad[0, 0, 0, 0], bad[1, 1, 2, 3], bad[2, 2, 5, 5] = float("nan"), float("inf"), float("-inf")
    assert rel(E.compute_inception_score(bad, cuda=False), float(z["is_bad_1"])) < 1e-5
    assert rel(E.compute_fid(gold["real"], bad, cuda=False), float(z["fid_real_bad"])) < 1e-5
    out = capsys.readouterr().out
    assert "Warning: Input images contain NaN or Inf values for IS calculation" in out
    assert "Warning: Input images contain NaN or Inf values\n" in out


def test_single_pass_is_fid_equals_the_two_functions(gold):
    i, f = E.compute_is_fid(gold["fake"], gold["real"], cuda=False, batch_size=64)
    assert rel(i, float(gold["z"]["is_fake_1"])) < 1e-5
    assert rel(f, float(gold["z"]["fid_real_fake"])) < 1e-5


def test_one_resident_network_per_device(gold):
    a = E.inception(torch.device("cpu"))
    E.compute_inception_score(gold["fake"][:8], cuda=False)
    E.compute_fid(gold["real"][:32], gold["fake"][:32], cuda=False)
    assert E.inception(torch.device("cpu")) is a


def test_statistics_against_the_numpy_scipy_restatement():
    r = np.random.RandomState(3)
    logits = r.standard_normal((70, 33)) * 2
    probs = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    for splits in (1, 2, 5):
        assert rel(E.inception_score_from_probs(torch.from_numpy(probs.astype(np.float32)), splits),
                   M.inception_score(probs.astype(np.float32), splits)) < 1e-9
    mix1, mix2 = r.standard_normal((16, 16)), r.standard_normal((16, 16))
    a1 = (r.standard_normal((200, 16)) @ mix1 + 0.3).astype(np.float32)
    a2 = (r.standard_normal((150, 16)) @ mix2 - 0.2).astype(np.float32)
    assert rel(E.fid_from_activations(torch.from_numpy(a1), torch.from_numpy(a2)), M.fid(a1, a2)) < 1e-6


def test_rank_deficient_covariances_stay_finite_and_consistent():
    # fewer samples than feature dimensions, the trainer's regime (768 images, 2048 features): S1 S2 is singular.
    # The symmetric form is still exact: identical sets give 0, and the value is symmetric in its arguments.
    r = np.random.RandomState(4)
    a1 = torch.from_numpy(r.standard_normal((20, 64)).astype(np.float32))
    a2 = torch.from_numpy((r.standard_normal((24, 64)) * 1.3 + 0.1).astype(np.float32))
    assert abs(E.fid_from_activations(a1, a1)) < 1e-5       # sqrt of the eps-sized eigenvalue noise of the null space
    f12, f21 = E.fid_from_activations(a1, a2), E.fid_from_activations(a2, a1)
    assert np.isfinite(f12) and f12 > 0 and rel(f12, f21) < 1e-6


def test_failure_values(gold, capsys):
    assert E.fid_from_activations(torch.tensor([[float("nan"), 1.0], [0.0, 1.0]]), torch.zeros(2, 2)) == float("inf")
    E.set_inception_factory(lambda: (_ for _ in ()).throw(RuntimeError("no weights offline")))
    try:
        assert E.compute_inception_score(gold["fake"][:8], cuda=False) == 1.0          # evaluation/metrics.py:59-61
        assert E.compute_fid(gold["real"][:8], gold["fake"][:8], cuda=False) == float("inf")   # :113-115
        assert E.compute_is_fid(gold["fake"][:8], gold["real"][:8], cuda=False) == (1.0, float("inf"))
    finally:
        E.set_inception_factory(lambda: M.StandInInception(gold["w"]))
    out = capsys.readouterr().out
    assert "Error in IS calculation: no weights offline" in out and "Exception in FID calculation" in out


def test_real_inception_architecture_pooled_features_and_logits():
    """torchvision's Inception-v3 layer graph with random weights (the pretrained file cannot be downloaded): the
    resident wrapper's pooled features equal what the reference's `model.fc = Identity()` network returns (:70-72),
    and its logits equal the unmodified network's (:12, 31) -- one pass instead of two."""
    import copy
    from torchvision.models import inception_v3
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(0)
        net = inception_v3(weights=None, aux_logits=True, transform_input=False, init_weights=False).eval()
    ref_logits_net = copy.deepcopy(net)
    ref_feat_net = copy.deepcopy(net)
    ref_feat_net.fc = torch.nn.Identity()
    ref_feat_net.aux_logits = False
    E.set_inception_factory(lambda: net)
    try:
        x = torch.rand(3, 3, 299, 299, generator=torch.Generator().manual_seed(1))
        logits, pooled = E.inception(torch.device("cpu"))(x)
        with torch.no_grad():
            assert pooled.shape == (3, 2048) and torch.equal(pooled, ref_feat_net(x))
            assert logits.shape == (3, 1000) and torch.equal(logits, ref_logits_net(x))
    finally:
        E.set_inception_factory(None)
