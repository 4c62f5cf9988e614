"""CPU tests of the Hungarian-matcher oracle: oracle/lsap_oracle.c against scipy itself,
oracle/matcher_port.py against golden vectors the reference's own HungarianMatcher produced
(tests/golden/matcher/make_matcher_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import matcher_port

scipy_opt = pytest.importorskip("scipy.optimize")
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "matcher", "*.npz")))


def _random_costs():
    rng = np.random.default_rng(7)
    for nr, nc in [(1, 1), (1, 9), (9, 1), (5, 5), (12, 900), (900, 12), (40, 16), (16, 40), (97, 300), (64, 64)]:
        yield f"gauss_{nr}x{nc}", rng.standard_normal((nr, nc))
        yield f"ties_{nr}x{nc}", rng.integers(0, 3, (nr, nc)).astype(np.float64)       # small integers: ties everywhere
    yield "constant", np.full((7, 11), 2.5)
    yield "zeros_tall", np.zeros((11, 7))
    c = rng.standard_normal((8, 20))
    c[rng.random((8, 20)) < 0.3] = np.inf
    yield "some_inf", c
    yield "f32_values", rng.standard_normal((30, 200)).astype(np.float32)


@pytest.mark.parametrize("name,cost", list(_random_costs()), ids=lambda v: v if isinstance(v, str) else "")
def test_lsap_restatement_equals_scipy(name, cost):
    want_i, want_j = scipy_opt.linear_sum_assignment(cost)
    got_i, got_j = matcher_port.assign(cost)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_j, want_j), name


def test_lsap_error_classes_follow_scipy():
    bad = np.zeros((3, 3))
    bad[1, 1] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        matcher_port.assign(bad)
    with pytest.raises(ValueError):
        scipy_opt.linear_sum_assignment(bad)
    inf = np.full((2, 2), np.inf)
    with pytest.raises(ValueError, match="infeasible"):
        matcher_port.assign(inf)
    with pytest.raises(ValueError, match="infeasible"):
        scipy_opt.linear_sum_assignment(inf)
    i, j = matcher_port.assign(np.zeros((0, 4)))
    assert len(i) == 0 and len(j) == 0


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_matcher_port_matches_reference(path):
    g = np.load(path)
    w = g["weights"]
    c = matcher_port.cost_matrix(torch.from_numpy(g["pred_logits"]), torch.from_numpy(g["pred_boxes"]),
                                 torch.from_numpy(g["tgt_boxes"]), torch.from_numpy(g["label_map"][g["labels"]]),
                                 cost_class=float(w[0]), cost_bbox=float(w[1]), cost_giou=float(w[2]),
                                 focal_alpha=float(g["focal_alpha"]))
    # same torch ops in the same order (bit-identical on the authoring host; gemv blocking may differ by CPU)
    assert np.allclose(c.numpy(), g["cost"], rtol=0, atol=2e-6 * np.abs(g["cost"]).max())
    if "ties" not in path:      # exact ties make the pairing depend on the last bit of the matrix
        i, j = matcher_port.assign(c.numpy())
        assert np.array_equal(i, g["index_i"]) and np.array_equal(j, g["index_j"])
    # and on the reference's recorded matrix
    i2, j2 = matcher_port.assign(g["cost"])
    assert np.array_equal(i2, g["index_i"]) and np.array_equal(j2, g["index_j"])


def test_golden_set_is_present():
    assert len(GOLDEN) >= 6
