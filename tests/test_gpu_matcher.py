"""GPU parity tests of the on-device Hungarian matcher (csrc/hm_match.cu through the C ABI
of include/hungarian_b200.h).

Bars: assignment indices BIT-EXACT against scipy (when importable) and against oracle/lsap_oracle.c on the same cost
matrix -- including tie-heavy matrices, where the answer depends on scipy's tie-breaking; cost matrix within 1e-5
normwise (max|a-b| / max|b|) of the matrix the reference's own HungarianMatcher built (golden, fp32 on a CPU)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import matcher_port

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "matcher", "*.npz")))
COST_TOL = 1e-5

try:
    from scipy.optimize import linear_sum_assignment as scipy_lsa
except Exception:  # pragma: no cover
    scipy_lsa = None


def _dev():
    return torch.device("cuda:0")


def _check_pairs(pairs, cost_np, what):
    i, j = pairs[0].cpu().numpy(), pairs[1].cpu().numpy()
    oi, oj = matcher_port.assign(cost_np)
    assert np.array_equal(i, oi) and np.array_equal(j, oj), f"{what}: differs from the oracle"
    if scipy_lsa is not None:
        si, sj = scipy_lsa(cost_np)
        assert np.array_equal(i, si) and np.array_equal(j, sj), f"{what}: differs from scipy"


def _cost_cases():
    rng = np.random.default_rng(11)
    for Q, T in [(900, 12), (900, 1), (900, 97), (900, 300), (64, 64), (16, 40), (5, 900), (1, 1), (300, 299), (33, 1)]:
        yield f"gauss_{Q}x{T}", rng.standard_normal((Q, T)).astype(np.float32)
        yield f"ties_{Q}x{T}", rng.integers(0, 3, (Q, T)).astype(np.float32)
    yield "constant_7x11", np.full((7, 11), 2.5, np.float32)
    yield "zeros_900x20", np.zeros((900, 20), np.float32)
    c = rng.standard_normal((40, 8)).astype(np.float32)
    c[rng.random((40, 8)) < 0.3] = np.inf
    yield "some_inf_40x8", c
    yield "big_range", (rng.standard_normal((200, 50)) * 10.0 ** rng.integers(-6, 6, (200, 50))).astype(np.float32)


@pytest.mark.parametrize("name,cost", list(_cost_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_assignment_bit_exact(name, cost):
    from matcher_b200 import hungarian_assign
    pairs, status = hungarian_assign([torch.from_numpy(cost).to(_dev())])
    assert int(status[0]) == 0
    _check_pairs(pairs[0], cost, name)


def test_assignment_batch_of_problems_one_launch():
    from matcher_b200 import hungarian_assign
    rng = np.random.default_rng(5)
    blocks = [rng.standard_normal((900, t)).astype(np.float32) for t in (3, 0, 41, 12, 1, 97, 12)]
    blocks[3] = np.round(blocks[3])            # ties
    pairs, status = hungarian_assign([torch.from_numpy(b).to(_dev()) for b in blocks])
    assert status.cpu().tolist() == [0] * len(blocks)
    for k, b in enumerate(blocks):
        if b.shape[1] == 0:
            assert pairs[k][0].numel() == 0 and pairs[k][1].numel() == 0
        else:
            _check_pairs(pairs[k], b, f"block {k}")


def test_assignment_error_classes():
    from matcher_b200 import hungarian_assign
    bad = np.zeros((6, 3), np.float32)
    bad[2, 1] = np.nan
    inf = np.full((4, 2), np.inf, np.float32)
    ok = np.arange(12, dtype=np.float32).reshape(4, 3)
    for blocks, want in (([bad], [-2]), ([inf], [-1])):
        pairs, status = hungarian_assign([torch.from_numpy(b).to(_dev()) for b in blocks])
        assert status.cpu().tolist() == want
        assert (pairs[0][0] == -1).all() and (pairs[0][1] == -1).all()
    pairs, status = hungarian_assign([torch.from_numpy(ok).to(_dev())])
    assert int(status[0]) == 0
    _check_pairs(pairs[0], ok, "ok")


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_matcher_against_reference_golden(path):
    """HungarianMatcher.forward, called like the criterion does (one image), against what the reference produced."""
    from matcher_b200 import HungarianMatcher, hungarian_match
    g = np.load(path)
    w = g["weights"]
    dev = _dev()
    m = HungarianMatcher(cost_class=float(w[0]), cost_bbox=float(w[1]), cost_giou=float(w[2]), focal_alpha=float(g["focal_alpha"]),
                         indices_device=None)
    outputs = {"pred_logits": torch.from_numpy(g["pred_logits"])[None].to(dev), "pred_boxes": torch.from_numpy(g["pred_boxes"])[None].to(dev)}
    targets = [{"labels": torch.from_numpy(g["labels"]).to(dev), "boxes": torch.from_numpy(g["tgt_boxes"]).to(dev)}]
    label_map = torch.from_numpy(g["label_map"])               # on the CPU, like create_positive_map's output
    (i, j), = m(outputs, targets, label_map)
    assert i.is_cuda and i.dtype == torch.int64 and j.dtype == torch.int64
    # the cost matrix our kernel built
    rows = label_map.to(dev)[targets[0]["labels"]]
    pairs, costs, status = hungarian_match(outputs["pred_logits"], outputs["pred_boxes"], targets[0]["boxes"], rows,
                                           [(0, len(g["labels"]))], return_cost=True, **m._weights())
    c = costs[0].cpu().numpy()
    want = g["cost"]
    err = np.abs(c.astype(np.float64) - want).max() / np.abs(want).max()
    assert err <= COST_TOL, err
    assert int(status[0]) == 0 and torch.equal(pairs[0][0], i) and torch.equal(pairs[0][1], j)
    # bit-exact assignment for OUR matrix ...
    _check_pairs((i, j), c, "own cost")
    # ... and the reference's pairs whenever its optimum is not decided by exact ties in the last bit
    if "ties" not in path:
        assert np.array_equal(i.cpu().numpy(), g["index_i"]) and np.array_equal(j.cpu().numpy(), g["index_j"])


def test_match_sets_equals_per_image_calls_and_oracle():
    """All seven output sets of a training step (main + 5 aux + interm) x 4 images in ONE call."""
    from matcher_b200 import HungarianMatcher
    dev = _dev()
    gen = torch.Generator().manual_seed(21)
    bs, Q, K, sets = 4, 900, 256, 7
    sizes = [9, 0, 23, 1]
    targets, label_maps = [], []
    for t in sizes:
        n_cat = 5
        lm = torch.zeros(n_cat, K)
        for c in range(n_cat):
            lm[c, 1 + 3 * c:1 + 3 * c + 2] = 1.0
        label_maps.append(lm)
        targets.append({"labels": torch.randint(0, n_cat, (t,), generator=gen).to(dev),
                        "boxes": torch.cat([torch.rand(t, 2, generator=gen), torch.rand(t, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)})
    output_sets = [{"pred_logits": (torch.randn(bs, Q, K, generator=gen) * 2).to(dev),
                    "pred_boxes": torch.cat([torch.rand(bs, Q, 2, generator=gen), torch.rand(bs, Q, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)}
                   for _ in range(sets)]
    m = HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, indices_device=None)
    got = m.match_sets(output_sets, targets, label_maps)
    assert len(got) == sets and all(len(g) == bs for g in got)
    for s in range(sets):
        for b in range(bs):
            single = m({"pred_logits": output_sets[s]["pred_logits"][b:b + 1], "pred_boxes": output_sets[s]["pred_boxes"][b:b + 1]},
                       [targets[b]], label_maps[b])[0]
            assert torch.equal(single[0], got[s][b][0]) and torch.equal(single[1], got[s][b][1])
            assert got[s][b][0].numel() == sizes[b]
        # the whole reference pipeline restated on the CPU (fp32 torch ops + scipy's algorithm)
        want = matcher_port.match(output_sets[s]["pred_logits"], output_sets[s]["pred_boxes"],
                                  [t["boxes"] for t in targets], [label_maps[b][targets[b]["labels"].cpu()] for b in range(bs)],
                                  cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal_alpha=0.25)
        for b in range(bs):
            assert torch.equal(got[s][b][0].cpu(), want[b][0]) and torch.equal(got[s][b][1].cpu(), want[b][1]), (s, b)


def test_edge_inputs_and_errors():
    from matcher_b200 import HungarianMatcher, build_matcher
    dev = _dev()
    m = HungarianMatcher(indices_device=None)
    K, Q = 32, 50
    lm = torch.zeros(3, K)
    lm[0, 1:3] = 1
    lm[1, 4:5] = 1                      # category 2 has an all-zero map: 0/0 -> NaN -> cost 0 like the reference
    logits = torch.randn(1, Q, K)
    logits[0, 3, :] = float("nan")      # a NaN prediction row is zeroed, not propagated (matcher.py:102)
    logits[0, 4, 0] = float("inf")
    boxes = torch.rand(1, Q, 4) * 0.5 + 0.1
    tgt = {"labels": torch.tensor([0, 1, 2, 2]), "boxes": torch.rand(4, 4) * 0.5 + 0.1}
    (i, j), = m({"pred_logits": logits.to(dev), "pred_boxes": boxes.to(dev)}, [{k: v.to(dev) for k, v in tgt.items()}], lm)
    want = matcher_port.match(logits, boxes, [tgt["boxes"]], [lm[tgt["labels"]]])
    assert torch.equal(i.cpu(), want[0][0]) and torch.equal(j.cpu(), want[0][1])
    # no targets at all
    (i, j), = m({"pred_logits": logits.to(dev), "pred_boxes": boxes.to(dev)},
                [{"labels": torch.zeros(0, dtype=torch.int64, device=dev), "boxes": torch.zeros(0, 4, device=dev)}], lm)
    assert i.numel() == 0 and j.numel() == 0
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        m({"pred_logits": logits, "pred_boxes": boxes}, [tgt], lm)
    with pytest.raises(AssertionError):
        HungarianMatcher(0, 0, 0)

    class Args:
        matcher_type, set_cost_class, set_cost_bbox, set_cost_giou, focal_alpha = "HungarianMatcher", 1.0, 5.0, 2.0, 0.25
    assert isinstance(build_matcher(Args), HungarianMatcher)


def _random_problem(gen, Q, K, T, n_cat, span):
    lm = torch.zeros(n_cat, K)
    for c in range(n_cat):
        start = int(torch.randint(0, K - span, (1,), generator=gen))
        lm[c, start:start + span] = torch.rand(span, generator=gen) + 0.1 if c % 2 else 1.0
    labels = torch.randint(0, n_cat, (T,), generator=gen)
    tb = torch.cat([torch.rand(T, 2, generator=gen), torch.rand(T, 2, generator=gen) * 0.4 + 0.02], -1)
    logits = torch.randn(1, Q, K, generator=gen) * 2
    pb = torch.cat([torch.rand(1, Q, 2, generator=gen), torch.rand(1, Q, 2, generator=gen) * 0.4 + 0.02], -1)
    return logits, pb, tb, lm, labels


@pytest.mark.parametrize("span,K,T", [(3, 256, 12), (16, 256, 40), (17, 256, 12), (60, 256, 33), (5, 100, 7), (40, 300, 9)])
def test_cost_sparse_and_dense_paths(span, K, T, monkeypatch):
    """Rows with <= 16 non-zero tokens take the per-pair sparse path, wider rows (or HM_COST_DENSE) the warp-per-row dense
    path; both must reproduce the reference's fp32 cost within 1e-5 and give the oracle's pairs on their own matrix."""
    from matcher_b200 import hungarian_match
    dev = _dev()
    gen = torch.Generator().manual_seed(100 + span)
    Q = 300
    logits, pb, tb, lm, labels = _random_problem(gen, Q, K, T, 6, span)
    rows = lm[labels]
    want = matcher_port.cost_matrix(logits[0], pb[0], tb, rows, 2.0, 5.0, 2.0, 0.25).numpy()
    got = {}
    for mode in ("auto", "dense"):
        if mode == "dense":
            monkeypatch.setenv("HM_COST_DENSE", "1")
        pairs, costs, status = hungarian_match(logits.to(dev), pb.to(dev), tb.to(dev), rows.to(dev), [(0, T)],
                                               cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, return_cost=True)
        c = costs[0].cpu().numpy()
        assert np.abs(c.astype(np.float64) - want).max() / np.abs(want).max() <= COST_TOL, mode
        _check_pairs(pairs[0], c, mode)
        got[mode] = c
    assert np.abs(got["auto"] - got["dense"]).max() <= 2e-6 * np.abs(want).max()


def test_nan_logit_zeroes_the_whole_row_on_both_paths(monkeypatch):
    """NaN * 0 = NaN in the reference's dense product: one NaN logit, even at a token no target uses, zeroes the row."""
    from matcher_b200 import hungarian_match
    dev = _dev()
    gen = torch.Generator().manual_seed(9)
    logits, pb, tb, lm, labels = _random_problem(gen, 40, 64, 6, 3, 2)
    logits[0, 7, 63] = float("nan")
    rows = lm[labels]
    want = matcher_port.cost_matrix(logits[0], pb[0], tb, rows).numpy()
    assert (want[7] == 0).all()
    for mode in ("auto", "dense"):
        if mode == "dense":
            monkeypatch.setenv("HM_COST_DENSE", "1")
        _, costs, _ = hungarian_match(logits.to(dev), pb.to(dev), tb.to(dev), rows.to(dev), [(0, 6)], return_cost=True)
        c = costs[0].cpu().numpy()
        assert (c[7] == 0).all() and np.abs(c - want).max() <= COST_TOL * np.abs(want).max()


def test_one_hot_targets_equal_the_reference_host_loop():
    """groundingdino.py:561-567 on the host vs the device builder, from the same matches."""
    from matcher_b200 import HungarianMatcher, one_hot_targets
    dev = _dev()
    gen = torch.Generator().manual_seed(33)
    bs, Q, K = 3, 120, 48
    sizes = [5, 0, 9]
    label_maps, targets = [], []
    for t in sizes:
        lm = torch.zeros(4, K)
        for c in range(4):
            lm[c, 2 + 5 * c:4 + 5 * c] = 1.0
        label_maps.append(lm)
        targets.append({"labels": torch.randint(0, 4, (t,), generator=gen).to(dev),
                        "boxes": torch.cat([torch.rand(t, 2, generator=gen), torch.rand(t, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)})
    outputs = {"pred_logits": torch.randn(bs, Q, K, generator=gen).to(dev),
               "pred_boxes": torch.cat([torch.rand(bs, Q, 2, generator=gen), torch.rand(bs, Q, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)}
    m = HungarianMatcher(2.0, 5.0, 2.0)
    pairs = m.match_sets([outputs], targets, label_maps)[0]
    got = one_hot_targets(pairs, targets, label_maps, Q, K)
    want = torch.zeros(bs, Q, K, dtype=torch.int64)
    for i in range(bs):                                   # the reference's loop, on the host
        ids = targets[i]["labels"].cpu()[pairs[i][1].cpu()]
        want[i, pairs[i][0].cpu()] = label_maps[i][ids].to(torch.long)
    assert got.is_cuda and torch.equal(got.cpu(), want) and int(want.sum()) > 0


def test_default_returns_cpu_indices_like_the_reference_and_its_host_loop_runs():
    """Unmodified-criterion mode: CPU int64 index tensors (matcher.py:119), usable in the reference's host-side loop
    (groundingdino.py:561-567) exactly as written there."""
    from matcher_b200 import HungarianMatcher
    dev = _dev()
    gen = torch.Generator().manual_seed(12)
    Q, K, T = 80, 40, 6
    label_map = torch.zeros(3, K)
    for c in range(3):
        label_map[c, 2 + 4 * c:4 + 4 * c] = 1.0
    targets = [{"labels": torch.randint(0, 3, (T,), generator=gen).to(dev),
                "boxes": torch.cat([torch.rand(T, 2, generator=gen), torch.rand(T, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)}]
    outputs = {"pred_logits": torch.randn(1, Q, K, generator=gen).to(dev),
               "pred_boxes": torch.cat([torch.rand(1, Q, 2, generator=gen), torch.rand(1, Q, 2, generator=gen) * 0.4 + 0.02], -1).to(dev)}
    indices = HungarianMatcher(2.0, 5.0, 2.0)(outputs, targets, label_map)
    assert all((not i.is_cuda) and i.dtype == torch.int64 and (not j.is_cuda) for i, j in indices)
    same = HungarianMatcher(2.0, 5.0, 2.0, indices_device=None)(outputs, targets, label_map)
    assert torch.equal(same[0][0].cpu(), indices[0][0]) and torch.equal(same[0][1].cpu(), indices[0][1])
    one_hot = torch.zeros(1, Q, K, dtype=torch.int64)             # the reference's loop, verbatim in structure
    tgt_ids = [v["labels"].cpu() for v in targets]
    for i in range(len(indices)):
        tgt_ids[i] = tgt_ids[i][indices[i][1]]
        one_hot[i, indices[i][0]] = label_map[tgt_ids[i]].to(torch.long)
    assert int(one_hot.sum()) == 2 * T
