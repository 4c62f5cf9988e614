#!/usr/bin/env python
"""Generate tests/golden/matcher/*.npz by running the REFERENCE's own HungarianMatcher.

Run in the authoring container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/matcher/make_matcher_golden.py

It loads /root/reference/models/GroundingDINO/matcher.py by file path (importing the package would drag in the whole
model) with /root/reference on sys.path for its ``util.box_ops`` import, calls ``HungarianMatcher.forward`` the way the
criterion does (one image per call, groundingdino.py:548-554), and records per case the inputs, the cost matrix the
reference built (captured by wrapping ``scipy.optimize.linear_sum_assignment`` as seen by that module) and the index
pairs it returned.  Nothing of the reference's source is copied: only numbers it produced.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF_ROOT = "/root/reference"
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def load_reference_matcher():
    sys.path.insert(0, REF_ROOT)
    spec = importlib.util.spec_from_file_location("_ref_matcher", os.path.join(REF_ROOT, "models/GroundingDINO/matcher.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def positive_map(gen, n_cat, K, max_span=4):
    """Rows like create_positive_map's: a short run of ones per category inside the first tokens of the caption."""
    m = torch.zeros(n_cat, K)
    pos = 1
    for c in range(n_cat):
        span = int(torch.randint(1, max_span + 1, (1,), generator=gen))
        m[c, pos:pos + span] = 1.0
        pos += span + 1
    return m


def run_case(mod, name, Q, K, T, n_cat, seed, weights=(2.0, 5.0, 2.0), box_mode="random", logit_scale=2.0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(1, Q, K, generator=g) * logit_scale
    if box_mode == "random":
        pb = torch.cat([torch.rand(1, Q, 2, generator=g), torch.rand(1, Q, 2, generator=g) * 0.5 + 0.01], -1)
        tb = torch.cat([torch.rand(T, 2, generator=g), torch.rand(T, 2, generator=g) * 0.5 + 0.01], -1)
    elif box_mode == "lattice":      # few distinct boxes: many exactly tied costs
        cells = torch.tensor([[0.25, 0.25, 0.5, 0.5], [0.75, 0.25, 0.5, 0.5], [0.25, 0.75, 0.5, 0.5], [0.75, 0.75, 0.5, 0.5]])
        pb = cells[torch.randint(0, 4, (Q,), generator=g)][None]
        tb = cells[torch.randint(0, 4, (T,), generator=g)]
        logits = torch.zeros(1, Q, K)
    else:
        raise ValueError(box_mode)
    labels = torch.randint(0, n_cat, (T,), generator=g)
    label_map = positive_map(g, n_cat, K)
    matcher = mod.HungarianMatcher(cost_class=weights[0], cost_bbox=weights[1], cost_giou=weights[2], focal_alpha=0.25)

    seen = {}
    real = mod.linear_sum_assignment

    def spy(c):
        seen["cost"] = c.detach().numpy().astype(np.float32).copy() if torch.is_tensor(c) else np.asarray(c, dtype=np.float32).copy()
        return real(c)

    mod.linear_sum_assignment = spy
    try:
        (i, j), = matcher({"pred_logits": logits, "pred_boxes": pb}, [{"labels": labels, "boxes": tb}], label_map)
    finally:
        mod.linear_sum_assignment = real
    np.savez_compressed(
        os.path.join(OUT_DIR, f"{name}.npz"),
        pred_logits=logits[0].numpy(), pred_boxes=pb[0].numpy(), tgt_boxes=tb.numpy(), labels=labels.numpy(),
        label_map=label_map.numpy(), weights=np.array(weights, dtype=np.float64), focal_alpha=np.float64(0.25),
        cost=seen.get("cost", np.zeros((Q, 0), np.float32)), index_i=i.numpy(), index_j=j.numpy())
    print(f"{name}: Q={Q} K={K} T={T} -> {len(i)} pairs")


def main():
    mod = load_reference_matcher()
    run_case(mod, "q900_t12", 900, 256, 12, 6, seed=1)                      # the model's shape (cfg_odvg.py: 900 queries, 256 tokens)
    run_case(mod, "q900_t1", 900, 256, 1, 1, seed=2)
    run_case(mod, "q900_t97", 900, 256, 97, 20, seed=3)
    run_case(mod, "q64_t5_unit", 64, 32, 5, 3, seed=4, weights=(1.0, 1.0, 1.0))
    run_case(mod, "q16_t40_wide", 16, 32, 40, 7, seed=5)                    # more targets than queries: no transpose inside scipy
    run_case(mod, "q100_t30_ties", 100, 16, 30, 4, seed=6, box_mode="lattice")


if __name__ == "__main__":
    main()
