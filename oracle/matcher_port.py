"""CPU restatement of the reference's Hungarian matcher.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package never imports it.

What it follows (paths under /root/reference/models/GroundingDINO/):
  * ``cost_matrix``  -- matcher.py:66-101, the matching cost of one image: token-sigmoid focal class cost against the
    normalised positive map of each target (:74-91), L1 box cost (``torch.cdist(p=1)``, :82) and negative generalised
    IoU (:98, /root/reference/util/box_ops.py:9-14,25-62), combined with the three weights and with NaN / inf entries
    zeroed (:100-103);
  * ``assign``       -- matcher.py:105-107: ``scipy.optimize.linear_sum_assignment`` on the CPU copy (oracle/lsap_oracle.c
    restates scipy's solver; ``assign_scipy`` calls scipy itself where it is installed);
  * ``match``        -- HungarianMatcher.forward for a batch (matcher.py:47-119): block c[i] of image i, int64 index pairs.

Parity status: pinned -- tests/golden/matcher/*.npz were produced by importing the reference's own ``HungarianMatcher``
in this container (tests/golden/matcher/make_matcher_golden.py); tests/test_matcher_oracle_cpu.py checks this file against them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblsap_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lsap_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liblsap_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _lsap():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.lsap_solve.restype = ctypes.c_int
        _lib.lsap_solve.argtypes = [ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 3
    return _lib


def assign(cost) -> tuple[np.ndarray, np.ndarray]:
    """(row_ind, col_ind) of the minimum-cost assignment of a 2-d cost matrix, scipy conventions."""
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
    nr, nc = c.shape
    k = min(nr, nc)
    a = np.empty(k, dtype=np.int64)
    b = np.empty(k, dtype=np.int64)
    rc = _lsap().lsap_solve(nr, nc, c.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                            b.ctypes.data_as(ctypes.c_void_p))
    if rc == -1:
        raise ValueError("cost matrix is infeasible")
    if rc == -2:
        raise ValueError("matrix contains invalid numeric entries")
    return a, b


def assign_scipy(cost):
    from scipy.optimize import linear_sum_assignment
    return linear_sum_assignment(np.asarray(cost))


def _xyxy(b: torch.Tensor) -> torch.Tensor:
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def _giou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    iou = inter / (union + 1e-6)
    lt = torch.min(b1[:, None, :2], b2[:, :2])
    rb = torch.max(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    hull = wh[:, :, 0] * wh[:, :, 1]
    return iou - (hull - union) / (hull + 1e-6)


def cost_matrix(pred_logits: torch.Tensor, pred_boxes: torch.Tensor, tgt_boxes: torch.Tensor,
                tgt_label_map: torch.Tensor, cost_class: float = 1.0, cost_bbox: float = 1.0,
                cost_giou: float = 1.0, focal_alpha: float = 0.25) -> torch.Tensor:
    """pred_logits (Q, K), pred_boxes (Q, 4) cxcywh, tgt_boxes (T, 4), tgt_label_map (T, K) = label_map[tgt_ids]
    (un-normalised positive-map rows)  ->  (Q, T) cost, NaN / inf zeroed."""
    prob = pred_logits.sigmoid()
    gamma = 2.0
    neg = (1 - focal_alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
    pos = focal_alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
    l1 = torch.cdist(pred_boxes, tgt_boxes, p=1)
    cls = []
    for row in tgt_label_map:
        row = row / row.sum()
        cls.append(pos @ row - neg @ row)
    cls = torch.stack(cls, dim=0).T if cls else torch.zeros_like(l1)
    giou = -_giou(_xyxy(pred_boxes), _xyxy(tgt_boxes))
    c = cost_bbox * l1 + cost_class * cls + cost_giou * giou
    c = c.clone()
    c[torch.isnan(c)] = 0.0
    c[torch.isinf(c)] = 0.0
    return c


def match(pred_logits: torch.Tensor, pred_boxes: torch.Tensor, tgt_boxes: list, tgt_label_maps: list,
          solver=assign, **weights):
    """Batch form: pred_logits (B, Q, K), pred_boxes (B, Q, 4); per image a (T_b, 4) box tensor and a (T_b, K) map.
    Returns [(index_i, index_j)] * B as int64 tensors, like HungarianMatcher.forward."""
    out = []
    for b in range(pred_logits.shape[0]):
        c = cost_matrix(pred_logits[b].float().cpu(), pred_boxes[b].float().cpu(), tgt_boxes[b].float().cpu(),
                        tgt_label_maps[b].float().cpu(), **weights)
        i, j = solver(c.numpy())
        out.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
    return out
