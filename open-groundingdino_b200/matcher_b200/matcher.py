"""Host side of the on-device Hungarian matcher.

The reference builds the matching cost of ONE image with ~40 small PyTorch launches, copies it to the host
(``C.cpu()``, matcher.py:101 -- a full device synchronisation) and runs scipy on the CPU, seven times per image and
step (main, five auxiliary and the interm output: groundingdino.py:548-554,587-593,613-619).  Here the cost matrices
of every (output set, image) problem and their assignments come from two kernel launches (csrc/hm_match.cu) and
nothing is read back: the index tensors stay on the device; their lengths are known on the host from the target lists.

Python only validates, allocates with torch and passes raw pointers plus the current stream to the C ABI.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from msda_b200 import _lib


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def hungarian_match(pred_logits: torch.Tensor, pred_boxes: torch.Tensor, tgt_boxes: torch.Tensor,
                    tgt_label_map: torch.Tensor, problems: Sequence[tuple], cost_class: float = 1.0,
                    cost_bbox: float = 1.0, cost_giou: float = 1.0, focal_alpha: float = 0.25, return_cost: bool = False):
    """Match P problems in one call.

    pred_logits (P, Q, K) and pred_boxes (P, Q, 4) cxcywh, fp32 CUDA; tgt_boxes (Ttot, 4); tgt_label_map (Ttot, K), row t
    = ``label_map[tgt_ids[t]]`` un-normalised (matcher.py:78); ``problems[p] = (first target row, T_p)`` (host ints --
    the reference also takes the sizes from ``len(v["boxes"])``, matcher.py:105).

    Returns ``[(index_i, index_j)] * P`` int64 CUDA tensors (views of two flat buffers) with scipy's conventions
    (predictions ascending), plus -- with ``return_cost`` -- the list of (Q, T_p) cost matrices and the (P,) int32 status
    tensor (0 solved; the matcher zeroes NaN / inf, so nothing else can occur on this path).
    """
    _require(pred_logits.dim() == 3 and pred_boxes.dim() == 3 and pred_boxes.size(-1) == 4, "pred_logits (P,Q,K), pred_boxes (P,Q,4)")
    if not pred_logits.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    P, Q, K = pred_logits.shape
    _require(tuple(pred_boxes.shape) == (P, Q, 4), "pred_boxes does not match pred_logits")
    _require(len(problems) == P, "one (first_target_row, num_targets) pair per problem")
    dev = pred_logits.device
    logits = pred_logits.detach().float().contiguous()
    boxes = pred_boxes.detach().float().contiguous()
    tb = tgt_boxes.detach().to(dev, torch.float32).contiguous().view(-1, 4)
    lm = tgt_label_map.detach().to(dev, torch.float32).contiguous().view(-1, K)
    _require(lm.size(0) == tb.size(0), "one positive-map row per target box")

    table, cost_off, pair_off, t_max = [], 0, 0, 0
    for first, t in problems:
        first, t = int(first), int(t)
        _require(t >= 0 and first >= 0 and first + t <= tb.size(0), "problem refers to targets outside tgt_boxes")
        table.append((first, t, cost_off, pair_off))
        cost_off += Q * t
        pair_off += min(Q, t)
        t_max = max(t_max, t)
    _require(cost_off < 2 ** 31, "cost matrices exceed 32-bit element offsets")
    lib = _lib.lib()
    _require(max(Q, t_max) <= lib.hm_max_columns(), f"max(Q, T) above {lib.hm_max_columns()} is not supported")
    prob = torch.tensor(table, dtype=torch.int32).view(P, 4).to(dev, non_blocking=True)
    cost = torch.empty(cost_off, dtype=torch.float32, device=dev)
    index_i = torch.empty(pair_off, dtype=torch.int64, device=dev)
    index_j = torch.empty(pair_off, dtype=torch.int64, device=dev)
    status = torch.empty(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        rc = lib.hm_match(_ptr(logits), _ptr(boxes), _ptr(tb), _ptr(lm), _ptr(prob), _ptr(cost), _ptr(index_i),
                          _ptr(index_j), _ptr(status), P, Q, K, t_max, float(cost_class), float(cost_bbox),
                          float(cost_giou), float(focal_alpha), stream)
    _lib.check(rc, "hm_match")
    counts = [min(Q, t) for (_, t, _, _) in table]
    pairs = list(zip(index_i.split(counts), index_j.split(counts)))     # two calls instead of 2P slicing ops
    if return_cost:
        costs = [cost[c:c + Q * t].view(Q, t) for (_, t, c, _) in table]
        return pairs, costs, status
    return pairs


def hungarian_assign(cost_blocks: Sequence[torch.Tensor]):
    """The assignment step alone (hm_assign): ``cost_blocks[p]`` is a (Q, T_p) fp32 CUDA matrix, Q shared.  Returns
    (pairs, status) like ``hungarian_match``; unlike the matcher path the blocks may hold +inf (forbidden pairs), and a
    NaN / -inf entry or an infeasible block is reported in ``status`` (-2 / -1, scipy raises for both) with -1 pairs."""
    _require(len(cost_blocks) > 0, "no problems")
    if not cost_blocks[0].is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    dev = cost_blocks[0].device
    Q = cost_blocks[0].size(0)
    table, cost_off, pair_off, t_max = [], 0, 0, 0
    for c in cost_blocks:
        _require(c.dim() == 2 and c.size(0) == Q and c.dtype == torch.float32, "blocks must be (Q, T_p) fp32")
        t = c.size(1)
        table.append((0, t, cost_off, pair_off))
        cost_off += Q * t
        pair_off += min(Q, t)
        t_max = max(t_max, t)
    _require(cost_off < 2 ** 31, "cost matrices exceed 32-bit element offsets")
    lib = _lib.lib()
    _require(max(Q, t_max) <= lib.hm_max_columns(), f"max(Q, T) above {lib.hm_max_columns()} is not supported")
    cost = torch.cat([c.contiguous().view(-1) for c in cost_blocks]) if cost_off else torch.empty(0, device=dev)
    prob = torch.tensor(table, dtype=torch.int32).view(-1, 4).to(dev, non_blocking=True)
    index_i = torch.empty(pair_off, dtype=torch.int64, device=dev)
    index_j = torch.empty(pair_off, dtype=torch.int64, device=dev)
    status = torch.empty(len(table), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.hm_assign(_ptr(cost), _ptr(prob), _ptr(index_i), _ptr(index_j), _ptr(status), len(table), Q, t_max,
                           torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "hm_assign")
    return [(index_i[o:o + min(Q, t)], index_j[o:o + min(Q, t)]) for (_, t, _, o) in table], status


class HungarianMatcher(nn.Module):
    """Drop-in for the reference class of the same name (matcher.py:25-119): same constructor, same
    ``forward(outputs, targets, label_map)`` and the same return structure.

    ``indices_device``: the reference returns CPU index tensors (matcher.py:119) and its criterion indexes CPU tensors
    with them (``tgt_ids[i][indices[i][1]]``, groundingdino.py:561-567), so the default "cpu" keeps the unmodified
    criterion working -- at the price of ONE device->host read per call (instead of the reference's ~40 launches, the
    cost-matrix copy and scipy).  ``None`` leaves them on the predictions' device: no synchronisation at all; use it
    together with ``one_hot_targets`` / ``match_sets``."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25,
                 indices_device="cpu"):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.focal_alpha = focal_alpha
        self.indices_device = indices_device

    def _weights(self):
        return dict(cost_class=self.cost_class, cost_bbox=self.cost_bbox, cost_giou=self.cost_giou,
                    focal_alpha=self.focal_alpha)

    @torch.no_grad()
    def forward(self, outputs, targets, label_map):
        """outputs: {"pred_logits": (bs, Q, K), "pred_boxes": (bs, Q, 4)}; targets: list of {"labels": (T_b,), "boxes":
        (T_b, 4)}; label_map: (num_categories, K) positive map shared by the call (matcher.py:78)."""
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        dev = logits.device
        sizes = [len(v["boxes"]) for v in targets]
        tgt_ids = torch.cat([v["labels"] for v in targets]).to(dev)
        tgt_boxes = torch.cat([v["boxes"] for v in targets]).to(dev)
        rows = label_map.to(dev)[tgt_ids]
        problems, first = [], 0
        for t in sizes:
            problems.append((first, t))
            first += t
        pairs = hungarian_match(logits, boxes, tgt_boxes, rows, problems, **self._weights())
        if self.indices_device is not None and torch.device(self.indices_device) != dev:
            flat_i = torch.cat([p[0] for p in pairs]).to(self.indices_device)      # one copy for the whole call
            flat_j = torch.cat([p[1] for p in pairs]).to(self.indices_device)
            counts = [p[0].numel() for p in pairs]
            pairs = list(zip(flat_i.split(counts), flat_j.split(counts)))
        return pairs

    @torch.no_grad()
    def match_sets(self, output_sets, targets, label_maps):
        """All output sets of a step in one call.  ``output_sets``: list of dicts like ``outputs`` above (main, aux...,
        interm); ``label_maps``: one (num_categories_b, K) map per image (groundingdino.py:538-547).  Returns
        ``result[s][b] = (index_i, index_j)``, always on the predictions' device (no synchronisation)."""
        dev = output_sets[0]["pred_logits"].device
        bs = len(targets)
        sizes = [len(v["boxes"]) for v in targets]
        tgt_boxes = torch.cat([v["boxes"] for v in targets]).to(dev)
        rows = torch.cat([label_maps[b].to(dev)[targets[b]["labels"].to(dev)] for b in range(bs)])
        logits = torch.cat([o["pred_logits"] for o in output_sets])       # (sets * bs, Q, K)
        boxes = torch.cat([o["pred_boxes"] for o in output_sets])
        firsts, first = [], 0
        for t in sizes:
            firsts.append(first)
            first += t
        problems = [(firsts[b], sizes[b]) for _ in output_sets for b in range(bs)]
        flat = hungarian_match(logits, boxes, tgt_boxes, rows, problems, **self._weights())
        return [flat[s * bs:(s + 1) * bs] for s in range(len(output_sets))]


def one_hot_targets(pairs, targets, label_maps, num_queries: int, num_tokens: int) -> torch.Tensor:
    """The (bs, Q, K) int64 classification target the criterion builds from the matches on the host
    (groundingdino.py:561-567: ``one_hot[i, indices[i][0]] = label_map_list[i][tgt_ids[i][indices[i][1]]]``), built on
    the device from device index pairs -- index arithmetic only, no synchronisation."""
    dev = pairs[0][0].device
    one_hot = torch.zeros(len(pairs), num_queries, num_tokens, dtype=torch.int64, device=dev)
    for b, (idx_i, idx_j) in enumerate(pairs):
        if idx_i.numel():
            rows = label_maps[b].to(dev)[targets[b]["labels"].to(dev)[idx_j]]
            one_hot[b, idx_i] = rows.to(torch.int64)
    return one_hot


def build_matcher(args):
    """matcher.py:204-217.  Only the Hungarian matcher has a device implementation."""
    assert args.matcher_type in ["HungarianMatcher", "SimpleMinsumMatcher"], \
        "Unknown args.matcher_type: {}".format(args.matcher_type)
    if args.matcher_type == "HungarianMatcher":
        return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                                cost_giou=args.set_cost_giou, focal_alpha=args.focal_alpha,
                                indices_device=getattr(args, "matcher_indices_device", "cpu"))
    raise NotImplementedError("matcher_b200 implements HungarianMatcher only (got {})".format(args.matcher_type))
