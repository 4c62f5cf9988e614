"""Seeded input generators and comparison helpers shared by the GPU parity tests and bench.py.

Distributions:
  "U"  value~N(0,1), loc~U[0,1), attn = softmax(N(0,1)) over L*P           (kernel-level worst case)
  "W"  like U with loc~U[-0.25,1.25): every border and the reject test      (edge stress)
  "I"  what a default-initialised MSDeformAttn feeds the op: reference points at the pixel centres of
       the query's own level, offsets = head direction * (p+1) pixels, uniform attention
       (ops/modules/ms_deform_attn.py:62-76, models/GroundingDINO/transformer.py:470-485)
"""
import math

import numpy as np

SWIN_T_800x1333 = [(100, 167), (50, 84), (25, 42), (13, 21)]
SWIN_B_1024 = [(128, 128), (64, 64), (32, 32), (16, 16)]


def level_starts(shapes):
    areas = [h * w for h, w in shapes]
    return np.concatenate([[0], np.cumsum(areas)[:-1]]).astype(np.int64)


def encoder_reference_points(shapes):
    """(S, 2) pixel-centre (x, y) of every pyramid cell, normalised by its own level (valid_ratio = 1)."""
    pts = []
    for h, w in shapes:
        ys, xs = np.meshgrid((np.arange(h, dtype=np.float32) + 0.5) / np.float32(h),
                             (np.arange(w, dtype=np.float32) + 0.5) / np.float32(w), indexing="ij")
        pts.append(np.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    return np.concatenate(pts, 0).astype(np.float32)


def default_init_offsets(M, L, P):
    """(M, L, P, 2) sampling-offset bias of a freshly constructed module."""
    ang = np.arange(M, dtype=np.float32) * np.float32(2.0 * math.pi / M)
    d = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)
    d = d / np.abs(d).max(-1, keepdims=True)
    steps = np.arange(1, P + 1, dtype=np.float32)
    return (d[:, None, None, :] * steps[None, None, :, None] * np.ones((1, L, 1, 1), np.float32)).astype(np.float32)


def make_inputs(shapes, N, M, D, Lq, P, dist="U", seed=3):
    """Returns dict of float32 numpy arrays + int64 tables.  Lq=None means one query per pixel."""
    rng = np.random.default_rng(seed)
    L = len(shapes)
    S = int(sum(h * w for h, w in shapes))
    if Lq is None:
        Lq = S
    value = rng.standard_normal((N, S, M, D), dtype=np.float32)
    if dist in ("U", "W"):
        loc = rng.random((N, Lq, M, L, P, 2), dtype=np.float32)
        if dist == "W":
            loc = loc * np.float32(1.5) - np.float32(0.25)
        logits = rng.standard_normal((N, Lq, M, L * P), dtype=np.float32)
        e = np.exp(logits - logits.max(-1, keepdims=True))
        attn = (e / e.sum(-1, keepdims=True)).reshape(N, Lq, M, L, P).astype(np.float32)
    elif dist == "I":
        if Lq == S:
            ref = encoder_reference_points(shapes)                                   # (S, 2)
        else:
            ref = (1.0 / (1.0 + np.exp(-rng.standard_normal((Lq, 2))))).astype(np.float32)
        wh = np.asarray([(w, h) for h, w in shapes], dtype=np.float32)               # (L, 2)
        off = default_init_offsets(M, L, P) / wh[None, :, None, :]                   # (M, L, P, 2)
        loc = (ref[None, :, None, None, None, :] + off[None, None]).astype(np.float32)
        loc = np.broadcast_to(loc, (N, Lq, M, L, P, 2)).copy()
        attn = np.full((N, Lq, M, L, P), 1.0 / (L * P), dtype=np.float32)
    else:
        raise ValueError(dist)
    grad_out = rng.standard_normal((N, Lq, M * D), dtype=np.float32)
    return dict(shapes=np.asarray(shapes, dtype=np.int64), starts=level_starts(shapes), value=value, loc=loc,
                attn=attn, grad_out=grad_out, dims=(N, S, M, D, L, Lq, P))


def algorithmic_bytes(N, S, M, D, L, Lq, P, e_v, e_l=4):
    """every input read once, every output written once."""
    fwd = N * (S * M * D * e_v + Lq * M * L * P * 3 * e_l + Lq * M * D * e_v)
    bwd = N * (2 * S * M * D * e_v + Lq * M * D * e_v + 2 * Lq * M * L * P * 3 * e_l)
    return fwd, bwd
