"""Torch restatement of the reference's CPU path for multi-scale deformable attention.

TEST INFRASTRUCTURE ONLY (see oracle/msda_oracle.c for the rule).  The reference has no CPU
kernel: on a host it runs ``ms_deform_attn_core_pytorch``
(models/GroundingDINO/ops/functions/ms_deform_attn_func.py:41-61 == models/GroundingDINO/
ms_deform_attn.py:94-134), i.e. one ``F.grid_sample(bilinear, zeros, align_corners=False)`` per
pyramid level followed by an attention-weighted sum.  This file re-expresses that algorithm so it
can travel to the GPU box (``/root/reference`` does not exist there) and be timed on the host
cores as bench.py's ``cpu_baseline`` / ``--impl reference`` leg ("kind": "port").

Parity status: pinned -- tests/test_oracle_cpu.py compares it with golden outputs/gradients that
tests/golden/make_golden.py produced by calling the reference function itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_grid_sample(value: torch.Tensor, spatial_shapes, sampling_locations: torch.Tensor,
                     attention_weights: torch.Tensor) -> torch.Tensor:
    """value (N,S,M,D), spatial_shapes [(H,W)]*L, sampling_locations (N,Lq,M,L,P,2) as (x,y) in
    [0,1], attention_weights (N,Lq,M,L,P)  ->  (N, Lq, M*D)."""
    n, _, heads, ch = value.shape
    lq, levels, points = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]

    # head-major images, one per (batch, head): (N*M, D, S)
    planes = value.permute(0, 2, 3, 1).reshape(n * heads, ch, -1)
    # grid_sample wants [-1, 1]; (N*M, Lq, L, P, 2)
    grid = (sampling_locations * 2 - 1).permute(0, 2, 1, 3, 4, 5).reshape(n * heads, lq, levels, points, 2)

    taps = []
    first = 0
    for lvl, (h, w) in enumerate(hw):
        image = planes[:, :, first:first + h * w].reshape(n * heads, ch, h, w)
        first += h * w
        taps.append(F.grid_sample(image, grid[:, :, lvl], mode="bilinear", padding_mode="zeros",
                                  align_corners=False))           # (N*M, D, Lq, P)
    taps = torch.cat(taps, dim=-1)                                  # (N*M, D, Lq, L*P)
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * heads, 1, lq, levels * points)
    mixed = (taps * weights).sum(dim=-1)                            # (N*M, D, Lq)
    return mixed.reshape(n, heads * ch, lq).permute(0, 2, 1).contiguous()


def fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """One forward + backward through the port; returns (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    s = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = msda_grid_sample(v, spatial_shapes, s, a)
    out.backward(grad_output.reshape(out.shape))
    return out.detach(), v.grad, s.grad, a.grad
