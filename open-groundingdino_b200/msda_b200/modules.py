"""nn.Module layer of the op.

``MSDeformAttn`` mirrors ops/modules/ms_deform_attn.py:30-126 of the reference: same constructor
arguments, parameter names (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` --
reference checkpoints load unchanged), initialisation scheme (:62-76), positional ``forward`` and return
value (``output_proj(out)``; no residual, no dropout).

``MultiScaleDeformableAttention`` is the keyword-style adapter for the model code, which constructs the
attention as ``(embed_dims=, num_levels=, num_heads=, num_points=, batch_first=True)`` and calls it with
``query=, reference_points=, value=, spatial_shapes=, level_start_index=, key_padding_mask=``
(models/GroundingDINO/transformer.py:761-767,794-801,827-833,922-929) and then ``init_weights()`` (:195-196).
Its semantics follow the in-repo module of that name (models/GroundingDINO/ms_deform_attn.py:137-363).
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from .functions import MSDeformAttnFunction, MSDeformAttnFusedFunction, fused_mask_supported, fused_supported


def _power_of_two(n: int) -> bool:
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


def sampling_locations_from(reference_points, sampling_offsets, spatial_shapes, n_points):
    """Normalised sampling positions (N, Lq, M, L, P, 2) from reference points and raw offsets.
    2-d references: offsets are pixels of each level (ops/modules/ms_deform_attn.py:102-105);
    4-d references (cx, cy, w, h boxes): offsets are fractions of half the box (:106-108)."""
    last = reference_points.shape[-1]
    if last == 2:
        wh = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        return reference_points[:, :, None, :, None, :] + sampling_offsets / wh[None, None, None, :, None, :]
    if last == 4:
        return (reference_points[:, :, None, :, None, :2]
                + sampling_offsets / n_points * reference_points[:, :, None, :, None, 2:] * 0.5)
    raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {last} instead.")


class MSDeformAttn(nn.Module):
    #: the reference compares sum(H*W) with the value length on the host every call
    #: (ops/modules/ms_deform_attn.py:92), which is a device sync; opt in when debugging.
    validate_shapes = False
    #: fold softmax + sampling-location arithmetic into the kernels when the library supports the shape
    #:; False reproduces the reference's op-by-op graph (forward :98-124).
    fuse_pre_op = True
    #: fp16 values (the reference's --amp mode) run on native fp16-row kernels instead of being widened to fp32 around
    #: the op; False restores the reference's branch (:114-120) literally.
    native_fp16 = True
    #: hand input_padding_mask to the fused kernels (padded rows gather as zero, zero grad_value rows) instead of running
    #: value.masked_fill over the whole value tensor (scope row f4); False = the reference's masked_fill (:95-96).
    #: Used when the queries are few against the value rows (decoder cross-attention: 1.50 -> 1.35 ms per module call at
    #: cfg-3).  With one query per value row (encoder) the kernels would test 64 mask bytes per query and head where the
    #: masked_fill pass touches each row once: measured 2.80 vs 2.70 ms at cfg-2, so the pass is kept there.
    mask_in_kernel = True
    #: call the kernels through torch.ops.msda_b200.* (msda_b200/torch_ops.py) so that selective activation
    #: checkpointing can keep the op's output instead of recomputing it
    use_dispatcher_ops = False

    def __init__(self, d_model: int = 256, n_levels: int = 4, n_heads: int = 8, n_points: int = 4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        if not _power_of_two(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head "
                          "a power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self) -> None:
        # offsets: zero weight; bias = the head's compass direction (L-inf normalised) times (point + 1)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, self.n_points + 1, dtype=torch.float32)
        bias = direction[:, None, None, :] * steps[None, None, :, None]            # (M, 1, P, 2)
        bias = bias.expand(self.n_heads, self.n_levels, self.n_points, 2).reshape(-1)
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            # (the reference rebinds the bias to a fresh CPU tensor here; keeping device/dtype makes
            # init_weights() safe to call after .cuda() as well)
            self.sampling_offsets.bias = nn.Parameter(bias.to(self.sampling_offsets.weight).clone())
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """query (N,Lq,C); reference_points (N,Lq,L,2|4) in [0,1]; input_flatten (N,S,C); input_spatial_shapes
        (L,2) int64 (H,W); input_level_start_index (L) int64; input_padding_mask (N,S) bool, True = padding.
        Returns (N, Lq, C)."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if self.validate_shapes:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

        value = self.value_proj(input_flatten).view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        weights = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if (self.fuse_pre_op and (value.dtype != torch.float16 or self.native_fp16)
                and value.is_contiguous() and offsets.is_contiguous() and weights.is_contiguous()
                and fused_supported(value, offsets, weights, reference_points, self.n_levels, self.n_points)):
            mask = None
            if input_padding_mask is not None:
                if self.mask_in_kernel and 4 * Len_q <= Len_in and fused_mask_supported(self.n_heads, self.n_levels):
                    mask = input_padding_mask                        # folded into the gather / scatter
                else:
                    value = value.masked_fill(input_padding_mask[..., None, None], float(0))
            if self.use_dispatcher_ops:
                from . import torch_ops  # noqa: F401  (registers torch.ops.msda_b200)
                out = torch.ops.msda_b200.fused_forward(value, input_spatial_shapes, input_level_start_index, offsets,
                                                        weights, reference_points.detach().float().contiguous(), mask)
            else:
                out = MSDeformAttnFusedFunction.apply(value, input_spatial_shapes, input_level_start_index, offsets,
                                                      weights, reference_points, mask)
            return self.output_proj(out)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None, None], float(0))
        weights = F.softmax(weights, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        locations = sampling_locations_from(reference_points, offsets, input_spatial_shapes, self.n_points)

        if value.dtype == torch.float16 and not self.native_fp16:
            # the reference's AMP branch verbatim (:114-120): fp16 is widened around the op
            out = MSDeformAttnFunction.apply(value.to(torch.float32), input_spatial_shapes, input_level_start_index,
                                             locations.to(torch.float32), weights, self.im2col_step)
            return self.output_proj(out.to(torch.float16))
        # fp32 / fp64 as in the reference; bf16 / fp16 values go to the native 16-bit-row kernels (same numbers as
        # the widened path: the widening is exact, accumulation is fp32, the output is rounded once)
        if self.use_dispatcher_ops:
            from . import torch_ops  # noqa: F401
            out = torch.ops.msda_b200.forward(value, input_spatial_shapes, input_level_start_index,
                                              locations.float().contiguous(), weights.float().contiguous(), self.im2col_step)
        else:
            out = MSDeformAttnFunction.apply(value, input_spatial_shapes, input_level_start_index, locations, weights,
                                             self.im2col_step)
        return self.output_proj(out)


class MultiScaleDeformableAttention(MSDeformAttn):
    """Keyword adapter (see module docstring).  ``embed_dims`` (mmcv spelling) and ``embed_dim`` (in-repo
    spelling) are both accepted.  By default the result is ``output_proj(out)`` like the in-repo module;
    ``add_identity=True`` reproduces the mmcv class the fork imports at transformer.py:29, which returns
    ``dropout(output_proj(out)) + identity`` with ``identity`` defaulting to the (position-added) query."""

    def __init__(self, embed_dims: Optional[int] = None, num_heads: int = 8, num_levels: int = 4,
                 num_points: int = 4, img2col_step: int = 64, batch_first: bool = False,
                 embed_dim: Optional[int] = None, im2col_step: Optional[int] = None, dropout: float = 0.0,
                 add_identity: bool = False, **_ignored):
        dim = embed_dims if embed_dims is not None else (embed_dim if embed_dim is not None else 256)
        super().__init__(d_model=dim, n_levels=num_levels, n_heads=num_heads, n_points=num_points)
        self.im2col_step = im2col_step if im2col_step is not None else img2col_step
        self.batch_first = batch_first
        self.embed_dim = self.embed_dims = dim
        self.num_heads, self.num_levels, self.num_points = num_heads, num_levels, num_points
        self.add_identity = add_identity
        self.dropout = nn.Dropout(dropout)

    def init_weights(self) -> None:
        self._reset_parameters()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **_ignored):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:          # (n, bs, c) -> (bs, n, c)
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        out = super().forward(query, reference_points, value, spatial_shapes, level_start_index, key_padding_mask)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        if self.add_identity:
            out = self.dropout(out) + identity
        return out
