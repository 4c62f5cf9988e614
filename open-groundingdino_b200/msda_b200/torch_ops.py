"""Dispatcher-level registration of the op (`torch.ops.msda_b200.*`), for what needs to see it as an operator:
selective activation checkpointing and graph capture tools.

The reference checkpoints every encoder layer (`use_transformer_ckpt`, models/GroundingDINO/transformer.py:580-589,
config/cfg_odvg.py:42), so each layer's MSDeformAttn forward runs twice per training step.
With the op visible to the dispatcher a selective-checkpoint policy can keep just its output (bf16, 45 MB per layer at
the cfg-2 shape) and skip the recomputation:

    from torch.utils.checkpoint import checkpoint, create_selective_checkpoint_contexts
    from msda_b200.torch_ops import save_msda_policy
    y = checkpoint(layer, *args, use_reentrant=False,
                   context_fn=lambda: create_selective_checkpoint_contexts(save_msda_policy))

`MSDeformAttn.use_dispatcher_ops = True` routes the module through these operators (same kernels, same numbers).
"""
from __future__ import annotations

import torch
from torch.utils.checkpoint import CheckpointPolicy

from . import functions as F

_lib_def = torch.library.Library("msda_b200", "DEF")
_lib_def.define("forward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_loc, "
                "Tensor attn_weight, int im2col_step) -> Tensor")
_lib_def.define("backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_loc, "
                "Tensor attn_weight, Tensor grad_output, int im2col_step) -> (Tensor, Tensor, Tensor)")
_lib_def.define("fused_forward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_offsets, "
                "Tensor attn_logits, Tensor reference_points, Tensor? padding_mask=None) -> Tensor")
_lib_def.define("fused_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_offsets, "
                "Tensor attn_logits, Tensor reference_points, Tensor grad_output, Tensor? padding_mask=None) -> (Tensor, Tensor, Tensor)")


@torch.library.impl(_lib_def, "forward", "CUDA")
def _forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    return F.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)


@torch.library.impl(_lib_def, "backward", "CUDA")
def _backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    gv, gl, ga = F.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                           grad_output.contiguous(), im2col_step)
    return gv, gl, ga


@torch.library.impl(_lib_def, "fused_forward", "CUDA")
def _fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, padding_mask=None):
    return F.fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, padding_mask)


@torch.library.impl(_lib_def, "fused_backward", "CUDA")
def _fused_backward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, grad_output,
                    padding_mask=None):
    return F.fused_backward(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points,
                            grad_output, padding_mask)


@torch.library.register_fake("msda_b200::forward")
def _forward_fake(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    return value.new_empty((value.shape[0], sampling_loc.shape[1], value.shape[2] * value.shape[3]))


@torch.library.register_fake("msda_b200::backward")
def _backward_fake(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    return torch.empty_like(value), torch.empty_like(sampling_loc), torch.empty_like(attn_weight)


@torch.library.register_fake("msda_b200::fused_forward")
def _fused_forward_fake(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, padding_mask=None):
    return value.new_empty((value.shape[0], sampling_offsets.shape[1], value.shape[2] * value.shape[3]))


@torch.library.register_fake("msda_b200::fused_backward")
def _fused_backward_fake(value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points, grad_output,
                         padding_mask=None):
    return torch.empty_like(value), torch.empty_like(sampling_offsets), torch.empty_like(attn_logits)


def _setup(ctx, inputs, output):
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step = inputs
    ctx.im2col_step = im2col_step
    ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


def _bwd(ctx, grad_output):
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight = ctx.saved_tensors
    gv, gl, ga = torch.ops.msda_b200.backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                              grad_output.contiguous(), ctx.im2col_step)
    return gv, None, None, gl, ga, None


def _fused_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs[:6])
    ctx.padding_mask = inputs[6] if len(inputs) > 6 else None


def _fused_bwd(ctx, grad_output):
    value, spatial_shapes, level_start_index, sampling_offsets, attn_logits, reference_points = ctx.saved_tensors
    gv, go, gg = torch.ops.msda_b200.fused_backward(value, spatial_shapes, level_start_index, sampling_offsets,
                                                    attn_logits, reference_points, grad_output.contiguous(), ctx.padding_mask)
    return gv, None, None, go, gg, None, None


torch.library.register_autograd("msda_b200::forward", _bwd, setup_context=_setup)
torch.library.register_autograd("msda_b200::fused_forward", _fused_bwd, setup_context=_fused_setup)

_SAVE = {torch.ops.msda_b200.forward.default, torch.ops.msda_b200.fused_forward.default}


def save_msda_policy(ctx, op, *args, **kwargs):
    """Selective-checkpoint policy: keep the op's output, recompute everything else."""
    return CheckpointPolicy.MUST_SAVE if op in _SAVE else CheckpointPolicy.PREFER_RECOMPUTE
