"""Autograd boundary of the op, mirroring ops/functions/ms_deform_attn_func.py of the reference.

``ms_deform_attn_forward`` / ``ms_deform_attn_backward`` have the signatures of the reference's pybind
functions (ops/src/vision.cpp:13-16, ops/src/ms_deform_attn.h:20-61) and the same preconditions and
error behaviour (RuntimeError for non-contiguous or CPU tensors, ops/src/cuda/ms_deform_attn_cuda.cu:
28-38,52 and ms_deform_attn.h:38).  They validate, allocate outputs with torch, fetch the current
CUDA stream and hand raw pointers to the C ABI (include/msda_b200.h).  Nothing is computed in Python.

dtype contract: value fp32 (reference), fp64 (reference), bf16 or fp16 (new, native 16-bit rows with fp32
accumulation); sampling locations and attention weights are fp32 (fp64 with fp64 value) -- bf16/fp16
coordinates are widened, because a 16-bit pixel coordinate is useless.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_DTYPE_CODE = {torch.float32: _lib.MSDA_F32, torch.bfloat16: _lib.MSDA_BF16, torch.float64: _lib.MSDA_F64,
               torch.float16: _lib.MSDA_F16}

# "atomic" (default) or "deterministic"; read per call so tests can flip it.
_BWD_MODES = {"atomic": _lib.MSDA_BWD_ATOMIC, "deterministic": _lib.MSDA_BWD_DETERMINISTIC}


def backward_mode() -> int:
    name = os.environ.get("MSDA_BWD_MODE", "atomic").lower()
    if name not in _BWD_MODES:
        raise RuntimeError(f"MSDA_BWD_MODE must be one of {sorted(_BWD_MODES)}, got {name!r}")
    return _BWD_MODES[name]


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _check_inputs(named):
    for name, t in named:
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    for name, t in named:
        if not t.is_cuda:
            # ops/src/ms_deform_attn.h:38,60 -- the reference has no CPU implementation either
            raise RuntimeError("Not implemented on the CPU" if name == "value" else f"{name} must be a CUDA tensor")


def _dims(value, spatial_shapes, sampling_loc, attn_weight):
    _require(value.dim() == 4, "value must be (N, S, M, D)")
    _require(sampling_loc.dim() == 6 and sampling_loc.size(5) == 2, "sampling_loc must be (N, Lq, M, L, P, 2)")
    _require(spatial_shapes.dim() == 2 and spatial_shapes.size(1) == 2, "spatial_shapes must be (L, 2)")
    N, S, M, D = value.shape
    L = spatial_shapes.size(0)
    Lq, P = sampling_loc.size(1), sampling_loc.size(4)
    _require(tuple(sampling_loc.shape) == (N, Lq, M, L, P, 2), "sampling_loc does not match value/spatial_shapes")
    _require(tuple(attn_weight.shape) == (N, Lq, M, L, P), "attn_weight does not match sampling_loc")
    return N, S, M, D, L, Lq, P


def _prepare(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    _require(value.dtype in _DTYPE_CODE, f"value dtype {value.dtype} not supported (fp32, fp64, bf16, fp16)")
    coord_dtype = torch.float64 if value.dtype == torch.float64 else torch.float32
    if sampling_loc.dtype != coord_dtype:
        sampling_loc = sampling_loc.to(coord_dtype)
    if attn_weight.dtype != coord_dtype:
        attn_weight = attn_weight.to(coord_dtype)
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64")
    _check_inputs((("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)))
    dims = _dims(value, spatial_shapes, sampling_loc, attn_weight)
    _require(level_start_index.numel() == dims[4], "level_start_index must have one entry per level")
    step = min(dims[0], int(im2col_step))
    # ops/src/cuda/ms_deform_attn_cuda.cu:50-52 -- kept although one launch covers the whole batch
    _require(step > 0 and dims[0] % step == 0, f"batch({dims[0]}) must divide im2col_step({step})")
    return sampling_loc, attn_weight, dims


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step,
                           variant: int = 0):
    """(N,S,M,D), (L,2) i64, (L) i64, (N,Lq,M,L,P,2), (N,Lq,M,L,P), int -> (N, Lq, M*D)."""
    sampling_loc, attn_weight, (N, S, M, D, L, Lq, P) = _prepare(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    with torch.cuda.device(value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = _lib.lib().msda_forward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), output.data_ptr(), N, S, M, D, L, Lq, P, _DTYPE_CODE[value.dtype], variant,
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_forward")
    return output


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step, mode: int | None = None, variant: int = 0):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight] shaped/typed like the three inputs."""
    loc_dtype, attn_dtype = sampling_loc.dtype, attn_weight.dtype
    sampling_loc, attn_weight, (N, S, M, D, L, Lq, P) = _prepare(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
    if grad_output.dtype != value.dtype:
        grad_output = grad_output.to(value.dtype)
    _check_inputs((("grad_output", grad_output),))
    _require(grad_output.numel() == N * Lq * M * D, "grad_output must be (N, Lq, M*D)")
    mode = backward_mode() if mode is None else mode
    code = _DTYPE_CODE[value.dtype]
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_loc = torch.empty_like(sampling_loc)
        grad_attn = torch.empty_like(attn_weight)
        ws_bytes = _lib.lib().msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, code, mode)
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=value.device) if ws_bytes else None
        rc = _lib.lib().msda_backward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
            grad_attn.data_ptr(), N, S, M, D, L, Lq, P, code, mode, variant,
            workspace.data_ptr() if workspace is not None else None, ws_bytes,
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ms_deform_attn_backward")
    if grad_loc.dtype != loc_dtype:
        grad_loc = grad_loc.to(loc_dtype)
    if grad_attn.dtype != attn_dtype:
        grad_attn = grad_attn.to(attn_dtype)
    return [grad_value, grad_loc, grad_attn]


class MSDeformAttnFunction(Function):
    """Same name, argument order and gradient arity as the reference class
    (ops/functions/ms_deform_attn_func.py:21-38): grads for value, sampling_locations and
    attention_weights; None for the two tables and im2col_step."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


# ---- fused pre-op ----------------------------------------------------------
# variant bits handed to the fused entry points (include/msda_b200.h); tests set it to reach the alternative kernels
fused_variant = 0


def fused_supported(value, sampling_offsets, attn_logits, reference_points, n_levels, n_points) -> bool:
    """True when libmsda_b200 can take the raw offsets / logits directly (see include/msda_b200.h)."""
    if not (value.is_cuda and value.dtype in (torch.float32, torch.bfloat16, torch.float16)):
        return False
    if sampling_offsets.dtype != attn_logits.dtype or sampling_offsets.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return False
    if reference_points.requires_grad or reference_points.shape[-1] not in (2, 4):
        return False
    return bool(_lib.lib().msda_fused_supported(value.shape[-1], n_levels, n_points, _DTYPE_CODE[value.dtype],
                                                _DTYPE_CODE[sampling_offsets.dtype]))


def _fused_dims(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points):
    """Shape / dtype validation of the fused entry points: a mismatch must be a RuntimeError, not an out-of-bounds
    device read in the table prologue or the reference-point indexing."""
    _check_inputs((("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                   ("sampling_offsets", sampling_offsets), ("attention_logits", attention_logits),
                   ("reference_points", reference_points)))
    _require(value.dtype in _DTYPE_CODE and sampling_offsets.dtype in _DTYPE_CODE, "unsupported dtype")
    _require(sampling_offsets.dtype == attention_logits.dtype, "sampling_offsets and attention_logits must share a dtype")
    _require(reference_points.dtype == torch.float32, "reference_points must be fp32")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64")
    _require(value.dim() == 4, "value must be (N, S, M, D)")
    _require(sampling_offsets.dim() == 6 and sampling_offsets.size(5) == 2, "sampling_offsets must be (N, Lq, M, L, P, 2)")
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_offsets.shape
    _require(tuple(sampling_offsets.shape) == (N, Lq, M, L, P, 2), "sampling_offsets does not match value")
    _require(tuple(spatial_shapes.shape) == (L, 2) and level_start_index.numel() == L,
             "spatial_shapes must be (L, 2) and level_start_index (L,) with L taken from sampling_offsets")
    _require(attention_logits.numel() == N * Lq * M * L * P and attention_logits.size(0) == N,
             "attention_logits must be (N, Lq, M, L*P)")
    _require(reference_points.dim() == 4 and tuple(reference_points.shape[:3]) == (N, Lq, L)
             and reference_points.shape[-1] in (2, 4), "reference_points must be (N, Lq, L, 2|4)")
    return N, S, M, D, L, Lq, P


def _mask_ptr(padding_mask, N, S):
    """input_padding_mask (N, S) bool / uint8, True = padding (ops/modules/ms_deform_attn.py:95-96) -> byte pointer or None."""
    if padding_mask is None:
        return None, None
    _require(padding_mask.is_cuda and tuple(padding_mask.shape) == (N, S), "padding mask must be a CUDA tensor of shape (N, S)")
    _require(padding_mask.dtype in (torch.bool, torch.uint8), "padding mask must be bool or uint8")
    m = padding_mask.contiguous()
    m = m.view(torch.uint8) if m.dtype == torch.bool else m
    return m, m.data_ptr()


def fused_mask_supported(n_heads: int, n_levels: int) -> bool:
    """True when the fused kernels can take input_padding_mask themselves for this geometry."""
    return bool(_lib.lib().msda_fused_mask_supported(n_heads, n_levels))


def fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points,
                  padding_mask=None):
    """(N,S,M,D), tables, raw offsets (N,Lq,M,L,P,2), raw logits (N,Lq,M,L*P), reference points fp32 -> (N, Lq, M*D).
    With ``padding_mask`` the rows it flags gather as zero (no masked_fill pass over ``value`` needed)."""
    N, S, M, D, L, Lq, P = _fused_dims(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits,
                                       reference_points)
    mask_keep, mask_ptr = _mask_ptr(padding_mask, N, S)
    with torch.cuda.device(value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        rc = _lib.lib().msda_fused_forward_masked(
            value.data_ptr(), mask_ptr, spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_offsets.data_ptr(),
            attention_logits.data_ptr(), reference_points.data_ptr(), output.data_ptr(), N, S, M, D, L, Lq, P,
            reference_points.shape[-1], _DTYPE_CODE[value.dtype], _DTYPE_CODE[sampling_offsets.dtype], fused_variant,
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "msda_fused_forward")
    return output


def fused_backward(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points,
                   grad_output, padding_mask=None):
    """-> (grad_value, grad_sampling_offsets, grad_attention_logits); grad_value rows flagged by ``padding_mask`` are zero."""
    grad_output = grad_output.contiguous().to(value.dtype)
    N, S, M, D, L, Lq, P = _fused_dims(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits,
                                       reference_points)
    _check_inputs((("grad_output", grad_output),))
    _require(grad_output.numel() == N * Lq * M * D, "grad_output must be (N, Lq, M*D)")
    mask_keep, mask_ptr = _mask_ptr(padding_mask, N, S)
    code = _DTYPE_CODE[value.dtype]
    with torch.cuda.device(value.device):
        grad_value = torch.empty_like(value)
        grad_offsets = torch.empty_like(sampling_offsets)
        grad_logits = torch.empty_like(attention_logits)
        ws_bytes = _lib.lib().msda_backward_workspace_bytes(N, S, M, D, L, Lq, P, code, _lib.MSDA_BWD_ATOMIC)
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=value.device) if ws_bytes else None
        rc = _lib.lib().msda_fused_backward_masked(
            value.data_ptr(), mask_ptr, spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_offsets.data_ptr(),
            attention_logits.data_ptr(), reference_points.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(),
            grad_offsets.data_ptr(), grad_logits.data_ptr(), N, S, M, D, L, Lq, P, reference_points.shape[-1], code,
            _DTYPE_CODE[sampling_offsets.dtype], fused_variant, workspace.data_ptr() if workspace is not None else None, ws_bytes,
            torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "msda_fused_backward")
    return grad_value, grad_offsets, grad_logits


class MSDeformAttnFusedFunction(Function):
    """softmax + sampling-location arithmetic + multi-scale deformable attention in one kernel each way.
    forward(value (N,S,M,D), spatial_shapes, level_start_index, sampling_offsets (N,Lq,M,L,P,2),
            attention_logits (N,Lq,M,L*P), reference_points (N,Lq,L,2|4), padding_mask (N,S) bool or None) -> (N, Lq, M*D)"""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points,
                padding_mask=None):
        reference_points = reference_points.detach().to(torch.float32).contiguous()
        output = fused_forward(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points,
                               padding_mask)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points)
        ctx.padding_mask = padding_mask
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points = ctx.saved_tensors
        grad_value, grad_offsets, grad_logits = fused_backward(value, spatial_shapes, level_start_index, sampling_offsets,
                                                               attention_logits, reference_points, grad_output, ctx.padding_mask)
        return grad_value, None, None, grad_offsets, grad_logits, None, None
