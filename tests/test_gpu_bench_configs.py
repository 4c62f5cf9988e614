"""Parity at EXACTLY the configurations bench.py measures (BASELINE.json configs[1..3]), through the default
dispatch (variant 0), against two independent judges:

  * the fp64 CPU oracle (oracle/msda_oracle.c) on bf16-rounded inputs, normwise <= 1e-2 (north_star's bf16 bound);
    grad_sampling_loc where it is continuous (samples farther than 1e-4 px from an integer coordinate);
  * the REFERENCE's own CUDA op (oracle/_ref) wrapped the way the reference wraps 16-bit values
    (ops/modules/ms_deform_attn.py:114-120: widen value to fp32, run the fp32 op, narrow the output) -- same inputs,
    same GPU; grad_sampling_loc compared EVERYWHERE because the pixel coordinates are bit-identical.

Configurations: cfg-2 (SwinT 800x1333 encoder, bs 4, bf16 value, 22,223 queries) on distributions I and U;
cfg-4 shapes (SwinB 1024x1024: 128^2..16^2) encoder bs 4 and decoder (900 queries) bs 4, bf16 value;
cfg-3 (decoder, bs 8, fp32 and bf16).  Nothing here may skip on a GPU box: a missing reference build is a failure.
"""
import numpy as np
import pytest
import torch

from conftest import generic_coordinate_mask
from helpers import SWIN_B_1024, SWIN_T_800x1333, make_inputs
from oracle import oracle, ref_cuda

pytestmark = pytest.mark.gpu

NAMES = ("output", "grad_value", "grad_sampling_loc", "grad_attn_weight")


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _ours(g, dtype, variant=0):
    from msda_b200 import functions as F
    value, gout = _dev(g["value"], dtype), _dev(g["grad_out"], dtype)
    shapes, starts, loc, attn = _dev(g["shapes"]), _dev(g["starts"]), _dev(g["loc"]), _dev(g["attn"])
    out = F.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64, variant=variant)
    gv, gl, ga = F.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout, 64, variant=variant)
    torch.cuda.synchronize()
    assert out.dtype == dtype and gv.dtype == dtype and gl.dtype == torch.float32 and ga.dtype == torch.float32
    return [t.double().cpu().numpy() for t in (out, gv, gl, ga)]


def _reference_op_wrapped(g, dtype):
    """The reference CUDA op around 16-bit tensors, as the reference module does for fp16 (:114-120)."""
    assert ref_cuda.available(), ("oracle/_ref/msda_ref_cuda.so is missing: build it with `python oracle/build_ref.py` "
                                  "where /root/reference is mounted -- it must travel to the GPU box")
    ref = ref_cuda.load()
    value, gout = _dev(g["value"], dtype).float(), _dev(g["grad_out"], dtype).float()
    shapes, starts, loc, attn = _dev(g["shapes"]), _dev(g["starts"]), _dev(g["loc"]), _dev(g["attn"])
    out = ref.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64)
    gv, gl, ga = ref.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout.view(out.shape), 64)
    if dtype != torch.float32:
        out, gv = out.to(dtype), gv.to(dtype)
    torch.cuda.synchronize()
    return [t.double().cpu().numpy() for t in (out, gv, gl, ga)]


def _round_inputs(g, dtype):
    if dtype == torch.bfloat16:
        g["value"] = oracle.bf16_round(g["value"])
        g["grad_out"] = oracle.bf16_round(g["grad_out"])
    return g


def _check(g, dtype, tol, small_tol, small_vs_oracle=4e-5):
    res = _ours(g, dtype)
    # judge 1: fp64 oracle
    args = (g["value"], g["shapes"], g["starts"], g["loc"], g["attn"])
    truth = (oracle.forward(*args, np.float64),) + oracle.backward(*args, g["grad_out"], np.float64)
    m = generic_coordinate_mask(g["loc"], g["shapes"], 1e-4)
    for name, got, want in zip(NAMES, res, truth):
        got = got.reshape(want.shape)
        if name == "grad_sampling_loc":
            got, want = np.where(m, got, 0), np.where(m, want, 0)
        # fp32 pixel coordinates at W = 167 alone cost ~1e-5 against fp64 truth (profiles/r01_parity_notes.md)
        bound = tol if name in ("output", "grad_value") else small_vs_oracle
        err = oracle.normwise(got, want)
        assert err <= bound, f"{name} vs fp64 oracle: {err:.3e} > {bound:.1e}"
    # judge 2: the reference CUDA op, everywhere
    theirs = _reference_op_wrapped(g, dtype)
    for name, got, want in zip(NAMES, res, theirs):
        bound = tol if name in ("output", "grad_value") else small_tol
        err = oracle.normwise(got.reshape(want.shape), want)
        assert err <= bound, f"{name} vs reference CUDA op: {err:.3e} > {bound:.1e}"


@pytest.mark.parametrize("dist", ["I", "U"])
def test_cfg2_bf16_bs4_full_size(dist):
    """BASELINE.json configs[1] / bench.py default workload: the headline configuration itself."""
    g = _round_inputs(make_inputs(SWIN_T_800x1333, N=4, M=8, D=32, Lq=None, P=4, dist=dist, seed=41), torch.bfloat16)
    # the two small gradients are fp32 sums over bf16 products: far tighter than the bf16 bound
    _check(g, torch.bfloat16, 1e-2, 2e-5)


@pytest.mark.parametrize("Lq,dist", [(None, "I"), (None, "U"), (900, "I"), (900, "W")])
def test_cfg4_swinb_shapes_bf16_bs4(Lq, dist):
    """configs[3]'s op shapes: SwinB 1024x1024 -> 128^2, 64^2, 32^2, 16^2; encoder and 900-query decoder."""
    g = _round_inputs(make_inputs(SWIN_B_1024, N=4, M=8, D=32, Lq=Lq, P=4, dist=dist, seed=43), torch.bfloat16)
    _check(g, torch.bfloat16, 1e-2, 2e-5)


@pytest.mark.parametrize("dtype,tol,small", [(torch.float32, 1e-5, 1e-5), (torch.bfloat16, 1e-2, 2e-5)])
def test_cfg3_decoder_bs8(dtype, tol, small):
    """configs[2]: 900 queries over the 4-level 800x1333 memory, bs 8."""
    g = _round_inputs(make_inputs(SWIN_T_800x1333, N=8, M=8, D=32, Lq=900, P=4, dist="I", seed=47), dtype)
    if dtype == torch.float32:
        # fp32 vs fp64 truth at W = 167 costs ~1e-5 by itself (profiles/r01_parity_notes.md): judge fp32 by the fp32 oracle
        res = _ours(g, dtype)
        args = (g["value"], g["shapes"], g["starts"], g["loc"], g["attn"])
        o32 = (oracle.forward(*args, np.float32),) + oracle.backward(*args, g["grad_out"], np.float32)
        theirs = _reference_op_wrapped(g, dtype)
        for name, got, a, b in zip(NAMES, res, o32, theirs):
            assert oracle.normwise(got.reshape(b.shape), b) <= tol, name + " vs reference CUDA op"
            if name != "grad_sampling_loc":      # on distribution I the oracle is no judge of grad_loc (integer coordinates)
                assert oracle.normwise(got.reshape(a.shape), a) <= tol, name + " vs fp32 oracle"
    else:
        _check(g, dtype, tol, small)


@pytest.mark.parametrize("variant", [3 | 0x800, 3 | 0x2000])
@pytest.mark.parametrize("dist", ["I", "U"])
def test_cfg2_tile_backward_bf16_bs4(dist, variant):
    """The alternative backward kernels of the query-tile order at the headline size -- MSDA_VARIANT_TILE_BWD (counting sort +
    row-stationary pass: every window level, directly handled samples (U: nearly all of them) and the sort capacity limits
    are exercised only at full size) and MSDA_VARIANT_ROWGROUP_BWD (the row-group kernel; the default is the warp-scoped one)."""
    g = _round_inputs(make_inputs(SWIN_T_800x1333, N=4, M=8, D=32, Lq=None, P=4, dist=dist, seed=59), torch.bfloat16)
    res = _ours(g, torch.bfloat16, variant=variant)
    theirs = _reference_op_wrapped(g, torch.bfloat16)
    for name, got, want in zip(NAMES, res, theirs):
        bound = 1e-2 if name in ("output", "grad_value") else 2e-5
        err = oracle.normwise(got.reshape(want.shape), want)
        assert err <= bound, f"{name} vs reference CUDA op: {err:.3e} > {bound:.1e}"


def test_cfg1_fp32_bs1_every_family_vs_reference_op():
    """configs[0] shape on the GPU (fp32, bs 1), default dispatch, distribution I: all four results against the reference op."""
    g = make_inputs(SWIN_T_800x1333, N=1, M=8, D=32, Lq=None, P=4, dist="I", seed=53)
    res = _ours(g, torch.float32)
    theirs = _reference_op_wrapped(g, torch.float32)
    for name, got, want in zip(NAMES, res, theirs):
        assert oracle.normwise(got.reshape(want.shape), want) <= 1e-5, name
