"""Our kernels vs the REFERENCE's own CUDA op (built from /root/reference by oracle/build_ref.py into
oracle/_ref/, executed here on the same B200, same inputs).  fp32, normwise 1e-5.  Because the pixel
coordinate arithmetic is bit-identical, grad_sampling_loc is compared EVERYWHERE -- including the
default-init distribution whose samples sit on integer coordinates, where the grid_sample oracle is
not a valid judge."""
import numpy as np
import pytest
import torch

from helpers import SWIN_T_800x1333, make_inputs
from oracle import oracle, ref_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reference_build_present():
    """The reference build is the strongest judge this suite has (grad_sampling_loc everywhere, index parity on integer
    coordinates): its absence on a GPU box is a FAILURE, never a skip.  It is built by `python oracle/build_ref.py`
    (or __graft_entry__.build()) where /root/reference is mounted and travels to the box in the snapshot."""
    assert ref_cuda.available(), "oracle/_ref/msda_ref_cuda.so is missing -- run `python oracle/build_ref.py` before gpurun"


def both(g, variant=0):
    from msda_b200 import functions as F
    ref = ref_cuda.load()
    t = lambda k: torch.from_numpy(g[k]).cuda()
    value, shapes, starts, loc, attn, gout = (t(k) for k in ("value", "shapes", "starts", "loc", "attn", "grad_out"))
    ours = [F.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64, variant=variant)]
    ours += F.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout, 64, variant=variant)
    theirs = [ref.ms_deform_attn_forward(value, shapes, starts, loc, attn, 64)]
    theirs += ref.ms_deform_attn_backward(value, shapes, starts, loc, attn, gout.view(theirs[0].shape), 64)
    torch.cuda.synchronize()
    return [x.double().cpu().numpy() for x in ours], [x.double().cpu().numpy() for x in theirs]


@pytest.mark.parametrize("dist", ["U", "W", "I"])
@pytest.mark.parametrize("variant", [1, 2, 3, 3 | 0x400, 3 | 0x800, 3 | 0x2000])
def test_matches_reference_cuda_full_size(dist, variant):
    g = make_inputs(SWIN_T_800x1333, N=1, M=8, D=32, Lq=None, P=4, dist=dist, seed=31)
    ours, theirs = both(g, variant)
    for name, a, b in zip(("output", "grad_value", "grad_sampling_loc", "grad_attn_weight"), ours, theirs):
        err = oracle.normwise(a.reshape(b.shape), b)
        assert err <= 1e-5, f"{name} ({dist}, variant {variant}): {err:.3e}"


@pytest.mark.parametrize("D", [30, 32, 64, 71, 16])
def test_matches_reference_cuda_channel_counts(D):
    """The reference's gradcheck loop covers D = 30, 32, 64, 71 (OPS/test.py:85): one kernel each there."""
    g = make_inputs([(13, 21), (7, 11), (4, 6)], N=2, M=3, D=D, Lq=77, P=4, dist="W", seed=D)
    ours, theirs = both(g)
    for name, a, b in zip(("output", "grad_value", "grad_sampling_loc", "grad_attn_weight"), ours, theirs):
        assert oracle.normwise(a.reshape(b.shape), b) <= 1e-5, name


def test_decoder_shape_matches_reference_cuda():
    g = make_inputs(SWIN_T_800x1333, N=8, M=8, D=32, Lq=900, P=4, dist="I", seed=7)
    ours, theirs = both(g)
    for name, a, b in zip(("output", "grad_value", "grad_sampling_loc", "grad_attn_weight"), ours, theirs):
        assert oracle.normwise(a.reshape(b.shape), b) <= 1e-5, name
