"""Sampling-index and level-offset computation must be BIT-EXACT (north_star): the device's
(h_low, w_low, accept flag, neighbour mask, slab row) per sample equals the oracle's restatement of
ms_deform_im2col_cuda.cuh:285-288,38-41,274-278, including the default-init distribution whose
coordinates sit on exact integers."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import SWIN_B_1024, SWIN_T_800x1333, make_inputs
from oracle import oracle

pytestmark = pytest.mark.gpu


def device_records(g):
    from msda_b200 import _lib
    N, S, M, D, L, Lq, P = g["dims"]
    shapes, starts = torch.from_numpy(g["shapes"]).cuda(), torch.from_numpy(g["starts"]).cuda()
    loc = torch.from_numpy(g["loc"]).cuda()
    rec = torch.empty((N, Lq, M, L, P, 4), dtype=torch.int32, device="cuda")
    rc = _lib.lib().msda_sample_index(shapes.data_ptr(), starts.data_ptr(), loc.data_ptr(), N, S, M, L, Lq, P,
                                      rec.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "msda_sample_index")
    return rec.cpu().numpy()


@pytest.mark.parametrize("shapes", [SWIN_T_800x1333, SWIN_B_1024])
@pytest.mark.parametrize("dist", ["U", "W", "I"])
def test_index_records_bit_exact(shapes, dist):
    g = make_inputs(shapes, N=1, M=8, D=1, Lq=None, P=4, dist=dist, seed=17)
    want = oracle.sample_index(g["shapes"], g["starts"], g["loc"], g["dims"][1])
    got = device_records(g)
    assert np.array_equal(got, want)
    if dist == "I":
        # the hazard is real: some nominally-integer coordinates land one ulp low (floor = k-1)
        pix = g["loc"][..., 0].astype(np.float64) * np.asarray([w for _, w in shapes])[None, None, None, :, None] - 0.5
        assert (np.abs(pix - np.round(pix)) < 1e-4).mean() > 0.1


def test_index_special_values():
    g = make_inputs([(7, 5), (3, 2)], N=1, M=2, D=1, Lq=64, P=4, dist="W", seed=1)
    loc = g["loc"]
    specials = np.array([0.0, 1.0, -0.0, 0.5, 1e-30, -1e-30, 1.0 - 2 ** -24, np.nan, np.inf, -np.inf, 1e30],
                        dtype=np.float32)
    flat = loc.reshape(-1, 2)
    flat[: len(specials), 0] = specials
    flat[len(specials): 2 * len(specials), 1] = specials
    want = oracle.sample_index(g["shapes"], g["starts"], g["loc"], g["dims"][1])
    got = device_records(g)
    assert np.array_equal(got, want)
