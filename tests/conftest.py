"""Shared test plumbing: marker registration, import paths, golden loader."""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "open-groundingdino_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(params=golden_names())
def golden(request):
    g = load_golden(request.param)
    g["name"] = request.param
    return g


def generic_coordinate_mask(loc, shapes, tol=1e-4):
    """True where a sample is farther than ``tol`` pixels from any integer pixel coordinate, i.e.
    where grad_sampling_loc is continuous (integer-coordinate hazard)."""
    loc = np.asarray(loc, dtype=np.float64)
    shapes = np.asarray(shapes)
    wh = np.stack([shapes[:, 1], shapes[:, 0]], -1).astype(np.float64)      # (L,2) as (W,H)
    pix = loc * wh[None, None, None, :, None, :] - 0.5
    dist = np.abs(pix - np.round(pix))
    ok = (dist > tol).all(-1)
    return np.broadcast_to(ok[..., None], loc.shape)
