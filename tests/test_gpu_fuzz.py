"""Seeded random shapes through every kernel family against the fp64 oracle: ragged pyramids (1-pixel-wide levels,
single-pixel levels), every level count 1..6, P in {1..8}, D in the vector set and outside it, tiny and odd query
counts, locations far outside the image.  The case list is fixed by the seed, so failures are reproducible."""
import os

import numpy as np
import pytest
import torch

from conftest import generic_coordinate_mask
from helpers import make_inputs
from oracle import oracle

pytestmark = pytest.mark.gpu


def _cases(n=36):
    rng = np.random.default_rng(20261018)
    out = []
    for i in range(n):
        L = int(rng.integers(1, 7))
        shapes = [(int(rng.integers(1, 12)), int(rng.integers(1, 12))) for _ in range(L)]
        D = int(rng.choice([16, 32, 64, 32, 32, 8, 24, 40]))
        P = int(rng.choice([4, 8, 4, 4, 1, 2, 3, 5]))
        M = int(rng.choice([1, 2, 3, 8]))
        N = int(rng.integers(1, 4))
        encoder = bool(rng.integers(0, 2))
        Lq = None if encoder else int(rng.integers(1, 70))
        dist = str(rng.choice(["U", "W", "W"]))
        out.append((i, shapes, N, M, D, Lq, P, dist))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c[0]}-L{len(c[1])}-D{c[4]}-P{c[6]}-{'enc' if c[5] is None else 'q%d' % c[5]}")
def test_random_shapes_all_families(case):
    from msda_b200 import functions as F
    i, shapes, N, M, D, Lq, P, dist = case
    g = make_inputs(shapes, N=N, M=M, D=D, Lq=Lq, P=P, dist=dist, seed=1000 + i)
    if dist == "W":                       # push a few samples far away and onto exact borders
        flat = g["loc"].reshape(-1)
        flat[::17] = -3.0
        flat[5::23] = 1.0
        flat[7::29] = 0.0
    args = (g["value"], g["shapes"], g["starts"], g["loc"], g["attn"])
    truth = (oracle.forward(*args, np.float64),) + oracle.backward(*args, g["grad_out"], np.float64)
    t = lambda k, dt=None: torch.from_numpy(np.ascontiguousarray(g[k])).cuda() if dt is None else torch.from_numpy(np.ascontiguousarray(g[k])).cuda().to(dt)
    m = generic_coordinate_mask(g["loc"], g["shapes"])
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        if dtype == torch.bfloat16:
            gb = dict(g, value=oracle.bf16_round(g["value"]), grad_out=oracle.bf16_round(g["grad_out"]))
            a2 = (gb["value"], g["shapes"], g["starts"], g["loc"], g["attn"])
            tr = (oracle.forward(*a2, np.float64),) + oracle.backward(*a2, gb["grad_out"], np.float64)
        else:
            tr = truth
        value, gout = t("value", dtype), t("grad_out", dtype)
        sh, st, loc, attn = t("shapes"), t("starts"), t("loc"), t("attn")
        for variant in (1, 2, 3, 3 | 0x400, 3 | 0x800, 3 | 0x2000):
            out = F.ms_deform_attn_forward(value, sh, st, loc, attn, 64, variant=variant)
            gv, gl, ga = F.ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64, variant=variant)
            res = [x.double().cpu().numpy() for x in (out, gv, gl, ga)]
            for name, got, want in zip(("output", "grad_value", "grad_attn"), (res[0], res[1], res[3]), (tr[0], tr[1], tr[3])):
                err = oracle.normwise(got.reshape(want.shape), want)
                assert err <= tol, f"{name} dtype={dtype} variant={variant}: {err:.3e}"
            err = oracle.normwise(np.where(m, res[2], 0), np.where(m, tr[2], 0))
            assert err <= max(tol, 2e-5), f"grad_loc dtype={dtype} variant={variant}: {err:.3e}"


def _spec_cases(n=int(os.environ.get("MSDA_FUZZ_SPEC_CASES", "12"))):      # more cases for a one-off soak run
    rng = np.random.default_rng(4242)
    out = []
    for i in range(n):
        shapes = [(int(rng.integers(1, 40 >> l) + 1), int(rng.integers(1, 60 >> l) + 1)) for l in range(4)]
        encoder = bool(rng.integers(0, 2))
        Lq = None if encoder else int(rng.integers(1, 400))
        out.append((i, shapes, int(rng.integers(1, 4)), Lq, str(rng.choice(["U", "W", "I"]))))
    return out


@pytest.mark.parametrize("case", _spec_cases(), ids=lambda c: f"{c[0]}-{'enc' if c[3] is None else 'q%d' % c[3]}-{c[4]}")
def test_random_pyramids_model_geometry(case):
    """The model's own geometry (L = 4, M = 8, D = 32, P = 4) selects the specialised kernels (window aggregation,
    pass-wise tail, staged linear forward): random ragged pyramids, encoder and decoder query sets, all three input
    distributions, both kernel orders, fp32 and bf16, against the fp64 oracle."""
    from msda_b200 import functions as F
    i, shapes, N, Lq, dist = case
    g = make_inputs(shapes, N=N, M=8, D=32, Lq=Lq, P=4, dist=dist, seed=2000 + i)
    if dist == "I":      # fractional chains: jitter the default-init rays
        rng = np.random.default_rng(i)
        wh = np.asarray([(w, h) for h, w in shapes], dtype=np.float32)
        g["loc"] = (g["loc"] + (rng.random(g["loc"].shape, dtype=np.float32) - 0.5) * 0.8 / wh[None, None, None, :, None, :]).astype(np.float32)
    m = generic_coordinate_mask(g["loc"], g["shapes"])
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).cuda() if dt is None else torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dt)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        v, go = (g["value"], g["grad_out"]) if dtype == torch.float32 else (oracle.bf16_round(g["value"]), oracle.bf16_round(g["grad_out"]))
        args = (v, g["shapes"], g["starts"], g["loc"], g["attn"])
        tr = (oracle.forward(*args, np.float64),) + oracle.backward(*args, go, np.float64)
        value, gout = t(v, dtype), t(go, dtype)
        sh, st, loc, attn = t(g["shapes"]), t(g["starts"]), t(g["loc"]), t(g["attn"])
        for variant in (2, 3):
            out = F.ms_deform_attn_forward(value, sh, st, loc, attn, 64, variant=variant)
            gv, gl, ga = F.ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64, variant=variant)
            res = [x.double().cpu().numpy() for x in (out, gv, gl, ga)]
            for name, got, want in zip(("output", "grad_value", "grad_attn"), (res[0], res[1], res[3]), (tr[0], tr[1], tr[3])):
                err = oracle.normwise(got.reshape(want.shape), want)
                assert err <= tol, f"{name} dtype={dtype} variant={variant}: {err:.3e}"
            err = oracle.normwise(np.where(m, res[2], 0), np.where(m, tr[2], 0))
            assert err <= max(tol, 2e-5), f"grad_loc dtype={dtype} variant={variant}: {err:.3e}"
