"""Seeded random shapes through every kernel family against the fp64 oracle: ragged pyramids (1-pixel-wide levels,
single-pixel levels), every level count 1..6, P in {1..8}, D in the vector set and outside it, tiny and odd query
counts, locations far outside the image.  The case list is fixed by the seed, so failures are reproducible."""
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
        for variant in (1, 2, 3):
            out = F.ms_deform_attn_forward(value, sh, st, loc, attn, 64, variant=variant)
            gv, gl, ga = F.ms_deform_attn_backward(value, sh, st, loc, attn, gout, 64, variant=variant)
            res = [x.double().cpu().numpy() for x in (out, gv, gl, ga)]
            for name, got, want in zip(("output", "grad_value", "grad_attn"), (res[0], res[1], res[3]), (tr[0], tr[1], tr[3])):
                err = oracle.normwise(got.reshape(want.shape), want)
                assert err <= tol, f"{name} dtype={dtype} variant={variant}: {err:.3e}"
            err = oracle.normwise(np.where(m, res[2], 0), np.where(m, tr[2], 0))
            assert err <= max(tol, 2e-5), f"grad_loc dtype={dtype} variant={variant}: {err:.3e}"
