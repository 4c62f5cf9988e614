"""Pin the CPU oracle (oracle/) to numbers produced by the reference's own function.

Golden files come from tests/golden/make_golden.py, which calls the reference's
ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:41-61) and torch autograd.
Tolerances: fp64 restatement vs fp64 reference 1e-12 normwise; fp32 restatement vs fp64 truth
1e-5 normwise (the north-star fp32 tolerance); same for the reference's own fp32 run.
"""
import numpy as np
import pytest
import torch

from conftest import generic_coordinate_mask
from oracle import oracle, torch_port


def _run(g, dtype):
    out = oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], dtype)
    gv, gl, ga = oracle.backward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], g["grad_out"], dtype)
    return out, gv, gl, ga


def test_fp64_restatement_matches_reference_fp64(golden):
    out, gv, gl, ga = _run(golden, np.float64)
    assert oracle.normwise(out, golden["out_f64"]) < 1e-12
    assert oracle.normwise(gv, golden["gv_f64"]) < 1e-12
    assert oracle.normwise(ga, golden["ga_f64"]) < 1e-12
    m = generic_coordinate_mask(golden["loc"], golden["shapes"], tol=1e-9)
    assert oracle.normwise(np.where(m, gl, 0), np.where(m, golden["gl_f64"], 0)) < 1e-10


def test_fp32_restatement_within_north_star_tolerance(golden):
    out, gv, gl, ga = _run(golden, np.float32)
    tol = 1e-5
    assert oracle.normwise(out, golden["out_f64"]) < tol
    assert oracle.normwise(gv, golden["gv_f64"]) < tol
    assert oracle.normwise(ga, golden["ga_f64"]) < tol
    m = generic_coordinate_mask(golden["loc"], golden["shapes"])
    assert oracle.normwise(np.where(m, gl, 0), np.where(m, golden["gl_f64"], 0)) < 5 * tol
    # and it is as close to truth as the reference's own fp32 run is
    assert oracle.normwise(out, golden["out_f32"]) < 2 * tol


def test_torch_port_matches_reference(golden):
    t = lambda k, dt: torch.from_numpy(golden[k]).to(dt)
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 1e-6)):
        out, gv, gl, ga = torch_port.fwd_bwd(t("value", dt), golden["shapes"].tolist(), t("loc", dt),
                                             t("attn", dt), t("grad_out", dt))
        assert oracle.normwise(out.numpy(), golden[f"out_{tag}"]) <= tol
        assert oracle.normwise(gv.numpy(), golden[f"gv_{tag}"]) <= tol
        assert oracle.normwise(gl.numpy(), golden[f"gl_{tag}"]) <= tol
        assert oracle.normwise(ga.numpy(), golden[f"ga_{tag}"]) <= tol


def test_reference_testpy_tolerances():
    """The reference's own acceptance check (OPS/test.py:40,56): allclose rtol 1e-5/atol 1e-8 in
    fp64 and rtol 1e-2/atol 1e-3 in fp32, on its N,M,D,Lq,L,P = 1,2,2,2,2,2 fixture."""
    from conftest import load_golden
    g = load_golden("testpy_D2")
    out64 = oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], np.float64)
    out32 = oracle.forward(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], np.float32)
    assert np.allclose(out64, g["out_f64"], rtol=1e-5, atol=1e-8)
    assert np.allclose(out32, g["out_f32"], rtol=1e-2, atol=1e-3)


def test_fp32_coordinate_is_single_rounded():
    """The reference AS COMPILED contracts `loc * size - 0.5` (cuh:285-286) into one FFMA
    (profiles/reference_sass_coordinate.txt), so the oracle's integer records must equal a float64
    evaluation rounded once -- and must NOT equal the unfused multiply-then-subtract, which differs in
    a measurable share of coordinates and flips floor() for some samples on the pixel lattice."""
    rng = np.random.default_rng(0)
    n = 100000
    loc = (rng.random((1, n, 1, 1, 1, 2), dtype=np.float32) * 1.5 - 0.25).astype(np.float32)
    loc[0, :1000, 0, 0, 0, 1] = (np.arange(1000, dtype=np.float32) % 100 + 0.5) / np.float32(100)   # pixel centres
    shapes = np.array([[100, 167]], dtype=np.int64)
    rec = oracle.sample_index(shapes, np.zeros(1, np.int64), loc, 100 * 167)[0, :, 0, 0, 0]
    h = (loc[0, :, 0, 0, 0, 1].astype(np.float64) * 100 - 0.5).astype(np.float32)     # one rounding
    w = (loc[0, :, 0, 0, 0, 0].astype(np.float64) * 167 - 0.5).astype(np.float32)
    inside = (h > -1) & (w > -1) & (h < 100) & (w < 167)
    assert np.array_equal(rec[:, 2] & 1, inside.astype(np.int32))
    assert np.array_equal(rec[inside, 0], np.floor(h[inside]).astype(np.int32))
    assert np.array_equal(rec[inside, 1], np.floor(w[inside]).astype(np.int32))
    unfused = (loc[0, :, 0, 0, 0, 0] * np.float32(167)).astype(np.float32) - np.float32(0.5)
    assert (unfused != w).mean() > 0.005


def test_sample_index_records():
    from conftest import load_golden
    g = load_golden("pyr4_wide_D32")
    S = g["value"].shape[1]
    rec = oracle.sample_index(g["shapes"], g["starts"], g["loc"], S)
    N, Lq, M, L, P, _ = g["loc"].shape
    H = g["shapes"][:, 0][None, None, None, :, None]
    W = g["shapes"][:, 1][None, None, None, :, None]
    inside = (rec[..., 2] & 1).astype(bool)
    assert inside.any() and (~inside).any()
    assert (rec[..., 0][inside] >= -1).all() and (rec[..., 0] < H)[inside].all()
    assert (rec[..., 1][inside] >= -1).all() and (rec[..., 1] < W)[inside].all()
    b = np.arange(N)[:, None, None, None, None]
    assert np.array_equal(rec[..., 3], np.broadcast_to(b * S + g["starts"][None, None, None, :, None], rec[..., 3].shape))


def test_bf16_round_matches_torch():
    x = np.random.default_rng(1).standard_normal(10000).astype(np.float32)
    assert np.array_equal(oracle.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).float().numpy())
