"""ctypes/numpy front-end of the CPU oracle (oracle/msda_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (open-groundingdino_b200/) never
imports this module; its ops raise when the CUDA library is missing instead of falling back here.

Parity status: pinned against golden vectors produced by the reference's own
``ms_deform_attn_core_pytorch`` (models/GroundingDINO/ops/functions/ms_deform_attn_func.py:41-61),
see tests/golden/make_golden.py and tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/msda_oracle.c with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "msda_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmsda_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    _, Lq, M2, L, P, two = loc.shape
    assert M2 == M and two == 2 and shapes.shape == (L, 2)
    return N, S, M, D, L, Lq, P


def _prep(value, shapes, starts, loc, attn, dtype):
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    attn = np.ascontiguousarray(attn, dtype=dtype)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    return value, shapes, starts, loc, attn


def level_starts(shapes) -> np.ndarray:
    """level_start_index as the callers build it (models/GroundingDINO/transformer.py:245-250)."""
    shapes = np.asarray(shapes, dtype=np.int64)
    areas = shapes[:, 0] * shapes[:, 1]
    return np.concatenate([[0], np.cumsum(areas)[:-1]]).astype(np.int64)


def forward(value, shapes, starts, loc, attn, dtype=np.float32) -> np.ndarray:
    """out (N, Lq, M*D) in ``dtype`` (np.float32 follows the fp32 CUDA arithmetic, np.float64 is truth)."""
    value, shapes, starts, loc, attn = _prep(value, shapes, starts, loc, attn, dtype)
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib().msda_oracle_forward_f32 if dtype == np.float32 else lib().msda_oracle_forward_f64
    rc = fn(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn),
            N, S, M, D, L, Lq, P, _ptr(out))
    assert rc == 0
    return out


def backward(value, shapes, starts, loc, attn, grad_out, dtype=np.float32):
    """(grad_value, grad_sampling_loc, grad_attn_weight) in ``dtype``."""
    value, shapes, starts, loc, attn = _prep(value, shapes, starts, loc, attn, dtype)
    grad_out = np.ascontiguousarray(grad_out, dtype=dtype)
    N, S, M, D, L, Lq, P = _dims(value, shapes, loc)
    assert grad_out.size == N * Lq * M * D
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    fn = lib().msda_oracle_backward_f32 if dtype == np.float32 else lib().msda_oracle_backward_f64
    rc = fn(_ptr(value), _ptr(shapes), _ptr(starts), _ptr(loc), _ptr(attn), _ptr(grad_out),
            N, S, M, D, L, Lq, P, _ptr(gv), _ptr(gl), _ptr(ga))
    assert rc == 0
    return gv, gl, ga


def sample_index(shapes, starts, loc, S: int) -> np.ndarray:
    """int32 (N,Lq,M,L,P,4) records [h_low, w_low, flags, slab_row] of the fp32 path."""
    loc = np.ascontiguousarray(loc, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    N, Lq, M, L, P, _ = loc.shape
    rec = np.empty((N, Lq, M, L, P, 4), dtype=np.int32)
    rc = lib().msda_oracle_sample_index_f32(_ptr(shapes), _ptr(starts), _ptr(loc),
                                            N, S, M, L, Lq, P, _ptr(rec))
    assert rc == 0
    return rec


def num_threads() -> int:
    return int(lib().msda_oracle_num_threads())


# ---- bf16 helpers (numpy has no bf16: carry it as uint16 bit patterns) ------------------------

def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bf16 (ties to even) and return it widened back to fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32).reshape(np.shape(x))
    if nan.any():
        out = np.where(nan, np.float32(np.nan), out)
    return out


def normwise(a, b) -> float:
    """The parity metric of SURVEY.md section 8(d): max|a-b| / max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    if denom == 0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / denom)
