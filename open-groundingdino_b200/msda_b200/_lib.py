"""ctypes binding of libmsda_b200.so -- the only place Python touches the C ABI (include/msda_b200.h).

Loading is strict: a missing or unloadable library raises; nothing here (or anywhere in the package)
substitutes a PyTorch or CPU implementation.
"""
from __future__ import annotations

import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# MSDA_B200_LIB: another build of the same library (A/B measurements of compile-time switches)
LIB_PATH = os.environ.get("MSDA_B200_LIB") or os.path.join(_PKG_DIR, "libmsda_b200.so")

MSDA_F32, MSDA_BF16, MSDA_F64, MSDA_F16 = 0, 1, 2, 3
MSDA_BWD_ATOMIC, MSDA_BWD_DETERMINISTIC = 0, 1
ABI_VERSION = 1
VARIANT_NO_ZERO, VARIANT_NO_NARROW = 0x100, 0x200

# every symbol include/msda_b200.h declares (tests/test_cpu_host.py checks the two lists agree)
EXPORTS = (
    "msda_abi_version", "msda_error_string", "msda_forward", "msda_backward_workspace_bytes", "msda_backward",
    "msda_backward_finish_bf16", "msda_fused_supported", "msda_fused_forward", "msda_fused_backward",
    "msda_fused_mask_supported", "msda_fused_forward_masked", "msda_fused_backward_masked",
    "msda_sample_index", "msda_host_create", "msda_host_destroy", "msda_host_h2d_bytes", "msda_host_d2h_bytes",
    "msda_host_fwd_bwd", "msda_host_submit", "msda_host_wait",
)
# include/hungarian_b200.h (the on-device Hungarian matcher, scope row f2)
HM_EXPORTS = ("hm_max_columns", "hm_cost_matrix", "hm_assign", "hm_match")

_lib = None


class MSDAError(RuntimeError):
    pass


def _declare(lib: ctypes.CDLL) -> None:
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.msda_abi_version.restype = i
    lib.msda_error_string.restype = ctypes.c_char_p
    lib.msda_error_string.argtypes = [i]
    lib.msda_forward.restype = i
    lib.msda_forward.argtypes = [vp] * 6 + [i] * 7 + [i, i, vp]
    lib.msda_backward_workspace_bytes.restype = sz
    lib.msda_backward_workspace_bytes.argtypes = [i] * 7 + [i, i]
    lib.msda_backward.restype = i
    lib.msda_backward.argtypes = [vp] * 9 + [i] * 7 + [i, i, i, vp, sz, vp]
    lib.msda_backward_finish_bf16.restype = i
    lib.msda_backward_finish_bf16.argtypes = [vp, vp, i, i, i, i, i, vp]
    lib.msda_fused_supported.restype = i
    lib.msda_fused_supported.argtypes = [i] * 5
    lib.msda_fused_forward.restype = i
    lib.msda_fused_forward.argtypes = [vp] * 7 + [i] * 7 + [i, i, i, i, vp]
    lib.msda_fused_backward.restype = i
    lib.msda_fused_backward.argtypes = [vp] * 10 + [i] * 7 + [i, i, i, i, vp, sz, vp]
    lib.msda_fused_mask_supported.restype = i
    lib.msda_fused_mask_supported.argtypes = [i, i]
    lib.msda_fused_forward_masked.restype = i
    lib.msda_fused_forward_masked.argtypes = [vp] * 8 + [i] * 7 + [i, i, i, i, vp]
    lib.msda_fused_backward_masked.restype = i
    lib.msda_fused_backward_masked.argtypes = [vp] * 11 + [i] * 7 + [i, i, i, i, vp, sz, vp]
    lib.msda_sample_index.restype = i
    lib.msda_sample_index.argtypes = [vp] * 3 + [i] * 6 + [vp, vp]
    lib.msda_host_create.restype = i
    lib.msda_host_create.argtypes = [ctypes.POINTER(vp)] + [i] * 7 + [i, i]
    lib.msda_host_destroy.restype = None
    lib.msda_host_destroy.argtypes = [vp]
    lib.msda_host_h2d_bytes.restype = sz
    lib.msda_host_h2d_bytes.argtypes = [vp]
    lib.msda_host_d2h_bytes.restype = sz
    lib.msda_host_d2h_bytes.argtypes = [vp]
    lib.msda_host_fwd_bwd.restype = i
    lib.msda_host_fwd_bwd.argtypes = [vp] * 11
    lib.msda_host_submit.restype = i
    lib.msda_host_submit.argtypes = [vp] * 11
    lib.msda_host_wait.restype = i
    lib.msda_host_wait.argtypes = [vp]
    f = ctypes.c_float
    lib.hm_max_columns.restype = i
    lib.hm_max_columns.argtypes = []
    lib.hm_cost_matrix.restype = i
    lib.hm_cost_matrix.argtypes = [vp] * 6 + [i] * 3 + [f] * 4 + [vp]
    lib.hm_assign.restype = i
    lib.hm_assign.argtypes = [vp] * 5 + [i] * 3 + [vp]
    lib.hm_match.restype = i
    lib.hm_match.argtypes = [vp] * 9 + [i] * 4 + [f] * 4 + [vp]


def lib() -> ctypes.CDLL:
    """The loaded library; raises MSDAError if it is missing (build with msda_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MSDAError(
                f"{LIB_PATH} not found: build it with `python open-groundingdino_b200/msda_b200/build.py` "
                "(or __graft_entry__.build()). There is no CPU/PyTorch fallback for this op.")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as exc:  # pragma: no cover - depends on the box
            raise MSDAError(f"cannot load {LIB_PATH}: {exc}") from exc
        _declare(handle)
        if handle.msda_abi_version() != ABI_VERSION:
            raise MSDAError(f"libmsda_b200 ABI {handle.msda_abi_version()} != expected {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MSDAError(f"{what} failed: {lib().msda_error_string(rc).decode()} (code {rc})")
