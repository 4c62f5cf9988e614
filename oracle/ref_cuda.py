"""Loader for the reference's own CUDA op built by oracle/build_ref.py (TEST/BENCH infrastructure only).

``load()`` returns a module object exposing ``ms_deform_attn_forward`` / ``ms_deform_attn_backward`` with
the reference's pybind signatures (ops/src/vision.cpp:13-16), or None when oracle/_ref/msda_ref_cuda.so
was not built (then the tests that need it skip and bench.py omits the comparison)."""
import importlib.machinery
import importlib.util
import os

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "msda_ref_cuda.so")
_mod = None


def available() -> bool:
    return os.path.exists(_SO)


def load():
    global _mod
    if _mod is None and available():
        import torch  # noqa: F401  (the extension links against libtorch)
        loader = importlib.machinery.ExtensionFileLoader("msda_ref_cuda", _SO)
        spec = importlib.util.spec_from_loader("msda_ref_cuda", loader)
        _mod = importlib.util.module_from_spec(spec)
        loader.exec_module(_mod)
    return _mod
