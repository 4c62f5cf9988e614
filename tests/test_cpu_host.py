"""CPU-side checks (no compute calls): the C-ABI library loads and exports every symbol the public
header declares, the Python mirror keeps the reference's names / init / error behaviour."""
import ctypes
import math
import os
import re

import pytest
import torch

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from msda_b200 import _lib
    names = _header_symbols()
    assert sorted(_lib.EXPORTS) == names
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    assert _lib.lib().msda_abi_version() == _lib.ABI_VERSION
    assert b"workspace" in _lib.lib().msda_error_string(-3)


def test_matcher_header_symbols_exported_and_validated():
    """include/hungarian_b200.h: every declared entry point is exported; rejected calls return before any CUDA work."""
    from msda_b200 import _lib
    text = open(os.path.join(ROOT, "include", "hungarian_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(hm_[a-z0-9_]+)\s*\(", text)))
    assert names == sorted(_lib.HM_EXPORTS)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    lib = _lib.lib()
    assert lib.hm_max_columns() >= 900
    assert lib.hm_cost_matrix(None, None, None, None, None, None, 1, 900, 256, 1.0, 1.0, 1.0, 0.25, None) == -1
    assert lib.hm_assign(None, None, None, None, None, 1, 900, 10, None) == -1
    assert lib.hm_assign(None, None, None, None, None, 0, 900, 10, None) == 0
    assert lib.hm_cost_matrix(None, None, None, None, None, None, 1, 900, 4096, 1.0, 1.0, 1.0, 0.25, None) in (-1, -2)
    import matcher_b200
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        matcher_b200.HungarianMatcher()({"pred_logits": torch.zeros(1, 4, 8), "pred_boxes": torch.zeros(1, 4, 4)},
                                        [{"labels": torch.zeros(1, dtype=torch.int64), "boxes": torch.zeros(1, 4)}],
                                        torch.ones(1, 8))


def test_argument_validation_without_a_gpu():
    """Rejected calls return negative codes before any CUDA work is attempted."""
    from msda_b200 import _lib
    lib = _lib.lib()
    assert lib.msda_forward(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 0, 0, None) == -1
    assert lib.msda_backward_workspace_bytes(1, 10, 2, 32, 1, 5, 4, _lib.MSDA_F32, 0) == 0
    assert lib.msda_backward_workspace_bytes(1, 10, 2, 32, 1, 5, 4, _lib.MSDA_BF16, 0) == 10 * 2 * 32 * 4
    assert lib.msda_backward_workspace_bytes(0, 10, 2, 32, 1, 5, 4, _lib.MSDA_F32, 0) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from msda_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmsda_b200.so"))
    with pytest.raises(_lib.MSDAError, match="no CPU/PyTorch fallback"):
        _lib.lib()


def test_cpu_tensors_raise_like_the_reference():
    import MultiScaleDeformableAttention as MSDA
    from msda_b200 import MSDeformAttnFunction
    v = torch.zeros(1, 30, 2, 2)
    sh, st = torch.tensor([[6, 4], [3, 2]]), torch.tensor([0, 24])
    loc, at = torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2)
    for fn in (lambda: MSDA.ms_deform_attn_forward(v, sh, st, loc, at, 2),
               lambda: MSDA.ms_deform_attn_backward(v, sh, st, loc, at, torch.zeros(1, 2, 4), 2),
               lambda: MSDeformAttnFunction.apply(v, sh, st, loc, at, 2)):
        with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
            fn()


def test_module_names_shapes_and_init():
    from msda_b200 import MSDeformAttn, MultiScaleDeformableAttention
    m = MSDeformAttn(256, 4, 8, 4)
    assert sorted(m.state_dict()) == sorted(
        f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj") for p in ("weight", "bias"))
    assert m.sampling_offsets.weight.shape == (256, 256) and m.attention_weights.weight.shape == (128, 256)
    assert m.im2col_step == 64
    assert not m.sampling_offsets.weight.any() and not m.attention_weights.weight.any() and not m.attention_weights.bias.any()
    bias = m.sampling_offsets.bias.detach().view(8, 4, 4, 2)
    for h in range(8):
        d = torch.tensor([math.cos(h * math.pi / 4), math.sin(h * math.pi / 4)])
        d = d / d.abs().max()
        for p in range(4):
            assert torch.allclose(bias[h, :, p], (d * (p + 1)).expand(4, 2), atol=1e-6)
    assert not m.value_proj.bias.any() and not m.output_proj.bias.any()
    lim = math.sqrt(6.0 / 512)
    assert m.value_proj.weight.abs().max() <= lim and m.value_proj.weight.std() > 0.5 * lim / math.sqrt(3)
    with pytest.raises(ValueError):
        MSDeformAttn(250, 4, 8, 4)
    with pytest.warns(UserWarning):
        MSDeformAttn(240, 4, 8, 4)
    a = MultiScaleDeformableAttention(embed_dims=128, num_levels=3, num_heads=4, num_points=2, batch_first=True)
    b = MultiScaleDeformableAttention(embed_dim=128, num_levels=3, num_heads=4, num_points=2, img2col_step=32)
    assert a.d_model == b.d_model == 128 and b.im2col_step == 32 and a.batch_first and not b.batch_first
    a.init_weights()


def test_sampling_locations_forms():
    from msda_b200.modules import sampling_locations_from
    sh = torch.tensor([[4, 6], [2, 3]])
    off = torch.randn(1, 5, 2, 2, 3, 2)
    ref2 = torch.rand(1, 5, 2, 2)
    loc = sampling_locations_from(ref2, off, sh, 3)
    assert torch.allclose(loc[0, 1, 1, 1, 2], ref2[0, 1, 1] + off[0, 1, 1, 1, 2] / torch.tensor([3.0, 2.0]))
    ref4 = torch.rand(1, 5, 2, 4)
    loc = sampling_locations_from(ref4, off, sh, 3)
    assert torch.allclose(loc[0, 2, 0, 1, 0], ref4[0, 2, 1, :2] + off[0, 2, 0, 1, 0] / 3 * ref4[0, 2, 1, 2:] * 0.5)
    with pytest.raises(ValueError):
        sampling_locations_from(torch.rand(1, 5, 2, 3), off, sh, 3)


def test_reference_arm_prints_one_schema_complete_json_line():
    """`bench.py --impl reference` is the CPU arm the driver runs; it must print exactly one JSON line with the
    contract's keys (the GPU arm's line is checked on the box)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/models/GroundingDINO/ops"), reason="reference tree not mounted")
def test_reference_code_imports_against_the_shim():
    """INTEGRATION.md route 1: with this repo's package directory on sys.path the reference's own
    `import MultiScaleDeformableAttention as MSDA` (ops/functions/ms_deform_attn_func.py:18) resolves to the stand-in
    module, and its unmodified MSDeformAttnFunction / MSDeformAttn run through it (on a CPU tensor that means the
    reference's own 'Not implemented on the CPU' error, raised by our library binding)."""
    import importlib
    import sys
    import types
    sys.modules.pop("MultiScaleDeformableAttention", None)
    ops = "/root/reference/models/GroundingDINO/ops"
    pkg = types.ModuleType("refops")          # the ops directory as a package, without importing models/__init__
    pkg.__path__ = [ops]
    sys.modules["refops"] = pkg
    try:
        func = importlib.import_module("refops.functions.ms_deform_attn_func")
        import MultiScaleDeformableAttention as shim
        assert func.MSDA is shim and shim.__file__.startswith(os.path.join(ROOT, "open-groundingdino_b200"))
        mod = importlib.import_module("refops.modules.ms_deform_attn").MSDeformAttn(64, 2, 2, 2)
        sh = torch.tensor([[6, 4], [3, 2]])
        st = torch.tensor([0, 24])
        with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
            mod(torch.zeros(1, 5, 64), torch.rand(1, 5, 2, 2), torch.zeros(1, 30, 64), sh, st)
        # and the reference module's parameters load into ours unchanged (same names, same shapes)
        from msda_b200 import MSDeformAttn
        ours = MSDeformAttn(64, 2, 2, 2)
        ours.load_state_dict(mod.state_dict(), strict=True)
        assert torch.equal(ours.sampling_offsets.bias, mod.sampling_offsets.bias)
    finally:
        for name in [k for k in sys.modules if k == "refops" or k.startswith("refops.")]:
            sys.modules.pop(name)
