"""The "bit-exact sampling index" contract is defined against the reference op AS COMPILED: nvcc 12.9 -O3 contracts
`loc * size - 0.5` (ms_deform_im2col_cuda.cuh:285-286) into one FFMA with the -0.5 addend in every kernel, and this
repo's locate() (csrc/msda_common.cuh) therefore uses fmaf(loc, size, -0.5f).  A different toolchain could un-contract
the reference, silently invalidating the contract; this test disassembles oracle/_ref/msda_ref_cuda.so and fails if
the fused form is gone (or an unfused FMUL-then-FADD(-0.5) form appears) in any of the reference kernels.  CPU-only:
needs cuobjdump and the reference build, both present in the authoring container."""
import os
import re
import shutil
import subprocess

import pytest

from oracle import ref_cuda

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass():
    if not ref_cuda.available():
        if os.path.isdir("/root/reference/models/GroundingDINO/ops/src"):
            from oracle import build_ref
            build_ref.build()
        else:
            pytest.skip("no reference sources and no prebuilt oracle/_ref here")
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not installed")
    so = os.path.join(os.path.dirname(ref_cuda.__file__), "_ref", "msda_ref_cuda.so")
    return subprocess.run([CUOBJDUMP, "-sass", so], check=True, capture_output=True, text=True).stdout


def test_reference_coordinate_is_a_fused_multiply_add():
    text = _sass()
    kernels = re.split(r"\n\s*Function : ", text)[1:]
    # fp32 instantiations: template argument list starts with `f` (IfE... / IfLj32EE...)
    fp32 = [k for k in kernels if re.search(r"ms_deformable_\w*?kernel\w*?If(E|Lj)", k.split("\n", 1)[0])]
    assert len(fp32) >= 16, f"expected the forward + 15 backward fp32 kernels, found {len(fp32)}"
    for k in fp32:
        name = k.split("\n", 1)[0]
        fused = re.findall(r"FFMA[.\w]* R\d+, R\d+, R\d+, -0\.5 ?;", k)
        unfused = re.findall(r"FADD[.\w]* R\d+, R\d+, -0\.5 ?;", k)
        assert len(fused) >= 2, f"{name}: no `FFMA R, loc, size, -0.5` pair -- the reference coordinate is no longer fused"
        assert not unfused, f"{name}: unfused `FADD R, R, -0.5` coordinate form present"
