"""Library hygiene, checked on the compiled code (CPU-only: needs cuobjdump and the in-tree libmsda_b200.so).

The performance notes of DESIGN.md rest on properties of the SASS that a toolchain change could silently undo: the default
kernels of the benchmarked shape must not spill, must gather with predicated 16-byte read-only loads, accumulate with packed
FFMA2, reduce with 16-byte vector reductions, use the fused-multiply-add coordinate of the bit-exact index contract, and
form their gather / reduction addresses with one IMAD.WIDE (the fenced lane pointers of msda_common.cuh as_given())."""
import os
import re
import shutil
import subprocess

import pytest

from msda_b200 import build as lib_build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def kernels():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not installed")
    so = lib_build.LIB_PATH
    assert os.path.exists(so), "libmsda_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    text = subprocess.run([CUOBJDUMP, "-sass", so], check=True, capture_output=True, text=True).stdout
    out = {}
    for k in re.split(r"\n\s*Function : ", text)[1:]:
        name, body = k.split("\n", 1)
        out[name.strip()] = body
    return out


def _one(kernels, pattern):
    hits = [n for n in kernels if re.search(pattern, n)]
    assert len(hits) == 1, f"{pattern}: {hits}"
    return kernels[hits[0]]


# mangled-name patterns of the kernels the cfg-2 benchmark launches (bf16 rows, op boundary) and of the decoder's
DEFAULTS = {
    "fwd_tile": r"fwd_tile_kernelI13__nv_bfloat16Li0ELi4EE",
    "bwd_warp": r"bwd_warp_kernelI13__nv_bfloat16Li0ELi3EE",
    "bwd_rowgroup": r"bwd_vec_kernelI13__nv_bfloat16Li32ELi4ELi2ELi0ELi0ELb1ELb0EE",
}


@pytest.mark.parametrize("which", sorted(DEFAULTS))
def test_default_kernels_do_not_spill(kernels, which):
    body = _one(kernels, DEFAULTS[which])
    local = len(re.findall(r"\b(LDL|STL)\b", body))
    assert local <= 8, f"{which}: {local} local-memory instructions (register spills)"


@pytest.mark.parametrize("which", sorted(DEFAULTS))
def test_default_kernels_use_the_intended_instructions(kernels, which):
    body = _one(kernels, DEFAULTS[which])
    assert re.search(r"LDG\.E\.128\.CONSTANT", body), "16-byte read-only gathers"
    assert len(re.findall(r"\bFFMA2\b", body)) >= 16, "packed fp32x2 accumulation"
    assert re.search(r"FFMA[.\w]* R\d+, R\d+, R\d+, -0\.5 ?;", body), "coordinate = fma(loc, size, -0.5)"
    assert len(re.findall(r"IMAD\.WIDE", body)) >= 8, "one IMAD.WIDE per gather / reduction address"
    if which.startswith("bwd"):
        assert re.search(r"REDG\.E\.ADD\.F32x4", body), "16-byte vector reductions"
        assert not re.search(r"\bATOMG\b", body), "no returning global atomics"
    else:
        assert not re.search(r"\b(REDG|ATOMG)\b", body), "the forward writes every output once"


def test_warp_scoped_kernels_have_no_block_barrier_in_the_unit_loop(kernels):
    """fwd_tile_kernel and bwd_warp_kernel synchronise with WARPSYNC only; the two BAR.SYNC they contain belong to the table
    prologue (load_tables)."""
    for which in ("fwd_tile", "bwd_warp"):
        body = _one(kernels, DEFAULTS[which])
        assert len(re.findall(r"BAR\.SYNC", body)) <= 2, which
