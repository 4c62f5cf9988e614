#!/usr/bin/env python
"""Build the REFERENCE's own CUDA op into oracle/_ref/ (git-ignored; travels to the GPU box).

TEST/BENCH INFRASTRUCTURE ONLY -- the product never loads it.  Used by tests/test_vs_reference_cuda.py
and bench.py to put the reference kernels next to ours on the same B200.

Nothing is copied: three 3-line wrapper translation units are generated under oracle/_ref/src/ that
``#include`` the reference sources where they lie (/root/reference/models/GroundingDINO/ops/src/...).
The reference does not compile verbatim against torch 2.11 (``AT_DISPATCH_FLOATING_TYPES(value.type(), ...)``
at src/cuda/ms_deform_attn_cuda.cu:64,134 no longer converts); the CUDA wrapper re-defines that macro to
accept a DeprecatedTypeProperties before including the file, which is equivalent to the 2-token
``.type()`` -> ``.scalar_type()`` patch without touching the reference text.

Only runs where /root/reference exists (the authoring container).  On the GPU box the prebuilt
oracle/_ref/msda_ref_cuda.so is loaded as is (oracle/ref_cuda.py).
"""
import os
import shutil
import sys

REF_SRC = "/root/reference/models/GroundingDINO/ops/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "msda_ref_cuda"

WRAPPERS = {
    "wrap_vision.cpp": f'#include "{REF_SRC}/vision.cpp"\n',
    "wrap_cpu.cpp": f'#include "{REF_SRC}/cpu/ms_deform_attn_cpu.cpp"\n',
    "wrap_cuda.cu": (
        "#include <ATen/ATen.h>\n#include <ATen/Dispatch.h>\n"
        "static inline at::ScalarType msda_ref_st(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }\n"
        "static inline at::ScalarType msda_ref_st(at::ScalarType t) { return t; }\n"
        "#undef AT_DISPATCH_FLOATING_TYPES\n"
        "#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \\\n"
        "  AT_DISPATCH_SWITCH(msda_ref_st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))\n"
        f'#include "{REF_SRC}/cuda/ms_deform_attn_cuda.cu"\n'),
}


def build(force: bool = False) -> str:
    target = os.path.join(OUT, NAME + ".so")
    if os.path.exists(target) and not force:
        return target
    if not os.path.isdir(REF_SRC):
        raise RuntimeError(f"{REF_SRC} not present; the reference op can only be built in the authoring container")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils.cpp_extension import load
    src_dir = os.path.join(OUT, "src")
    build_dir = os.path.join(OUT, "build")
    os.makedirs(src_dir, exist_ok=True)
    os.makedirs(build_dir, exist_ok=True)
    for fname, text in WRAPPERS.items():
        with open(os.path.join(src_dir, fname), "w") as f:
            f.write(text)
    load(name=NAME, sources=[os.path.join(src_dir, f) for f in WRAPPERS], extra_include_paths=[REF_SRC],
         extra_cflags=["-DWITH_CUDA", "-O2"], extra_cuda_cflags=["-DWITH_CUDA", "-O3", "-lineinfo"],
         with_cuda=True, build_directory=build_dir, is_python_module=False, verbose=False)
    shutil.copy(os.path.join(build_dir, NAME + ".so"), target)
    shutil.rmtree(build_dir, ignore_errors=True)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
