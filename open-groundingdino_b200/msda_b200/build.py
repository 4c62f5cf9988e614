"""In-tree build of libmsda_b200.so (nvcc, sm_100a only).  The .so is git-ignored but travels to the
GPU box with the repo snapshot; nothing is compiled at import time."""
from __future__ import annotations

import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DIST_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libmsda_b200.so")


def _sources():
    files = glob.glob(os.path.join(DIST_DIR, "csrc", "*.cu")) + glob.glob(os.path.join(DIST_DIR, "csrc", "*.cuh"))
    files += glob.glob(os.path.join(DIST_DIR, "csrc", "*.h"))
    files.append(os.path.join(os.path.dirname(DIST_DIR), "include", "msda_b200.h"))
    files.append(os.path.join(os.path.dirname(DIST_DIR), "include", "hungarian_b200.h"))
    files.append(os.path.join(DIST_DIR, "build_lib.sh"))
    return files


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        subprocess.run(["bash", os.path.join(DIST_DIR, "build_lib.sh")], check=True,
                       stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
