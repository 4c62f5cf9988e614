#!/usr/bin/env bash
# Builds libmsda_b200.so in-tree (sm_100a only).  Called by msda_b200/build.py and __graft_entry__.build().
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="$here/msda_b200/libmsda_b200.so"
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared -Xptxas -v ${MSDA_NVCC_EXTRA:-} \
  -I "$here/../include" -I "$here/csrc" \
  -ccbin /usr/bin/g++ \
  -o "$out" "$here"/csrc/msda_api.cu "$here"/csrc/msda_fwd.cu "$here"/csrc/msda_bwd.cu \
  "$here"/csrc/msda_generic.cu "$here"/csrc/msda_det.cu 2> "$here/msda_b200/build.log" || { cat "$here/msda_b200/build.log" >&2; exit 1; }
echo "$out"
