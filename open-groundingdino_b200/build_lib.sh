#!/usr/bin/env bash
# Builds libmsda_b200.so in-tree (sm_100a only).  Called by msda_b200/build.py and __graft_entry__.build().
# One nvcc per translation unit, run in parallel, then one link step.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="$here/msda_b200/libmsda_b200.so"
obj="$here/build"
log="$here/msda_b200/build.log"
mkdir -p "$obj"
: > "$log"
flags=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v
       -I "$here/../include" -I "$here/csrc" -ccbin /usr/bin/g++ ${MSDA_NVCC_EXTRA:-})
pids=()
for src in "$here"/csrc/*.cu; do
  base="$(basename "$src" .cu)"
  ( "$NVCC" "${flags[@]}" -c -o "$obj/$base.o" "$src" > "$obj/$base.log" 2>&1 ) &
  pids+=($!)
done
fail=0
for pid in "${pids[@]}"; do wait "$pid" || fail=1; done
cat "$obj"/*.log >> "$log"
if [ "$fail" -ne 0 ]; then cat "$log" >&2; exit 1; fi
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o "$out" "$obj"/*.o >> "$log" 2>&1 || { cat "$log" >&2; exit 1; }
echo "$out"
