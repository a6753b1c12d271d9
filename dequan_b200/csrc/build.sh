#!/bin/bash
# Builds dequan_b200/lib/libdequan_b200.so for sm_100a (in-tree, travels with gpurun snapshots).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       -I"$HERE/../../include" -I"$HERE")
if [ "${DQ_PTXAS_V:-0}" = "1" ]; then FLAGS+=(-Xptxas -v); fi

"$NVCC" "${FLAGS[@]}" -shared -o "$OUT/libdequan_b200.so" "$HERE/dq_api.cu" "$HERE/dq_compile.cpp" "$HERE/dq_formats.cpp"
echo "built $OUT/libdequan_b200.so"
