#!/usr/bin/env bash
# Builds libcvb200.so (sm_100a only) next to the Python package.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libcvb200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --shared
       -I"$HERE/../../include" -lcudart)
"$NVCC" "${FLAGS[@]}" "$@" -o "$OUT" "$HERE/api.cu" "$HERE/conv_tc.cu" "$HERE/aux_kernels.cu" "$HERE/nms.cu"
echo "built $OUT"
