#!/usr/bin/env bash
# Builds libcvb200.so (sm_100a only) next to the Python package.  Usage: build.sh [diag] [extra nvcc flags]
#   build.sh diag ...  builds libcvb200_diag.so instead: same sources with -DCVB_DIAG=1 (per-role cycle counters + CVB_DBG switches of the
#   conv kernel; used by tools/conv_pipeline_profile.py and tools/head_profile.py through CVB_DIAG_LIB=1, never by the product path)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libcvb200.so"
EXTRA=()
if [[ "${1:-}" == "diag" ]]; then
  shift
  OUT="$HERE/../libcvb200_diag.so"
  EXTRA=(-DCVB_DIAG=1)
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --shared
       -I"$HERE/../../include" -lcudart)
"$NVCC" "${FLAGS[@]}" "${EXTRA[@]}" "$@" -o "$OUT" "$HERE/api.cu" "$HERE/conv_tc.cu" "$HERE/aux_kernels.cu" "$HERE/nms.cu" "$HERE/train_kernels.cu"
echo "built $OUT"
