#!/usr/bin/env bash
# Builds libquadgym.so (sm_100a) in-tree: quadruped_gym_b200/libquadgym.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libquadgym.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -shared ${QG_NVCC_EXTRA:-} \
    -o "$OUT" "$HERE/qg_api.cu"
echo "built $OUT"
