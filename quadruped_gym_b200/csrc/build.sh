#!/usr/bin/env bash
# Builds libquadgym.so (sm_100a) in-tree: quadruped_gym_b200/libquadgym.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${QG_OUT:-$HERE/../libquadgym.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# -prec-div/-prec-sqrt=false, -ftz: float divisions and square roots become MUFU.RCP/RSQ sequences (<= 2 ulp) instead
# of the IEEE-rounded multi-instruction forms.  They sit on the serial dependency chains of the per-environment
# factorisations, which bound this latency-limited kernel: measured 1.91 -> 1.67 ms per launch.  (Full
# -use_fast_math also swaps sincosf for the MUFU forms and fails the open-loop parity test: not used.)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -prec-div=false -prec-sqrt=false -ftz=true \
    -Xcompiler -fPIC -shared ${QG_NVCC_EXTRA:-} \
    -o "$OUT" "$HERE/qg_api.cu"
echo "built $OUT"
