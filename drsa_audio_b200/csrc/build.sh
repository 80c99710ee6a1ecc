#!/usr/bin/env bash
# Builds libdrsa_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libdrsa_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=(api.cu sgemm.cu drsa_fp32.cu drsa_tc.cu retract.cu retract_fused.cu misc.cu lrp.cu conv_tc.cu subspace_filter.cu logmel.cu)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
OBJS=()
mkdir -p "$HERE/build"
pids=()
for s in "${SRCS[@]}"; do
  [ -f "$HERE/$s" ] || continue
  o="$HERE/build/${s%.cu}.o"
  OBJS+=("$o")
  if [ ! -f "$o" ] || [ "$HERE/$s" -nt "$o" ] || [ -n "$(find "$HERE" "$HERE/../../include" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null)" ]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$s" -o "$o" > "$HERE/build/${s%.cu}.ptxas.log" 2>&1 || { cat "$HERE/build/${s%.cu}.ptxas.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo "built $OUT"
