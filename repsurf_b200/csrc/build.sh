#!/usr/bin/env bash
# Builds librepsurf_b200.so (C-ABI, no torch dependency) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../librepsurf_b200.so
SRCS=$(ls *.cu)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -ccbin /usr/bin/g++"
mkdir -p obj
pids=()
for s in $SRCS; do
  o=obj/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name "*.cuh" -o -name "*.h" \) -newer "$o")" ] || [ "${FORCE:-0}" = 1 ]; then
    $NVCC $FLAGS ${RSB_EXTRA_FLAGS:-} ${PTXAS_V:+-Xptxas -v} -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT obj/*.o -lcudart -ccbin /usr/bin/g++
echo "built $(readlink -f $OUT)"
