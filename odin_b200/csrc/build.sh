#!/usr/bin/env bash
# Builds odin_b200/lib/libodin_b200.so for sm_100a (in-tree, so it travels with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/.obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr
       -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v)
pids=()
for f in capi fe_kernels fe_frame5 cmvn_kernels gmm_kernels gmm_tc gmm_h tmat_kernels feat_kernels sig_kernels; do
  if [ ! -f "$HERE/.obj/$f.o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.cu' -o -name '*.cuh' \) -newer "$HERE/.obj/$f.o" 2>/dev/null | head -1)" ] \
     || [ "$HERE/../../include/odin_b200.h" -nt "$HERE/.obj/$f.o" ]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/.obj/$f.o" > "$HERE/.obj/$f.log" 2>&1 || { cat "$HERE/.obj/$f.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libodin_b200.so" "$HERE"/.obj/{capi,fe_kernels,fe_frame5,cmvn_kernels,gmm_kernels,gmm_tc,gmm_h,tmat_kernels,feat_kernels,sig_kernels}.o -lcudart -ldl
echo "built $OUT/libodin_b200.so"
