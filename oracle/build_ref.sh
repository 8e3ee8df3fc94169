#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the UNMODIFIED reference rasterizer (sources stay under /root/reference) plus
# oracle/ref_shim.cu into oracle/_ref/libgslidar_ref.so for sm_100a.  The only deviation
# from the reference's own JIT build (gaussian_renderer/diff_gaussian_rasterization_2d.py:14-24)
# is the flag-only fix `-include cstdint` (gcc 13 rejects rasterizer_impl.h:21 without it)
# and dropping `-g`.  Effective device flags match torch's cpp_extension defaults:
# -O3, -fmad=true, no --use_fast_math, --expt-relaxed-constexpr, -std=c++17.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSL_REFERENCE_DIR:-/root/reference}/diff-gaussian-rasterization-2d"
OUT="$HERE/_ref"
if [ ! -d "$REF/cuda_rasterizer" ]; then
  echo "[build_ref] reference not present at $REF; keeping prebuilt $OUT (if any)" >&2
  exit 0
fi
mkdir -p "$OUT"
SO="$OUT/libgslidar_ref.so"
STAMP="$OUT/.stamp"
NEW_STAMP="$(cat "$HERE/ref_shim.cu" "$HERE/build_ref.sh" "$REF"/cuda_rasterizer/*.cu "$REF"/cuda_rasterizer/*.h | sha1sum | cut -d' ' -f1)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"

# ---- the reference's Chamfer kernels (chamfer/chamfer3D/chamfer3D.cu), unmodified, against the stub ATen header ----
CH_SRC="${GSL_REFERENCE_DIR:-/root/reference}/chamfer/chamfer3D/chamfer3D.cu"
CH_SO="$OUT/libchamfer_ref.so"
CH_STAMP="$OUT/.stamp_chamfer"
if [ -f "$CH_SRC" ]; then
  CH_NEW="$(cat "$CH_SRC" "$HERE/ref_chamfer_shim.cu" "$HERE/ref_stub/ATen/ATen.h" "$HERE/build_ref.sh" | sha1sum | cut -d' ' -f1)"
  if [ -f "$CH_SO" ] && [ -f "$CH_STAMP" ] && [ "$(cat "$CH_STAMP")" = "$CH_NEW" ]; then
    echo "[build_ref] up to date: $CH_SO"
  else
    CH_TMP="$(mktemp -d)"
    CH_FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I"$HERE/ref_stub")
    "$NVCC" "${CH_FLAGS[@]}" -c "$CH_SRC" -o "$CH_TMP/chamfer3D.o"
    "$NVCC" "${CH_FLAGS[@]}" -c "$HERE/ref_chamfer_shim.cu" -o "$CH_TMP/shim.o"
    "$NVCC" -shared -o "$CH_SO" "$CH_TMP/chamfer3D.o" "$CH_TMP/shim.o" -lcudart
    rm -rf "$CH_TMP"
    echo "$CH_NEW" > "$CH_STAMP"
    echo "[build_ref] built $CH_SO"
  fi
fi

if [ -f "$SO" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$NEW_STAMP" ]; then
  echo "[build_ref] up to date: $SO"
  exit 0
fi
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr
       -include cstdint -Xcompiler -fPIC -I"$REF/third_party/glm" -I"$REF" -I"$REF/cuda_rasterizer")
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
pids=()
for f in forward backward rasterizer_impl; do
  "$NVCC" "${FLAGS[@]}" -c "$REF/cuda_rasterizer/$f.cu" -o "$TMP/$f.o" &
  pids+=($!)
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_shim.cu" -o "$TMP/ref_shim.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$SO" "$TMP"/forward.o "$TMP"/backward.o "$TMP"/rasterizer_impl.o "$TMP"/ref_shim.o -lcudart
echo "$NEW_STAMP" > "$STAMP"
echo "[build_ref] built $SO"
