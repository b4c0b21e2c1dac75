#!/bin/bash
# TEST INFRASTRUCTURE. Compiles the reference's own hot-path sources, where they lie under
# $LM_REFERENCE (default /root/reference), plus the shims in this directory into
# oracle/_ref/liblightmetrica.so. Nothing is copied into the repo: the only transformation is
# stripping "#pragma region/endregion" lines on the fly (GCC 13 rejects them inside NSDMI lambdas),
# into a scratch dir under oracle/_ref/build (git-ignored).
# Flags define the oracle: -ffp-contract=off (no FMA fusion => TriAccel results are pure IEEE
# single ops and bit-reproducible across -march choices), -DLM_USE_SINGLE_PRECISION (the
# reference default, cmake/LMBuildOptions.cmake:32), -DLM_EXPORTS (LM_EXPORTED_F is a direct call).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LM_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
BUILD="$OUT/build"
if [ ! -d "$REF/src/liblightmetrica" ]; then
  echo "build_ref: reference not found at $REF (prebuilt oracle/_ref is used as is)"; exit 0
fi
mkdir -p "$BUILD"
CXX="${LMB200_CXX:-/usr/bin/g++}"
CXXFLAGS="-std=c++14 -O2 -msse4.2 -ffp-contract=off -fPIC -DLM_EXPORTS -DLM_USE_SINGLE_PRECISION -DDSFMT_MEXP=19937 -DNDEBUG -Wno-deprecated -Wno-deprecated-declarations -I$HERE/shim -I$REF/include -I$REF/external-src/dSFMT-src-2.2.3 -I$HERE"
SRCS="
src/liblightmetrica/accel/accel_naive.cpp
src/liblightmetrica/accel/accel_bvh_sahbin.cpp
src/liblightmetrica/accel/accel_qbvh.cpp
src/liblightmetrica/scene3.cpp
src/liblightmetrica/renderer/renderer_pt.cpp
src/liblightmetrica/renderer/renderer_ptdirect.cpp
src/liblightmetrica/renderer/renderer_ptmis.cpp
src/liblightmetrica/asset/bsdf/bsdf_diffuse.cpp
src/liblightmetrica/asset/bsdf/bsdf_cooktorrance.cpp
src/liblightmetrica/asset/bsdf/bsdf_null.cpp
src/liblightmetrica/asset/bsdf/bsdf_reflectall.cpp
src/liblightmetrica/asset/bsdf/bsdf_refractall.cpp
src/liblightmetrica/asset/bsdf/bsdf_flesnel.cpp
src/liblightmetrica/asset/light/light_area.cpp
src/liblightmetrica/asset/light/light_point.cpp
src/liblightmetrica/asset/light/light_directional.cpp
src/liblightmetrica/asset/light/light_env.cpp
src/liblightmetrica/asset/sensor/sensor_pinhole.cpp
src/liblightmetrica/asset/sensor/sensor_thinlens.cpp
src/liblightmetrica/asset/trianglemesh/trianglemesh_raw.cpp
src/liblightmetrica/asset/trianglemesh/trianglemesh_obj.cpp
src/liblightmetrica/random.cpp
plugin/texture_checker/texture_checker.cpp
"
OBJS=""
pids=()
for s in $SRCS; do
  n="$(basename "$s" .cpp)"
  o="$BUILD/$n.o"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    # sources from the reference's plugin/ tree include <lightmetrica/lightmetrica.h> instead of the pch: give them the std headers
    EXTRA=""; case "$s" in plugin/*) EXTRA="-include $HERE/shim/pch.h" ;; esac
    ( sed -E 's/^\s*#pragma (region|endregion).*$//' "$REF/$s" > "$BUILD/$n.cpp" && $CXX $CXXFLAGS $EXTRA -c "$BUILD/$n.cpp" -o "$o" ; rm -f "$BUILD/$n.cpp" ) &
    pids+=($!)
  fi
done
for s in runtime host_shims harness nanort_bench; do
  o="$BUILD/shim_$s.o"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$HERE/$s.cpp" -nt "$o" ] || [ "$HERE/host_shims.h" -nt "$o" ]; then
    ( $CXX $CXXFLAGS -c "$HERE/$s.cpp" -o "$o" ) &
    pids+=($!)
  fi
done
if [ ! -f "$BUILD/dSFMT.o" ]; then
  ( gcc -O2 -fPIC -DDSFMT_MEXP=19937 -msse2 -DHAVE_SSE2 -I"$REF/external-src/dSFMT-src-2.2.3" -c "$REF/external-src/dSFMT-src-2.2.3/dSFMT.c" -o "$BUILD/dSFMT.o" ) &
  pids+=($!)
fi
OBJS="$OBJS $BUILD/dSFMT.o"
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$CXX -shared -o "$OUT/liblightmetrica.so" -Wl,-soname,liblightmetrica.so $OBJS -ldl -pthread
echo "build_ref: built $OUT/liblightmetrica.so"
