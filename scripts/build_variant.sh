#!/bin/bash
# Tuning aid: builds lightmetrica-v2_b200/lib/variants/liblmb200_<name>.so with extra -D flags.
# usage: build_variant.sh <name> -DLMB_REFILL_BELOW=26 ...
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
SRC="$ROOT/lightmetrica-v2_b200/csrc"; OUT="$ROOT/lightmetrica-v2_b200/lib/variants"; mkdir -p "$OUT/obj_$NAME"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2,-ffp-contract=off --fmad=false $*"
for f in accel render bvh_build_gpu service; do $NVCC $FLAGS -c "$SRC/$f.cu" -o "$OUT/obj_$NAME/$f.o" & done
/usr/bin/g++ -O2 -std=c++17 -fPIC -ffp-contract=off -pthread "$@" -c "$SRC/bvh_build.cpp" -o "$OUT/obj_$NAME/bvh_build.o" &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "$OUT/liblmb200_$NAME.so" "$OUT/obj_$NAME/accel.o" "$OUT/obj_$NAME/render.o" "$OUT/obj_$NAME/bvh_build_gpu.o" "$OUT/obj_$NAME/service.o" "$OUT/obj_$NAME/bvh_build.o" -lpthread -ldl
rm -rf "$OUT/obj_$NAME"
echo "$OUT/liblmb200_$NAME.so"
