#!/bin/bash
# Builds the two Lightmetrica plugins against the reference headers (only where the reference tree
# exists: this container). Same ABI flags as the host build (SURVEY.md §8b "Build coupling"):
# -std=c++14 -DLM_USE_SINGLE_PRECISION -msse4.2, system g++/libstdc++. No -DLM_EXPORTS: plugins
# resolve the core's C exports by dlsym on liblightmetrica.so (static.h:228-280).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${LM_REFERENCE:-/root/reference}"
if [ ! -d "$REF/include/lightmetrica" ]; then
  echo "build_plugins: reference headers not found at $REF (prebuilt plugins are used as is)"; exit 0
fi
CXX=/usr/bin/g++
FLAGS="-std=c++14 -O2 -msse4.2 -ffp-contract=off -fPIC -DLM_USE_SINGLE_PRECISION -DNDEBUG -Wno-deprecated -Wno-deprecated-declarations -I$REF/include -I$ROOT/include -I$HERE -include $HERE/prelude.h"
for p in accel_lmb200 renderer_lmb200pt; do
  $CXX $FLAGS -shared -o "$HERE/$p.so" "$HERE/$p.cpp" -L"$ROOT/lightmetrica-v2_b200/lib" -llmb200 -Wl,-rpath,'$ORIGIN/../lib' -ldl -pthread
done
echo "build_plugins: built $HERE/accel_lmb200.so $HERE/renderer_lmb200pt.so"
