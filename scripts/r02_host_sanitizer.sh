#!/bin/bash
# Host-side code under AddressSanitizer + UndefinedBehaviorSanitizer, no GPU needed: the library's host paths (host-only
# accel, binned SAH builder, unit flattening, argument checks: tests/test_host.py) and the renderer plugin's scene
# extraction through the reference host (flatten.h, BSDF / light / sensor conversion, texture baking, scene dump).
# Builds sanitized copies into a scratch directory; nothing in the tree is touched.
# Usage: scripts/r02_host_sanitizer.sh [scratch dir]        (needs /root/reference for the plugin part)
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-/tmp/lmb200_asan}"
SRC="$ROOT/lightmetrica-v2_b200/csrc"; PLG="$ROOT/lightmetrica-v2_b200/plugin"; REF="${LM_REFERENCE:-/root/reference}"
mkdir -p "$OUT/lib" "$OUT/plug"
SAN="-fsanitize=address -fsanitize=undefined -fno-omit-frame-pointer"
NVF="-O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ --fmad=false -I$ROOT/include -Xcompiler -fPIC,-O1,-g,-ffp-contract=off,${SAN// /,}"
for f in accel render bvh_build_gpu service; do /usr/local/cuda/bin/nvcc $NVF -c "$SRC/$f.cu" -o "$OUT/lib/$f.o" & done
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off $SAN -pthread -I"$ROOT/include" -c "$SRC/bvh_build.cpp" -o "$OUT/lib/bvh_build.o" &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "$OUT/lib/liblmb200.so" "$OUT"/lib/*.o -Xcompiler ${SAN// /,} -lpthread -ldl
export LD_PRELOAD="$(g++ -print-file-name=libasan.so) $(g++ -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 LMB200_LIB="$OUT/lib/liblmb200.so"
cd "$ROOT"
python -m pytest tests/test_host.py -q -s 2>&1 | tee "$OUT/host.log" | grep -iE "runtime error|AddressSanitizer|passed|failed" | sort | uniq -c
if [ -d "$REF/include/lightmetrica" ]; then
  FLAGS="-std=c++14 -O1 -g -msse4.2 -ffp-contract=off -fPIC $SAN -DLM_USE_SINGLE_PRECISION -DNDEBUG -Wno-deprecated -Wno-deprecated-declarations -I$REF/include -I$ROOT/include -I$PLG -include $PLG/prelude.h"
  for p in accel_lmb200 renderer_lmb200pt; do
    LD_PRELOAD= g++ $FLAGS -shared -o "$OUT/plug/$p.so" "$PLG/$p.cpp" -L"$OUT/lib" -llmb200 -Wl,-rpath,"$OUT/lib" -ldl -pthread
  done
  LMB200_ASAN_PLUG="$OUT/plug" LMB200_DUMP_SCENE="$OUT/scene.bin" python - <<'PY' 2>&1 | tee "$OUT/plugin.log" | grep -vi "no CUDA device" | grep -iE "runtime error|AddressSanitizer|^ok|Traceback|Error" | sort | uniq -c
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
import numpy as np
from oracle import bindings as ob
from lmb200py import scenedesc, scenes
P = os.environ["LMB200_ASAN_PLUG"]
L = ob.ref()
assert L.ref_load_plugin((P + "/accel_lmb200").encode()) == 1 and L.ref_load_plugin((P + "/renderer_lmb200pt").encode()) == 1
ball_c = np.array([0.4, 0.9, 0.3], np.float32)
ball = scenes.sphere(ball_c, 0.25, 10, 6)
n = ball.reshape(-1, 3) - ball_c
n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
with_normals = scenedesc.cornell_box(16, 16)
with_normals.add_mesh_tris(ball, "red", normals=n)
for name, sc in (("cornell", scenedesc.cornell_box(16, 16, glossy_block=True)), ("textured", scenedesc.textured_box(16, 16)),
                 ("outdoor", scenedesc.outdoor_scene(32, 18, light="both", thinlens=True)), ("config2-200k", scenedesc.config2_scene(200000, 64, 36)),
                 ("vertex normals", with_normals)):
    R = ob.RefScene(sc, accel="qbvh")
    R.render("lmb200pt", 10, extra={"mode": "ptdirect", "texture_resolution": getattr(sc, "tex_res", 64)}, in_tree=True)   # dumps, then fails on the missing device
    print("ok", name, os.path.getsize(os.environ["LMB200_DUMP_SCENE"]), "bytes dumped")
PY
fi
# ---- ThreadSanitizer over the multi-threaded host builder (record preparation and the subtree work queue of bvh_build.cpp) ----
mkdir -p "$OUT/tsan"
TS="-fsanitize=thread -fno-omit-frame-pointer"
NVT="-O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ --fmad=false -I$ROOT/include -Xcompiler -fPIC,-O1,-g,-ffp-contract=off,${TS// /,}"
for f in accel render bvh_build_gpu service; do LD_PRELOAD= /usr/local/cuda/bin/nvcc $NVT -c "$SRC/$f.cu" -o "$OUT/tsan/$f.o" 2>/dev/null & done
LD_PRELOAD= g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off $TS -pthread -I"$ROOT/include" -c "$SRC/bvh_build.cpp" -o "$OUT/tsan/bvh_build.o" &
wait
LD_PRELOAD= /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o "$OUT/tsan/liblmb200.so" "$OUT"/tsan/*.o -Xcompiler -fsanitize=thread -lpthread -ldl
LD_PRELOAD="$(g++ -print-file-name=libtsan.so)" TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0" LMB200_LIB="$OUT/tsan/liblmb200.so" \
  python -m pytest tests/test_host.py -q -s -k "builder or wide_tree" 2>&1 | tee "$OUT/tsan.log" | grep -iE "WARNING: ThreadSanitizer|passed|failed" | sort | uniq -c

