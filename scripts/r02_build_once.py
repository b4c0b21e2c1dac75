import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
verts = scenes.soup(4000000, seed=42)
A = capi.Accel(0)
for _ in range(2):
    t0 = time.perf_counter(); st = A.build(verts, builder=b); print("build", st["build_seconds"], "wall", time.perf_counter() - t0, "nodes", st["num_nodes"], "depth", st["max_depth"])
