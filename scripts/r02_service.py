"""Per-ray service: latency (1 thread) and aggregate rate (2..32 host threads) on the soup and the mesh scene."""
import os, sys, ctypes as C, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import numpy as np
from lmb200py import capi, scenes
L = capi.lib()
for name, verts in [("soup4M", scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)), ("mesh1M", scenes.mesh_scene(1000000, seed=42)[0])]:
    lo, hi = scenes.bounds(verts)
    A = capi.Accel(0); A.build(verts)
    n = 100000
    rays = scenes.random_rays(n, lo, hi, seed=7)
    batch = A.trace_closest(rays)
    hits = np.zeros(n, capi.HIT_DTYPE)
    sec = C.c_double()
    out = {}
    capi.check(L.lmb200_trace_closest_one_mt(A.h, rays.ctypes.data, hits.ctypes.data, 5000, 4, C.byref(sec)))
    for th in (1, 2, 4, 8, 16, 32, 64):
        m = min(n, 8000 * th)
        capi.check(L.lmb200_trace_closest_one_mt(A.h, rays.ctypes.data, hits.ctypes.data, m, th, C.byref(sec)))
        assert np.array_equal(hits[:m].view(np.uint32), batch[:m].view(np.uint32))
        out[th] = dict(mrays=round(m / sec.value / 1e6, 3), us_per_ray_per_thread=round(sec.value / m * th * 1e6, 1))
    print(name, json.dumps(out), flush=True)
    A.close()
