"""Tree-quality figures without a GPU: nodes / triangle records fetched per ray by an ordered CPU walk of the product's
flattened structure (oracle orc_wide_count), on the bench scenes. usage: cpu_tree_quality.py [soup_tris] [mesh_tris] [lib.so]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
import numpy as np
from lmb200py import capi, scenes
from oracle import bindings as ob
if len(sys.argv) > 3:
    capi.LIB_PATH = sys.argv[3]
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
for name, verts in (("soup", scenes.soup(ns, seed=42, extent=100.0, edge=0.2) if ns else None), ("mesh", scenes.mesh_scene(nm, seed=42)[0] if nm else None)):
    if verts is None:
        continue
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(200_000, lo, hi, seed=7)
    A = capi.Accel(host_only=True)
    t = time.time(); st = A.build(verts); bt = time.time() - t
    units, nn, nt, grid = A.host_layout()
    N = units.view(np.dtype([("k", "<u2", 3), ("counts", "<u2"), ("e", "u1", 3), ("imask", "u1"), ("base", "<u4"), ("q", "u1", 48)])).reshape(-1)
    npr, tpr = ob.wide_count(units, grid, rays)
    print(f"{name}: {len(verts)} tris, build {bt:.1f} s, {nn} nodes ({len(verts)/nn:.2f} tris/node), depth {st['max_depth']}, sah {st['sah_cost']:.2f}: "
          f"{npr:.2f} nodes/ray, {tpr:.2f} tris/ray, bytes/ray {48 + 64*npr + 48*tpr:.0f}")
    sn, stt, sd = ob.wide_count_sorted(units, grid, rays)
    for mode, what in ((1, "octant order + one entry-distance bound per stacked group"), (2, "octant order + entry distance per child")):
        cn, ct, cd = ob.wide_count_cull(units, grid, rays, mode)
        print(f"    {what}: {cn:.2f} nodes/ray, {ct:.2f} tris/ray, {cd:.2f} children dropped unfetched")
    print(f"    exact front-to-back order with entry-distance cull: {sn:.2f} nodes/ray, {stt:.2f} tris/ray, {sd:.2f} entries dropped unfetched")
