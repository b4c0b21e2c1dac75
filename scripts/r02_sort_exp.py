"""Experiment: how much faster does the traversal kernel run when the ray batch is sorted for coherence (origin cell + direction
octant)? Upper bound on what an internal ray-sorting pass can buy (sort cost and indirection not included)."""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import numpy as np, torch
from lmb200py import capi, scenes
import bench
L = capi.lib()
dev = torch.device('cuda')

def timed(A, rays, n, reps=4):
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): capi.check(L.lmb200_trace_closest_dev(A.h, rays.data_ptr(), hits.data_ptr(), n, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): capi.check(L.lmb200_trace_closest_dev(A.h, rays.data_ptr(), hits.data_ptr(), n, st))
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / reps) / 1e3

def spread(x, bits):
    # interleave-ready: put bit i of x at position 3 i
    r = torch.zeros_like(x)
    for i in range(bits):
        r |= ((x >> i) & 1) << (3 * i)
    return r

for name, verts in [("soup4M", scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)), ("mesh1M", scenes.mesh_scene(1000000, seed=42)[0])]:
    lo, hi = scenes.bounds(verts)
    A = capi.Accel(0); A.build(verts)
    n = 1 << 24
    rays = bench.gen_rays_device(torch, n, lo.tolist(), hi.tolist(), 7, dev)
    base = timed(A, rays, n)
    res = {"unsorted": base}
    lo_t, hi_t = torch.tensor(lo, device=dev), torch.tensor(hi, device=dev)
    octant = ((rays[:, 4] < 0).long() | ((rays[:, 5] < 0).long() << 1) | ((rays[:, 6] < 0).long() << 2))
    for bits in (3, 4, 5, 6, 7):
        c = ((rays[:, 0:3] - lo_t) / (hi_t - lo_t) * (1 << bits)).long().clamp(0, (1 << bits) - 1)
        m = spread(c[:, 0], bits) | (spread(c[:, 1], bits) << 1) | (spread(c[:, 2], bits) << 2)
        for label, key in ((f"cell{bits}+oct", (m << 3) | octant), (f"oct+cell{bits}", (octant << (3 * bits)) | m)):
            order = torch.argsort(key)
            sr = rays[order].contiguous()
            res[label] = timed(A, sr, n)
            del sr, order
    # direction-major: octant + 2 more direction bits per axis, then cell
    d = ((rays[:, 4:7] * 0.5 + 0.5) * 8).long().clamp(0, 7)
    dm = spread(d[:, 0], 3) | (spread(d[:, 1], 3) << 1) | (spread(d[:, 2], 3) << 2)
    c = ((rays[:, 0:3] - lo_t) / (hi_t - lo_t) * 32).long().clamp(0, 31)
    m = spread(c[:, 0], 5) | (spread(c[:, 1], 5) << 1) | (spread(c[:, 2], 5) << 2)
    order = torch.argsort((m << 9) | dm)
    res["cell5+dir9"] = timed(A, rays[order].contiguous(), n)
    print(name, json.dumps({k: round(v, 1) for k, v in res.items()}), flush=True)
    A.close()
