"""Primary-ray throughput vs how the rays are grouped (what a tile-sorted primary queue could reach):
random raster positions (as the path tracer draws them), the same rays sorted by T x T pixel tile, and tile-coherent
groups of G rays shuffled globally."""
import sys, os, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
d, keep = sc.flatten()
A = capi.Accel(0); A.build(keep['verts'])
c = sc.camera
W, H = 1920, 1080
g = np.random.default_rng(1)
n = 1 << 22
# random raster positions -> rays (pinhole)
u = g.random((n, 2), dtype=np.float32)
vx, vy, vz = (np.asarray(v, np.float32) for v in scenes.lookat(c['eye'], c['center'], c['up']))
tanf = np.float32(np.tan(np.radians(np.float32(c['fov'])) * 0.5)); asp = np.float32(W / H)
e = np.stack([asp * tanf * (2 * u[:, 0] - 1), tanf * (2 * u[:, 1] - 1), -np.ones(n, np.float32)], axis=1)
e /= np.linalg.norm(e, axis=1, keepdims=True)
dirs = e[:, 0:1] * vx + e[:, 1:2] * vy + e[:, 2:3] * vz
rays = np.zeros((n, 8), np.float32); rays[:, 0:3] = np.asarray(c['eye'], np.float32); rays[:, 3] = 1e-4; rays[:, 4:7] = dirs; rays[:, 7] = 3.4e38
px = np.minimum((u[:, 0] * W).astype(np.int64), W - 1); py = np.minimum((u[:, 1] * H).astype(np.int64), H - 1)
L = capi.lib()
def rate(r, label):
    m = len(r); dr = torch.from_numpy(np.ascontiguousarray(r)).cuda(); dh = torch.empty((m, 4), dtype=torch.float32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): capi.check(L.lmb200_trace_closest_dev(A.h, dr.data_ptr(), dh.data_ptr(), m, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): capi.check(L.lmb200_trace_closest_dev(A.h, dr.data_ptr(), dh.data_ptr(), m, s))
    e1.record(); torch.cuda.synchronize()
    print("%-44s %7.0f Mrays/s" % (label, m / (e0.elapsed_time(e1) / 5) / 1e3), flush=True)
rate(rays, "random raster positions (today)")
for T in (64, 32, 16, 8):
    tile = (py // T) * ((W + T - 1) // T) + (px // T)
    order = np.argsort(tile, kind='stable')
    rate(rays[order], f"sorted by {T}x{T} tile")
    if T in (32, 16):
        for G in (16, 32, 128):
            k = (n // G) * G
            grp = order[:k].reshape(-1, G)
            grp = grp[g.permutation(len(grp))]
            rate(rays[grp.reshape(-1)], f"  {T}x{T} tile, groups of {G} shuffled")
rate(rays[np.argsort(py * W + px, kind='stable')], "sorted by pixel")
