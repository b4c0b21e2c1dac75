"""How much faster are the same primary rays when traced in pixel order vs shuffled? (potential of sample sorting)"""
import sys, os, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
d, keep = sc.flatten()
A = capi.Accel(0); A.build(keep['verts'])
c = sc.camera
rays = scenes.camera_rays(c['eye'], c['center'], c['up'], c['fov'], 3840, 2160)
L = capi.lib()
def rate(r, label):
    n = len(r); dr = torch.from_numpy(r).cuda(); dh = torch.empty((n, 4), dtype=torch.float32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): capi.check(L.lmb200_trace_closest_dev(A.h, dr.data_ptr(), dh.data_ptr(), n, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): capi.check(L.lmb200_trace_closest_dev(A.h, dr.data_ptr(), dh.data_ptr(), n, s))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    npr, tpr = C.c_double(), C.c_double()
    capi.check(L.lmb200_trace_count_dev(A.h, dr.data_ptr(), min(n, 1 << 22), C.byref(npr), C.byref(tpr)))
    print("%-28s %7.0f Mrays/s  nodes/ray %.1f tris/ray %.2f hit %.2f" % (label, n / ms / 1e3, npr.value, tpr.value, float((dh[:, 3].view(torch.int32) != -1).float().mean())), flush=True)
rate(rays, "primary, pixel order")
g = np.random.default_rng(1); perm = g.permutation(len(rays))
rate(rays[perm], "primary, shuffled")
# tile-sorted-within-groups: shuffle, then sort groups of 16 by pixel (what warp-level refill gives)
k = (len(rays) // 16) * 16
grp = perm[:k].reshape(-1, 16); grp = np.sort(grp, axis=1)
rate(rays[grp.reshape(-1)], "primary, 16-ray sorted groups")
srt = np.sort(perm[:k].reshape(-1, 1 << 16), axis=1) if k % (1 << 16) == 0 else None
lo, hi = scenes.bounds(keep['verts'])
rate(scenes.random_rays(1 << 23, lo, hi, seed=7), "uniform random rays")
