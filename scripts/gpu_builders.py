"""Host SAH builder vs device LBVH builder: build time and traversal rate on both scene families."""
import sys, os, time, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes
import bench
if os.environ.get("LMB200_LIB"):
    capi.LIB_PATH = os.environ["LMB200_LIB"]
L = capi.lib()
ONLY = os.environ.get("BUILDERS", "host_sah,gpu_lbvh,gpu_ploc").split(",")
for name, verts in [("soup4M", scenes.soup(4000000, seed=42)), ("mesh1M", scenes.mesh_scene(1000000, seed=42)[0])]:
    lo, hi = scenes.bounds(verts)
    n = 1 << 24
    d_rays = bench.gen_rays_device(torch, n, lo.tolist(), hi.tolist(), 7, torch.device('cuda'))
    d_hits = torch.empty((n, 4), dtype=torch.float32, device='cuda')
    for bname, b in [("host_sah", capi.BUILD_HOST_SAH), ("gpu_lbvh", capi.BUILD_GPU_LBVH), ("gpu_ploc", capi.BUILD_GPU_PLOC)]:
        if bname not in ONLY:
            continue
        A = capi.Accel(0)
        A.build(verts, builder=b)           # warm-up (allocator, module load)
        t0 = time.perf_counter(); st = A.build(verts, builder=b); wall = time.perf_counter() - t0
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3): capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, s))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, s))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        npr, tpr = C.c_double(), C.c_double()
        capi.check(L.lmb200_trace_count_dev(A.h, d_rays.data_ptr(), 1 << 22, C.byref(npr), C.byref(tpr)))
        print("%s %s: build %.3f s (wall %.3f), nodes %d, %.0f Mrays/s, %.1f nodes/ray %.2f tris/ray" % (name, bname, st['build_seconds'] + st['upload_seconds'], wall, st['num_nodes'], n / ms / 1e3, npr.value, tpr.value), flush=True)
        A.close()
