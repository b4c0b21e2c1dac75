"""First GPU contact: parity of the CUDA closest-hit kernel vs the C oracle + a rough rate."""
import sys, time, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from oracle import bindings as ob
from lmb200py import capi, scenes
import torch

for ntri, nray, extent in [(20000, 200000, 10.0), (1000000, 4000000, 100.0)]:
    verts = scenes.soup(ntri, seed=42, extent=extent, edge=0.2)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(nray, lo, hi, seed=7)
    A = capi.Accel(0)
    st = A.build(verts); print(st, flush=True)
    t0 = time.time(); hits = A.trace_closest(rays); t1 = time.time()
    print("host-buffer trace: %.3fs = %.1f Mrays/s" % (t1 - t0, nray / (t1 - t0) / 1e6), flush=True)
    nchk = min(nray, 300000)
    P = ob.PortScene(verts)
    tuv, tri = P.closest(rays[:nchk])
    gtri = hits['tri'][:nchk].astype(np.int64); gtri[gtri == 0xFFFFFFFF] = -1
    print("idx mismatch:", np.count_nonzero(gtri != tri), "of", nchk, "hits", (tri >= 0).sum())
    g = np.stack([hits['t'][:nchk], hits['u'][:nchk], hits['v'][:nchk]], axis=1)
    print("tuv bit mismatch:", np.count_nonzero(g.view(np.uint32) != tuv.view(np.uint32)))
    occ = A.trace_any(rays[:nchk]); print("any mismatch:", np.count_nonzero(occ != P.any(rays[:nchk])))
    # device-resident timing
    d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.empty((nray, 4), dtype=torch.float32, device='cuda')
    L = capi.lib()
    for _ in range(3):
        capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), nray, None))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), nray, torch.cuda.current_stream().cuda_stream))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    npr, tpr = C.c_double(), C.c_double()
    capi.check(L.lmb200_trace_count_dev(A.h, d_rays.data_ptr(), nray, C.byref(npr), C.byref(tpr)))
    print("device trace: %.3f ms = %.1f Mrays/s; nodes/ray %.2f tris/ray %.2f" % (ms, nray / ms / 1e3, npr.value, tpr.value), flush=True)
