"""Checks the streaming host-buffer path (accel.cu trace_host_stream) against the device-pointer calls on the same rays.
Run with LMB200_E2E_CHUNK_LOG2=16 to get many small chunks (ring wrap-around, kernel outrunning the uploads)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import ctypes as C
import numpy as np, torch
from lmb200py import capi, scenes
L = capi.lib()
ntri = int(os.environ.get("CHECK_TRIS", "200000"))
verts = scenes.soup(ntri, seed=5, extent=20.0, edge=0.3)
lo, hi = scenes.bounds(verts)
A = capi.Accel(0); A.build(verts)
ok = True
for n in [int(x) for x in os.environ.get("CHECK_N", "1500001,262144,700000,65537").split(",")]:
    rays = scenes.random_rays(n, lo, hi, seed=9 + n % 7)
    rays[:, 3] = 1e-4; rays[:, 7] = 3.0e38
    d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.empty((n, 4), dtype=torch.float32, device='cuda'); d_occ = torch.zeros(n, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, st))
    capi.check(L.lmb200_trace_any_dev(A.h, d_rays.data_ptr(), d_occ.data_ptr(), n, st))
    torch.cuda.synchronize()
    want = d_hits.cpu().numpy().view(np.uint32); want_occ = d_occ.cpu().numpy()
    for rep in range(3):
        t0 = time.perf_counter()
        got = A.trace_closest(rays)
        t1 = time.perf_counter()
        occ = A.trace_any(rays)
        r24 = np.ascontiguousarray(rays[:, [0, 1, 2, 4, 5, 6]])
        h2 = np.zeros(n, capi.HIT_DTYPE)
        capi.check(L.lmb200_trace_closest_compact(A.h, r24.ctypes.data, 1e-4, 3.0e38, h2.ctypes.data, n))
        o2 = np.zeros(n, np.uint8)
        capi.check(L.lmb200_trace_any_compact(A.h, r24.ctypes.data, 1e-4, 3.0e38, o2.ctypes.data, n))
        e = [np.array_equal(got.view(np.uint32).reshape(-1, 4), want), np.array_equal(occ.astype(np.uint8), want_occ),
             np.array_equal(h2.view(np.uint32).reshape(-1, 4), want), np.array_equal(o2, want_occ)]
        ok = ok and all(e)
        print(f"n={n} rep {rep}: closest {e[0]} any {e[1]} compact {e[2]} compact-any {e[3]}  ({n / (t1 - t0) / 1e6:.1f} Mrays/s closest, pageable host arrays)", flush=True)
print("STREAM_CHECK_OK" if ok else "STREAM_CHECK_FAILED")
