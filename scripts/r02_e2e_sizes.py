"""Host-buffer call (pinned buffers, compact form) at several batch sizes: Mrays/s, device-resident launch beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import torch
from lmb200py import capi, scenes
import bench
L = capi.lib()
verts = scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)
lo, hi = scenes.bounds(verts)
A = capi.Accel(0); A.build(verts)
out = []
for n in (1 << 18, 1 << 20, 1 << 22, 1 << 24, 3 << 23):
    d = bench.gen_rays_device(torch, n, lo.tolist(), hi.tolist(), 7, torch.device('cuda'))
    h24 = torch.empty((n, 6), dtype=torch.float32, pin_memory=True); h24.copy_(d[:, [0, 1, 2, 4, 5, 6]])
    hh = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    dh = torch.empty((n, 4), dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    def dev(): capi.check(L.lmb200_trace_closest_dev(A.h, d.data_ptr(), dh.data_ptr(), n, st)); torch.cuda.synchronize()
    def host(): capi.check(L.lmb200_trace_closest_compact(A.h, h24.data_ptr(), 1e-4, 3.4028234663852886e38, hh.data_ptr(), n))
    res = []
    for f in (dev, host):
        f(); f()
        t0 = time.perf_counter(); reps = 5
        for _ in range(reps): f()
        res.append(n * reps / (time.perf_counter() - t0) / 1e6)
    same = bool(torch.equal(hh.cuda().view(torch.int32), dh.view(torch.int32)))
    out.append(f"{n >> 10}Ki: device {res[0]:.0f} host {res[1]:.0f} equal {same}")
print(os.environ.get("LMB200_E2E_STREAM", "1"), "; ".join(out))
