"""Where does the host-buffer call lose against one device-resident launch? Device-resident rays, 64 Mi in total:
one launch / 8 launches of 8 Mi on one stream / 8 launches alternating between two streams / the same with the copies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import torch
from lmb200py import capi, scenes
import bench
L = capi.lib()
verts = scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)
lo, hi = scenes.bounds(verts)
A = capi.Accel(0); A.build(verts)
n = 1 << 26
rays = bench.gen_rays_device(torch, n, lo.tolist(), hi.tolist(), 7, torch.device('cuda'))
hits = torch.empty((n, 4), dtype=torch.float32, device='cuda')
s = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(chunks, nstreams):
    m = n // chunks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for st in s[:nstreams]: st.wait_event(e0)
    for c in range(chunks):
        st = s[c % nstreams]
        capi.check(L.lmb200_trace_closest_dev(A.h, rays.data_ptr() + c * m * 32, hits.data_ptr() + c * m * 16, m, st.cuda_stream))
    for st in s[:nstreams]:
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for chunks, ns in ((1, 1), (8, 1), (8, 2), (16, 2), (32, 2), (1, 1)):
    run(chunks, ns)
    ms = min(run(chunks, ns) for _ in range(3))
    print(f"{chunks:2d} launches on {ns} stream(s): {ms:7.2f} ms = {n / ms / 1e3:7.1f} Mrays/s", flush=True)
