"""Tuning aid: times device-resident closest-hit on (a) the 4M-tri soup and (b) the 1M-tri mesh scene for every
library variant under lightmetrica-v2_b200/lib/variants (each in a fresh process via LMB200_LIB)."""
import glob, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, ctypes as C, numpy as np, torch, json
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes
sys.path.insert(0, ROOT)
import bench
L = capi.lib()
out = {}
for name, verts in [("soup4M", scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)), ("mesh1M", scenes.mesh_scene(1000000, seed=42)[0])]:
    lo, hi = scenes.bounds(verts)
    B = os.environ.get('SWEEP_BUILDER'); A = capi.Accel(0); A.build(verts, builder=int(B) if B else None)
    n = 1 << 24
    d_rays = bench.gen_rays_device(torch, n, lo.tolist(), hi.tolist(), 7, torch.device('cuda'))
    d_hits = torch.empty((n, 4), dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    npr, tpr = C.c_double(), C.c_double()
    capi.check(L.lmb200_trace_count_dev(A.h, d_rays.data_ptr(), 1 << 22, C.byref(npr), C.byref(tpr)))
    out[name] = dict(mrays=n / ms / 1e3, nodes=npr.value, tris=tpr.value, hit=float((d_hits[:, 3].view(torch.int32) != -1).float().mean()))
    A.close()
if os.environ.get("SWEEP_PT", "1") != "0":
    from lmb200py import scenedesc
    sc = scenedesc.config2_scene(1000000, 1920, 1080)
    S = capi.Scene(sc, builder=int(os.environ.get('SWEEP_BUILDER', capi.BUILD_DEFAULT)))
    N = 1920 * 1080 * 32
    S.render(capi.MODE_PTDIRECT, N // 4, seed=1)
    best = 0
    for _ in range(2):
        img, st = S.render(capi.MODE_PTDIRECT, N, seed=1)
        best = max(best, N / st["seconds"] / 1e6)
    out["ptdirect"] = dict(msamples=best, mean=float(img.mean()))
    S.close()
out = {k: {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()} for k, v in out.items()}
print("RESULT", json.dumps(out))
''' % ROOT
libs = [os.path.join(ROOT, 'lightmetrica-v2_b200', 'lib', 'liblmb200.so')] + sorted(glob.glob(os.path.join(ROOT, 'lightmetrica-v2_b200', 'lib', 'variants', '*.so')))
if len(sys.argv) > 1:
    libs = [l for l in libs if any(a in l for a in sys.argv[1:])]
for lib in libs:
    env = dict(os.environ, LMB200_LIB=lib)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    res = [l for l in r.stdout.splitlines() if l.startswith('RESULT')]
    print(os.path.basename(lib), res[0][7:] if res else ('FAILED ' + r.stderr[-300:]), flush=True)
