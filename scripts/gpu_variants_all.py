"""Tuning aid: for each library variant (fresh process each): 4M-tri soup Mrays/s, 1M-tri mesh random-ray Mrays/s (+ nodes / tris
per ray) and ptdirect / pt Msamples/s on the configs[2] scene."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, ctypes as C, numpy as np, torch
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes, scenedesc
L = capi.lib()
def rate(acc, rays, reps=5):
    n = len(rays); dr = torch.from_numpy(rays).cuda(); dh = torch.empty((n, 4), dtype=torch.float32, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3): capi.check(L.lmb200_trace_closest_dev(acc, dr.data_ptr(), dh.data_ptr(), n, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): capi.check(L.lmb200_trace_closest_dev(acc, dr.data_ptr(), dh.data_ptr(), n, s))
    e1.record(); torch.cuda.synchronize()
    npr, tpr = C.c_double(), C.c_double()
    capi.check(L.lmb200_trace_count_dev(acc, dr.data_ptr(), min(n, 1 << 22), C.byref(npr), C.byref(tpr)))
    return n / (e0.elapsed_time(e1) / reps) / 1e3, npr.value, tpr.value
out = []
if "--no-soup" not in sys.argv:
    verts = scenes.soup(4000000, seed=42, extent=100.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    A = capi.Accel(0); A.build(verts)
    r = rate(A.h, scenes.random_rays(1 << 24, lo, hi, seed=7))
    out.append("soup4M %%.0f (%%.1f n, %%.2f t)" %% r)
    A.close()
sc = scenedesc.config2_scene(1000000, 1920, 1080)
S = capi.Scene(sc)
acc = L.lmb200_scene_accel(S.h_)
lo, hi = scenes.bounds(S.keep["verts"])
r = rate(acc, scenes.random_rays(1 << 23, lo, hi, seed=7))
out.append("mesh1M %%.0f (%%.1f n, %%.2f t)" %% r)
c = sc.camera
r = rate(acc, scenes.camera_rays(c['eye'], c['center'], c['up'], c['fov'], 1920, 1080)[np.random.default_rng(1).permutation(1920 * 1080)])
out.append("primary(shuffled) %%.0f (%%.1f n, %%.2f t)" %% r)
W, H, spp = 1920, 1080, 32
N = W * H * spp
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda'); st = capi.RenderStats()
for mode, name in ((capi.MODE_PTDIRECT, "ptdirect"), (capi.MODE_PT, "pt")):
    best = 0
    for rep in range(3):
        film.zero_()
        p = S.params(mode, N, seed=1)
        capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)))
        best = max(best, N / st.seconds / 1e6)
    out.append("%%s %%.0f" %% (name, best))
print(os.path.basename(os.environ.get("LMB200_LIB", "default")), " | ".join(out), flush=True)
''' % ROOT
args = [a for a in sys.argv[1:] if not a.startswith("--")]
flags = [a for a in sys.argv[1:] if a.startswith("--")]
for v in args or ["default"]:
    env = dict(os.environ)
    if v != "default":
        env["LMB200_LIB"] = os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "variants", f"liblmb200_{v}.so")
    subprocess.run([sys.executable, "-c", CHILD] + flags, env=env)
