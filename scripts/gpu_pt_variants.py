"""Tuning aid: ptdirect / pt rate on the configs[2] scene for each library variant given on the command line
(each in a fresh process: LMB200_LIB selects the .so)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, ctypes as C, torch
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
S = capi.Scene(sc)
W, H, spp = 1920, 1080, 32
N = W * H * spp
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
L = capi.lib(); st = capi.RenderStats()
out = []
for mode in (capi.MODE_PTDIRECT, capi.MODE_PT, capi.MODE_PTMIS):
    best = 0
    for rep in range(3):
        film.zero_()
        p = S.params(mode, N, seed=1)
        capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)))
        best = max(best, N / st.seconds / 1e6)
    out.append("mode %%d: %%.0f Msamples/s" %% (mode, best))
print(os.environ.get("LMB200_LIB", "default"), " | ".join(out), flush=True)
''' % ROOT
for v in sys.argv[1:] or ["default"]:
    env = dict(os.environ)
    if v != "default":
        env["LMB200_LIB"] = os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "variants", f"liblmb200_{v}.so")
    subprocess.run([sys.executable, "-c", CHILD], env=env)
