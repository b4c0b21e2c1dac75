"""Tuning aid: ptdirect rate on the configs[2] scene vs wavefront pool size."""
import sys, os, ctypes as C, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
S = capi.Scene(sc)
W, H, spp = 1920, 1080, 16
N = W * H * spp
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
L = capi.lib(); st = capi.RenderStats()
for mode in (capi.MODE_PTDIRECT, capi.MODE_PT):
    for pool in (1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25):
        for rep in range(2):
            film.zero_()
            p = S.params(mode, N, seed=1, pool=pool)
            capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)))
        print("mode", mode, "pool", pool, "%.1f ms  %.0f Msamples/s  %.0f Mrays/s  iters %d" % (st.seconds * 1e3, N / st.seconds / 1e6, (st.extend_rays + st.shadow_rays) / st.seconds / 1e6, st.iterations), flush=True)
