import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
from lmb200py import capi, scenedesc
L = capi.lib()
print("nccl lib:", os.environ.get("LMB200_NCCL_LIB"))
sc = scenedesc.cornell_box(48, 48, glossy_block=True)
N = 48 * 48 * 64
S = [capi.Scene(sc, device=g) for g in range(2)]
one, _ = S[0].render(capi.MODE_PTDIRECT, N, seed=3)
oneb, _ = S[1].render(capi.MODE_PTDIRECT, N, seed=3)
print("one mean", one.mean(), "on gpu1", oneb.mean(), "max diff", np.abs(one - oneb).max())
halves = [S[g].render(capi.MODE_PTDIRECT, N, seed=3, begin=N * g // 2, end=N * (g + 1) // 2)[0] for g in range(2)]
print("halves", halves[0].mean(), halves[1].mean(), "sum diff", np.abs(halves[0] + halves[1] - one).max())
arr = (C.c_void_p * 2)(*[s.h_ for s in S])
for trial in range(3):
    p = S[0].params(capi.MODE_PTDIRECT, N, seed=3)
    film = np.zeros((48, 48, 4), np.float32)
    st = capi.RenderStats()
    rc = L.lmb200_render_multi(arr, 2, C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(st))
    print("multi rc", rc, "mean", film[..., :3].mean(), "max diff", np.abs(film[..., :3] - one).max(), "samples", st.samples, "ext", st.extend_rays, "reduce_s", st.reduce_seconds)
    for g in range(2):
        print("   vs half", g, np.abs(film[..., :3] - halves[g]).max())
