"""ptdirect Msamples/s on the configs[2] scene for each BVH builder (32 spp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
N = 1920 * 1080 * 32
for name, b in [("host_sah", capi.BUILD_HOST_SAH), ("gpu_lbvh", capi.BUILD_GPU_LBVH), ("gpu_lbvh_sah", capi.BUILD_GPU_LBVH_SAH), ("gpu_ploc", capi.BUILD_GPU_PLOC)]:
    S = capi.Scene(sc, builder=b)
    S.render(capi.MODE_PTDIRECT, N // 4, seed=1)
    best = 0
    for _ in range(3):
        img, st = S.render(capi.MODE_PTDIRECT, N, seed=1)
        best = max(best, N / st["seconds"] / 1e6)
    p = S.params(capi.MODE_PTDIRECT, N // 4, seed=1); 
    print(f"{name}: ptdirect {best:.1f} Msamples/s, mean {img.mean():.5f}", flush=True)
    S.close()
