"""ptdirect Msamples/s on the configs[2] scene: pool size sweep at the automatic primary tile (64 spp, default builder)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
N = 1920 * 1080 * 64
S = capi.Scene(sc)
S.render(capi.MODE_PTDIRECT, N // 8, seed=1)
print("auto tile:", capi.lib().lmb200_default_primary_tile(1920, 1080, N))
def run(**kw):
    best = 0
    for _ in range(2):
        img, st = S.render(capi.MODE_PTDIRECT, N, seed=1, **kw)
        best = max(best, N / st["seconds"] / 1e6)
    return best, st
for pool in (1 << 22, 1 << 23, 1 << 24, 3 << 22):
    r, st = run(pool=pool)
    print(f"pool {pool / (1 << 20):.0f} Mi: {r:7.1f} Msamples/s", flush=True)
for mode, name in ((capi.MODE_PT, "pt"), (capi.MODE_PTMIS, "ptmis")):
    best = 0
    for _ in range(2):
        img, st = S.render(mode, N, seed=1)
        best = max(best, N / st["seconds"] / 1e6)
    print(f"{name}: {best:7.1f} Msamples/s", flush=True)
