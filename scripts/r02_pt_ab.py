"""A/B of library variants on the renderer: ptdirect / pt / ptmis Msamples/s on configs[2] at 64 spp and configs[0]-like Cornell box
(each library in a fresh process via LMB200_LIB)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenedesc
out = []
for name, sc, N in (("configs[2]", scenedesc.config2_scene(1000000, 1920, 1080), 1920 * 1080 * 64), ("cornell 1024^2", scenedesc.cornell_box(1024, 1024, glossy_block=True), 1024 * 1024 * 256)):
    S = capi.Scene(sc)
    S.render(capi.MODE_PTDIRECT, N // 8, seed=1)
    for mode, mname in ((capi.MODE_PTDIRECT, "ptdirect"), (capi.MODE_PT, "pt"), (capi.MODE_PTMIS, "ptmis")):
        best = 0
        for _ in range(2):
            img, st = S.render(mode, N, seed=1)
            best = max(best, N / st["seconds"] / 1e6)
        out.append("%%s %%s %%.0f" %% (name, mname, best))
    S.close()
print("; ".join(out))
''' % ROOT
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "lightmetrica-v2_b200", "lib", "variants", "*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib: env["LMB200_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib) if lib else "liblmb200.so", (r.stdout.strip().splitlines() or [r.stderr[-400:]])[-1], flush=True)
