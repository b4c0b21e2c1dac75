"""First GPU contact of the wavefront renderer: GPU vs C-port images on the same samples."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
from oracle import bindings as ob
from lmb200py import capi, scenedesc
sc = scenedesc.cornell_box(64, 64, glossy_block=True)
N = 64 * 64 * 64
P = ob.PortPT(sc)
G = capi.Scene(sc)
rel = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))
for mode, name in [(1, 'ptdirect'), (0, 'pt')]:
    fp, cnt = P.render(mode, N, seed=1)
    t0 = time.time(); fg, st = G.render(mode, N, seed=1, pool=1 << 16); t1 = time.time()
    print(name, "gpu %.3fs" % (t1 - t0), st)
    print("  mean port", fp.mean(axis=(0, 1)), "mean gpu", fg.mean(axis=(0, 1)), "relRMSE gpu-vs-port %.2e" % rel(fg, fp), "port rays", cnt)
    print("  max abs diff", np.abs(fg - fp).max(), "nan", np.isnan(fg).sum())
fg, st = G.render(2, 0)
fn, tri = P.render_normal()
print("normal: max abs diff", np.abs(fg - fn).max(), st)
# big run for rate
sc2 = scenedesc.cornell_box(512, 512, glossy_block=True)
G2 = capi.Scene(sc2)
for mode, name in [(1, 'ptdirect'), (0, 'pt')]:
    N2 = 512 * 512 * 64
    fg, st = G2.render(mode, N2, seed=1)
    print(name, "512x512x64: %.3fs device = %.1f Msamples/s" % (st['seconds'], N2 / st['seconds'] / 1e6), st, "mean", fg.mean(axis=(0, 1)))
