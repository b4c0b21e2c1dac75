import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import bindings as ob
from lmb200py import scenedesc, capi
PLUG = os.path.join(ROOT, "lightmetrica-v2_b200", "plugin")
L = ob.ref()
assert L.ref_load_plugin(os.path.join(PLUG, "accel_lmb200").encode()) == 1
assert L.ref_load_plugin(os.path.join(PLUG, "renderer_lmb200pt").encode()) == 1
sc = scenedesc.textured_box(32, 32)
N = 32 * 32 * 2048
R = ob.RefScene(sc, accel="qbvh")
ours, _ = R.render("lmb200pt", N, seed=1, extra={"mode": "ptdirect", "texture_resolution": 256}, in_tree=True)
direct, _ = capi.Scene(sc).render(capi.MODE_PTDIRECT, N, seed=1)
ra, _ = R.render("ptdirect", N, seed=1, threads=os.cpu_count() or 1)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dbg_tex.npz"), ours=ours, direct=direct, ra=ra)
print("means", ours.mean(axis=(0,1)), direct.mean(axis=(0,1)), ra.mean(axis=(0,1)))
