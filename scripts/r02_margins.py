"""Prints how far the statistical configs[2]-vs-reference test is from its bars (run on a GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
from lmb200py import capi, scenedesc
g = np.load(os.path.join(ROOT, "tests", "golden", "config2_480x270_blockmeans.npz"))
ra, rb = g["ptdirect_clamped_a"], g["ptdirect_clamped_b"]
rr = lambda a, b: float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))
bm = lambda im: im.reshape(45, 6, 80, 6, 3).mean(axis=(1, 3))
floor = rr(ra, rb)
S = capi.Scene(scenedesc.config2_scene(1_000_000, 480, 270))
N = 480 * 270 * 64
for base in (31, 131, 231):
    imgs = [S.render(capi.MODE_PTDIRECT, N, seed=base + k)[0] for k in range(16)]
    cl = [bm(np.minimum(im, 2.0)) for im in imgs]
    ref = 0.5 * (ra + rb); avg = np.mean(cl, axis=0)
    raw_ref = 0.5 * (g["ptdirect_a"] + g["ptdirect_b"]).mean(axis=(0, 1))
    print("seed base", base, "floor %.4f" % floor, "single/floor %.3f (bar 1.25)" % (max(rr(c, ra) for c in cl) / floor),
          "avg/floor %.3f (bar 0.7)" % (rr(avg, ref) / floor), "clamped mean ratio", (avg.mean(axis=(0, 1)) / ref.mean(axis=(0, 1))).round(4),
          "raw mean ratio", (np.mean(imgs, axis=0).mean(axis=(0, 1)) / raw_ref).round(4))
