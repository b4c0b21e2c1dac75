"""Feasibility probe: two wavefront loops on ONE GPU at the same time (two scenes, two host threads, each half of the sample
range) against one loop over the whole range. If the pair is faster, the tails of one loop's persistent traversal kernels and
its half-idle shading kernels are being filled by the other loop."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'lightmetrica-v2_b200'))
import numpy as np
from lmb200py import capi, scenedesc
sc = scenedesc.config2_scene(1000000, 1920, 1080)
spp = int(os.environ.get("SPP", "256"))
N = 1920 * 1080 * spp
scenes = [capi.Scene(sc) for _ in range(3)]
for S in scenes: S.render(capi.MODE_PTDIRECT, N // 16, seed=1)
def one(S, b, e, pool, out, k):
    img, st = S.render(capi.MODE_PTDIRECT, N, seed=1, begin=b, end=e, pool=pool)
    out[k] = (img, st)
for lanes, pool in ((1, 1 << 23), (2, 1 << 22), (2, 1 << 23), (3, 1 << 22), (1, 1 << 23)):
    best = 1e9
    for rep in range(2):
        out = [None] * lanes
        th = [threading.Thread(target=one, args=(scenes[k], N * k // lanes, N * (k + 1) // lanes, pool, out, k)) for k in range(lanes)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        best = min(best, time.perf_counter() - t0)
    img = sum(o[0] for o in out)
    print(f"{lanes} loop(s), pool {pool >> 20} Mi each: {N / best / 1e6:7.1f} Msamples/s wall (incl. film readback), mean {img.mean():.5f}", flush=True)
