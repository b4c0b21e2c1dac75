"""Generates the golden vectors under tests/golden/ from the REFERENCE ITSELF (oracle/_ref,
compiled from /root/reference by oracle/ref/build_ref.sh). Run here, commit the .npz files:
they pin the C oracle (and through it the CUDA path) on machines without /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
from oracle import bindings as ob  # noqa: E402
from lmb200py import scenes  # noqa: E402


def cornell_obj():
    """configs[0] says "Cornell box (tinyobjloader)": the Cornell box of scenedesc.cornell_box as Wavefront OBJ files, one
    per mesh, for the reference's own trianglemesh::obj loader (tests/test_gpu_plugin.py::test_config0_through_the_obj_loader)."""
    from lmb200py import scenedesc
    sc = scenedesc.cornell_box(512, 512, glossy_block=True)
    paths = sc.write_obj(os.path.join(HERE, "cornell_obj"))
    print("cornell_obj:", len(paths), "files,", sum(os.path.getsize(p) for p in paths), "bytes")


def accel_golden():
    # (1) random soup, the StubTriangleMesh_Random recipe (test_accel3.cpp:191-224)
    verts = scenes.soup(3000, seed=11, extent=4.0, edge=0.25)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(20000, lo, hi, seed=5)
    # range-limited and tmin=0 variants, as the reference tests use (test_accel3.cpp:299)
    rays[5000:10000, 7] = 1.5
    rays[10000:12000, 3] = 0.0
    recs = ob.ref_triaccel_records(verts)
    out = {}
    for accel in ("qbvh", "naive", "bvh_sahbin"):
        R = ob.RefSoup(verts, accel)
        r = R.intersect(rays, threads=1)
        out[accel] = r
    for a in ("naive", "bvh_sahbin"):
        assert np.array_equal(out["qbvh"]["face"], out[a]["face"]), a
        assert np.array_equal(out["qbvh"]["tuv"].view(np.uint32), out[a]["tuv"].view(np.uint32)), a
    q = out["qbvh"]
    np.savez_compressed(os.path.join(HERE, "accel_soup.npz"), verts=verts, rays=rays, records=recs[:, :10],
                        face=q["face"], tuv=q["tuv"], geom=q["geom"])
    print("accel_soup.npz:", int((q["face"] >= 0).sum()), "hits of", len(rays))




def pt_golden():
    """Reference images of the Cornell-style box (configs[0] geometry at 48x48) rendered by the
    reference's own renderer::pt / renderer::ptdirect / renderer::ptmis with accel::qbvh and dSFMT, two seeds each
    (the second seed measures the Monte-Carlo noise floor)."""
    from lmb200py import scenedesc
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    R = ob.RefScene(sc, "qbvh")
    spp = 16384
    N = 48 * 48 * spp
    out = {}
    for name in ("ptdirect", "pt", "ptmis"):
        a, _ = R.render(name, N, seed=1, threads=8)
        b, _ = R.render(name, N, seed=2, threads=8)
        out[name + "_a"] = a
        out[name + "_b"] = b
        rel = float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(a))
        print(name, "mean", a.mean(axis=(0, 1)), "two-seed relRMSE", rel)
    np.savez_compressed(os.path.join(HERE, "pt_cornell.npz"), spp=spp, **out)


def config0_golden():
    """BASELINE configs[0]: Cornell box, reference path tracers at 512x512, 64 spp on the CPU (the PR1
    equivalence reference). Stored as 8x8 block means (64x64x3) so the fixture stays small; block means are
    what the GPU test compares (per-pixel noise at 64 spp is ~40 %)."""
    from lmb200py import scenedesc
    sc = scenedesc.cornell_box(512, 512, glossy_block=True)
    R = ob.RefScene(sc, "qbvh")
    N = 512 * 512 * 64
    out = {}
    for name in ("ptdirect", "pt"):
        for seed, tag in ((1, "a"), (2, "b")):
            img, sec = R.render(name, N, seed=seed, threads=8)
            out[f"{name}_{tag}"] = img.reshape(64, 8, 64, 8, 3).mean(axis=(1, 3)).astype(np.float32)
            print(name, seed, "%.1f s" % sec, img.mean(axis=(0, 1)))
    np.savez_compressed(os.path.join(HERE, "config0_cornell512_blockmeans.npz"), spp=64, **out)


def config2_golden():
    """BASELINE configs[2]: the 1M-triangle diffuse + glossy scene with area-light NEE, reference renderer::ptdirect +
    accel::qbvh on the CPU, 480x270 view, 64 spp, two seeds. Stored as 6x6 block means (45x80x3) of the image and of the image with pixels clamped at 2."""
    from lmb200py import scenedesc
    sc = scenedesc.config2_scene(1_000_000, 480, 270)
    R = ob.RefScene(sc, "qbvh")
    N = 480 * 270 * 64
    out = {}
    for seed, tag in ((1, "a"), (2, "b")):
        img, sec = R.render("ptdirect", N, seed=seed, threads=16)
        out[f"ptdirect_{tag}"] = img.reshape(45, 6, 80, 6, 3).mean(axis=(1, 3)).astype(np.float32)
        # the glossy surfaces throw fireflies at 64 spp: the pixel-clamped image is the robust statistic the tests compare
        out[f"ptdirect_clamped_{tag}"] = np.minimum(img, 2.0).reshape(45, 6, 80, 6, 3).mean(axis=(1, 3)).astype(np.float32)
        print("config2", seed, "%.1f s" % sec, img.mean(axis=(0, 1)), np.minimum(img, 2.0).mean(axis=(0, 1)))
    np.savez_compressed(os.path.join(HERE, "config2_480x270_blockmeans.npz"), spp=64, **out)


OUTDOOR_CASES = [("directional", False, "ptdirect"), ("env", False, "ptdirect"), ("both", True, "ptdirect"),
                 ("directional", True, "ptmis"), ("directional", True, "pt"), ("cornell", True, "pt"),
                 ("textured", False, "ptdirect")]


def outdoor_golden():
    """Reference images of the open scene lit by light::directional / light::env and seen through sensor::pinhole or
    sensor::thinlens (scenedesc.outdoor_scene at 48x27), two dSFMT seeds each. renderer::pt sees nothing of a directional
    light (delta direction, never hit): its golden image pins exactly that (all zero)."""
    from lmb200py import scenedesc
    spp = 8192
    out = {}
    for light, thin, name in OUTDOOR_CASES:
        sc = scenedesc.outdoor_scene(48, 27, light, thin)
        R = ob.RefScene(sc, "qbvh")
        N = 48 * 27 * spp
        key = f"{light}_{'thinlens' if thin else 'pinhole'}_{name}"
        a, _ = R.render(name, N, seed=1, threads=8)
        b, _ = R.render(name, N, seed=2, threads=8)
        out[key + "_a"], out[key + "_b"] = a, b
        print(key, "mean", a.mean(axis=(0, 1)), "two-seed relRMSE", float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(a), 1e-30)))
    np.savez_compressed(os.path.join(HERE, "pt_outdoor.npz"), spp=spp, **out)


if __name__ == "__main__":
    cornell_obj()
    which = sys.argv[1:] or ["accel", "pt"]
    if "accel" in which:
        accel_golden()
    if "pt" in which:
        pt_golden()
    if "config0" in which:
        config0_golden()
    if "outdoor" in which:
        outdoor_golden()
    if "config2" in which:
        config2_golden()
