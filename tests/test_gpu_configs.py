"""BASELINE.json's configurations as parity cases at (or near) their stated sizes. configs[3] (the bench workload)
is covered by tests/test_gpu_trace.py::test_full_size_properties."""
import os

import numpy as np
import pytest

from oracle import bindings as ob
from lmb200py import capi, scenedesc, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


def block_means(img, b=8):
    h, w, c = img.shape
    return img.reshape(h // b, b, w // b, b, c).mean(axis=(1, 3))


@pytest.mark.parametrize("mode,name", [(capi.MODE_PTDIRECT, "ptdirect"), (capi.MODE_PT, "pt")])
def test_config0_cornell_512x512_64spp(mode, name):
    """configs[0]: Cornell box, 512x512, 64 spp, against the reference path tracer's CPU render of the same size
    (golden: 8x8 block means of two reference seeds). Bar: block-mean relRMSE <= 1.25x the reference's own
    two-seed floor at equal spp; a 1024-spp GPU render must be closer to the mean of the two references than they
    are to each other."""
    g = np.load(os.path.join(GOLD, "config0_cornell512_blockmeans.npz"))
    ra, rb = g[name + "_a"], g[name + "_b"]
    floor = rel_rmse(ra, rb)
    sc = scenedesc.cornell_box(512, 512, glossy_block=True)
    S = capi.Scene(sc)
    img64, st = S.render(mode, 512 * 512 * 64, seed=21)
    assert st["samples"] == 512 * 512 * 64
    assert rel_rmse(block_means(img64), ra) < 1.25 * floor
    img1k, _ = S.render(mode, 512 * 512 * 1024, seed=22)
    ref = 0.5 * (ra + rb)
    assert rel_rmse(block_means(img1k), ref) < 0.85 * floor
    assert np.allclose(img1k.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)


def test_config1_primary_rays_256k_tris_1080p():
    """configs[1]: primary-ray normal renderer on a procedurally generated 256k-triangle mesh, 1920x1080, 1 spp.
    Per-pixel triangle index exact vs the oracle, image max-abs <= 1e-6 (bar was 1e-4)."""
    verts, _ = scenes.mesh_scene(256_000, seed=42, half=10.0, n_objects=60)
    sc = scenedesc.Scene()
    sc.add_bsdf("w", "diffuse", (0.7, 0.7, 0.7))
    sc.add_mesh_tris(verts, "w")
    sc.set_camera((0.0, 6.0, 13.0), (0.0, 1.0, 0.0), (0, 1, 0), 45.0, 1920, 1080)
    img, st = capi.Scene(sc).render(capi.MODE_NORMAL, 0)
    ref, tri = ob.PortPT(sc).render_normal()
    assert st["extend_rays"] == 1920 * 1080
    assert (tri >= 0).mean() > 0.5
    assert np.abs(img - ref).max() <= 1e-6
    # the same rays through the batch API give the same triangle per pixel
    A = capi.Accel(0)
    A.build(verts)
    rays = scenes.camera_rays((0.0, 6.0, 13.0), (0.0, 1.0, 0.0), (0, 1, 0), 45.0, 1920, 1080)
    hits = A.trace_closest(rays)
    t2 = hits["tri"].astype(np.int64)
    t2[t2 == capi.MISS] = -1
    # camera_rays (numpy) and the renderer's ray generation differ in the last ulp of the direction: allow silhouettes
    assert (t2.reshape(1080, 1920) != tri).mean() < 2e-3


def test_config2_scene_reduced_vs_oracle_and_full_size_smoke():
    """configs[2]: 1M-triangle scene, diffuse + glossy, area-light NEE. The oracle is too slow at 1080p x 1024 spp,
    so: (a) same-sample parity with the oracle on a 480x270 view at 4 spp on the FULL 1M-triangle scene,
    (b) at full resolution the image is finite, non-negative and energy-consistent across sample-range shards."""
    sc = scenedesc.config2_scene(1_000_000, 480, 270)
    N = 480 * 270 * 4
    gpu, st = capi.Scene(sc).render(capi.MODE_PTDIRECT, N, seed=5)
    port, counts = ob.PortPT(sc).render(capi.MODE_PTDIRECT, N, seed=5)
    assert abs(st["extend_rays"] - counts[0]) <= 1e-4 * counts[0] + 4
    assert rel_rmse(gpu, port) < 2e-3
    sc_full = scenedesc.config2_scene(1_000_000, 1920, 1080)
    S = capi.Scene(sc_full)
    Nf = 1920 * 1080 * 8
    full, st = S.render(capi.MODE_PTDIRECT, Nf, seed=1)
    assert np.isfinite(full).all() and (full >= 0).all() and st["samples"] == Nf
    halves = [S.render(capi.MODE_PTDIRECT, Nf, seed=1, begin=Nf * g // 2, end=Nf * (g + 1) // 2)[0] for g in range(2)]
    assert np.allclose(halves[0] + halves[1], full, rtol=1e-3, atol=1e-4)


def test_config2_scene_against_the_reference_renderer():
    """configs[2] against the REFERENCE (not the port): renderer::ptdirect + accel::qbvh rendered the full 1M-triangle
    scene on the CPU at 480x270, 64 spp, two seeds (golden: 6x6 block means, tests/golden/make_golden.py config2).
    The glossy surfaces throw fireflies at 64 spp (two-seed relRMSE of the raw block means: 0.65), so the statistic is
    the block means of the 64-spp image with pixels clamped at 2 (two-seed floor 0.18). Bars: one GPU image is within
    1.25x the floor of a reference image; the average of 16 independent clamped GPU images is within 0.7x the floor of
    the mean of the two references (noise alone gives 0.53x) and the global means agree to 1.5 %; the unclamped global
    means agree to 3 %."""
    g = np.load(os.path.join(GOLD, "config2_480x270_blockmeans.npz"))
    ra, rb = g["ptdirect_clamped_a"], g["ptdirect_clamped_b"]
    floor = rel_rmse(ra, rb)
    S = capi.Scene(scenedesc.config2_scene(1_000_000, 480, 270))
    N = 480 * 270 * 64
    imgs = [S.render(capi.MODE_PTDIRECT, N, seed=31 + k)[0] for k in range(16)]
    clamped = [block_means(np.minimum(im, 2.0), 6) for im in imgs]
    assert rel_rmse(clamped[0], ra) < 1.25 * floor
    ref = 0.5 * (ra + rb)
    avg = np.mean(clamped, axis=0)
    assert rel_rmse(avg, ref) < 0.7 * floor
    assert np.allclose(avg.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.015)
    raw_ref = 0.5 * (g["ptdirect_a"] + g["ptdirect_b"]).mean(axis=(0, 1))
    assert np.allclose(np.mean(imgs, axis=0).mean(axis=(0, 1)), raw_ref, rtol=0.03)


def test_config4_10m_instanced_triangles_hbm_sizing():
    """configs[4]: 10M-triangle "instanced" scene (flattened to world space like the reference, accel_qbvh.cpp:161-194),
    4K film. Checks the HBM-resident sizing: device build of 10M triangles, 3840x2160 film, a short ptdirect run, and
    closest-hit parity with the oracle on a ray sample."""
    sc, verts = scenedesc.config4_scene()
    assert 9_500_000 < len(verts) < 10_500_000
    S = capi.Scene(sc, builder=capi.BUILD_GPU_LBVH)
    st = capi.AccelStats()
    capi.check(capi.lib().lmb200_accel_get_stats(capi.lib().lmb200_scene_accel(S.h_), st))
    assert st.num_valid_triangles == len(verts) + 2
    assert st.node_bytes + st.tri_bytes < 2.5e9
    img, rs = S.render(capi.MODE_PTDIRECT, 3840 * 2160 * 2, seed=1)
    assert np.isfinite(img).all() and img.mean() > 0
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(20000, lo, hi, seed=9)
    A = capi.Accel(0)
    A.build(np.concatenate([verts, np.zeros((0, 9), np.float32)]), builder=capi.BUILD_GPU_LBVH)
    hits = A.trace_closest(rays)
    tuv, tri = ob.PortScene(verts).closest(rays)
    gt = hits["tri"].astype(np.int64)
    gt[gt == capi.MISS] = -1
    assert np.array_equal(gt, tri)
    assert np.array_equal(np.stack([hits["t"], hits["u"], hits["v"]], axis=1).view(np.uint32), tuv.view(np.uint32))
