"""The path-tracing oracle (oracle/lm_oracle_pt.c) against images rendered by the reference itself.
The reference has no renderer tests, so the pin is statistical: tests/golden/pt_cornell.npz
holds renderer::pt / renderer::ptdirect images of the Cornell-style box from the compiled reference
(two dSFMT seeds each; their mutual relRMSE is the Monte-Carlo noise floor). CPU only."""
import os

import numpy as np
import pytest

from oracle import bindings as ob
from lmb200py import scenedesc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "pt_cornell.npz"))


@pytest.mark.parametrize("mode,name", [(1, "ptdirect"), (0, "pt"), (3, "ptmis")])
def test_port_matches_reference_images(gold, mode, name):
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    spp = 1024
    img, counts = ob.PortPT(sc).render(mode, 48 * 48 * spp, seed=5)
    ref_a, ref_b = gold[name + "_a"], gold[name + "_b"]
    ref = 0.5 * (ref_a + ref_b)
    # mean radiance: all estimators converge to the same value (SURVEY.md §6 probe); 1.5 % covers the
    # noise of pt at 1024 spp (ptdirect is ~10x tighter)
    m, mr = img.mean(axis=(0, 1)), ref.mean(axis=(0, 1))
    assert np.allclose(m, mr, rtol=0.04 if mode == 0 else 0.015), (m, mr)
    # per-pixel: error vs the reference must look like Monte-Carlo noise at this spp, i.e. about
    # floor * sqrt(spp_ref/spp) (+ the reference's own floor), not like a bias
    floor = rel_rmse(ref_a, ref_b)                      # two seeds at 16384 spp
    expected = floor / np.sqrt(2) * np.sqrt(int(gold["spp"]) / spp)
    got = rel_rmse(img, ref)
    assert got < 1.35 * expected, (got, expected)
    assert counts[0] > 0 and (counts[1] > 0) == (mode != 0)


def test_port_is_deterministic_and_shardable():
    sc = scenedesc.cornell_box(32, 32)
    P = ob.PortPT(sc)
    N = 32 * 32 * 16
    a, _ = P.render(1, N, seed=3)
    b, _ = P.render(1, N, seed=3)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)      # thread-private films are summed in arbitrary order
    h1, _ = P.render(1, N, seed=3, begin=0, end=N // 3)
    h2, _ = P.render(1, N, seed=3, begin=N // 3, end=N)
    assert np.allclose(h1 + h2, a, rtol=1e-4, atol=1e-5)


def test_port_max_vertices():
    sc = scenedesc.cornell_box(32, 32)
    P = ob.PortPT(sc)
    N = 32 * 32 * 8
    img1, c1 = P.render(1, N, seed=1, max_verts=1)      # loop-top test fires immediately (renderer_pt.cpp:120-123)
    assert c1[0] == 0 and img1.max() == 0
    img2, c2 = P.render(1, N, seed=1, max_verts=2)      # camera vertex only: one extend ray per sample
    assert c2[0] == N
    img3, c3 = P.render(1, N, seed=1)
    assert c3[0] > c2[0]


def test_normal_renderer_port():
    sc = scenedesc.cornell_box(32, 32)
    img, tri = ob.PortPT(sc).render_normal()
    assert (tri >= 0).mean() > 0.5
    hit = tri >= 0
    assert np.allclose(np.linalg.norm(img[hit], axis=-1), 1.0, atol=1e-5)
    assert (img[~hit] == 0).all()


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_vs_reference_live_glossy_scene():
    """A second scene (mesh objects, glossy + diffuse, several lights) against the live reference."""
    sc = scenedesc.config2_scene(4000, 32, 18, n_objects=12)
    N = 32 * 18 * 2048
    img, _ = ob.PortPT(sc).render(1, N, seed=1)
    R = ob.RefScene(sc)
    ra, _ = R.render("ptdirect", N, seed=1, threads=4)
    rb, _ = R.render("ptdirect", N, seed=2, threads=4)
    floor = rel_rmse(ra, rb)
    assert rel_rmse(img, ra) < 1.3 * floor, (rel_rmse(img, ra), floor)
    assert np.allclose(img.mean(axis=(0, 1)), ra.mean(axis=(0, 1)), rtol=0.02)


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mode,name", [(1, "ptdirect"), (3, "ptmis")])
def test_port_vs_reference_live_delta_bsdfs_and_point_light(mode, name):
    """bsdf::flesnel / reflect_all / refract_all and light::point against the live reference
    (bsdf_flesnel.cpp, bsdf_reflectall.cpp, bsdf_refractall.cpp, light_point.cpp)."""
    sc = scenedesc.specular_box(32, 32)
    N = 32 * 32 * 2048
    img, _ = ob.PortPT(sc).render(mode, N, seed=1)
    R = ob.RefScene(sc)
    ra, _ = R.render(name, N, seed=1, threads=4)
    rb, _ = R.render(name, N, seed=2, threads=4)
    floor = rel_rmse(ra, rb)
    assert not np.isnan(img).any()
    assert rel_rmse(img, ra) < 1.3 * floor, (rel_rmse(img, ra), floor)
    assert np.allclose(img.mean(axis=(0, 1)), 0.5 * (ra + rb).mean(axis=(0, 1)), rtol=0.03)


OUTDOOR_CASES = [("directional", False, 1, "ptdirect"), ("env", False, 1, "ptdirect"), ("both", True, 1, "ptdirect"),
                 ("directional", True, 3, "ptmis"), ("directional", True, 0, "pt"), ("cornell", True, 0, "pt"),
                 ("textured", False, 1, "ptdirect")]


@pytest.mark.parametrize("light,thin,mode,name", OUTDOOR_CASES)
def test_port_matches_reference_outdoor_images(light, thin, mode, name):
    """light::directional, light::env (constant Le) and sensor::thinlens: the port against images the reference itself
    rendered (tests/golden/pt_outdoor.npz, two dSFMT seeds at 8192 spp; make_golden.py outdoor)."""
    gold = np.load(os.path.join(GOLD, "pt_outdoor.npz"))
    key = f"{light}_{'thinlens' if thin else 'pinhole'}_{name}"
    ra, rb = gold[key + "_a"], gold[key + "_b"]
    sc = scenedesc.outdoor_scene(48, 27, light, thin)
    spp = 1024
    img, counts = ob.PortPT(sc).render(mode, 48 * 27 * spp, seed=5)
    if light == "directional" and name == "pt":
        # a delta-direction light is never hit by BSDF sampling: renderer::pt renders black, and so must we
        assert ra.max() == 0 and img.max() == 0 and counts[1] == 0
        return
    ref = 0.5 * (ra + rb)
    floor = rel_rmse(ra, rb)
    expected = floor / np.sqrt(2) * np.sqrt(int(gold["spp"]) / spp)
    assert rel_rmse(img, ref) < 1.35 * expected, (rel_rmse(img, ref), expected)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.05 if mode == 0 else 0.02)


def test_thinlens_blurs_out_of_focus_only():
    """Sanity of the lens model: with the focal plane on the sphere, the thin-lens image equals the pinhole image in
    mean radiance (same importance normalisation) but differs per pixel away from the focal plane."""
    a = scenedesc.outdoor_scene(32, 18, "directional", False)
    b = scenedesc.outdoor_scene(32, 18, "directional", True)
    N = 32 * 18 * 512
    ia, _ = ob.PortPT(a).render(1, N, seed=2)
    ib, _ = ob.PortPT(b).render(1, N, seed=2)
    assert np.allclose(ia.mean(), ib.mean(), rtol=0.03)
    assert rel_rmse(ia, ib) > 0.05


def test_tile_partitioning_port_is_statistically_the_whole_image():
    """Optional tile partitioning: camera samples of shard k drawn inside strip k of the raster. The whole
    tile (0,0,1,1) is the same sample set as no tile; four strips with a quarter of the samples each sum to an image with the
    same expectation as whole-image sampling (ptdirect's camera-vertex light splats still land anywhere)."""
    sc = scenedesc.cornell_box(32, 32)
    P = ob.PortPT(sc)
    N = 32 * 32 * 256
    full, _ = P.render(1, N, seed=3)
    same, _ = P.render(1, N, seed=3, tile=(0, 0, 1, 1))
    assert np.allclose(full, same, rtol=1e-5, atol=1e-6)      # same paths; thread-private films are summed in arbitrary order
    strips = sum(P.render(1, N, seed=3, begin=N * k // 4, end=N * (k + 1) // 4, tile=(0, k / 4, 1, (k + 1) / 4))[0] for k in range(4))
    other, _ = P.render(1, N, seed=4)
    floor = rel_rmse(full, other)
    assert rel_rmse(strips, full) < 1.25 * floor
    assert np.allclose(strips.mean(axis=(0, 1)), full.mean(axis=(0, 1)), rtol=0.02)
    # a strip's own camera rays stay inside it: with pt (no light sampling from the camera vertex) nothing lands outside
    top, _ = P.render(0, N, seed=3, begin=0, end=N // 4, tile=(0, 0, 1, 0.25))
    assert top[8:].max() == 0 and top[:8].max() > 0


@pytest.mark.skipif(not ob.have_ref(), reason="needs oracle/_ref")
def test_reference_obj_loader_route_equals_in_memory_meshes():
    """configs[0] route: the committed Cornell OBJ files through the reference's trianglemesh::obj (tinyobjloader) give the
    reference renderer exactly the scene the in-memory meshes give it (same seed, one thread => bit-identical image)."""
    obj_dir = os.path.join(os.path.dirname(__file__), "golden", "cornell_obj")
    big = scenedesc.cornell_box(512, 512, glossy_block=True)          # what the fixtures were written from
    sc = scenedesc.cornell_box(24, 24, glossy_block=True)
    assert len(sc.meshes) == len(big.meshes) and all(np.array_equal(a["verts"], b["verts"]) for a, b in zip(sc.meshes, big.meshes))
    paths = [os.path.join(obj_dir, f"mesh{i}.obj") for i in range(len(sc.meshes))]
    N = 24 * 24 * 32
    a, _ = ob.RefScene(sc, accel="qbvh", obj_paths=paths).render("ptdirect", N, seed=4, threads=1)
    b, _ = ob.RefScene(sc, accel="qbvh").render("ptdirect", N, seed=4, threads=1)
    assert np.array_equal(a, b) and a.mean() > 0


def test_port_matches_the_reference_on_the_config2_scene():
    """configs[2] (1M triangles, diffuse + glossy, area-light NEE) at 480x270, 64 spp: the C port against the reference's
    renderer::ptdirect + accel::qbvh golden (tests/golden/make_golden.py config2). Statistic: 6x6 block means of the image
    with pixels clamped at 2 (the raw means are firefly-dominated); bar: 1.25x the reference's own two-seed floor, global
    mean within 2 %."""
    from lmb200py import capi
    g = np.load(os.path.join(GOLD, "config2_480x270_blockmeans.npz"))
    ra, rb = g["ptdirect_clamped_a"], g["ptdirect_clamped_b"]
    floor = float(np.sqrt(np.mean((ra - rb) ** 2)) / np.mean(rb))
    img, _ = ob.PortPT(scenedesc.config2_scene(1_000_000, 480, 270)).render(capi.MODE_PTDIRECT, 480 * 270 * 64, seed=9)
    bm = np.minimum(img, 2.0).reshape(45, 6, 80, 6, 3).mean(axis=(1, 3))
    assert float(np.sqrt(np.mean((bm - ra) ** 2)) / np.mean(ra)) < 1.25 * floor
    assert np.allclose(bm.mean(axis=(0, 1)), 0.5 * (ra + rb).mean(axis=(0, 1)), rtol=0.02)


def test_coherent_sample_groups_are_stratified_over_the_tiles():
    """lmb200_render_params::primary_tile in the port (the GPU renderer is checked against it sample for sample,
    tests/test_gpu_render.py::test_coherent_camera_sample_groups): groups of 32 samples share a tile, consecutive groups visit
    ALL tiles once per round in a keyed pseudo-random order. Counted with a scene in which every camera ray hits an emitter of
    radiance 1 (pt, two vertices), so the film is the per-pixel sample count: complete rounds put exactly the same number of
    samples into every tile (the tile order is a bijection), a partial round at most one group more, any split of the sample
    range adds up to the same film, and independent positions (primary_tile < 0) do none of this."""
    W = H = 32
    s = scenedesc.Scene()
    s.add_bsdf("black", "diffuse", (0, 0, 0))
    s.add_light("wall", (1.0, 1.0, 1.0))
    s.add_quad((-10, -10, 0), (10, -10, 0), (10, 10, 0), (-10, 10, 0), "black", "wall")      # fills the view, faces the camera
    s.set_camera((0, 0, 3), (0, 0, 0), (0, 1, 0), 40.0, W, H)
    P = ob.PortPT(s)

    def counts(N, tp, seed=5, **kw):
        img, _ = P.render(0, N, seed=seed, max_verts=2, primary_tile=tp, **kw)
        c = img[..., 0] * (N / (W * H))
        assert np.allclose(c, np.round(c), atol=1e-3)
        return np.round(c)

    for tp in (4, 8):
        tiles = (W // tp) * (H // tp)
        rounds = 4
        N = rounds * 32 * tiles
        c = counts(N, tp)
        per_tile = c.reshape(H // tp, tp, W // tp, tp).sum(axis=(1, 3))
        assert c.sum() == N and (per_tile == rounds * 32).all()
        # a partial round: every tile has its complete rounds, some one group more
        N2 = N + 32 * (tiles // 3) + 7
        c2 = counts(N2, tp)
        pt2 = c2.reshape(H // tp, tp, W // tp, tp).sum(axis=(1, 3))
        assert c2.sum() == N2 and pt2.min() >= rounds * 32 and pt2.max() <= (rounds + 1) * 32
        # sharding the sample range (what the GPUs of one box do) gives the same counts
        parts = sum(P.render(0, N, seed=5, max_verts=2, primary_tile=tp, begin=N * k // 3, end=N * (k + 1) // 3)[0][..., 0] for k in range(3))
        assert np.allclose(parts * (N / (W * H)), c, atol=1e-3)
        # another seed: another order and other positions, the same stratification
        c3 = counts(N, tp, seed=6)
        assert (c3 != c).any() and (c3.reshape(H // tp, tp, W // tp, tp).sum(axis=(1, 3)) == rounds * 32).all()
    free = counts(4 * 32 * 64, -1)
    ft = free.reshape(8, 4, 8, 4).sum(axis=(1, 3))
    assert free.sum() == 4 * 32 * 64 and ft.min() < 128 < ft.max()
