"""GPU parity of the wavefront path tracer, through the C ABI, against the C oracle (same counter-based
random numbers => same paths) and against images rendered by the reference itself (statistical)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import bindings as ob
from lmb200py import capi, scenedesc, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


def same_image(a, b):
    """Same samples, different fp32 summation order (atomic splats, per-GPU partial films): equal up to rounding of the
    accumulation, which scales with the brightest pixels (two runs on ONE GPU differ by ~1e-5 of the image maximum)."""
    return np.allclose(a, b, rtol=2e-4, atol=2e-5 * max(float(np.max(b)), 1.0))


@pytest.mark.parametrize("mode", [capi.MODE_PTDIRECT, capi.MODE_PT, capi.MODE_PTMIS])
@pytest.mark.parametrize("scene_name", ["cornell", "config2", "specular"])
def test_same_samples_as_oracle(mode, scene_name):
    """Same seed, same sample indices: the GPU image equals the oracle's up to libm-vs-CUDA ulps in
    sin/cos (tolerance: relRMSE 1e-3, three orders below the Monte-Carlo noise) and traces the same rays."""
    sc = {"cornell": lambda: scenedesc.cornell_box(64, 64, glossy_block=True),
          "config2": lambda: scenedesc.config2_scene(30000, 64, 36, n_objects=40),
          "specular": lambda: scenedesc.specular_box(64, 64)}[scene_name]()     # delta BSDFs + light::point
    w, h = sc.camera["w"], sc.camera["h"]
    N = w * h * 32
    port, counts = ob.PortPT(sc).render(mode, N, seed=7)
    gpu, st = capi.Scene(sc).render(mode, N, seed=7, pool=1 << 15)
    assert not np.isnan(gpu).any()
    # per pixel: fp32 summation order differs (device atomics vs per-thread partial films), which matters only for
    # pixels that collect thousands of equal splats (the camera-vertex NEE of a point light lands in ONE pixel);
    # at most 0.1 % of the pixels may differ beyond that (paths flipped by a sin/cos ulp)
    close = np.abs(gpu - port) <= 2e-3 + 1e-3 * np.abs(port)
    assert close.all(axis=2).mean() >= 0.999
    dim = port.max(axis=2) < 50 * port.mean()
    assert rel_rmse(gpu[dim], port[dim]) < 1e-3, rel_rmse(gpu[dim], port[dim])
    assert abs(st["extend_rays"] - counts[0]) <= max(4, 1e-4 * counts[0])
    assert abs(st["shadow_rays"] - counts[1]) <= max(4, 1e-4 * counts[1])
    assert st["samples"] == N


@pytest.mark.parametrize("primary_tile", [-1, 4, 16])
def test_coherent_camera_sample_groups(primary_tile):
    """lmb200_render_params::primary_tile: groups of 32 samples share a random tile of the image (coherent primary rays). Same
    samples as the oracle for every setting (off, 4 and 16 pixels; the default of 8 is what every other test runs), and the
    setting does not change the expected image: a grouped render agrees with an ungrouped one of another seed within the
    ungrouped two-seed noise."""
    sc = scenedesc.cornell_box(48, 40, glossy_block=True)
    N = 48 * 40 * 64
    S = capi.Scene(sc)
    port, counts = ob.PortPT(sc).render(capi.MODE_PTDIRECT, N, seed=7, primary_tile=primary_tile)
    gpu, st = S.render(capi.MODE_PTDIRECT, N, seed=7, primary_tile=primary_tile)
    assert rel_rmse(gpu, port) < 1e-3 and st["extend_rays"] == counts[0] and st["shadow_rays"] == counts[1]
    a, _ = S.render(capi.MODE_PTDIRECT, 4 * N, seed=1, primary_tile=-1)
    b, _ = S.render(capi.MODE_PTDIRECT, 4 * N, seed=2, primary_tile=-1)
    g, _ = S.render(capi.MODE_PTDIRECT, 4 * N, seed=3, primary_tile=primary_tile)
    assert rel_rmse(g, a) < 1.25 * rel_rmse(b, a)
    assert np.allclose(g.mean(axis=(0, 1)), a.mean(axis=(0, 1)), rtol=0.02)


def test_pool_size_and_sharding_invariance():
    """The image depends only on (seed, sample index): any pool size and any split of the sample range
    (= any GPU count) gives the same film up to fp32 summation order."""
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    S = capi.Scene(sc)
    N = 48 * 48 * 64
    a, _ = S.render(capi.MODE_PTDIRECT, N, seed=3, pool=1 << 12)
    b, _ = S.render(capi.MODE_PTDIRECT, N, seed=3, pool=1 << 17)
    assert np.allclose(a, b, rtol=2e-4, atol=1e-5)
    parts = [S.render(capi.MODE_PTDIRECT, N, seed=3, begin=N * g // 4, end=N * (g + 1) // 4)[0] for g in range(4)]
    assert np.allclose(sum(parts), a, rtol=2e-4, atol=1e-5)
    c, _ = S.render(capi.MODE_PTDIRECT, N, seed=4)
    assert rel_rmse(c, a) > 1e-2          # a different seed is a different sample set


def test_render_calls_from_two_threads_on_one_scene_take_turns():
    """A scene owns one path pool, film and set of counters: concurrent lmb200_render calls on it are serialised inside the
    library and each returns the image of its own parameters."""
    import threading
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    S = capi.Scene(sc)
    N = 48 * 48 * 64
    want = [S.render(capi.MODE_PTDIRECT, N, seed=11 + k)[0] for k in range(3)]
    got = [None] * 3

    def work(k):
        for _ in range(3):
            got[k] = S.render(capi.MODE_PTDIRECT, N, seed=11 + k)[0]
    th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    for t in th: t.start()
    for t in th: t.join()
    for k in range(3):
        assert same_image(got[k], want[k])


def test_max_num_vertices_and_empty_range():
    sc = scenedesc.cornell_box(32, 32)
    S = capi.Scene(sc)
    P = ob.PortPT(sc)
    N = 32 * 32 * 8
    for mv in (1, 2, 3):
        g, st = S.render(capi.MODE_PTDIRECT, N, seed=1, max_verts=mv)
        p, counts = P.render(capi.MODE_PTDIRECT, N, seed=1, max_verts=mv)
        assert st["extend_rays"] == counts[0]
        assert np.allclose(g, p, rtol=1e-3, atol=1e-5)
    g, st = S.render(capi.MODE_PT, N, seed=1, begin=5, end=5)
    assert st["samples"] == 0 and g.max() == 0


@pytest.mark.parametrize("mode,name", [(capi.MODE_PTDIRECT, "ptdirect"), (capi.MODE_PT, "pt"), (capi.MODE_PTMIS, "ptmis")])
def test_matches_reference_images(mode, name):
    """Converged renders against the reference's own renderer (golden images from oracle/_ref):
    stated bar = relRMSE at equal spp no larger than 1.25x the reference's two-seed noise floor, and
    mean radiance within 0.5 % (ptdirect, ptmis) / 2 % (pt)."""
    gold = np.load(os.path.join(GOLD, "pt_cornell.npz"))
    spp = int(gold["spp"])
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    img, st = capi.Scene(sc).render(mode, 48 * 48 * spp, seed=11)
    ra, rb = gold[name + "_a"], gold[name + "_b"]
    floor = rel_rmse(ra, rb)
    assert rel_rmse(img, ra) < 1.25 * floor, (rel_rmse(img, ra), floor)
    assert rel_rmse(img, rb) < 1.25 * floor, (rel_rmse(img, rb), floor)
    ref_mean = 0.5 * (ra + rb).mean(axis=(0, 1))
    assert np.allclose(img.mean(axis=(0, 1)), ref_mean, rtol=0.02 if mode == capi.MODE_PT else 0.005)


def test_normal_renderer_equals_oracle():
    sc = scenedesc.config2_scene(30000, 160, 90, n_objects=40)
    g, st = capi.Scene(sc).render(capi.MODE_NORMAL, 0)
    p, tri = ob.PortPT(sc).render_normal()
    assert st["extend_rays"] == 160 * 90
    assert np.abs(g - p).max() <= 1e-6


def test_render_dev_with_caller_owned_film():
    """lmb200_render_dev accumulates UNSCALED splats into a caller-owned device film (the multi-GPU path:
    every rank renders its slice, films are summed by NCCL, then lmb200_film_rescale_dev)."""
    import torch
    sc = scenedesc.cornell_box(32, 32)
    S = capi.Scene(sc)
    N = 32 * 32 * 16
    ref, _ = S.render(capi.MODE_PTDIRECT, N, seed=2)
    film = torch.zeros((32, 32, 4), dtype=torch.float32, device="cuda")
    L = capi.lib()
    st = capi.RenderStats()
    for g in range(2):
        p = S.params(capi.MODE_PTDIRECT, N, seed=2, begin=N * g // 2, end=N * (g + 1) // 2)
        capi.check(L.lmb200_render_dev(S.h_, C.byref(p), film.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)))
    capi.check(L.lmb200_film_rescale_dev(film.data_ptr(), 32 * 32, float(32 * 32) / N, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.allclose(film.cpu().numpy()[..., :3], ref, rtol=2e-4, atol=1e-5)


def test_render_multi_nccl_reduce():
    """Single-process multi-GPU entry point (what renderer::lmb200pt calls with num_gpus > 1): per-GPU films are
    summed with ncclReduce; the image must equal the 1-GPU image of the same seed."""
    L = capi.lib()
    if L.lmb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    N = 48 * 48 * 64
    scenes_ = [capi.Scene(sc, device=g) for g in range(2)]
    one, _ = scenes_[0].render(capi.MODE_PTDIRECT, N, seed=3)
    arr = (C.c_void_p * 2)(*[s.h_ for s in scenes_])
    p = scenes_[0].params(capi.MODE_PTDIRECT, N, seed=3)
    film = np.zeros((48, 48, 4), np.float32)
    st = capi.RenderStats()
    capi.check(L.lmb200_render_multi(arr, 2, C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert st.samples == N
    assert same_image(film[..., :3], one)


def test_time_budget_and_progress():
    """Scheduler_'s render_time / progress_image_update_interval semantics (scheduler.cpp:108,191-255): passes until
    the time budget is spent, progress images rescaled by W*H/processed, final image rescaled by the processed count."""
    sc = scenedesc.cornell_box(32, 32)
    S = capi.Scene(sc)
    L = capi.lib()
    arr = (C.c_void_p * 1)(S.h_)
    film = np.zeros((32, 32, 4), np.float32)
    st = capi.RenderStats()
    ticks = []

    def on_progress(user, rgba, done, tick):
        img = np.ctypeslib.as_array(rgba, shape=(32, 32, 4))
        ticks.append((int(done), int(tick), float(img[..., :3].mean())))
        return 0
    cb = capi.PROGRESS_FN(on_progress)
    p = S.params(capi.MODE_PTDIRECT, 1, seed=9)
    capi.check(L.lmb200_render_timed(arr, 1, C.byref(p), 0.5, 200000, 0.1, cb, None, film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert st.samples >= 200000 and st.samples % 200000 == 0
    assert len(ticks) >= 2 and [t[1] for t in ticks] == list(range(1, len(ticks) + 1))
    assert all(ticks[i][0] < ticks[i + 1][0] for i in range(len(ticks) - 1))
    ref, _ = S.render(capi.MODE_PTDIRECT, 32 * 32 * 2048, seed=1)
    assert abs(film[..., :3].mean() - ref.mean()) / ref.mean() < 0.05
    assert all(abs(t[2] - ref.mean()) / ref.mean() < 0.25 for t in ticks)
    # without a time budget the call renders exactly [sample_begin, sample_end) and equals lmb200_render
    N = 32 * 32 * 16
    p = S.params(capi.MODE_PTDIRECT, N, seed=3)
    capi.check(L.lmb200_render_timed(arr, 1, C.byref(p), -1.0, 5000, -1.0, capi.PROGRESS_FN(0), None, film.ctypes.data_as(C.c_void_p), C.byref(st)))
    one, _ = S.render(capi.MODE_PTDIRECT, N, seed=3)
    assert st.samples == N and same_image(film[..., :3], one)


def test_gpu_built_scene_renders_the_same_image():
    """The renderer on a device-built BVH: same samples => same image (hits do not depend on the tree)."""
    sc = scenedesc.config2_scene(30000, 64, 36, n_objects=40)
    N = 64 * 36 * 32
    a, sa = capi.Scene(sc, builder=capi.BUILD_HOST_SAH).render(capi.MODE_PTDIRECT, N, seed=2)
    b, sb = capi.Scene(sc, builder=capi.BUILD_GPU_LBVH).render(capi.MODE_PTDIRECT, N, seed=2)
    assert sa["extend_rays"] == sb["extend_rays"] and sa["shadow_rays"] == sb["shadow_rays"]
    assert np.allclose(a, b, rtol=2e-4, atol=1e-5)


def test_scene_sharing_a_prebuilt_accel():
    """lmb200_scene_create_shared: the renderer borrows an accel built over the same triangle list (what
    renderer::lmb200pt does when the YAML also selected accel::lmb200) and renders the same image."""
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    L = capi.lib()
    own = capi.Scene(sc)
    N = 48 * 48 * 32
    a, _ = own.render(capi.MODE_PTDIRECT, N, seed=4)
    desc, keep = sc.flatten()
    A = capi.Accel(0)
    A.build(keep["verts"], builder=capi.BUILD_GPU_LBVH)
    h = L.lmb200_scene_create_shared(C.byref(desc), A.h)
    assert h, L.lmb200_last_error()
    assert L.lmb200_scene_accel(h) == A.h
    film = np.zeros((48, 48, 4), np.float32)
    st = capi.RenderStats()
    p = own.params(capi.MODE_PTDIRECT, N, seed=4)
    capi.check(L.lmb200_render(h, C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(st)))
    L.lmb200_scene_destroy(h)
    assert np.allclose(film[..., :3], a, rtol=2e-4, atol=1e-5)
    assert len(A.trace_closest(np.array([[0, 1, 4, 1e-4, 0, 0, -1, 3e38]], np.float32))) == 1     # the accel outlives the scene
    # mismatching triangle list is refused
    B = capi.Accel(0)
    B.build(keep["verts"][:10])
    assert not L.lmb200_scene_create_shared(C.byref(desc), B.h)
    # registry round trip
    L.lmb200_registry_put(12345, A.h)
    assert L.lmb200_registry_get(12345) == A.h
    L.lmb200_registry_put(12345, None)
    assert not L.lmb200_registry_get(12345)


OUTDOOR_CASES = [("directional", False, capi.MODE_PTDIRECT, "ptdirect"), ("env", False, capi.MODE_PTDIRECT, "ptdirect"),
                 ("both", True, capi.MODE_PTDIRECT, "ptdirect"), ("directional", True, capi.MODE_PTMIS, "ptmis"),
                 ("directional", True, capi.MODE_PT, "pt"), ("cornell", True, capi.MODE_PT, "pt"),
                 ("textured", False, capi.MODE_PTDIRECT, "ptdirect")]


@pytest.mark.parametrize("light,thin,mode,name", OUTDOOR_CASES)
def test_directional_env_thinlens_same_samples_as_oracle_and_reference_images(light, thin, mode, name):
    """light::directional / light::env (bounding-sphere emitter shapes) and sensor::thinlens on the device:
    (1) same seed => the oracle's image and ray counts; (2) converged => the reference's own image
    (tests/golden/pt_outdoor.npz) within 1.25x its two-seed noise floor."""
    sc = scenedesc.outdoor_scene(48, 27, light, thin)
    S = capi.Scene(sc)
    N = 48 * 27 * 32
    port, counts = ob.PortPT(sc).render(mode, N, seed=7)
    gpu, st = S.render(mode, N, seed=7, pool=1 << 14)
    assert not np.isnan(gpu).any()
    assert abs(st["extend_rays"] - counts[0]) <= max(4, 1e-4 * counts[0])
    assert abs(st["shadow_rays"] - counts[1]) <= max(4, 1e-4 * counts[1])
    gold = np.load(os.path.join(GOLD, "pt_outdoor.npz"))
    key = f"{light}_{'thinlens' if thin else 'pinhole'}_{name}"
    ra, rb = gold[key + "_a"], gold[key + "_b"]
    if light == "directional" and name == "pt":
        assert gpu.max() == 0 and ra.max() == 0 and st["shadow_rays"] == 0     # a delta-direction light is never hit
        return
    close = np.abs(gpu - port) <= 2e-3 + 1e-3 * np.abs(port)
    assert close.all(axis=2).mean() >= 0.999
    assert rel_rmse(gpu, port) < 1e-3, rel_rmse(gpu, port)
    spp = int(gold["spp"])
    img, _ = S.render(mode, 48 * 27 * spp, seed=11)
    floor = rel_rmse(ra, rb)
    assert rel_rmse(img, ra) < 1.25 * floor, (rel_rmse(img, ra), floor)
    assert rel_rmse(img, rb) < 1.25 * floor, (rel_rmse(img, rb), floor)
    ref_mean = 0.5 * (ra + rb).mean(axis=(0, 1))
    assert np.allclose(img.mean(axis=(0, 1)), ref_mean, rtol=0.02 if mode == capi.MODE_PT else 0.006)


def test_env_light_is_rejected_outside_ptdirect():
    """renderer::pt / ptmis of the reference crash on light::env (null primitive on escape); we refuse instead."""
    sc = scenedesc.outdoor_scene(16, 9, "env", False)
    S = capi.Scene(sc)
    for mode in (capi.MODE_PT, capi.MODE_PTMIS):
        with pytest.raises(capi.LmbError, match="light::env"):
            S.render(mode, 100)
    img, _ = S.render(capi.MODE_PTDIRECT, 16 * 9 * 16)
    assert img.mean() > 0


def test_normal_renderer_thinlens_uses_lens_centre():
    a = scenedesc.outdoor_scene(64, 36, "directional", False)
    b = scenedesc.outdoor_scene(64, 36, "directional", True)
    ga, _ = capi.Scene(a).render(capi.MODE_NORMAL, 0)
    gb, _ = capi.Scene(b).render(capi.MODE_NORMAL, 0)
    pb, _ = ob.PortPT(b).render_normal()
    assert np.abs(gb - pb).max() <= 1e-6
    assert (np.abs(ga - gb).max(axis=2) > 1e-3).mean() < 0.02      # same picture up to rounding at silhouettes


def test_tile_partitioning_matches_oracle_and_whole_image_sampling():
    """Optional tile partitioning (include/lmb200.h lmb200_render_params::tile): camera samples drawn inside a raster
    rectangle. Same samples as the oracle per tile; the whole tile is the same sample set as no tile; strips with
    their share of the samples sum to the whole-image estimate within Monte-Carlo noise (diffuse box: no fireflies)."""
    sc = scenedesc.cornell_box(48, 48)
    S = capi.Scene(sc)
    P = ob.PortPT(sc)
    N = 48 * 48 * 128
    full, _ = S.render(capi.MODE_PTDIRECT, N, seed=3)
    same, _ = S.render(capi.MODE_PTDIRECT, N, seed=3, tile=(0, 0, 1, 1))
    assert np.allclose(full, same, rtol=2e-4, atol=1e-5)
    parts = []
    for k in range(4):
        t = (0.0, k / 4, 1.0, (k + 1) / 4)
        g, st = S.render(capi.MODE_PTDIRECT, N, seed=3, begin=N * k // 4, end=N * (k + 1) // 4, tile=t)
        p, counts = P.render(capi.MODE_PTDIRECT, N, seed=3, begin=N * k // 4, end=N * (k + 1) // 4, tile=t)
        assert abs(st["extend_rays"] - counts[0]) <= 4 and abs(st["shadow_rays"] - counts[1]) <= 4
        assert rel_rmse(g, p) < 1e-3
        parts.append(g)
    other, _ = S.render(capi.MODE_PTDIRECT, N, seed=4)
    floor = rel_rmse(full, other)
    assert rel_rmse(sum(parts), full) < 1.25 * floor
    assert np.allclose(sum(parts).mean(axis=(0, 1)), full.mean(axis=(0, 1)), rtol=0.02)
    # pt has no camera-vertex light splats: a strip's film is empty outside the strip
    top, _ = S.render(capi.MODE_PT, N, seed=3, begin=0, end=N // 4, tile=(0, 0, 1, 0.25))
    assert top[12:].max() == 0 and top[:12].max() > 0
    with pytest.raises(capi.LmbError, match="tile"):
        S.render(capi.MODE_PT, N, tile=(0.5, 0, 0.25, 1))


def test_render_multi_tile_partitioning():
    """lmb200_render_multi with tile_partition = 1: GPU g samples strip g; the summed film matches whole-image sampling
    statistically."""
    L = capi.lib()
    if L.lmb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = scenedesc.cornell_box(48, 48, glossy_block=True)
    N = 48 * 48 * 256
    scenes_ = [capi.Scene(sc, device=g) for g in range(2)]
    one, _ = scenes_[0].render(capi.MODE_PTDIRECT, N, seed=3)
    two, _ = scenes_[0].render(capi.MODE_PTDIRECT, N, seed=4)
    arr = (C.c_void_p * 2)(*[s.h_ for s in scenes_])
    p = scenes_[0].params(capi.MODE_PTDIRECT, N, seed=3, tile_partition=True)
    film = np.zeros((48, 48, 4), np.float32)
    st = capi.RenderStats()
    capi.check(L.lmb200_render_multi(arr, 2, C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert st.samples == N
    assert rel_rmse(film[..., :3], one) < 1.25 * rel_rmse(two, one)
    assert np.allclose(film[..., :3].mean(axis=(0, 1)), one.mean(axis=(0, 1)), rtol=0.02)


def _two_scenes(sc):
    L = capi.lib()
    if L.lmb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scenes_ = [capi.Scene(sc, device=g) for g in range(2)]
    return L, scenes_, (C.c_void_p * 2)(*[s.h_ for s in scenes_])


def test_render_multi_normal_mode_equals_single_gpu():
    """MODE_NORMAL through lmb200_render_multi: every GPU renders its share of the pixels, the film sum is the image
    (round 1 summed N full images: N times too bright)."""
    sc = scenedesc.cornell_box(64, 40, glossy_block=True)
    L, scenes_, arr = _two_scenes(sc)
    one, _ = scenes_[0].render(capi.MODE_NORMAL, 64 * 40)
    p = scenes_[0].params(capi.MODE_NORMAL, 64 * 40)
    film = np.zeros((40, 64, 4), np.float32)
    st = capi.RenderStats()
    capi.check(L.lmb200_render_multi(arr, 2, C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert st.samples == 64 * 40 and st.extend_rays == 64 * 40
    assert np.array_equal(film[..., :3], one)
    # the time-budget entry point forwards MODE_NORMAL to the same path
    film[:] = 0
    capi.check(L.lmb200_render_timed(arr, 2, C.byref(p), -1.0, 0, -1.0, capi.PROGRESS_FN(0), None, film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert np.array_equal(film[..., :3], one)


def test_render_timed_two_gpus_progress_and_exact_range():
    """lmb200_render_timed on 2 GPUs: progress images come from a periodic NCCL reduce of the per-GPU films
    (scheduler.cpp:221-255, 280-288); without a time budget the call renders exactly the sample range and equals the
    1-GPU image; stats report the reduce time."""
    sc = scenedesc.cornell_box(32, 32)
    L, scenes_, arr = _two_scenes(sc)
    film = np.zeros((32, 32, 4), np.float32)
    st = capi.RenderStats()
    ticks = []

    def on_progress(user, rgba, done, tick):
        img = np.ctypeslib.as_array(rgba, shape=(32, 32, 4))
        ticks.append((int(done), int(tick), float(img[..., :3].mean())))
        return 0
    cb = capi.PROGRESS_FN(on_progress)
    p = scenes_[0].params(capi.MODE_PTDIRECT, 1, seed=9)
    capi.check(L.lmb200_render_timed(arr, 2, C.byref(p), 0.6, 200000, 0.1, cb, None, film.ctypes.data_as(C.c_void_p), C.byref(st)))
    assert st.samples >= 200000 and st.samples % 200000 == 0
    assert len(ticks) >= 2 and [t[1] for t in ticks] == list(range(1, len(ticks) + 1))
    assert st.reduce_seconds > 0
    ref, _ = scenes_[0].render(capi.MODE_PTDIRECT, 32 * 32 * 2048, seed=1)
    assert abs(film[..., :3].mean() - ref.mean()) / ref.mean() < 0.05
    assert all(abs(t[2] - ref.mean()) / ref.mean() < 0.25 for t in ticks)
    N = 32 * 32 * 16
    p = scenes_[0].params(capi.MODE_PTDIRECT, N, seed=3)
    capi.check(L.lmb200_render_timed(arr, 2, C.byref(p), -1.0, 5000, -1.0, capi.PROGRESS_FN(0), None, film.ctypes.data_as(C.c_void_p), C.byref(st)))
    one, _ = scenes_[0].render(capi.MODE_PTDIRECT, N, seed=3)
    assert st.samples == N and same_image(film[..., :3], one)


def test_render_multi_rejects_two_scenes_on_one_device():
    sc = scenedesc.cornell_box(16, 16)
    L = capi.lib()
    scenes_ = [capi.Scene(sc, device=0) for _ in range(2)]
    arr = (C.c_void_p * 2)(*[s.h_ for s in scenes_])
    p = scenes_[0].params(capi.MODE_PTDIRECT, 1000)
    film = np.zeros((16, 16, 4), np.float32)
    with pytest.raises(capi.LmbError, match="same device"):
        capi.check(L.lmb200_render_multi(arr, 2, C.byref(p), film.ctypes.data_as(C.c_void_p), None))


def test_accel_replicate_on_second_gpu():
    """lmb200_accel_replicate: device-to-device copy of the BVH; the replica traces the same hits and a scene sharing
    it renders the same image."""
    L = capi.lib()
    if L.lmb200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    verts = scenes.soup(50000, seed=3, extent=10.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(100000, lo, hi, seed=5)
    a = capi.Accel(0)
    a.build(verts)
    h0 = a.trace_closest(rays)
    r = L.lmb200_accel_replicate(a.h, 1)
    assert r and L.lmb200_accel_device(r) == 1 and L.lmb200_accel_device(a.h) == 0
    h1 = np.zeros(len(rays), dtype=capi.HIT_DTYPE)
    capi.check(L.lmb200_trace_closest(r, rays.ctypes.data_as(C.c_void_p), h1.ctypes.data_as(C.c_void_p), len(rays)))
    assert np.array_equal(h0.view(np.uint32), h1.view(np.uint32))
    L.lmb200_accel_destroy(r)
