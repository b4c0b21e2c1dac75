"""GPU parity of the traversal path, through the C ABI (include/lmb200.h), against the oracle:
closest-hit triangle index bit-exact, t/u/v bit-exact (bar: 1e-5 relative; we hold the stronger one)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

from oracle import bindings as ob
from lmb200py import capi, scenes

from test_oracle import SIMPLE_PS, SIMPLE_FS, SIMPLE2_PS, SIMPLE2_FS, TS, tri_verts, simple_rays, simple2_rays, interp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FLT_MAX = np.float32(3.4028234663852886e38)


def gpu_closest(accel, rays):
    hits = accel.trace_closest(rays)
    tri = hits["tri"].astype(np.int64)
    tri[tri == capi.MISS] = -1
    tuv = np.stack([hits["t"], hits["u"], hits["v"]], axis=1)
    return tuv, tri


def assert_bit_exact(tuv_g, tri_g, tuv_o, tri_o):
    assert np.array_equal(tri_g, tri_o), f"{np.count_nonzero(tri_g != tri_o)} index mismatches"
    assert np.array_equal(tuv_g.view(np.uint32), tuv_o.view(np.uint32)), "t/u/v bits differ"


def test_known_answers_accel3test():
    """Accel3Test.Simple / Simple2 (test_accel3.cpp:272-345), tolerance EpsLarge = 1e-3 as in the reference."""
    for ps, fs, ts, (rays, exp), zexp in [
        (SIMPLE_PS, SIMPLE_FS, TS[[0, 1, 2, 3, 0, 1, 2, 3]], simple_rays(), lambda e: 0 * e[:, 0]),
        (SIMPLE2_PS, SIMPLE2_FS, TS, simple2_rays(), lambda e: -e[:, 0]),
    ]:
        A = capi.Accel(0)
        A.build(tri_verts(ps, fs))
        tuv, tri = gpu_closest(A, rays)
        assert (tri >= 0).all()
        p = rays[:, 0:3] + rays[:, 4:7] * tuv[:, 0:1]
        assert np.allclose(p[:, :2], exp, atol=1e-3) and np.allclose(p[:, 2], zexp(exp), atol=1e-3)
        uv = interp(ps, fs, ts, tri, tuv[:, 1], tuv[:, 2])
        assert np.allclose(uv, exp, atol=1e-3)


def test_golden_vectors_from_reference():
    g = np.load(os.path.join(GOLD, "accel_soup.npz"))
    A = capi.Accel(0)
    A.build(g["verts"])
    tuv, tri = gpu_closest(A, g["rays"])
    assert_bit_exact(tuv, tri, g["tuv"], g["face"])
    assert np.array_equal(A.trace_any(g["rays"]).astype(bool), g["face"] >= 0)


@pytest.mark.parametrize("builder", [None, capi.BUILD_HOST_SAH])       # the default (device builder) and the host SAH builder
@pytest.mark.parametrize("ntri,extent,edge,nray", [(1, 1.0, 0.4, 2000), (9, 1.0, 0.4, 5000), (2000, 3.0, 0.3, 50000), (200000, 30.0, 0.2, 400000)])
def test_random_soup_vs_oracle(ntri, extent, edge, nray, builder):
    verts = scenes.soup(ntri, seed=100 + ntri, extent=extent, edge=edge)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(nray, lo - 0.5, hi + 0.5, seed=3)
    rays[: nray // 4, 7] = extent * 0.3          # bounded ranges
    rays[nray // 4: nray // 3, 3] = 0.0          # tmin = 0 as in the reference tests
    A = capi.Accel(0)
    A.build(verts, builder=builder)
    P = ob.PortScene(verts)
    tuv_o, tri_o = P.closest(rays)
    assert_bit_exact(*gpu_closest(A, rays), tuv_o, tri_o)
    assert np.array_equal(A.trace_any(rays), P.any(rays))


def test_mesh_scene_vs_oracle():
    verts, _ = scenes.mesh_scene(120000, seed=5, half=20.0, n_objects=30)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(300000, lo, hi, seed=11)
    cam = scenes.camera_rays((0, 9, 21), (0, 2, 0), (0, 1, 0), 45.0, 320, 180)    # coherent primary rays too
    rays = np.concatenate([rays, cam])
    tuv_o, tri_o = ob.PortScene(verts).closest(rays)
    for builder in (None, capi.BUILD_HOST_SAH):
        A = capi.Accel(0)
        A.build(verts, builder=builder)
        assert_bit_exact(*gpu_closest(A, rays), tuv_o, tri_o)


def test_edge_cases():
    A = capi.Accel(0)
    # empty scene, empty batch
    A.build(np.zeros((0, 9), np.float32))
    r = np.array([[0, 0, 1, 0, 0, 0, -1, FLT_MAX]], np.float32)
    assert gpu_closest(A, r)[1][0] == -1 and A.trace_any(r)[0] == 0
    assert len(A.trace_closest(np.zeros((0, 8), np.float32))) == 0
    # degenerate / NaN triangles are never hit; coincident triangles: larger index wins (accel::naive order)
    t = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    verts = np.array([t] * 25 + [[0, 0, 0, 1, 1, 1, 2, 2, 2]] + [[np.nan] * 9], np.float32)
    A.build(verts)
    rays = np.array([[0.2, 0.2, 1, 0, 0, 0, -1, FLT_MAX],
                     [0.2, 0.2, 1, 0, 0, 0, -1, 0.5],        # range ends before the plane
                     [0.2, 0.2, 1, 0, 0, 0, -1, 1.0],        # t == tmax is accepted (triaccel.h:137)
                     [0.2, 0.2, 1, 1.0, 0, 0, -1, FLT_MAX],  # t == tmin is accepted
                     [0.2, 0.2, 1, 0, 0, 0, 1, FLT_MAX],     # pointing away
                     [0.2, 0.2, 1, 0, 0, 0, -0.0, FLT_MAX],  # zero direction
                     [0.25, 0.25, 0.0, 0, 1, 0, 0, FLT_MAX]], np.float32)   # in-plane ray, zero components
    tuv, tri = gpu_closest(A, rays)
    tuv_o, tri_o = ob.PortScene(verts).closest(rays)
    assert_bit_exact(tuv, tri, tuv_o, tri_o)
    assert tri[0] == 24 and tri[1] == -1 and tri[2] == 24 and tri[3] == 24 and tri[4] == -1
    # rays with NaN or infinite components are misses, in the batch calls and through the per-ray service, and do not walk the
    # whole tree (NaN slab distances drop out of the min / max, so nothing would be culled): 100k of them return at once
    bad = np.array([[np.nan, 0.2, 1, 0, 0, 0, -1, FLT_MAX], [0.2, 0.2, 1, 0, np.nan, 0, -1, FLT_MAX],
                    [0.2, 0.2, 1, 0, 0, 0, -np.inf, FLT_MAX], [np.inf, 0.2, 1, 0, 0, 0, -1, FLT_MAX]], np.float32)
    assert (gpu_closest(A, bad)[1] == -1).all() and (A.trace_any(bad) == 0).all()
    assert (ob.PortScene(verts).closest(bad[:2])[1] == -1).all()
    big = scenes.soup(200_000, seed=3, extent=10.0, edge=0.2)
    A.build(big)
    many = np.repeat(bad, 25_000, axis=0)
    many[::7] = [5, 5, 5, 0, 0.3, 0.5, 0.8, FLT_MAX]
    t0 = time.perf_counter()
    tri_many = gpu_closest(A, many)[1]
    assert time.perf_counter() - t0 < 5.0
    assert (tri_many[np.arange(len(many)) % 7 != 0] == -1).all() and (tri_many[::7] == tri_many[0]).all()
    L = capi.lib()
    one = np.zeros(1, capi.HIT_DTYPE)
    for k in range(8):        # both service traversals (the worker warp alternates between them while it is sampling)
        r = np.ascontiguousarray(bad[k % 4:k % 4 + 1])
        capi.check(L.lmb200_trace_closest_one(A.h, r.ctypes.data_as(C.c_void_p), one.ctypes.data_as(C.c_void_p)))
        assert one["tri"][0] == capi.MISS


def test_axis_aligned_rays_and_boxes():
    """Zero direction components against axis-aligned geometry (the reference substitutes EpsLarge/Inf
    for 1/0, accel_qbvh.cpp:411-416); results must still equal the linear-scan oracle."""
    verts = tri_verts(SIMPLE_PS, SIMPLE_FS)
    g = np.linspace(0.05, 0.95, 19, dtype=np.float32)
    X, Y = np.meshgrid(g, g)
    n = X.size
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0], rays[:, 1], rays[:, 2] = X.ravel(), Y.ravel(), 1
    rays[:, 6] = -1
    rays[:, 7] = FLT_MAX
    A = capi.Accel(0)
    A.build(verts)
    tuv_o, tri_o = ob.PortScene(verts).closest(rays, use_bvh=False)
    assert_bit_exact(*gpu_closest(A, rays), tuv_o, tri_o)


def test_device_pointer_api_and_single_ray():
    import torch
    verts = scenes.soup(5000, seed=8, extent=5.0, edge=0.3)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(20000, lo, hi, seed=2)
    A = capi.Accel(0)
    A.build(verts)
    host = A.trace_closest(rays)
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.zeros((len(rays), 4), dtype=torch.float32, device="cuda")
    L = capi.lib()
    capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), len(rays), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(d_hits.cpu().numpy().view(np.uint32), host.view(np.uint32).reshape(-1, 4))
    d_occ = torch.zeros(len(rays), dtype=torch.uint8, device="cuda")
    capi.check(L.lmb200_trace_any_dev(A.h, d_rays.data_ptr(), d_occ.data_ptr(), len(rays), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(d_occ.cpu().numpy().astype(bool), host["tri"] != capi.MISS)
    one = np.zeros(1, capi.HIT_DTYPE)
    for i in range(0, 200):
        r = np.ascontiguousarray(rays[i])
        capi.check(L.lmb200_trace_closest_one(A.h, r.ctypes.data_as(C.c_void_p), one.ctypes.data_as(C.c_void_p)))
        assert one.view(np.uint32).tolist() == host[i:i + 1].view(np.uint32).tolist()


def test_compact_wire_form_equals_full_form():
    """lmb200_trace_closest_compact / lmb200_trace_any_compact: 24-byte rays + one [tmin, tmax] for the batch give exactly the
    hits of the 32-byte form (Accel3::Intersect's own argument shape: Ray + minT + maxT, accel3.h:68)."""
    L = capi.lib()
    verts = scenes.soup(40000, seed=12, extent=9.0, edge=0.3)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(70001, lo, hi, seed=6)
    A = capi.Accel(0)
    A.build(verts)
    for tmin, tmax in ((1e-4, 3.4028234663852886e38), (0.0, 2.5)):
        rays[:, 3] = tmin
        rays[:, 7] = tmax
        full, occ = A.trace_closest(rays), A.trace_any(rays)
        r24 = np.ascontiguousarray(rays[:, [0, 1, 2, 4, 5, 6]])
        hits = np.zeros(len(rays), capi.HIT_DTYPE)
        capi.check(L.lmb200_trace_closest_compact(A.h, r24.ctypes.data, tmin, tmax, hits.ctypes.data, len(rays)))
        assert np.array_equal(hits.view(np.uint32), full.view(np.uint32))
        o2 = np.zeros(len(rays), np.uint8)
        capi.check(L.lmb200_trace_any_compact(A.h, r24.ctypes.data, tmin, tmax, o2.ctypes.data, len(rays)))
        assert np.array_equal(o2, occ)
    assert L.lmb200_trace_closest_compact(A.h, None, 0.0, 1.0, None, 5) == -1


def test_host_buffer_calls_from_several_threads_on_one_accel():
    """The host-buffer batch calls of one accel share its staging buffers: concurrent calls from several host threads take
    turns and every one of them gets the hits of its own rays."""
    import threading
    verts = scenes.soup(60000, seed=14, extent=9.0, edge=0.3)
    lo, hi = scenes.bounds(verts)
    A = capi.Accel(0)
    A.build(verts)
    batches = [scenes.random_rays(150_000 + 1000 * k, lo, hi, seed=20 + k) for k in range(6)]
    want = [A.trace_closest(b) for b in batches]
    got = [None] * len(batches)
    occ = [None] * len(batches)

    def work(k):
        for _ in range(3):
            got[k] = A.trace_closest(batches[k])
            occ[k] = A.trace_any(batches[k])
    th = [threading.Thread(target=work, args=(k,)) for k in range(len(batches))]
    for t in th: t.start()
    for t in th: t.join()
    for k in range(len(batches)):
        assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32))
        assert np.array_equal(occ[k].astype(bool), want[k]["tri"] != capi.MISS)


STREAM_CHILD = r"""
import os, sys, threading
import numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'lightmetrica-v2_b200'))
from lmb200py import capi, scenes
L = capi.lib()
verts = scenes.soup(120000, seed=5, extent=20.0, edge=0.3)
lo, hi = scenes.bounds(verts)
A = capi.Accel(0); A.build(verts)
bad = 0
def check(n, seed):
    global bad
    rays = scenes.random_rays(n, lo, hi, seed=seed)
    rays[:, 3] = 1e-4; rays[:, 7] = 3.0e38
    rays[::1001, 4] = np.nan          # a few rays that finish at once
    d_rays = torch.from_numpy(rays).cuda(); d_hits = torch.empty((n, 4), dtype=torch.float32, device='cuda'); d_occ = torch.zeros(n, dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    capi.check(L.lmb200_trace_closest_dev(A.h, d_rays.data_ptr(), d_hits.data_ptr(), n, st))
    capi.check(L.lmb200_trace_any_dev(A.h, d_rays.data_ptr(), d_occ.data_ptr(), n, st))
    torch.cuda.synchronize()
    want = d_hits.cpu().numpy().view(np.uint32); want_occ = d_occ.cpu().numpy()
    for rep in range(2):
        got = A.trace_closest(rays); occ = A.trace_any(rays)
        r24 = np.ascontiguousarray(rays[:, [0, 1, 2, 4, 5, 6]])
        h2 = np.zeros(n, capi.HIT_DTYPE); o2 = np.zeros(n, np.uint8)
        capi.check(L.lmb200_trace_closest_compact(A.h, r24.ctypes.data, 1e-4, 3.0e38, h2.ctypes.data, n))
        capi.check(L.lmb200_trace_any_compact(A.h, r24.ctypes.data, 1e-4, 3.0e38, o2.ctypes.data, n))
        ok = (np.array_equal(got.view(np.uint32).reshape(-1, 4), want) and np.array_equal(occ.astype(np.uint8), want_occ)
              and np.array_equal(h2.view(np.uint32).reshape(-1, 4), want) and np.array_equal(o2, want_occ))
        if not ok: bad += 1
for n, seed in ((1500001, 3), (262144, 4), (700000, 5), (65537, 6), (300000, 7)):
    check(n, seed)
# several host threads on one accel, each a streaming call
th = [threading.Thread(target=check, args=(400000 + 1000 * k, 10 + k)) for k in range(3)]
for t in th: t.start()
for t in th: t.join()
print("STREAM_OK" if bad == 0 else "STREAM_BAD %%d" %% bad)
"""


def test_streaming_host_buffer_path_with_small_chunks():
    """The host-buffer calls of four chunks and more run as ONE persistent launch that takes the rays chunk by chunk as their
    uploads land (accel.cu trace_host_stream, traverse.cuh StreamGate). With 64 Ki-ray chunks (LMB200_E2E_CHUNK_LOG2=16, read
    once per process: hence the child process) a 1.5 M-ray call has ~40 chunks, wraps around the staging ring six times and
    keeps outrunning its uploads; all four call forms must equal the device-pointer launch bit for bit. The per-chunk
    launches (LMB200_E2E_STREAM=0) stay available and are checked the same way."""
    import subprocess
    import sys
    for stream in ("1", "0"):
        env = dict(os.environ, LMB200_E2E_CHUNK_LOG2="16", LMB200_E2E_STREAM=stream)
        r = subprocess.run([sys.executable, "-c", STREAM_CHILD % (ROOT, ROOT)], env=env, capture_output=True, text=True, timeout=600)
        assert "STREAM_OK" in r.stdout, (stream, r.stdout[-500:], r.stderr[-1500:])


def test_per_ray_service_many_threads_and_restart():
    """The per-ray Accel3::Intersect path (persistent service kernel + mailboxes): 16 host threads posting rays
    concurrently get bit-identical hits to the batch call; the service survives going idle (it exits after 2 ms without
    requests and is restarted by the next call), a rebuild of the accel, and more threads than mailboxes."""
    import threading
    import time
    L = capi.lib()
    verts = scenes.soup(30000, seed=8, extent=8.0, edge=0.3)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(32000, lo, hi, seed=2)
    A = capi.Accel(0)
    A.build(verts)
    batch = A.trace_closest(rays)
    hits = np.zeros(len(rays), capi.HIT_DTYPE)
    sec = C.c_double()
    capi.check(L.lmb200_trace_closest_one_mt(A.h, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), len(rays), 16, C.byref(sec)))
    assert np.array_equal(hits.view(np.uint32), batch.view(np.uint32))
    time.sleep(0.05)                                  # the service kernel has left by now; the next call restarts it
    hits[:] = 0
    capi.check(L.lmb200_trace_closest_one_mt(A.h, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), 2000, 3, C.byref(sec)))
    assert np.array_equal(hits[:2000].view(np.uint32), batch[:2000].view(np.uint32))
    # 80 threads > 64 mailboxes (shared under a lock), through ctypes from Python threads
    out = np.zeros(80 * 20, capi.HIT_DTYPE)

    def work(t):
        for i in range(20 * t, 20 * t + 20):
            capi.check(L.lmb200_trace_closest_one(A.h, rays[i:i + 1].ctypes.data_as(C.c_void_p), out[i:i + 1].ctypes.data_as(C.c_void_p)))
    th = [threading.Thread(target=work, args=(t,)) for t in range(80)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert np.array_equal(out.view(np.uint32), batch[:1600].view(np.uint32))
    # rebuild on another scene while the service may still be resident, then a device-wide synchronisation
    verts2 = scenes.soup(2000, seed=9, extent=3.0, edge=0.3)
    A.build(verts2)
    import torch
    torch.cuda.synchronize()
    lo2, hi2 = scenes.bounds(verts2)
    rays2 = scenes.random_rays(3000, lo2, hi2, seed=4)
    b2 = A.trace_closest(rays2)
    h2 = np.zeros(len(rays2), capi.HIT_DTYPE)
    capi.check(L.lmb200_trace_closest_one_mt(A.h, rays2.ctypes.data_as(C.c_void_p), h2.ctypes.data_as(C.c_void_p), len(rays2), 8, C.byref(sec)))
    assert np.array_equal(h2.view(np.uint32), b2.view(np.uint32))
    A.close()


def test_full_size_properties():
    """BASELINE sizes (4 M triangles, 16 Mi rays here) are beyond the oracle's reach in seconds, so
    parity is checked through size-independent properties plus an oracle spot check."""
    import torch
    ntri, nray = 4_000_000, 1 << 24
    verts = scenes.soup(ntri, seed=42, extent=100.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(nray, lo, hi, seed=7)
    A = capi.Accel(0)
    st = A.build(verts)
    assert st["num_valid_triangles"] == ntri
    hits = A.trace_closest(rays)
    hit = hits["tri"] != capi.MISS
    assert 0.2 < hit.mean() < 0.9
    # (1) every reported t lies in the ray's range
    assert (hits["t"][hit] >= rays[hit, 3]).all() and (hits["t"][hit] <= rays[hit, 7]).all()
    # (2) re-evaluating the reference triangle test on the reported triangle reproduces t,u,v bit for bit
    idx = np.flatnonzero(hit)[:: max(1, hit.sum() // 20000)]
    P = ob.PortScene(verts[hits["tri"][idx]])
    sub = rays[idx]
    # evaluate each sampled ray against its own reported triangle with the oracle's TriAccel test
    tuv_chk = np.zeros((len(idx), 3), np.float32)
    L = ob.port()
    rec = P.records()
    u, v, t = C.c_float(), C.c_float(), C.c_float()
    for k in range(len(idx)):
        ok = L.orc_triaccel_intersect(rec[k].ctypes.data_as(C.c_void_p), sub[k, 0:3].ctypes.data_as(C.c_void_p), np.ascontiguousarray(sub[k, 4:7]).ctypes.data_as(C.c_void_p),
                                      C.c_float(sub[k, 3]), C.c_float(sub[k, 7]), C.byref(u), C.byref(v), C.byref(t))
        assert ok
        tuv_chk[k] = (t.value, u.value, v.value)
    got = np.stack([hits["t"][idx], hits["u"][idx], hits["v"][idx]], axis=1)
    assert np.array_equal(got.view(np.uint32), tuv_chk.view(np.uint32))
    # (3) closest means closest: shrinking tmax to just below the reported t leaves no hit at all
    r2 = rays[hit].copy()
    r2[:, 7] = np.nextafter(hits["t"][hit], np.float32(-1))
    assert (A.trace_closest(r2)["tri"] == capi.MISS).all()
    # (4) any-hit agrees with closest-hit
    assert np.array_equal(A.trace_any(rays).astype(bool), hit)
    # (5) oracle spot check on the first 100k rays
    tuv_o, tri_o = ob.PortScene(verts).closest(rays[:100000])
    tri_g = hits["tri"][:100000].astype(np.int64)
    tri_g[tri_g == capi.MISS] = -1
    assert np.array_equal(tri_g, tri_o)
    assert np.array_equal(np.stack([hits["t"], hits["u"], hits["v"]], axis=1)[:100000].view(np.uint32), tuv_o.view(np.uint32))


@pytest.mark.parametrize("builder", [capi.BUILD_GPU_LBVH, capi.BUILD_GPU_LBVH_SAH, capi.BUILD_GPU_PLOC])
@pytest.mark.parametrize("ntri,extent,edge", [(0, 1.0, 0.3), (1, 1.0, 0.3), (2, 1.0, 0.3), (3, 1.0, 0.3), (4, 1.0, 0.3), (5, 1.0, 0.3), (9, 1.0, 0.3), (300, 2.0, 0.3), (50000, 12.0, 0.2)])
def test_gpu_builder_same_hits_and_valid_structure(ntri, extent, edge, builder):
    """The device builders (Morton radix tree, or PLOC clustering with the SAH-optimal collapse) produce other trees of the same format:
    hits must be bit-identical to the oracle's (the closest hit does not depend on the tree), every triangle must
    be referenced once, and every quantised child box must contain what is below it."""
    from test_host import check_structure
    verts = scenes.soup(ntri, seed=77, extent=extent, edge=edge) if ntri else np.zeros((0, 9), np.float32)
    lo, hi = (scenes.bounds(verts) if ntri else (np.zeros(3, np.float32), np.ones(3, np.float32)))
    rays = scenes.random_rays(40000, lo - 0.3, hi + 0.3, seed=5)
    A = capi.Accel(0)
    st = A.build(verts, builder=builder)
    assert st["num_valid_triangles"] == ntri
    tuv_o, tri_o = ob.PortScene(verts).closest(rays)
    assert_bit_exact(*gpu_closest(A, rays), tuv_o, tri_o)
    assert np.array_equal(A.trace_any(rays).astype(bool), tri_o >= 0)
    units, num_nodes, num_tris, grid = A.host_layout()
    assert num_tris == ntri and st["num_nodes"] == num_nodes
    idx = check_structure(units, num_nodes, num_tris, grid, verts, ob.PortScene(verts).records() if ntri else np.zeros((0, 12), np.uint32))
    assert sorted(idx.tolist()) == list(range(ntri))


def _spiral_scene():
    parts = []
    for i in range(21):
        c = np.float32(100.0 * 2.0 ** -i)
        base = scenes.soup(24, seed=50 + i, extent=float(c) * 0.2, edge=float(c) * 0.05).reshape(-1, 3) + c
        parts.append(base.reshape(-1, 9))
    return np.ascontiguousarray(np.concatenate(parts), np.float32)


def test_deep_device_tree_falls_back_to_the_host_builder():
    """The traversal stack holds 16 levels; a device-built tree deeper than that is handed to the binned SAH builder (which
    falls back to median splits), and a tree that is still too deep is refused - never walked. Exercised with
    LMB200_DEPTH_LIMIT in a child process: clusters on an exponential spiral give a 9-level radix tree and an 8-level SAH tree."""
    import subprocess
    import sys
    code = """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from lmb200py import capi, scenes
from oracle import bindings as ob
from test_gpu_trace import _spiral_scene, gpu_closest, assert_bit_exact
verts = _spiral_scene()
A, G, H = capi.Accel(0), capi.Accel(0), capi.Accel(host_only=True)
try:
    sh = H.build(verts)
except capi.LmbError:
    sh = {'num_nodes': -1}
try:
    G.build(verts, builder=capi.BUILD_HOST_SAH); print('HOST', 'ok')
except capi.LmbError as e:
    print('HOST', 'refused' if 'exceeds the traversal stack' in str(e) else e)
try:
    st = A.build(verts)
    print('DEFAULT', st['max_depth'], st['num_nodes'] == sh['num_nodes'])
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(20000, lo, hi, seed=3)
    assert_bit_exact(*gpu_closest(A, rays), *ob.PortScene(verts).closest(rays, use_bvh=False))
    print('HITS ok')
except capi.LmbError as e:
    print('DEFAULT', 'refused' if 'exceeds the traversal stack' in str(e) else e)
""" % (ROOT, os.path.join(ROOT, "tests"))
    out = {}
    for limit in ("16", "8", "7"):
        env = dict(os.environ, LMB200_DEPTH_LIMIT=limit, PYTHONPATH=os.path.join(ROOT, "lightmetrica-v2_b200"))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-1500:]
        out[limit] = r.stdout
    assert "DEFAULT 9 False" in out["16"] and "HITS ok" in out["16"]              # the device tree itself (9 levels)
    assert "DEFAULT 8 True" in out["8"] and "HITS ok" in out["8"] and "HOST ok" in out["8"]      # fell back to the host tree
    assert "DEFAULT refused" in out["7"] and "HOST refused" in out["7"]            # nothing fits: an error, not a walk


def test_gpu_builder_large_and_degenerate():
    """1 M triangles incl. duplicates, degenerate and NaN triangles: both builders give identical hits."""
    verts = scenes.mesh_scene(1_000_000, seed=42)[0]
    verts = np.concatenate([verts, verts[:1000], np.zeros((10, 9), np.float32), np.full((3, 9), np.nan, np.float32)])
    lo, hi = scenes.bounds(verts[:-3])
    rays = scenes.random_rays(1 << 20, lo, hi, seed=3)
    A, B, P = capi.Accel(0), capi.Accel(0), capi.Accel(0)
    sa = A.build(verts, builder=capi.BUILD_HOST_SAH)
    sb = B.build(verts, builder=capi.BUILD_GPU_LBVH)
    sp = P.build(verts, builder=capi.BUILD_GPU_PLOC)
    assert sa["num_valid_triangles"] == sb["num_valid_triangles"] == sp["num_valid_triangles"] == len(verts) - 13
    ha, hb, hp = A.trace_closest(rays), B.trace_closest(rays), P.trace_closest(rays)
    assert np.array_equal(ha.view(np.uint32), hb.view(np.uint32)) and np.array_equal(ha.view(np.uint32), hp.view(np.uint32))
    assert sb["build_seconds"] < sa["build_seconds"] and sp["build_seconds"] < sa["build_seconds"]
