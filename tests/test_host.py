"""Host-side logic of the product without a GPU: the C-ABI library loads and exports what
include/lmb200.h declares, and the BVH builder's flattened output is complete and conservative
(checked by walking it with the oracle's scalar checker)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import bindings as ob
from lmb200py import capi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "lmb200.h")).read()
    declared = sorted(set(re.findall(r"\b(lmb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, f"declared in include/lmb200.h but not exported: {missing}"
    assert sorted(capi.EXPORTS) == declared


def test_struct_sizes():
    assert capi.RAY_DTYPE.itemsize == 32 and capi.HIT_DTYPE.itemsize == 16


def test_ctypes_mirrors_match_the_header(tmp_path):
    """The ctypes structures of lmb200py/capi.py (which the oracle port shares through Scene.flatten) have the sizes and field
    offsets a C compiler gives the structs of include/lmb200.h."""
    import subprocess
    pairs = [("lmb200_bsdf", capi.Bsdf), ("lmb200_texture", capi.Texture), ("lmb200_primitive", capi.Primitive), ("lmb200_light", capi.Light),
             ("lmb200_camera", capi.Camera), ("lmb200_scene_desc", capi.SceneDesc), ("lmb200_render_params", capi.RenderParams),
             ("lmb200_render_stats", capi.RenderStats), ("lmb200_accel_stats", capi.AccelStats)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lmb200.h"', 'int main(void) {']
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "sizes.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    for (cname, ct), line in zip(pairs, out):
        toks = line.split()
        assert toks[0] == cname
        assert int(toks[1]) == C.sizeof(ct), (cname, toks[1], C.sizeof(ct))
        offs = [int(t) for t in toks[2:]]
        assert offs == [getattr(ct, f).offset for f, _ in ct._fields_], cname


def test_no_device_fails_loudly(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(capi.LmbError):
        capi.Accel(0)


NODE_DT = np.dtype([("k", "<u2", 3), ("counts", "<u2"), ("e", "u1", 3), ("imask", "u1"), ("base", "<u4"),
                    ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
assert NODE_DT.itemsize == 64


def check_structure(units, num_nodes, num_tris, grid, verts, records):
    """Walks the flattened 64-byte units from the root (csrc/bvh.h): every unit is referenced exactly once, every valid
    triangle sits in exactly one leaf slot with its bit-exact TriAccel record, every quantised child box contains what is
    below it (triangles padded by the reference's 1e-4; a child node's own boxes up to one step of ITS grid).
    Returns the input indices of the triangles in unit order."""
    N = units.view(NODE_DT).reshape(-1)
    W = units.view(np.uint32).reshape(-1, 16)
    glo, gstep = grid[0].astype(np.float64), grid[1].astype(np.float64)
    seen = np.zeros(len(units), np.int32)
    seen[0] = 1
    tri_ids = []
    pad = 1e-4
    stack = [0]
    n_nodes = 0

    def boxes(nd):
        org = glo + gstep * nd["k"].astype(np.float64)
        sc = np.ldexp(1.0, nd["e"].astype(int) - 127)
        return org[:, None] + sc[:, None] * nd["qlo"], org[:, None] + sc[:, None] * nd["qhi"], sc

    while stack:
        ni = stack.pop()
        nd = N[ni]
        n_nodes += 1
        lo, hi, _ = boxes(nd)
        nint = bin(int(nd["imask"])).count("1")
        rel = toff = 0
        for s in range(8):
            cnt = (int(nd["counts"]) >> (2 * s)) & 3
            if (int(nd["imask"]) >> s) & 1:
                assert cnt == 0
                c = int(nd["base"]) + rel
                rel += 1
                seen[c] += 1
                ch = N[c]
                clo, chi, csc = boxes(ch)
                used = ch["qlo"][0] <= ch["qhi"][0]
                assert used.any()
                assert (clo[:, used].min(axis=1) >= lo[:, s] - csc - 1e-9).all() and (chi[:, used].max(axis=1) <= hi[:, s] + csc + 1e-9).all()
                stack.append(c)
            elif cnt:
                for k in range(cnt):
                    u = int(nd["base"]) + nint + toff + k
                    seen[u] += 1
                    tid = int(W[u, 10])
                    tri_ids.append(tid)
                    assert np.array_equal(W[u, :10], records[tid, :10])          # bit-exact TriAccel precompute
                    v = verts[tid].reshape(3, 3).astype(np.float64)
                    assert (v.min(axis=0) - pad >= lo[:, s] - 1e-9).all() and (v.max(axis=0) + pad <= hi[:, s] + 1e-9).all()
                toff += cnt
            else:
                assert (nd["qlo"][:, s] == 255).all() and (nd["qhi"][:, s] == 0).all()      # empty slot: inverted box
    assert (seen == 1).all()                      # no orphan and no shared unit
    assert n_nodes == num_nodes and len(tri_ids) == num_tris and n_nodes + len(tri_ids) == len(units)
    return np.array(tri_ids, np.int64)


@pytest.mark.parametrize("n,extent,edge", [(1, 1.0, 0.3), (2, 1.0, 0.3), (7, 1.0, 0.3), (300, 2.0, 0.3), (20000, 10.0, 0.2)])
def test_builder_structure(n, extent, edge):
    verts = scenes.soup(n, seed=3, extent=extent, edge=edge)
    A = capi.Accel(host_only=True)
    st = A.build(verts)
    units, num_nodes, num_tris, grid = A.host_layout()
    assert st["num_valid_triangles"] == n == num_tris and st["num_nodes"] == num_nodes
    assert st["node_bytes"] == 64 * num_nodes and st["tri_bytes"] == 64 * num_tris
    idx = check_structure(units, num_nodes, num_tris, grid, verts, ob.PortScene(verts).records())
    assert sorted(idx.tolist()) == list(range(n))          # every triangle referenced exactly once


def test_builder_far_from_origin_and_anisotropic():
    """The 16-bit scene grid of the node origins: a scene far from the origin (coarse fp32 there) and a very flat one."""
    for off, scale in [((5.0e4, -3.0e4, 1.0e3), (1.0, 1.0, 1.0)), ((0.0, 0.0, 0.0), (100.0, 1e-3, 10.0))]:
        verts = scenes.soup(3000, seed=5, extent=4.0, edge=0.2).reshape(-1, 3) * np.array(scale, np.float32) + np.array(off, np.float32)
        verts = np.ascontiguousarray(verts.reshape(-1, 9), np.float32)
        A = capi.Accel(host_only=True)
        A.build(verts)
        units, num_nodes, num_tris, grid = A.host_layout()
        check_structure(units, num_nodes, num_tris, grid, verts, ob.PortScene(verts).records())
        lo, hi = scenes.bounds(verts)
        rays = scenes.random_rays(20000, lo, hi, seed=7)
        tuv_w, tri_w = ob.wide_closest(units, grid, rays)
        # against the linear scan (accel::naive semantics): at coordinates of 5e4 the reference-style box test of the port's
        # own BVH (1e-4 pad, fp32 slabs) is itself no longer conservative
        tuv_p, tri_p = ob.PortScene(verts).closest(rays, use_bvh=False)
        assert np.array_equal(tri_w, tri_p) and np.array_equal(tuv_w.view(np.uint32), tuv_p.view(np.uint32))


def test_builder_closest_equals_oracle():
    verts = scenes.soup(30000, seed=42, extent=10.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(60000, lo, hi, seed=7)
    rays[:5000, 7] = 2.0
    A = capi.Accel(host_only=True)
    A.build(verts)
    units, _, _, grid = A.host_layout()
    tuv_w, tri_w = ob.wide_closest(units, grid, rays)
    tuv_p, tri_p = ob.PortScene(verts).closest(rays)
    assert np.array_equal(tri_w, tri_p)
    assert np.array_equal(tuv_w.view(np.uint32), tuv_p.view(np.uint32))


def test_builder_edge_cases():
    A = capi.Accel(host_only=True)
    st = A.build(np.zeros((0, 9), np.float32))
    assert st["num_nodes"] == 1 and st["num_valid_triangles"] == 0
    # degenerate + NaN triangles are dropped; coincident triangles (the reference's builder never
    # terminates on >= 10 of them, accel_qbvh.cpp:375-377) are handled
    t = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    verts = np.array([t] * 40 + [[0, 0, 0, 1, 1, 1, 2, 2, 2]] + [[np.nan] * 9], np.float32)
    st = A.build(verts)
    assert st["num_valid_triangles"] == 40
    units, _, _, grid = A.host_layout()
    rays = np.array([[0.2, 0.2, 1, 0, 0, 0, -1, 3.4e38]], np.float32)
    tuv, tri = ob.wide_closest(units, grid, rays)
    assert tri[0] == 39      # tie rule: larger index wins


def test_error_paths_without_a_device():
    """Error behaviour of the C ABI (nothing throws, codes + lmb200_last_error): bad arguments and wrong state."""
    L = capi.lib()
    A = capi.Accel(host_only=True)
    # not built yet
    st = capi.AccelStats()
    assert L.lmb200_accel_get_stats(A.h, C.byref(st)) == -3            # LMB200_E_STATE
    assert b"not built" in L.lmb200_last_error()
    # null arguments
    assert L.lmb200_accel_build(A.h, None, 5) == -1                    # LMB200_E_INVALID
    assert L.lmb200_accel_build(None, None, 0) == -1
    assert L.lmb200_accel_get_stats(None, C.byref(st)) == -1
    # unknown builder / GPU builder on a host-only accel
    verts = scenes.soup(10, seed=1, extent=1.0, edge=0.3)
    assert L.lmb200_accel_build_ex(A.h, verts.ctypes.data_as(C.c_void_p), 10, 7) == -1
    assert L.lmb200_accel_build_ex(A.h, verts.ctypes.data_as(C.c_void_p), 10, capi.BUILD_GPU_LBVH) == -3
    # too many triangles (2^27 limit of the leaf encoding): rejected before touching the data
    assert L.lmb200_accel_build(A.h, verts.ctypes.data_as(C.c_void_p), 1 << 27) == -1
    # a host-only accel cannot trace
    A.build(verts)
    rays = np.zeros((1, 8), np.float32)
    hits = np.zeros(1, capi.HIT_DTYPE)
    assert L.lmb200_trace_closest(A.h, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p), 1) == -3
    assert L.lmb200_trace_closest_one(A.h, rays.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p)) == -3
    assert L.lmb200_trace_closest_dev(A.h, None, None, 1, None) == -1
    # scene creation validates its description before any CUDA call
    assert L.lmb200_scene_create(0, None) is None
    from lmb200py import scenedesc
    d, keep = scenedesc.cornell_box(8, 8).flatten()
    d.num_lights = 1
    keep["ls"][0].primitive = 99
    assert L.lmb200_scene_create(0, C.byref(d)) is None
    assert L.lmb200_last_error() != b""


@pytest.mark.parametrize("family", ["mesh", "far_anisotropic", "axis_aligned_rays", "bounded_ranges", "duplicates_and_degenerates", "origins_on_surfaces"])
def test_wide_tree_on_the_pinned_families(family):
    """The 64-byte-unit tree (host builder) walked on the CPU in the device's format, on the scene / ray families the port is
    pinned on against the compiled reference (tests/test_oracle.py::test_port_vs_reference_live_families): the quantised
    child boxes never cull a hit, ranges and the tie rule hold, every hit bit for bit."""
    from test_oracle import _family
    verts, rays = _family(family)
    A = capi.Accel(host_only=True)
    st = A.build(verts)
    units, num_nodes, num_tris, grid = A.host_layout()
    P = ob.PortScene(verts)
    check_structure(units, num_nodes, num_tris, grid, verts, P.records())
    assert st["num_valid_triangles"] == num_tris <= len(verts)
    tuv_w, tri_w = ob.wide_closest(units, grid, rays)
    tuv_p, tri_p = P.closest(rays)
    assert np.array_equal(tri_w, tri_p)
    assert np.array_equal(tuv_w.view(np.uint32), tuv_p.view(np.uint32))
