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


def decode_nodes(nodes):
    dt = np.dtype([("p", "f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "u4"), ("tri_base", "u4"),
                   ("meta", "u1", 8), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
    assert dt.itemsize == 80
    return nodes.view(dt)


@pytest.mark.parametrize("n,extent,edge", [(1, 1.0, 0.3), (2, 1.0, 0.3), (7, 1.0, 0.3), (300, 2.0, 0.3), (20000, 10.0, 0.2)])
def test_builder_structure(n, extent, edge):
    verts = scenes.soup(n, seed=3, extent=extent, edge=edge)
    A = capi.Accel(host_only=True)
    st = A.build(verts)
    nodes, tris, idx = A.host_arrays()
    N = decode_nodes(nodes)
    assert st["num_valid_triangles"] == n and len(idx) == n
    assert sorted(idx.tolist()) == list(range(n))          # every triangle referenced exactly once
    # records are the bit-exact TriAccel precompute, in leaf order
    rec = tris.view(np.uint32).reshape(-1, 12)
    port = ob.PortScene(verts).records()
    assert np.array_equal(rec[:, :10], port[idx][:, :10]) and np.array_equal(rec[:, 10], idx)
    # every child box contains its triangles (padded by the reference's 1e-4) / its child node's boxes
    pad = 1e-4
    seen_nodes = np.zeros(len(N), bool)
    seen_nodes[0] = True
    for ni, nd in enumerate(N):
        sc = np.ldexp(1.0, nd["e"].astype(int) - 127)
        rel = 0
        for s in range(8):
            m = int(nd["meta"][s])
            if m == 0:
                continue
            lo = nd["p"].astype(np.float64) + sc * nd["qlo"][:, s]
            hi = nd["p"].astype(np.float64) + sc * nd["qhi"][:, s]
            if (nd["imask"] >> s) & 1:
                assert m == (0x20 | (24 + s))
                c = nd["child_base"] + rel
                rel += 1
                assert not seen_nodes[c]
                seen_nodes[c] = True
                ch = N[c]
                csc = np.ldexp(1.0, ch["e"].astype(int) - 127)
                used = ch["meta"] != 0
                clo = ch["p"].astype(np.float64)[:, None] + csc[:, None] * ch["qlo"][:, used]
                chi = ch["p"].astype(np.float64)[:, None] + csc[:, None] * ch["qhi"][:, used]
                # the child's own grid may round outward by < 1 step of ITS grid beyond the parent's slot box
                assert (clo.min(axis=1) >= lo - csc - 1e-9).all() and (chi.max(axis=1) <= hi + csc + 1e-9).all()
            else:
                cnt = {1: 1, 3: 2, 7: 3}[m >> 5]
                off = m & 31
                for k in range(cnt):
                    v = verts[idx[nd["tri_base"] + off + k]].reshape(3, 3).astype(np.float64)
                    assert (v.min(axis=0) - pad >= lo - 1e-9).all() and (v.max(axis=0) + pad <= hi + 1e-9).all()
    assert seen_nodes.all()


def test_builder_closest_equals_oracle():
    verts = scenes.soup(30000, seed=42, extent=10.0, edge=0.2)
    lo, hi = scenes.bounds(verts)
    rays = scenes.random_rays(60000, lo, hi, seed=7)
    rays[:5000, 7] = 2.0
    A = capi.Accel(host_only=True)
    A.build(verts)
    nodes, tris, idx = A.host_arrays()
    tuv_w, tri_w = ob.wide_closest(nodes, tris, rays)
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
    nodes, tris, idx = A.host_arrays()
    rays = np.array([[0.2, 0.2, 1, 0, 0, 0, -1, 3.4e38]], np.float32)
    tuv, tri = ob.wide_closest(nodes, tris, rays)
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
