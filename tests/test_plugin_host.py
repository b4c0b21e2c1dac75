"""The two Lightmetrica plugins against the reference host (oracle/_ref), CPU only: they load through
the reference's own ComponentFactory, register their keys, and fail loudly without a CUDA device."""
import os

import pytest

from oracle import bindings as ob
from lmb200py import capi, scenedesc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUG = os.path.join(ROOT, "lightmetrica-v2_b200", "plugin")

pytestmark = pytest.mark.skipif(
    not (ob.have_ref() and os.path.exists(os.path.join(PLUG, "accel_lmb200.so"))),
    reason="needs oracle/_ref and the built plugins (both need /root/reference to build)")


def load_plugins():
    L = ob.ref()
    assert L.ref_load_plugin(os.path.join(PLUG, "accel_lmb200").encode()) == 1
    assert L.ref_load_plugin(os.path.join(PLUG, "renderer_lmb200pt").encode()) == 1


def test_plugins_load_and_register():
    load_plugins()
    # a missing plugin is reported, not fatal (component.cpp:101-125)
    assert ob.ref().ref_load_plugin(os.path.join(PLUG, "does_not_exist").encode()) == 0


def test_no_gpu_fails_loudly():
    if capi.lib().lmb200_device_count() > 0:
        pytest.skip("GPU present")
    load_plugins()
    sc = scenedesc.cornell_box(16, 16)
    with pytest.raises(RuntimeError, match="accel init failed"):
        ob.RefScene(sc, accel="lmb200")
    R = ob.RefScene(sc, accel="qbvh")
    with pytest.raises(RuntimeError, match="renderer init failed"):
        R.render("lmb200pt", 100)


def _read_dump(path):
    import ctypes as C
    import numpy as np
    with open(path, "rb") as f:
        nt, npr, nb, nl, has_n = (int(x) for x in np.frombuffer(f.read(40), np.uint64))
        cam = capi.Camera.from_buffer_copy(f.read(C.sizeof(capi.Camera)))
        sphere = np.frombuffer(f.read(16), np.float32)
        verts = np.frombuffer(f.read(36 * nt), np.float32).reshape(-1, 9)
        tri_prim = np.frombuffer(f.read(4 * nt), np.uint32)
        prims = (capi.Primitive * npr).from_buffer_copy(f.read(C.sizeof(capi.Primitive) * npr))
        bsdfs = (capi.Bsdf * nb).from_buffer_copy(f.read(C.sizeof(capi.Bsdf) * nb))
        lights = (capi.Light * nl).from_buffer_copy(f.read(C.sizeof(capi.Light) * nl))
        ntex, has_uv = (int(x) for x in np.frombuffer(f.read(16), np.uint64))
        textures = []
        for _ in range(ntex):
            tw, th = (int(x) for x in np.frombuffer(f.read(8), np.int32))
            textures.append(np.frombuffer(f.read(12 * tw * th), np.float32).reshape(th, tw, 3))
        uvs = np.frombuffer(f.read(24 * nt), np.float32).reshape(-1, 6) if has_uv else None
        normals = np.frombuffer(f.read(36 * nt), np.float32).reshape(-1, 9) if has_n else None
        assert f.read(1) == b""
    return dict(normals=normals, nt=nt, npr=npr, nb=nb, nl=nl, has_n=has_n, cam=cam, sphere=sphere, verts=verts, tri_prim=tri_prim,
                prims=prims, bsdfs=bsdfs, lights=lights, textures=textures, uvs=uvs)


def test_renderer_plugin_scene_extraction(tmp_path, monkeypatch):
    """What renderer::lmb200pt reads through the reference's interfaces (Scene3::PrimitiveAt,
    TriangleMesh, BSDF::Reflectance/Glossiness, YAML eta/k, Light::Emittance, pinhole transform/fov)
    equals the scene that was described. Uses the plugin's LMB200_DUMP_SCENE aid, so no GPU is needed."""
    import numpy as np
    dump = str(tmp_path / "scene.bin")
    monkeypatch.setenv("LMB200_DUMP_SCENE", dump)
    load_plugins()
    sc = scenedesc.cornell_box(32, 24, glossy_block=True)
    R = ob.RefScene(sc, accel="qbvh")
    # Render returns void (renderer.h:81): without a device it logs the CUDA error and returns after the dump
    R.render("lmb200pt", 10, extra={"mode": "ptdirect"}, in_tree=True)
    D = _read_dump(dump)
    nt, npr, nl, has_n, cam, verts, tri_prim, prims, bsdfs, lights = (D[k] for k in ("nt", "npr", "nl", "has_n", "cam", "verts", "tri_prim", "prims", "bsdfs", "lights"))
    d, keep = sc.flatten()
    assert (nt, npr, nl, has_n) == (d.num_tris, d.num_prims, d.num_lights, 0)
    assert D["normals"] is None and D["uvs"] is None      # no mesh has any: the arrays are not even made (flatten.h)
    assert np.array_equal(verts, keep["verts"]) and np.array_equal(tri_prim, keep["tri_prim"])
    for i in range(npr):
        a, b = prims[i], keep["prims"][i]
        assert (a.light, a.first_tri, a.num_tris, a.has_normals) == (b.light, b.first_tri, b.num_tris, b.has_normals)
        ba, bb = bsdfs[a.bsdf], keep["bs"][b.bsdf]
        assert ba.type == bb.type
        if ba.type != capi.BSDF_NULL:
            assert list(ba.R) == list(bb.R)
        if ba.type == capi.BSDF_COOKTORRANCE:
            assert list(ba.eta) == list(bb.eta) and list(ba.k) == list(bb.k) and ba.roughness == bb.roughness
    assert list(lights[0].Le) == [17.0, 12.0, 4.0] and lights[0].primitive == keep["ls"][0].primitive
    # the camera basis comes from the reference's lookat (rsqrt-normalised, math.h:1872-1875): equal to 1e-3
    for k in ("position", "vx", "vy", "vz"):
        assert np.allclose(list(getattr(cam, k)), list(getattr(d.camera, k)), atol=1e-3)
    assert abs(cam.fov - d.camera.fov) < 1e-6 and (cam.width, cam.height) == (32, 24)
    assert cam.kind == capi.CAMERA_PINHOLE
    # Scene3::GetSphereBound as restated by Scene.flatten()
    assert np.allclose(D["sphere"][:3], list(d.sphere_center), atol=1e-5) and abs(D["sphere"][3] - d.sphere_radius) < 1e-4 * d.sphere_radius


def test_renderer_plugin_extracts_thinlens_directional_env(tmp_path, monkeypatch):
    """sensor::thinlens (lens radius / focal distance), light::directional (transformed, normalised direction) and light::env
    (constant Le) as the plugin reads them through the reference's Sensor / Light interfaces and the YAML tree."""
    import numpy as np
    dump = str(tmp_path / "scene.bin")
    monkeypatch.setenv("LMB200_DUMP_SCENE", dump)
    load_plugins()
    sc = scenedesc.outdoor_scene(32, 18, light="both", thinlens=True)
    R = ob.RefScene(sc, accel="qbvh")
    R.render("lmb200pt", 10, extra={"mode": "ptdirect"}, in_tree=True)
    D = _read_dump(dump)
    d, keep = sc.flatten()
    cam, lights = D["cam"], D["lights"]
    assert cam.kind == capi.CAMERA_THINLENS and cam.lens_radius == np.float32(0.25) and cam.focal_distance == np.float32(6.1)
    assert D["nl"] == 2
    kinds = {lights[i].kind: lights[i] for i in range(2)}
    sun, sky = kinds[capi.LIGHT_DIRECTIONAL], kinds[capi.LIGHT_ENV]
    want = [x for x in keep["ls"] if x.kind == capi.LIGHT_DIRECTIONAL][0]
    assert np.allclose(list(sun.direction), list(want.direction), atol=1e-3)      # rsqrt-normalised in the reference
    assert abs(np.linalg.norm(list(sun.direction)) - 1) < 1e-3
    assert np.allclose(list(sun.Le), [3.0, 2.8, 2.5]) and np.allclose(list(sky.Le), [0.5, 0.6, 0.8])
    assert sun.primitive == want.primitive
    assert np.allclose(D["sphere"][:3], list(d.sphere_center), atol=1e-5) and abs(D["sphere"][3] - d.sphere_radius) < 1e-4 * d.sphere_radius


def test_renderer_plugin_bakes_textures(tmp_path, monkeypatch):
    """TexR on bsdf::diffuse / bsdf::cook_torrance: the plugin resolves the texture asset through Scene::GetAssets, bakes
    Texture::Evaluate at texel centres (texture_resolution) and flattens TriangleMesh::Texcoords per triangle."""
    import numpy as np
    dump = str(tmp_path / "scene.bin")
    monkeypatch.setenv("LMB200_DUMP_SCENE", dump)
    load_plugins()
    sc = scenedesc.textured_box(16, 16)
    R = ob.RefScene(sc, accel="qbvh")
    R.render("lmb200pt", 10, extra={"mode": "ptdirect", "texture_resolution": sc.tex_res}, in_tree=True)
    D = _read_dump(dump)
    d, keep = sc.flatten()
    assert len(D["textures"]) == 2 and D["uvs"] is not None
    assert np.array_equal(D["uvs"], keep["uvs"])
    assert D["has_n"] == 0 and D["normals"] is None      # no vertex normals in this scene
    got_tex = {D["bsdfs"][D["prims"][i].bsdf].texR for i in range(D["npr"])}
    assert got_tex == {0, 1, 2}
    for i in range(D["npr"]):
        a, b = D["bsdfs"][D["prims"][i].bsdf], keep["bs"][keep["prims"][i].bsdf]
        assert (a.texR > 0) == (b.texR > 0)
        if a.texR > 0:
            assert np.array_equal(D["textures"][a.texR - 1], keep["baked"][b.texR - 1])


def test_renderer_plugin_flattens_vertex_normals_of_some_meshes(tmp_path, monkeypatch):
    """Meshes with and without vertex normals in one scene (intersectionutils.h:88-90 falls back to the geometric normal for
    the latter): the flattened normal array carries the given normals for the former and zeros for the latter, triangle for
    triangle in the reference accels' order."""
    import numpy as np
    from lmb200py import scenes
    dump = str(tmp_path / "scene.bin")
    monkeypatch.setenv("LMB200_DUMP_SCENE", dump)
    load_plugins()
    sc = scenedesc.cornell_box(16, 16)
    c = np.array([0.4, 0.9, 0.3], np.float32)
    ball = scenes.sphere(c, 0.25, 10, 6)
    n = ball.reshape(-1, 3) - c
    n = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    sc.add_mesh_tris(ball, "red", normals=n)
    sc.add_quad((-0.2, 0.01, 0.5), (-0.2, 0.01, 0.9), (0.2, 0.01, 0.9), (0.2, 0.01, 0.5), "green")      # a mesh without normals AFTER it
    R = ob.RefScene(sc, accel="qbvh")
    R.render("lmb200pt", 10, extra={"mode": "ptdirect"}, in_tree=True)
    D = _read_dump(dump)
    d, keep = sc.flatten()
    assert D["has_n"] == 1 and D["nt"] == d.num_tris
    assert np.array_equal(D["verts"], keep["verts"]) and np.array_equal(D["tri_prim"], keep["tri_prim"])
    assert np.array_equal(D["normals"], keep["norms"])
    with_n = np.array([D["prims"][int(p)].has_normals for p in D["tri_prim"]], bool)
    assert with_n.sum() == len(ball) and not with_n[-2:].any() and not with_n[:36].any()
    assert (D["normals"][~with_n] == 0).all() and np.allclose(np.linalg.norm(D["normals"][with_n].reshape(-1, 3), axis=1), 1, atol=1e-6)
