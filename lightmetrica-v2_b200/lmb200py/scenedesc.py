"""Scene model shared by the tests and the bench: one description, two consumers.

  to_yaml()  -> the Lightmetrica YAML the reference consumes, meshes passed
                through the harness's in-memory `trianglemesh::mem` asset
  flatten()  -> the POD arrays of include/lmb200.h (lmb200_scene_desc), in the reference's
                primitive order (scene3.cpp:136-389: nodes in document order, camera node included)
All meshes are given in world space (identity node transforms).
"""
import ctypes as C

import numpy as np

from . import capi, scenes


class Scene:
    def __init__(self):
        self.meshes = []      # dict(verts (nv,3) f32, faces (nf,3) u32, normals (nv,3) f32 or None)
        self.nodes = []       # dict(mesh=int, bsdf=str or None, light=str or None)
        self.bsdfs = {}       # name -> dict(type='diffuse'|'cook_torrance', R, eta, k, roughness)
        self.lights = {}      # name -> Le (3,)   (light::area)
        self.point_lights = {}   # name -> dict(Le, position)   (light::point)
        self.dir_lights = {}     # name -> dict(Le, direction)  (light::directional)
        self.env_lights = {}     # name -> Le                   (light::env, constant)
        self.textures = {}       # name -> dict(scale, color1, color2)   (texture::checker, plugin/texture_checker)
        self.tex_res = 256       # resolution textures are baked at for the POD scene (a multiple of every checker scale => exact)
        self.camera = None    # dict(eye, center, up, fov, w, h[, lens_radius, focal_distance])

    # ---- construction helpers ----
    def add_bsdf(self, name, type="diffuse", R=(0.8, 0.8, 0.8), roughness=0.1,
                 eta=(0.14, 0.129, 0.1585), k=(4.58625, 3.348125, 2.329375), eta1=1.0, eta2=2.0, texR=None):
        """type: diffuse | cook_torrance | reflect_all | refract_all | flesnel; texR = name of a texture replacing R
        (diffuse / cook_torrance, bsdf_diffuse.cpp:48-53)"""
        self.bsdfs[name] = dict(type=type, R=tuple(R), roughness=float(roughness), eta=tuple(eta), k=tuple(k),
                                eta1=float(eta1), eta2=float(eta2), texR=texR)

    def add_texture(self, name, scale=8.0, color1=(1.0, 0.0, 0.0), color2=(1.0, 1.0, 1.0)):
        """texture::checker (plugin/texture_checker/texture_checker.cpp:38-52)."""
        self.textures[name] = dict(scale=float(scale), color1=tuple(color1), color2=tuple(color2))

    def bake_texture(self, name):
        """(res, res, 3) float32: the texture evaluated at texel centres, the form lmb200_texture carries."""
        t = self.textures[name]
        res = self.tex_res
        c = ((np.arange(res, dtype=np.float32) + np.float32(0.5)) / np.float32(res)).astype(np.float32)
        iu = (c * np.float32(t["scale"])).astype(np.int32)
        even = ((iu[None, :] + iu[:, None]) % 2) == 0                      # [y, x]
        return np.where(even[..., None], np.asarray(t["color1"], np.float32), np.asarray(t["color2"], np.float32)).astype(np.float32)

    def add_light(self, name, Le):
        self.lights[name] = tuple(Le)

    def add_point_light(self, name, Le, position):
        """light::point (light_point.cpp:47-54) on a node of its own (no mesh)."""
        self.point_lights[name] = dict(Le=tuple(Le), position=tuple(float(x) for x in position))
        self.nodes.append(dict(mesh=None, bsdf=None, light=name))

    def add_directional_light(self, name, Le, direction):
        """light::directional (light_directional.cpp:100-105) on a node of its own; `direction` is where the light travels."""
        self.dir_lights[name] = dict(Le=tuple(Le), direction=tuple(float(x) for x in direction))
        self.nodes.append(dict(mesh=None, bsdf=None, light=name))

    def add_env_light(self, name, Le):
        """light::env with a constant Le (light_env.cpp:103-121) on a node of its own."""
        self.env_lights[name] = tuple(Le)
        self.nodes.append(dict(mesh=None, bsdf=None, light=name))

    def add_mesh_tris(self, tris9, bsdf=None, light=None, normals=None, uvs=None):
        """tris9: (n,9) world-space triangles, stored unshared (3 vertices per face); uvs: (3n,2) texture coordinates."""
        tris9 = np.ascontiguousarray(tris9, np.float32).reshape(-1, 9)
        n = tris9.shape[0]
        m = dict(verts=tris9.reshape(-1, 3).copy(), faces=np.arange(3 * n, dtype=np.uint32).reshape(-1, 3),
                 normals=None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3),
                 uvs=None if uvs is None else np.ascontiguousarray(uvs, np.float32).reshape(-1, 2))
        self.meshes.append(m)
        self.nodes.append(dict(mesh=len(self.meshes) - 1, bsdf=bsdf, light=light))

    def add_quad(self, a, b, c, d, bsdf=None, light=None, uv=False):
        """uv=True: texture coordinates (0,0),(1,0),(1,1),(0,1) at a,b,c,d."""
        a, b, c, d = (np.asarray(x, np.float32) for x in (a, b, c, d))
        uvs = np.array([[0, 0], [1, 0], [1, 1], [0, 0], [1, 1], [0, 1]], np.float32) if uv else None
        self.add_mesh_tris(np.stack([np.concatenate([a, b, c]), np.concatenate([a, c, d])]), bsdf, light, uvs=uvs)

    def set_camera(self, eye, center, up, fov, w, h, lens_radius=None, focal_distance=1.0):
        """sensor::pinhole, or sensor::thinlens when lens_radius is given (sensor_thinlens.cpp:44-68)."""
        self.camera = dict(eye=tuple(eye), center=tuple(center), up=tuple(up), fov=float(fov), w=int(w), h=int(h),
                           lens_radius=None if lens_radius is None else float(lens_radius), focal_distance=float(focal_distance))

    # ---- consumers ----
    def write_obj(self, directory):
        """One Wavefront OBJ file per mesh (a single shape each: trianglemesh::obj concatenates the shapes of a file in
        reverse, trianglemesh_obj.cpp:80-91), positions / normals printed with 9 significant digits. Returns the paths."""
        import os
        os.makedirs(directory, exist_ok=True)
        paths = []
        for i, m in enumerate(self.meshes):
            lines = [f"# mesh {i} of a lmb200py.scenedesc scene", f"o mesh{i}"]
            lines += ["v %.9g %.9g %.9g" % tuple(float(x) for x in v) for v in np.asarray(m["verts"], np.float32).reshape(-1, 3)]
            has_n = m["normals"] is not None
            if has_n:
                lines += ["vn %.9g %.9g %.9g" % tuple(float(x) for x in v) for v in np.asarray(m["normals"], np.float32).reshape(-1, 3)]
            for f in np.asarray(m["faces"], np.int64).reshape(-1, 3) + 1:
                lines.append(("f %d//%d %d//%d %d//%d" % (f[0], f[0], f[1], f[1], f[2], f[2])) if has_n else ("f %d %d %d" % tuple(f)))
            path = os.path.join(directory, f"mesh{i}.obj")
            with open(path, "w") as fh:
                fh.write("\n".join(lines) + "\n")
            paths.append(path)
        return paths

    def to_yaml(self, mesh_handles, accel="qbvh", renderer="ptdirect", renderer_params=None, obj_paths=None):
        """mesh_handles: handles of meshes registered with the oracle host (type `mem`); obj_paths: instead, one OBJ file
        per mesh, loaded by the reference's own trianglemesh::obj (tinyobjloader)."""
        def v3(x):
            return " ".join(repr(float(t)) for t in x)
        out = ["lightmetrica:", "  version: 1.1.0", "  assets:"]
        if obj_paths is not None:
            for i, path in enumerate(obj_paths):
                out += [f"    mesh{i}:", "      interface: trianglemesh", "      type: obj", "      params:", f"        path: {path}"]
            mesh_handles = []
        for i, h in enumerate(mesh_handles):
            out += [f"    mesh{i}:", "      interface: trianglemesh", "      type: mem", "      params:", f"        handle: {h}"]
        for name, t in self.textures.items():
            out += [f"    {name}:", "      interface: texture", "      type: checker", "      params:", f"        scale: {t['scale']!r}",
                    f"        color1: {v3(t['color1'])}", f"        color2: {v3(t['color2'])}"]
        for name, b in self.bsdfs.items():
            out += [f"    {name}:", "      interface: bsdf", f"      type: {b['type']}", "      params:",
                    f"        TexR: {b['texR']}" if b.get("texR") else f"        R: {v3(b['R'])}"]
            if b["type"] == "cook_torrance":
                out += [f"        eta: {v3(b['eta'])}", f"        k: {v3(b['k'])}", f"        roughness: {b['roughness']!r}"]
            if b["type"] in ("refract_all", "flesnel"):
                out += [f"        eta1: {b['eta1']!r}", f"        eta2: {b['eta2']!r}"]
        for name, le in self.lights.items():
            out += [f"    {name}:", "      interface: light", "      type: area", "      params:", f"        Le: {v3(le)}"]
        for name, pl in self.point_lights.items():
            out += [f"    {name}:", "      interface: light", "      type: point", "      params:", f"        Le: {v3(pl['Le'])}",
                    f"        position: {v3(pl['position'])}"]
        for name, dl in self.dir_lights.items():
            out += [f"    {name}:", "      interface: light", "      type: directional", "      params:", f"        Le: {v3(dl['Le'])}",
                    f"        direction: {v3(dl['direction'])}"]
        for name, le in self.env_lights.items():
            out += [f"    {name}:", "      interface: light", "      type: env", "      params:", f"        Le: {v3(le)}"]
        c = self.camera
        out += ["    film1:", "      interface: film", "      type: hdr", "      params:", f"        w: {c['w']}", f"        h: {c['h']}"]
        if c.get("lens_radius") is None:
            out += ["    cam:", "      interface: sensor", "      type: pinhole", "      params:", "        film: film1", f"        fov: {c['fov']!r}"]
        else:
            out += ["    cam:", "      interface: sensor", "      type: thinlens", "      params:", "        film: film1", f"        fov: {c['fov']!r}",
                    f"        lens_radius: {c['lens_radius']!r}", f"        focal_distance: {c['focal_distance']!r}"]
        out += ["  accel:", f"    type: {accel}"]
        out += ["  scene:", "    type: scene3", "    params:", "      sensor: n_cam", "      nodes:"]
        out += ["        - id: n_cam", "          sensor: cam", "          transform:", "            lookat:",
                f"              eye: {v3(c['eye'])}", f"              center: {v3(c['center'])}", f"              up: {v3(c['up'])}"]
        for nd in self.nodes:
            if nd["mesh"] is None:
                out += [f"        - light: {nd['light']}"]
                continue
            out += [f"        - mesh: mesh{nd['mesh']}"]
            if nd["bsdf"]:
                out += [f"          bsdf: {nd['bsdf']}"]
            if nd["light"]:
                out += [f"          light: {nd['light']}"]
        out += ["  renderer:", f"    type: {renderer}"]
        if renderer_params:
            out += ["    params:"] + [f"      {k}: {v}" for k, v in renderer_params.items()]
        return "\n".join(out) + "\n"

    def flatten(self):
        """Returns (SceneDesc, keepalive) in the reference's primitive order: primitive 0 is the camera node."""
        bs_names = list(self.bsdfs.keys())
        tex_names = list(self.textures.keys())
        bs = (capi.Bsdf * (len(bs_names) + 1))()
        for i, n in enumerate(bs_names):
            b = self.bsdfs[n]
            bs[i].type = {"diffuse": capi.BSDF_DIFFUSE, "cook_torrance": capi.BSDF_COOKTORRANCE, "reflect_all": capi.BSDF_REFLECT_ALL,
                          "refract_all": capi.BSDF_REFRACT_ALL, "flesnel": capi.BSDF_FLESNEL}[b["type"]]
            bs[i].eta1, bs[i].eta2 = b["eta1"], b["eta2"]
            bs[i].R = (C.c_float * 3)(*b["R"])
            bs[i].eta = (C.c_float * 3)(*b["eta"])
            bs[i].k = (C.c_float * 3)(*b["k"])
            bs[i].roughness = b["roughness"]
            bs[i].texR = (tex_names.index(b["texR"]) + 1) if b.get("texR") else 0
        null_idx = len(bs_names)          # nodes without a bsdf get bsdf::null (scene3.cpp:315-319)
        bs[null_idx].type = capi.BSDF_NULL
        prims = (capi.Primitive * (len(self.nodes) + 1))()
        prims[0].bsdf = null_idx
        prims[0].light = -1
        lights, verts, norms, tri_prim, uvs = [], [], [], [], []
        any_normals = any(nd["mesh"] is not None and self.meshes[nd["mesh"]]["normals"] is not None for nd in self.nodes)
        any_uvs = bool(tex_names)
        first = 0
        first_prim_of_light = {}
        for pi, nd in enumerate(self.nodes, start=1):
            if nd["mesh"] is None:      # point / directional / env light node: a primitive without geometry
                prims[pi].bsdf = null_idx
                prims[pi].first_tri = first
                prims[pi].light = len(lights)
                z3 = (0.0, 0.0, 0.0)
                if nd["light"] in self.point_lights:
                    pl = self.point_lights[nd["light"]]
                    lights.append((pl["Le"], pi, capi.LIGHT_POINT, pl["position"], z3))
                elif nd["light"] in self.dir_lights:
                    dl = self.dir_lights[nd["light"]]
                    dv = np.asarray(dl["direction"], np.float32)
                    dv = dv / np.sqrt(np.float32(dv @ dv))      # Math::Normalize at load (light_directional.cpp:103)
                    lights.append((dl["Le"], pi, capi.LIGHT_DIRECTIONAL, z3, tuple(float(x) for x in dv)))
                else:
                    lights.append((self.env_lights[nd["light"]], pi, capi.LIGHT_ENV, z3, z3))
                continue
            m = self.meshes[nd["mesh"]]
            t = m["verts"][m["faces"].reshape(-1)].reshape(-1, 9)
            verts.append(t)
            if any_normals:
                norms.append(m["normals"][m["faces"].reshape(-1)].reshape(-1, 9) if m["normals"] is not None else np.zeros_like(t))
            if any_uvs:
                mu = m.get("uvs")
                uvs.append(mu[m["faces"].reshape(-1)].reshape(-1, 6) if mu is not None else np.zeros((t.shape[0], 6), np.float32))
            tri_prim.append(np.full(t.shape[0], pi, np.uint32))
            prims[pi].bsdf = bs_names.index(nd["bsdf"]) if nd["bsdf"] else null_idx
            prims[pi].light = -1
            prims[pi].first_tri = first
            prims[pi].num_tris = t.shape[0]
            prims[pi].has_normals = 1 if m["normals"] is not None else 0
            if nd["light"]:
                # A light asset is loaded once, with the FIRST primitive that references it
                # (assets.cpp:50-122 caches by id; light_area.cpp:50-55 binds mesh/transform/area
                # distribution at Load): later primitives sharing the asset sample the first one's mesh.
                bound = first_prim_of_light.setdefault(nd["light"], pi)
                prims[pi].light = len(lights)
                lights.append((self.lights[nd["light"]], bound, capi.LIGHT_AREA, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0)))
            first += t.shape[0]
        verts = np.ascontiguousarray(np.concatenate(verts), np.float32) if verts else np.zeros((0, 9), np.float32)
        tri_prim = np.ascontiguousarray(np.concatenate(tri_prim), np.uint32) if tri_prim else np.zeros(0, np.uint32)
        norms = np.ascontiguousarray(np.concatenate(norms), np.float32) if any_normals else None
        ls = (capi.Light * max(1, len(lights)))()
        for i, (le, pi, kind, pos, dirn) in enumerate(lights):
            ls[i].Le = (C.c_float * 3)(*le)
            ls[i].primitive = pi
            ls[i].kind = kind
            ls[i].position = (C.c_float * 3)(*pos)
            ls[i].direction = (C.c_float * 3)(*dirn)
        c = self.camera
        vx, vy, vz = scenes.lookat(c["eye"], c["center"], c["up"])
        cam = capi.Camera()
        cam.position = (C.c_float * 3)(*c["eye"])
        cam.vx = (C.c_float * 3)(*vx)
        cam.vy = (C.c_float * 3)(*vy)
        cam.vz = (C.c_float * 3)(*vz)
        cam.fov = float(np.radians(np.float32(c["fov"])))
        cam.width, cam.height = c["w"], c["h"]
        if c.get("lens_radius") is not None:
            cam.kind = capi.CAMERA_THINLENS
            cam.lens_radius, cam.focal_distance = c["lens_radius"], c["focal_distance"]
        # Scene3::GetSphereBound (scene3.cpp:56-78): AABB of every mesh vertex and the sensor position; centre = mid point,
        # radius = |centre - max| * 1.01
        pts = [np.asarray(c["eye"], np.float32)[None, :]] + [self.meshes[nd["mesh"]]["verts"] for nd in self.nodes if nd["mesh"] is not None]
        allp = np.concatenate(pts).astype(np.float32)
        bmin, bmax = allp.min(axis=0), allp.max(axis=0)
        centre = ((bmax + bmin) * np.float32(0.5)).astype(np.float32)
        dd = (centre - bmax).astype(np.float32)
        radius = np.float32(np.sqrt(np.float32(dd @ dd))) * np.float32(1.01)
        d = capi.SceneDesc()
        d.sphere_center = (C.c_float * 3)(*[float(x) for x in centre])
        d.sphere_radius = float(radius)
        d.num_tris = verts.shape[0]
        d.verts = verts.ctypes.data_as(C.c_void_p)
        d.normals = norms.ctypes.data_as(C.c_void_p) if norms is not None else None
        d.tri_prim = tri_prim.ctypes.data_as(C.c_void_p)
        d.num_prims = len(self.nodes) + 1
        d.prims = C.cast(prims, C.POINTER(capi.Primitive))
        d.num_bsdfs = len(bs_names) + 1
        d.bsdfs = C.cast(bs, C.POINTER(capi.Bsdf))
        d.num_lights = len(lights)
        d.lights = C.cast(ls, C.POINTER(capi.Light))
        d.camera = cam
        uvs = np.ascontiguousarray(np.concatenate(uvs), np.float32) if (any_uvs and uvs) else None
        baked = [np.ascontiguousarray(self.bake_texture(n)) for n in tex_names]
        tx = (capi.Texture * max(1, len(baked)))()
        for i, b in enumerate(baked):
            tx[i].height, tx[i].width = b.shape[0], b.shape[1]
            tx[i].rgb = b.ctypes.data_as(C.c_void_p)
        d.uvs = uvs.ctypes.data_as(C.c_void_p) if uvs is not None else None
        d.num_textures = len(baked)
        d.textures = C.cast(tx, C.POINTER(capi.Texture))
        keep = dict(verts=verts, norms=norms, tri_prim=tri_prim, prims=prims, bs=bs, ls=ls, uvs=uvs, baked=baked, tx=tx)
        return d, keep


def _box(scene, center, half, angle_deg, bsdf):
    """Axis-aligned cuboid rotated about +y, 12 triangles, outward normals."""
    cx, cy, cz = center
    hx, hy, hz = half
    a = np.radians(angle_deg)
    ca, sa = np.cos(a), np.sin(a)

    def P(x, y, z):
        return (cx + ca * x + sa * z, cy + y, cz - sa * x + ca * z)
    v = [P(sx * hx, sy * hy, sz * hz) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    # vertex index = 4*ix + 2*iy + iz
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for q in quads:
        a_, b_, c_, d_ = (np.asarray(v[i], np.float32) for i in q)
        tris += [np.concatenate([a_, b_, c_]), np.concatenate([a_, c_, d_])]
    scene.add_mesh_tris(np.stack(tris), bsdf)


def cornell_box(w=512, h=512, glossy_block=False, thinlens=False):
    """Cornell-style box, 36 triangles: 5 walls, 2 blocks, 1 ceiling light (config 0 of BASELINE.json)."""
    s = Scene()
    s.add_bsdf("white", "diffuse", (0.75, 0.75, 0.75))
    s.add_bsdf("red", "diffuse", (0.75, 0.25, 0.25))
    s.add_bsdf("green", "diffuse", (0.25, 0.75, 0.25))
    s.add_bsdf("metal", "cook_torrance", (1.0, 1.0, 1.0), roughness=0.2)
    s.add_light("lamp", (17.0, 12.0, 4.0))
    s.add_quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1), "white")          # floor (normal +y)
    s.add_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), "white")          # ceiling (normal -y)
    s.add_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), "white")        # back wall (normal +z)
    s.add_quad((-1, 0, -1), (-1, 2, -1), (-1, 2, 1), (-1, 0, 1), "red")          # left wall (normal +x)
    s.add_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), "green")            # right wall (normal -x)
    _box(s, (-0.35, 0.6, -0.3), (0.3, 0.6, 0.3), 18.0, "white")                  # tall block
    _box(s, (0.4, 0.3, 0.3), (0.3, 0.3, 0.3), -17.0, "metal" if glossy_block else "white")   # short block
    s.add_quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), "white", "lamp")  # light, faces down
    if thinlens:   # sensor::thinlens focused on the front face of the short block
        s.set_camera((0, 1, 4.2), (0, 1, 0), (0, 1, 0), 40.0, w, h, lens_radius=0.2, focal_distance=3.6)
    else:
        s.set_camera((0, 1, 4.2), (0, 1, 0), (0, 1, 0), 40.0, w, h)
    return s


def config2_scene(target_tris=1_000_000, w=1920, h=1080, seed=42, half=50.0, n_objects=200, n_lights=6):
    """BASELINE.json configs[2]: ground + ~200 tessellated objects (about target_tris triangles) in
    [-half,half]^3, 70 % diffuse (R uniform in [0.2,0.8]^3) / 30 % cook_torrance (roughness in
    {0.05,0.1,0.3}, default eta/k), n_lights area-light quads facing down."""
    g = np.random.Generator(np.random.Philox(seed + 1000))
    verts, oid = scenes.mesh_scene(target_tris, seed=seed, half=half, n_objects=n_objects)
    s = Scene()
    n_pal = 24
    for i in range(n_pal):
        if i % 10 < 7:
            s.add_bsdf(f"m{i}", "diffuse", tuple(0.2 + 0.6 * g.random(3)))
        else:
            s.add_bsdf(f"m{i}", "cook_torrance", (1.0, 1.0, 1.0), roughness=[0.05, 0.1, 0.3][i % 3])
    s.add_bsdf("ground", "diffuse", (0.6, 0.6, 0.6))
    s.add_bsdf("lampb", "diffuse", (0.8, 0.8, 0.8))
    for k in range(n_lights):
        s.add_light(f"lamp{k}", (60.0, 55.0, 45.0))   # one asset per light primitive (see Scene.flatten)
    # objects are contiguous in oid
    bounds_idx = np.flatnonzero(np.diff(oid)) + 1
    starts = np.concatenate([[0], bounds_idx])
    ends = np.concatenate([bounds_idx, [len(oid)]])
    for k, (b, e) in enumerate(zip(starts, ends)):
        s.add_mesh_tris(verts[b:e], "ground" if oid[b] == 0 else f"m{int(g.integers(n_pal))}")
    for k in range(n_lights):
        cx = (k % 3 - 1) * half * 0.6
        cz = (k // 3 - 0.5) * half * 0.8
        y = half * 0.75
        r = half * 0.08
        # faces down: (b-a)x(c-a) = -y
        s.add_quad((cx - r, y, cz - r), (cx + r, y, cz - r), (cx + r, y, cz + r), (cx - r, y, cz + r), "lampb", f"lamp{k}")
    s.set_camera((0.0, half * 0.45, half * 1.05), (0.0, half * 0.1, 0.0), (0, 1, 0), 45.0, w, h)
    return s


def config4_scene(w=3840, h=2160, instances=100, asset_tris=100_000):
    """BASELINE.json configs[4]: a 10M-triangle "instanced" scene — `instances` copies of one `asset_tris`-triangle asset on a
    10 x 10 grid, flattened to world-space triangles exactly as the reference's accels do (accel_qbvh.cpp:161-194) — under one
    area light, 4K film."""
    base, _ = scenes.mesh_scene(asset_tris, seed=7, half=5.0, n_objects=20)
    side = int(np.ceil(np.sqrt(instances)))
    inst = []
    for k in range(instances):
        off = np.array([(k % side - (side - 1) / 2) * 10.0, 0.0, (k // side - (side - 1) / 2) * 10.0], np.float32)
        inst.append((base.reshape(-1, 3) + off).reshape(-1, 9))
    verts = np.ascontiguousarray(np.concatenate(inst), np.float32)
    sc = Scene()
    sc.add_bsdf("w", "diffuse", (0.6, 0.6, 0.6))
    sc.add_light("lamp", (40.0, 40.0, 40.0))
    sc.add_mesh_tris(verts, "w")
    sc.add_quad((-20, 30, -20), (20, 30, -20), (20, 30, 20), (-20, 30, 20), "w", "lamp")
    sc.set_camera((0.0, 25.0, 70.0), (0.0, 0.0, 0.0), (0, 1, 0), 45.0, w, h)
    return sc, verts


def specular_box(w=64, h=64, point_light=True):
    """Cornell-style box with a glass sphere (bsdf::flesnel), a mirror block (bsdf::reflect_all), a refract_all slab
    and, optionally, a light::point besides the ceiling area light: exercises the delta BSDFs and the delta light."""
    s = Scene()
    s.add_bsdf("white", "diffuse", (0.75, 0.75, 0.75))
    s.add_bsdf("red", "diffuse", (0.75, 0.25, 0.25))
    s.add_bsdf("green", "diffuse", (0.25, 0.75, 0.25))
    s.add_bsdf("glass", "flesnel", (1.0, 1.0, 1.0), eta1=1.0, eta2=1.5)
    s.add_bsdf("mirror", "reflect_all", (0.9, 0.9, 0.9))
    s.add_bsdf("slab", "refract_all", (0.95, 0.95, 1.0), eta1=1.0, eta2=1.3)
    s.add_light("lamp", (17.0, 12.0, 4.0))
    s.add_quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1), "white")
    s.add_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), "white")
    s.add_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), "white")
    s.add_quad((-1, 0, -1), (-1, 2, -1), (-1, 2, 1), (-1, 0, 1), "red")
    s.add_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), "green")
    s.add_mesh_tris(scenes.sphere((0.4, 0.45, 0.3), 0.4, 24, 16), "glass")
    _box(s, (-0.4, 0.6, -0.35), (0.3, 0.6, 0.3), 20.0, "mirror")
    _box(s, (-0.45, 0.25, 0.55), (0.25, 0.25, 0.05), -10.0, "slab")
    s.add_quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), "white", "lamp")
    if point_light:
        s.add_point_light("bulb", (1.5, 1.5, 2.0), (0.6, 1.5, 0.6))
    s.set_camera((0, 1, 4.2), (0, 1, 0), (0, 1, 0), 40.0, w, h)
    return s


def outdoor_scene(w=64, h=36, light="directional", thinlens=False):
    """Open scene (ground + a few objects, no enclosure) lit by light::directional, light::env (constant Le) or both,
    optionally seen through sensor::thinlens focused on the middle object: exercises the bounding-sphere emitter shapes
    (geom.infinite) and the lens sampling."""
    if light == "cornell":
        return cornell_box(w, h, glossy_block=True, thinlens=thinlens)
    if light == "textured":
        return textured_box(w, h)
    s = Scene()
    s.add_bsdf("ground", "diffuse", (0.6, 0.55, 0.5))
    s.add_bsdf("red", "diffuse", (0.7, 0.2, 0.2))
    s.add_bsdf("blue", "diffuse", (0.2, 0.3, 0.7))
    s.add_bsdf("metal", "cook_torrance", (1.0, 1.0, 1.0), roughness=0.3)
    s.add_quad((-6, 0, -6), (-6, 0, 6), (6, 0, 6), (6, 0, -6), "ground")
    s.add_mesh_tris(scenes.sphere((0.0, 0.8, 0.0), 0.8, 20, 14), "red")
    _box(s, (-1.9, 0.6, -1.2), (0.5, 0.6, 0.5), 25.0, "blue")
    s.add_mesh_tris(scenes.torus((1.9, 0.45, 1.0), 0.7, 0.25, 20, 10), "metal")
    if light in ("directional", "both"):
        s.add_directional_light("sun", (3.0, 2.8, 2.5), (-0.4, -1.0, -0.3))
    if light in ("env", "both"):
        s.add_env_light("sky", (0.5, 0.6, 0.8))
    if thinlens:
        s.set_camera((0.0, 2.0, 6.0), (0.0, 0.7, 0.0), (0, 1, 0), 35.0, w, h, lens_radius=0.25, focal_distance=6.1)
    else:
        s.set_camera((0.0, 2.0, 6.0), (0.0, 0.7, 0.0), (0, 1, 0), 35.0, w, h)
    return s


def textured_box(w=48, h=48):
    """Cornell-style box whose floor and back wall carry checker textures (TexR on bsdf::diffuse and on
    bsdf::cook_torrance); the checker scales divide Scene.tex_res, so the baked textures are exact."""
    s = Scene()
    s.add_texture("tiles", 8.0, (0.8, 0.2, 0.2), (0.9, 0.9, 0.9))
    s.add_texture("stripes", 4.0, (0.2, 0.3, 0.8), (0.8, 0.8, 0.3))
    s.add_bsdf("white", "diffuse", (0.75, 0.75, 0.75))
    s.add_bsdf("floor", "diffuse", texR="tiles")
    s.add_bsdf("back", "cook_torrance", roughness=0.3, texR="stripes")
    s.add_bsdf("red", "diffuse", (0.75, 0.25, 0.25))
    s.add_light("lamp", (17.0, 12.0, 4.0))
    s.add_quad((-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1), "floor", uv=True)
    s.add_quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), "white")
    s.add_quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), "back", uv=True)
    s.add_quad((-1, 0, -1), (-1, 2, -1), (-1, 2, 1), (-1, 0, 1), "red")
    s.add_quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), "white")
    _box(s, (0.3, 0.3, 0.2), (0.3, 0.3, 0.3), -17.0, "floor")       # textured BSDF on a mesh without texcoords: uv = 0
    s.add_quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), "white", "lamp")
    s.set_camera((0, 1, 4.2), (0, 1, 0), (0, 1, 0), 40.0, w, h)
    return s
