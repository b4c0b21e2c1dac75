"""TEST INFRASTRUCTURE: ctypes bindings for the oracles. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this module.

  liblmoracle.so            C restatement of the reference hot path (travels to the GPU box)
  _ref/liblightmetrica.so   the reference itself, compiled here from /root/reference (travels as a binary)
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(HERE, "liblmoracle.so")
REF_PATH = os.path.join(HERE, "_ref", "liblightmetrica.so")

_port = None
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def port():
    global _port
    if _port is None:
        L = C.CDLL(PORT_PATH)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_tris.restype = C.c_void_p
        L.orc_scene_tris.argtypes = [C.c_void_p]
        L.orc_closest.restype = C.c_longlong
        L.orc_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_any.restype = C.c_longlong
        L.orc_any.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.orc_wide_closest.restype = C.c_longlong
        L.orc_wide_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.orc_wide_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.orc_wide_count_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.orc_wide_count_cull.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
        L.orc_triaccel_load.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_triaccel_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _port = L
    return _port


def have_ref():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_PATH, mode=C.RTLD_GLOBAL)
        L.ref_last_error.restype = C.c_char_p
        L.ref_session_create.restype = C.c_void_p
        L.ref_session_create.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_session_destroy.argtypes = [C.c_void_p]
        L.ref_load_plugin.argtypes = [C.c_char_p]
        L.ref_register_mesh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_intersect_batch.restype = C.c_longlong
        L.ref_intersect_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_triaccel_load.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_world_triangles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_render.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_film_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_num_primitives.argtypes = [C.c_void_p]
        L.ref_nanort_create.restype = C.c_void_p
        L.ref_nanort_create.argtypes = [C.c_void_p, C.c_ulonglong, C.POINTER(C.c_double)]
        L.ref_nanort_destroy.argtypes = [C.c_void_p]
        L.ref_nanort_trace.restype = C.c_double
        L.ref_nanort_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p]
        _ref = L
    return _ref


class PortScene:
    """C restatement: TriAccel records + closest/any hit over a flat world-space triangle list."""

    def __init__(self, verts):
        self.verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        self.h = port().orc_scene_create(_p(self.verts), self.verts.shape[0])

    def __del__(self):
        if getattr(self, "h", None):
            port().orc_scene_destroy(self.h)
            self.h = None

    def records(self):
        n = self.verts.shape[0]
        ptr = port().orc_scene_tris(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(n, 12)).copy()

    def closest(self, rays, use_bvh=True):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        tuv = np.zeros((n, 3), np.float32)
        tri = np.zeros(n, np.int32)
        port().orc_closest(self.h, _p(rays), n, 1 if use_bvh else 0, _p(tuv), _p(tri))
        return tuv, tri

    def any(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], np.uint8)
        port().orc_any(self.h, _p(rays), rays.shape[0], _p(occ))
        return occ


def wide_closest(units, grid, rays):
    """Scalar walk of the product's flattened 64-byte units (host-logic checker, no GPU). grid = (lo[3], step[3])."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = rays.shape[0]
    tuv = np.zeros((n, 3), np.float32)
    tri = np.zeros(n, np.int32)
    g = np.ascontiguousarray(np.concatenate([grid[0], grid[1]]), np.float32)
    port().orc_wide_closest(_p(units), _p(g), _p(rays), n, _p(tuv), _p(tri))
    return tuv, tri


def wide_count(units, grid, rays):
    """(nodes per ray, triangle records per ray) of an ordered CPU walk of the product's structure: tree-quality figure."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    g = np.ascontiguousarray(np.concatenate([grid[0], grid[1]]), np.float32)
    out = np.zeros(2, np.float64)
    port().orc_wide_count(_p(units), _p(g), _p(rays), rays.shape[0], _p(out))
    return out[0] / rays.shape[0], out[1] / rays.shape[0]


def wide_count_sorted(units, grid, rays):
    """(nodes, triangle records, entries dropped at pop time) per ray of an EXACT front-to-back walk with entry-distance cull:
    what an ideally ordered traversal of the same tree would fetch (analysis only)."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    g = np.ascontiguousarray(np.concatenate([grid[0], grid[1]]), np.float32)
    out = np.zeros(3, np.float64)
    port().orc_wide_count_sorted(_p(units), _p(g), _p(rays), rays.shape[0], _p(out))
    return tuple(out / rays.shape[0])


def wide_count_cull(units, grid, rays, mode):
    """orc_wide_count's octant-ordered walk with an entry-distance cull (mode 1: one bound per stacked sibling group, mode 2: per
    child): (nodes, triangle records, children dropped unfetched) per ray. Analysis only."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    g = np.ascontiguousarray(np.concatenate([grid[0], grid[1]]), np.float32)
    out = np.zeros(3, np.float64)
    port().orc_wide_count_cull(_p(units), _p(g), _p(rays), rays.shape[0], int(mode), _p(out))
    return tuple(out / rays.shape[0])


def mesh_scene_yaml(handle, accel="qbvh", w=16, h=16):
    """A minimal Lightmetrica scene around one in-memory mesh (identity transform), as the
    reference's Stub_Scene does for accel tests (test_accel3.cpp:226-264)."""
    return f"""
lightmetrica:
  assets:
    mesh1:
      interface: trianglemesh
      type: mem
      params:
        handle: {handle}
    white:
      interface: bsdf
      type: diffuse
      params:
        R: 0.8 0.8 0.8
    film1:
      interface: film
      type: hdr
      params:
        w: {w}
        h: {h}
    cam:
      interface: sensor
      type: pinhole
      params:
        film: film1
        fov: 45
  accel:
    type: {accel}
  scene:
    sensor: n_cam
    nodes:
      - id: n_cam
        sensor: cam
        transform:
          lookat:
            eye: 0.5 0.5 3
            center: 0.5 0.5 0
            up: 0 1 0
      - mesh: mesh1
        bsdf: white
"""


class RefSoup:
    """The reference itself (oracle/_ref) over a flat triangle list: one primitive (index 1, after
    the camera node) whose mesh holds the triangles unshared, texcoords (0,0),(1,0),(0,1) so that
    the interpolated uv of the Intersection returns the barycentrics exactly."""

    def __init__(self, verts, accel="qbvh"):
        L = ref()
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        n = verts.shape[0]
        ps = verts.reshape(-1, 3)
        fs = np.arange(3 * n, dtype=np.uint32)
        ts = np.tile(np.array([0, 0, 1, 0, 0, 1], np.float32), n)
        h = L.ref_register_mesh(_p(ps), 3 * n, None, _p(ts), _p(fs), n)
        self.s = L.ref_session_create(mesh_scene_yaml(h, accel).encode(), accel.encode())
        if not self.s:
            raise RuntimeError(L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "s", None):
            ref().ref_session_destroy(self.s)
            self.s = None

    def intersect(self, rays, threads=1):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        face = np.zeros(n, np.int32)
        tuv = np.zeros((n, 3), np.float32)
        geom = np.zeros((n, 11), np.float32)
        sec = C.c_double()
        ref().ref_intersect_batch(self.s, n, _p(rays), threads, _p(prim), _p(face), _p(tuv), _p(geom), C.byref(sec))
        return dict(prim=prim, face=face, tuv=tuv, geom=geom, seconds=sec.value)


class RefNanort:
    """nanort (the reference's vendored third-party tracer) called as accel::nanort calls it — a THROUGHPUT baseline only:
    its triangle test is not TriAccel and the reference's own wrapper is broken (oracle/ref/nanort_bench.cpp)."""

    def __init__(self, verts):
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        sec = C.c_double()
        self.h = ref().ref_nanort_create(_p(v), v.shape[0], C.byref(sec))
        if not self.h:
            raise RuntimeError("nanort build failed")
        self.build_seconds = sec.value

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_nanort_destroy(self.h)
            self.h = None

    def trace(self, rays, threads=1, want_hits=True):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        face = np.zeros(n, np.int32) if want_hits else None
        t = np.zeros(n, np.float32) if want_hits else None
        sec = ref().ref_nanort_trace(self.h, _p(rays), n, threads, _p(face) if want_hits else None, _p(t) if want_hits else None)
        return dict(face=face, t=t, seconds=sec)


def ref_triaccel_records(verts):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    out = np.zeros((verts.shape[0], 12), np.uint32)
    L = ref()
    for i in range(verts.shape[0]):
        v = verts[i]
        L.ref_triaccel_load(_p(v[0:3].copy()), _p(v[3:6].copy()), _p(v[6:9].copy()), _p(out[i]))
    return out


# ---------------------------------------------------------------------------------------------
# Path tracing oracles


def _port_pt():
    L = port()
    if not hasattr(L, "_pt_ready"):
        L.orc_pt_scene_create.restype = C.c_void_p
        L.orc_pt_scene_create.argtypes = [C.c_void_p]
        L.orc_pt_scene_destroy.argtypes = [C.c_void_p]
        L.orc_pt_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_pt_render_tile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pt_render_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_render_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L._pt_ready = True
    return L


class PortPT:
    """C restatement of renderer::pt / renderer::ptdirect over a flattened scene (same POD layout as lmb200_scene_desc)."""

    def __init__(self, scene):
        self.desc, self.keep = scene.flatten()
        self.w, self.h = scene.camera["w"], scene.camera["h"]
        self.s = _port_pt().orc_pt_scene_create(C.byref(self.desc))

    def __del__(self):
        if getattr(self, "s", None):
            _port_pt().orc_pt_scene_destroy(self.s)
            self.s = None

    def render(self, mode, num_samples, seed=1, max_verts=-1, min_verts=0, begin=0, end=None, tile=None, primary_tile=0):
        """Returns (film (H,W,3) scaled by W*H/num_samples, counts [extend, shadow]). tile = (x0, y0, x1, y1): camera samples
        drawn inside that raster rectangle (tile partitioning, include/lmb200.h). primary_tile: lmb200_render_params::primary_tile
        (0 = the library's automatic choice, asked from lmb200_default_primary_tile; -1 = independent samples)."""
        end = num_samples if end is None else end
        film = np.zeros((self.h, self.w, 4), np.float32)
        counts = np.zeros(2, np.int64)
        t = np.asarray((0.0, 0.0, 1.0, 1.0) if tile is None else tile, np.float32)
        if primary_tile == 0:
            # the product's own rule for "automatic" (a function of the job size): asked from the C ABI so that port and
            # device draw the same camera samples
            from lmb200py import capi
            primary_tile = capi.lib().lmb200_default_primary_tile(self.w, self.h, int(num_samples))
        _port_pt().orc_pt_render_ex(self.s, mode, max_verts, min_verts, seed, begin, end, _p(t), int(primary_tile), _p(film), _p(counts))
        return film[..., :3] * np.float32(self.w * self.h / num_samples), counts

    def render_normal(self):
        film = np.zeros((self.h, self.w, 4), np.float32)
        tri = np.zeros((self.h, self.w), np.int32)
        _port_pt().orc_render_normal(self.s, _p(film), _p(tri))
        return film[..., :3], tri


class RefScene:
    """The reference itself (oracle/_ref) on a scenedesc.Scene: real scene3 + assets + renderers."""

    def __init__(self, scene, accel="qbvh", plugins=(), obj_paths=None):
        """obj_paths: load the meshes from these Wavefront OBJ files through the reference's trianglemesh::obj instead of
        registering the in-memory arrays with the host shim."""
        L = ref()
        for p in plugins:
            if not L.ref_load_plugin(p.encode()):
                raise RuntimeError(f"failed to load plugin {p}")
        handles = []
        self._keep = []
        self.obj_paths = obj_paths
        for m in (scene.meshes if obj_paths is None else []):
            ps = np.ascontiguousarray(m["verts"], np.float32)
            fs = np.ascontiguousarray(m["faces"], np.uint32)
            ns = None if m["normals"] is None else np.ascontiguousarray(m["normals"], np.float32)
            ts = None if m.get("uvs") is None else np.ascontiguousarray(m["uvs"], np.float32)
            self._keep.append((ps, fs, ns, ts))
            handles.append(L.ref_register_mesh(_p(ps), ps.shape[0], _p(ns) if ns is not None else None,
                                               _p(ts) if ts is not None else None, _p(fs), fs.shape[0]))
        self.scene = scene
        self.handles = handles
        self.accel = accel
        self.yaml = scene.to_yaml(handles, accel=accel, obj_paths=obj_paths)
        self.s = L.ref_session_create(self.yaml.encode(), accel.encode())
        if not self.s:
            raise RuntimeError(L.ref_last_error().decode())

    def __del__(self):
        if getattr(self, "s", None):
            ref().ref_session_destroy(self.s)
            self.s = None

    def render(self, renderer, num_samples, seed=1, threads=1, max_verts=-1, extra=None, in_tree=False):
        """in_tree=True: the renderer params live in the scene YAML itself (as with the real CLI), so a
        plugin can walk prop->Tree() to the assets; costs a fresh session."""
        w, h = self.scene.camera["w"], self.scene.camera["h"]
        out = np.zeros((h, w, 3), np.float32)
        pd = {"num_samples": int(num_samples), "max_num_vertices": int(max_verts), "min_num_vertices": 0}
        pd.update(extra or {})
        ow, oh, sec = C.c_int(), C.c_int(), C.c_double()
        if in_tree:
            y = self.scene.to_yaml(self.handles, accel=self.accel, renderer=renderer, renderer_params=pd, obj_paths=self.obj_paths)
            s = ref().ref_session_create(y.encode(), self.accel.encode())
            if not s:
                raise RuntimeError(ref().ref_last_error().decode())
            try:
                ok = ref().ref_render(s, None, None, seed, threads, _p(out), C.byref(ow), C.byref(oh), C.byref(sec))
            finally:
                ref().ref_session_destroy(s)
        else:
            params = "".join(f"{k}: {v}\n" for k, v in pd.items())
            ok = ref().ref_render(self.s, renderer.encode(), params.encode(), seed, threads, _p(out), C.byref(ow), C.byref(oh), C.byref(sec))
        if not ok:
            raise RuntimeError(ref().ref_last_error().decode())
        return out, sec.value

    def intersect(self, rays, threads=1):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        prim = np.zeros(n, np.int32)
        face = np.zeros(n, np.int32)
        tuv = np.zeros((n, 3), np.float32)
        geom = np.zeros((n, 11), np.float32)
        sec = C.c_double()
        ref().ref_intersect_batch(self.s, n, _p(rays), threads, _p(prim), _p(face), _p(tuv), _p(geom), C.byref(sec))
        return dict(prim=prim, face=face, tuv=tuv, geom=geom, seconds=sec.value)
