"""ctypes bindings for include/lmb200.h. Every call goes to liblmb200.so; there is no fallback."""
import ctypes as C
import os
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("LMB200_LIB", os.path.join(_ROOT, "lib", "liblmb200.so"))   # override: tuning variants only

MISS = 0xFFFFFFFF
MODE_PT, MODE_PTDIRECT, MODE_NORMAL, MODE_PTMIS = 0, 1, 2, 3
BSDF_NULL, BSDF_DIFFUSE, BSDF_COOKTORRANCE, BSDF_REFLECT_ALL, BSDF_REFRACT_ALL, BSDF_FLESNEL = 0, 1, 2, 3, 4, 5
LIGHT_AREA, LIGHT_POINT, LIGHT_DIRECTIONAL, LIGHT_ENV = 0, 1, 2, 3
CAMERA_PINHOLE, CAMERA_THINLENS = 0, 1
BUILD_HOST_SAH, BUILD_GPU_LBVH, BUILD_GPU_PLOC, BUILD_GPU_LBVH_SAH = 0, 1, 2, 3
BUILD_DEFAULT = BUILD_GPU_LBVH        # what lmb200_accel_build / lmb200_scene_create use on a device accel

RAY_DTYPE = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tmin", "f4"),
                      ("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("tmax", "f4")])
HIT_DTYPE = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("tri", "u4")])


class AccelStats(C.Structure):
    _fields_ = [("num_triangles", C.c_uint64), ("num_valid_triangles", C.c_uint64), ("num_nodes", C.c_uint64),
                ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64), ("build_seconds", C.c_double),
                ("upload_seconds", C.c_double), ("sah_cost", C.c_float), ("max_depth", C.c_int)]


class Bsdf(C.Structure):
    _fields_ = [("type", C.c_int32), ("R", C.c_float * 3), ("eta", C.c_float * 3), ("k", C.c_float * 3), ("roughness", C.c_float),
                ("eta1", C.c_float), ("eta2", C.c_float), ("texR", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.c_void_p)]


class Primitive(C.Structure):
    _fields_ = [("bsdf", C.c_int32), ("light", C.c_int32), ("first_tri", C.c_uint32), ("num_tris", C.c_uint32), ("has_normals", C.c_int32)]


class Light(C.Structure):
    _fields_ = [("Le", C.c_float * 3), ("primitive", C.c_int32), ("kind", C.c_int32), ("position", C.c_float * 3),
                ("direction", C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("vx", C.c_float * 3), ("vy", C.c_float * 3), ("vz", C.c_float * 3),
                ("fov", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("kind", C.c_int32), ("lens_radius", C.c_float), ("focal_distance", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("num_tris", C.c_uint64), ("verts", C.c_void_p), ("normals", C.c_void_p), ("tri_prim", C.c_void_p),
                ("num_prims", C.c_uint32), ("prims", C.POINTER(Primitive)),
                ("num_bsdfs", C.c_uint32), ("bsdfs", C.POINTER(Bsdf)),
                ("num_lights", C.c_uint32), ("lights", C.POINTER(Light)),
                ("camera", Camera), ("sphere_center", C.c_float * 3), ("sphere_radius", C.c_float),
                ("uvs", C.c_void_p), ("num_textures", C.c_uint32), ("textures", C.POINTER(Texture))]


class RenderParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("num_samples", C.c_int64), ("sample_begin", C.c_int64), ("sample_end", C.c_int64),
                ("max_num_vertices", C.c_int32), ("min_num_vertices", C.c_int32), ("seed", C.c_uint64), ("pool_size", C.c_int32),
                ("tile", C.c_float * 4), ("primary_tile", C.c_int32), ("count_work", C.c_int32), ("tile_partition", C.c_int32)]


class BvhLayout(C.Structure):
    _fields_ = [("units", C.c_void_p), ("num_units", C.c_uint64), ("num_nodes", C.c_uint64), ("num_triangles", C.c_uint64),
                ("grid_lo", C.c_float * 3), ("grid_step", C.c_float * 3)]


class RenderStats(C.Structure):
    _fields_ = [("samples", C.c_int64), ("extend_rays", C.c_int64), ("shadow_rays", C.c_int64), ("iterations", C.c_int64),
                ("launches", C.c_uint64), ("seconds", C.c_double), ("reduce_seconds", C.c_double), ("vertices", C.c_int64),
                ("extend_nodes", C.c_int64), ("extend_tris", C.c_int64), ("shadow_nodes", C.c_int64), ("shadow_tris", C.c_int64)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.c_int64)

EXPORTS = [
    "lmb200_last_error", "lmb200_device_count", "lmb200_accel_create", "lmb200_accel_destroy", "lmb200_accel_build", "lmb200_accel_build_ex",
    "lmb200_accel_get_stats", "lmb200_accel_device", "lmb200_accel_replicate", "lmb200_trace_closest", "lmb200_trace_closest_compact", "lmb200_trace_any_compact", "lmb200_trace_closest_one", "lmb200_trace_closest_one_mt", "lmb200_trace_closest_dev", "lmb200_trace_any", "lmb200_trace_any_dev",
    "lmb200_trace_count_dev", "lmb200_launch_count", "lmb200_accel_host_layout", "lmb200_accel_create_host_only",
    "lmb200_scene_create", "lmb200_scene_create_ex", "lmb200_scene_create_shared", "lmb200_registry_put", "lmb200_registry_get", "lmb200_scene_destroy", "lmb200_scene_accel", "lmb200_render_dev", "lmb200_film_rescale_dev",
    "lmb200_render", "lmb200_render_multi", "lmb200_render_timed", "lmb200_default_primary_tile",
]

_lib = None


def _point_at_torch_nccl():
    """lmb200_render_multi dlopens libnccl lazily. In a Python process that may also import torch, both must end up with the
    same libnccl.so.2 (objects are shared by soname; torch's libtorch_cuda.so needs the newer copy it ships with), so the
    library is told to use the wheel's copy when there is one."""
    if os.environ.get("LMB200_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["LMB200_NCCL_LIB"] = cand
                return
    except Exception:      # noqa: BLE001
        pass


def lib():
    """Loads liblmb200.so; raises (loudly) if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run ./build.sh (or __graft_entry__.build()). lmb200 has no CPU fallback.")
    _point_at_torch_nccl()
    L = C.CDLL(LIB_PATH)
    L.lmb200_last_error.restype = C.c_char_p
    L.lmb200_accel_create.restype = C.c_void_p
    L.lmb200_accel_create.argtypes = [C.c_int]
    L.lmb200_accel_create_host_only.restype = C.c_void_p
    L.lmb200_accel_destroy.argtypes = [C.c_void_p]
    L.lmb200_accel_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.lmb200_accel_build_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    L.lmb200_accel_get_stats.argtypes = [C.c_void_p, C.POINTER(AccelStats)]
    L.lmb200_accel_device.argtypes = [C.c_void_p]
    L.lmb200_accel_replicate.restype = C.c_void_p
    L.lmb200_accel_replicate.argtypes = [C.c_void_p, C.c_int]
    L.lmb200_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.lmb200_trace_closest_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_uint64]
    L.lmb200_trace_any_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_uint64]
    L.lmb200_trace_closest_one.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.lmb200_trace_closest_one_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    L.lmb200_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.lmb200_trace_closest_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.lmb200_trace_any_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    L.lmb200_trace_count_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.lmb200_launch_count.restype = C.c_uint64
    L.lmb200_accel_host_layout.argtypes = [C.c_void_p, C.POINTER(BvhLayout)]
    if hasattr(L, "lmb200_scene_create"):
        L.lmb200_scene_create.restype = C.c_void_p
        L.lmb200_scene_create.argtypes = [C.c_int, C.POINTER(SceneDesc)]
        L.lmb200_scene_create_ex.restype = C.c_void_p
        L.lmb200_scene_create_ex.argtypes = [C.c_int, C.POINTER(SceneDesc), C.c_int]
        L.lmb200_scene_create_shared.restype = C.c_void_p
        L.lmb200_scene_create_shared.argtypes = [C.POINTER(SceneDesc), C.c_void_p]
        L.lmb200_registry_put.argtypes = [C.c_void_p, C.c_void_p]
        L.lmb200_registry_get.restype = C.c_void_p
        L.lmb200_registry_get.argtypes = [C.c_void_p]
        L.lmb200_scene_destroy.argtypes = [C.c_void_p]
        L.lmb200_scene_accel.restype = C.c_void_p
        L.lmb200_scene_accel.argtypes = [C.c_void_p]
        L.lmb200_render_dev.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]
        L.lmb200_film_rescale_dev.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        L.lmb200_render.argtypes = [C.c_void_p, C.POINTER(RenderParams), C.c_void_p, C.POINTER(RenderStats)]
        L.lmb200_default_primary_tile.argtypes = [C.c_int, C.c_int, C.c_int64]
        L.lmb200_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RenderParams), C.c_void_p, C.POINTER(RenderStats)]
        L.lmb200_render_timed.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(RenderParams), C.c_double, C.c_int64, C.c_double,
                                          PROGRESS_FN, C.c_void_p, C.c_void_p, C.POINTER(RenderStats)]
    _lib = L
    return L


class LmbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise LmbError(f"lmb200 error {rc}: {lib().lmb200_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Accel:
    """Mirror of the reference's Accel interface (Initialize/Build/Intersect, accel.h:67-79, accel3.h:68) for batches."""

    def __init__(self, device=0, host_only=False):
        L = lib()
        self.h = L.lmb200_accel_create_host_only() if host_only else L.lmb200_accel_create(device)
        if not self.h:
            raise LmbError(L.lmb200_last_error().decode())
        self.host_only = host_only

    def close(self):
        if self.h:
            lib().lmb200_accel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build(self, verts, builder=None):
        """builder=None: the library default (device builder on a device accel, host SAH on a host-only one)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 9)
        self._verts = verts
        if builder is None:
            check(lib().lmb200_accel_build(self.h, _ptr(verts), verts.shape[0]))
        else:
            check(lib().lmb200_accel_build_ex(self.h, _ptr(verts), verts.shape[0], builder))
        return self.stats()

    def stats(self):
        s = AccelStats()
        check(lib().lmb200_accel_get_stats(self.h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in AccelStats._fields_}

    def host_layout(self):
        """The flattened structure as (units (N,64) uint8 copy, num_nodes, num_triangles, (grid_lo, grid_step))."""
        lay = BvhLayout()
        check(lib().lmb200_accel_host_layout(self.h, C.byref(lay)))
        units = np.ctypeslib.as_array(C.cast(lay.units, C.POINTER(C.c_uint8)), shape=(lay.num_units, 64)).copy()
        return units, int(lay.num_nodes), int(lay.num_triangles), (np.array(lay.grid_lo[:], np.float32), np.array(lay.grid_step[:], np.float32))

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        check(lib().lmb200_trace_closest(self.h, _ptr(rays), _ptr(hits), rays.shape[0]))
        return hits

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        check(lib().lmb200_trace_any(self.h, _ptr(rays), _ptr(occ), rays.shape[0]))
        return occ


class Scene:
    """Mirror of the reference's Renderer interface (Initialize/Render, renderer.h:68-81) over a flattened scene."""

    def __init__(self, scene, device=0, builder=BUILD_DEFAULT):
        self.desc, self.keep = scene.flatten()
        self.w, self.h = scene.camera["w"], scene.camera["h"]
        self.h_ = lib().lmb200_scene_create_ex(device, C.byref(self.desc), builder)
        if not self.h_:
            raise LmbError(lib().lmb200_last_error().decode())

    def close(self):
        if getattr(self, "h_", None):
            lib().lmb200_scene_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def params(self, mode, num_samples, seed=1, max_verts=-1, min_verts=0, begin=0, end=None, pool=0, tile=None, tile_partition=False, primary_tile=0):
        p = RenderParams()
        p.primary_tile = primary_tile
        if tile is not None:
            p.tile = (C.c_float * 4)(*tile)
        p.tile_partition = 1 if tile_partition else 0
        p.mode, p.num_samples, p.sample_begin = mode, num_samples, begin
        p.sample_end = num_samples if end is None else end
        p.max_num_vertices, p.min_num_vertices, p.seed, p.pool_size = max_verts, min_verts, seed, pool
        return p

    def render(self, mode, num_samples, **kw):
        """Host-buffer render through lmb200_render: returns (film (H,W,3) float32, stats dict)."""
        p = self.params(mode, num_samples, **kw)
        film = np.zeros((self.h, self.w, 4), np.float32)
        st = RenderStats()
        check(lib().lmb200_render(self.h_, C.byref(p), _ptr(film), C.byref(st)))
        return film[..., :3], {f: getattr(st, f) for f, _ in RenderStats._fields_}
