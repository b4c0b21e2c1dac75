"""Deterministic synthetic scenes and ray batches.

Two families, both seeded:
  (S) soup  — N independent small triangles, centres uniform in a cube (the
      StubTriangleMesh_Random recipe of the reference's tests, test_accel3.cpp:191-224, scaled up)
  (M) mesh  — closed tessellated surfaces (spheres, tori) scattered over a displaced ground grid.
All coordinates stay within +-100 so that float ulp << the reference's 1e-4 box padding.
"""
import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)
EPS_ISECT = np.float32(1e-4)   # Math::EpsIsect(), math.h:1666


def _rng(seed):
    return np.random.Generator(np.random.Philox(int(seed)))


def soup(n, seed=42, extent=100.0, edge=0.2):
    """N random small triangles: centre uniform in [0,extent]^3, vertices centre + U(-edge,edge)^3."""
    g = _rng(seed)
    c = g.random((n, 1, 3), dtype=np.float32) * np.float32(extent)
    off = (g.random((n, 3, 3), dtype=np.float32) * 2 - 1) * np.float32(edge)
    return np.ascontiguousarray((c + off).reshape(n, 9), dtype=np.float32)


def _grid_tris(P):
    """P: (nu, nv, 3) vertex grid -> ((nu-1)*(nv-1)*2, 9) triangles."""
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.stack([a, b, c], axis=2).reshape(-1, 9)
    t2 = np.stack([a, c, d], axis=2).reshape(-1, 9)
    return np.concatenate([t1, t2], axis=0).astype(np.float32)


def sphere(center, radius, nu, nv):
    u = np.linspace(0, 2 * np.pi, nu + 1, dtype=np.float64)
    v = np.linspace(0, np.pi, nv + 1, dtype=np.float64)
    U, V = np.meshgrid(u, v, indexing="ij")
    P = np.stack([np.cos(U) * np.sin(V), np.cos(V), np.sin(U) * np.sin(V)], axis=-1) * radius + np.asarray(center)
    t = _grid_tris(P.astype(np.float32))
    # drop the degenerate pole triangles
    a, b, c = t[:, 0:3], t[:, 3:6], t[:, 6:9]
    area = np.linalg.norm(np.cross(b - a, c - a), axis=1)
    return t[area > 1e-12]


def torus(center, R, r, nu, nv, tilt=0.0):
    u = np.linspace(0, 2 * np.pi, nu + 1, dtype=np.float64)
    v = np.linspace(0, 2 * np.pi, nv + 1, dtype=np.float64)
    U, V = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(V)) * np.cos(U)
    y = r * np.sin(V)
    z = (R + r * np.cos(V)) * np.sin(U)
    ct, st = np.cos(tilt), np.sin(tilt)
    P = np.stack([x, ct * y - st * z, st * y + ct * z], axis=-1) + np.asarray(center)
    t = _grid_tris(P.astype(np.float32))
    return t[:, [0, 1, 2, 6, 7, 8, 3, 4, 5]]   # wind so that normals point out of the tube


def ground(half, n, seed, amp=1.0):
    g = _rng(seed)
    xs = np.linspace(-half, half, n + 1, dtype=np.float64)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    ph = g.random(4) * 6.28
    Y = amp * (np.sin(X * 0.21 + ph[0]) * np.cos(Z * 0.17 + ph[1]) + 0.5 * np.sin(X * 0.53 + ph[2]) * np.sin(Z * 0.47 + ph[3]))
    P = np.stack([X, Y, Z], axis=-1)
    t = _grid_tris(P.astype(np.float32))
    # wind so that normals point up (+y): (b-a)x(c-a) must have +y
    return t[:, [0, 1, 2, 6, 7, 8, 3, 4, 5]]


def mesh_scene(target_tris, seed=42, half=50.0, n_objects=200, return_objects=False):
    """Ground + n_objects tessellated spheres/tori with about target_tris triangles in [-half,half]^3.

    Returns (verts (N,9) float32, object_id (N,) int32) — object 0 is the ground."""
    g = _rng(seed)
    n_ground = max(2, int(np.sqrt(target_tris * 0.1 / 2)))
    parts = [ground(half, n_ground, seed + 1)]
    ids = [np.zeros(parts[0].shape[0], np.int32)]
    per_obj = max(8, (target_tris - parts[0].shape[0]) // max(1, n_objects))
    for k in range(n_objects):
        c = np.array([(g.random() * 2 - 1) * half * 0.9, 2.0 + g.random() * half * 0.35, (g.random() * 2 - 1) * half * 0.9])
        rad = (0.02 + 0.06 * g.random()) * half
        res = max(3, int(np.sqrt(per_obj / 2)))
        if g.random() < 0.5:
            t = sphere(c, rad, res, res)
        else:
            t = torus(c, rad, rad * 0.35, res, res, tilt=g.random() * 3.14)
        parts.append(t)
        ids.append(np.full(t.shape[0], k + 1, np.int32))
    verts = np.ascontiguousarray(np.concatenate(parts, axis=0), dtype=np.float32)
    oid = np.concatenate(ids)
    return verts, oid


def bounds(verts):
    v = verts.reshape(-1, 3)
    return v.min(axis=0), v.max(axis=0)


def random_rays(n, lo, hi, seed=7, tmin=EPS_ISECT, tmax=FLT_MAX):
    """Incoherent batch: origins uniform in the AABB, directions uniform on the sphere
    (Sampler::UniformSampleSphere, sampler.h:79-85). Returns (n,8) float32 in lmb200_ray layout."""
    g = _rng(seed)
    lo = np.asarray(lo, np.float32)
    hi = np.asarray(hi, np.float32)
    rays = np.empty((n, 8), np.float32)
    rays[:, 0:3] = lo + g.random((n, 3), dtype=np.float32) * (hi - lo)
    u = g.random((n, 2), dtype=np.float32)
    z = 1 - 2 * u[:, 0]
    r = np.sqrt(np.maximum(0, 1 - z * z))
    phi = np.float32(2 * np.pi) * u[:, 1]
    rays[:, 4] = r * np.cos(phi)
    rays[:, 5] = r * np.sin(phi)
    rays[:, 6] = z
    rays[:, 3] = tmin
    rays[:, 7] = tmax
    return rays


def lookat(eye, center, up):
    """Camera basis as Lightmetrica's lookat transform yields it: vz points from center to eye."""
    eye, center, up = (np.asarray(a, np.float64) for a in (eye, center, up))
    vz = eye - center
    vz /= np.linalg.norm(vz)
    vx = np.cross(up, vz)
    vx /= np.linalg.norm(vx)
    vy = np.cross(vz, vx)
    return vx.astype(np.float32), vy.astype(np.float32), vz.astype(np.float32)


def camera_rays(eye, center, up, fov_deg, w, h):
    """Primary rays through pixel centres, raycast-style (renderer_raycast.cpp:77-83 with
    sensor_pinhole.cpp:79-90): raster = ((x+.5)/w, (y+.5)/h), row 0 = bottom."""
    vx, vy, vz = lookat(eye, center, up)
    tanf = np.tan(np.radians(np.float32(fov_deg)) * np.float32(0.5)).astype(np.float32)
    aspect = np.float32(w) / np.float32(h)
    xs = (np.arange(w, dtype=np.float32) + np.float32(0.5)) / np.float32(w)
    ys = (np.arange(h, dtype=np.float32) + np.float32(0.5)) / np.float32(h)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    rx = 2 * X - 1
    ry = 2 * Y - 1
    d = np.stack([aspect * tanf * rx, tanf * ry, -np.ones_like(rx)], axis=-1).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    wd = d[..., 0:1] * vx + d[..., 1:2] * vy + d[..., 2:3] * vz
    rays = np.empty((h * w, 8), np.float32)
    rays[:, 0:3] = np.asarray(eye, np.float32)
    rays[:, 4:7] = wd.reshape(-1, 3)
    rays[:, 3] = EPS_ISECT
    rays[:, 7] = FLT_MAX
    return rays
