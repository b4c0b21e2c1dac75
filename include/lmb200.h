/*
 * lmb200 — B200-native ray traversal + unidirectional path tracing behind a thin C ABI.
 *
 * This header is the drop-in boundary below Lightmetrica v2's plugin API: the two plugins
 * (accel::lmb200, renderer::lmb200pt) are host C++ that flatten a
 * `Scene3` into the POD arrays declared here and call these entry points; nothing in the
 * signatures is a torch or C++ type. Citations are paths under the reference tree
 * (hi2p-perim/lightmetrica-v2).
 *
 * Conventions
 *   - every function returning int returns 0 on success and a negative LMB200_E_* code on
 *     failure; lmb200_last_error() gives the message (thread-local). Nothing throws.
 *   - *_dev entry points take DEVICE pointers on the accel's device and enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = the legacy default stream) without synchronising.
 *     The plain entry points take HOST pointers and do the H2D/D2H copies themselves.
 *   - there is no CPU fallback: a build without a usable CUDA device fails with LMB200_E_CUDA.
 *   - threads: lmb200_trace_closest_one may be called from any number of host threads at once (one mailbox each);
 *     the host-buffer batch calls on one accel take turns (they share its staging buffers); *_dev calls on different
 *     streams may overlap (each gets its own work counter); build / destroy must not run beside other calls on the
 *     same object. A scene renders one job at a time: concurrent render calls on one scene take turns.
 */
#ifndef LMB200_H
#define LMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMB200_OK          0
#define LMB200_E_INVALID  -1   /* bad argument */
#define LMB200_E_CUDA     -2   /* CUDA runtime error or no device */
#define LMB200_E_STATE    -3   /* object not built / wrong state */
#define LMB200_E_NCCL     -4   /* NCCL not loadable or failed */

#define LMB200_MISS 0xffffffffu

/* Ray: origin, direction (NOT renormalised: t is in units of |d|, as in ray.h:39-43) and the
 * accepted parametric range [tmin, tmax] (accel3.h:68 minT/maxT; tmax may be FLT_MAX). 32 B. */
typedef struct lmb200_ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} lmb200_ray;

/* Closest hit. tri = index of the triangle in the order it was given to lmb200_accel_build
 * (the reference accels' `triangles_` order: primitive-major, face-minor,
 * accel_qbvh.cpp:161-194), LMB200_MISS if none. t,u,v are exactly the values
 * TriAccelTriangle::Intersect produces (triaccel.h:129-150): u weights vertex 2, v vertex 3. 16 B. */
typedef struct lmb200_hit {
    float t, u, v;
    uint32_t tri;
} lmb200_hit;

typedef struct lmb200_accel lmb200_accel;

typedef struct lmb200_accel_stats {
    uint64_t num_triangles;       /* triangles given to build */
    uint64_t num_valid_triangles; /* non-degenerate (TriAccel k != 3, triaccel.h:73-77) */
    uint64_t num_nodes;           /* 64-byte wide nodes */
    uint64_t node_bytes, tri_bytes; /* 64 bytes per node, 64 per triangle unit (48-byte record + pad) */
    double   build_seconds;       /* host build */
    double   upload_seconds;
    float    sah_cost;
    int      max_depth;
} lmb200_accel_stats;

const char* lmb200_last_error(void);
int  lmb200_device_count(void);

/* Replaces Accel::Initialize (accel.h:67). device = CUDA ordinal. */
lmb200_accel* lmb200_accel_create(int device);
void lmb200_accel_destroy(lmb200_accel* a);

/* Replaces Accel::Build (accel.h:79; accel_qbvh.cpp:152-396). verts = 9 floats per triangle:
 * world-space A,B,C exactly as the reference computes them
 * (Vec3(prim->transform * Vec4(p,1)), accel_qbvh.cpp:182-184). ntris may be 0. */
int lmb200_accel_build(lmb200_accel* a, const float* verts, uint64_t ntris);
int lmb200_accel_get_stats(const lmb200_accel* a, lmb200_accel_stats* out);
/* CUDA ordinal the accel lives on (-1: host-only accel). */
int lmb200_accel_device(const lmb200_accel* a);
/* A copy of a built accel on another GPU (device-to-device copy of the flattened arrays; no rebuild). The reference has
 * one Accel per Scene (scene.h:79); a multi-GPU render needs the same BVH resident on every device. The caller owns
 * the replica and destroys it with lmb200_accel_destroy. */
lmb200_accel* lmb200_accel_replicate(const lmb200_accel* a, int device);

/* Same, with a choice of builder. All produce the same unit format and therefore the same hits (the closest hit does not
 * depend on the tree); they differ in build time and tree quality (B200, 4 M-triangle soup / 1 M-triangle mesh scene,
 * incoherent rays, profiles/r02_sweep.md):
 *   LMB200_BUILD_HOST_SAH      multi-threaded binned-SAH build on the host + SAH-optimal 8-wide collapse: 3.5 s / 0.6 s,
 *                              the quality reference (100 % / 100 %). What a host-only accel uses.
 *   LMB200_BUILD_GPU_LBVH      Morton-order radix tree + greedy collapse on the device: 25 ms / 7 ms, 101 % / 99 % of the
 *                              SAH tree's traversal rate. THE DEFAULT of lmb200_accel_build / lmb200_scene_create.
 *   LMB200_BUILD_GPU_LBVH_SAH  the same tree with the SAH-optimal collapse: 103 % / 95 %
 *   LMB200_BUILD_GPU_PLOC      parallel locally-ordered clustering + SAH-optimal collapse: 90 % / 95 % */
#define LMB200_BUILD_HOST_SAH 0
#define LMB200_BUILD_GPU_LBVH 1
#define LMB200_BUILD_GPU_PLOC 2
#define LMB200_BUILD_GPU_LBVH_SAH 3
#define LMB200_BUILD_DEFAULT LMB200_BUILD_GPU_LBVH
int lmb200_accel_build_ex(lmb200_accel* a, const float* verts, uint64_t ntris, int builder);

/* Replaces Accel3::Intersect (accel3.h:68; accel_qbvh.cpp:398-497) for a batch of n rays.
 * Closest hit with the reference's acceptance rule (reject t<tmin or t>tmax, triaccel.h:137);
 * on exact ties in t the triangle with the larger index wins (= accel::naive's scan order,
 * accel_naive.cpp:92-124). A ray with a NaN or infinite origin / direction component reports a miss (every comparison
 * of the reference's triangle test fails on it) without walking the tree.
 * Host-pointer form: the call overlaps its own uploads, traversal and downloads; from 512 Ki rays on it is ONE persistent
 * launch that takes the rays chunk by chunk as their uploads land (pinned buffers make the copies asynchronous; pageable
 * ones work). Environment, read once per process: LMB200_E2E_STREAM=0 (one launch per chunk), LMB200_E2E_CHUNK_LOG2. */
int lmb200_trace_closest(lmb200_accel* a, const lmb200_ray* rays, lmb200_hit* hits, uint64_t n);
int lmb200_trace_closest_dev(lmb200_accel* a, const void* rays_dev, void* hits_dev, uint64_t n, void* stream);

/* Compact wire form of the host-buffer calls: 24 bytes per ray (o.xyz, d.xyz as 6 floats) and ONE [tmin, tmax] for the
 * whole batch - the shape of Accel3::Intersect's own arguments (a Ray of origin + direction, minT and maxT passed
 * separately, accel3.h:68; Scene3::Intersect always passes (Eps, Inf), scene3.cpp:461). Same hits as the 32-byte form;
 * a quarter less host-to-device traffic, which is what bounds these calls once several GPUs share a host memory path. */
int lmb200_trace_closest_compact(lmb200_accel* a, const float* rays24, float tmin, float tmax, lmb200_hit* hits, uint64_t n);
int lmb200_trace_any_compact(lmb200_accel* a, const float* rays24, float tmin, float tmax, uint8_t* occluded, uint64_t n);

/* One ray, synchronously, on the GPU — the exact shape of Accel3::Intersect (accel3.h:68). Safe to call concurrently
 * from many host threads (the reference's renderers do, scheduler.cpp:146-175). No kernel launch per ray: the first call
 * starts a persistent service kernel (one block) that polls per-thread mailboxes in mapped pinned host memory, traverses
 * with the same code as the batch kernels (bit-identical hits) and writes the hit back into the mailbox the caller spins
 * on. The kernel leaves by itself after 2 ms without a request (so it never blocks a device-wide synchronisation for
 * longer) and is restarted on demand. Latency is one PCIe round trip plus one single-lane traversal; batches should
 * still use lmb200_trace_closest*. */
int lmb200_trace_closest_one(lmb200_accel* a, const lmb200_ray* ray, lmb200_hit* hit);
/* The same for n rays from `threads` host threads at once, each calling the per-ray entry point in a loop over its
 * share (how Scheduler_::Process drives Accel3::Intersect); *seconds receives the wall time. For measurements. */
int lmb200_trace_closest_one_mt(lmb200_accel* a, const lmb200_ray* rays, lmb200_hit* hits, uint64_t n, int threads, double* seconds);

/* Replaces Scene3::Visible's query (scene3.h:107-116): occluded[i] = 1 iff ANY triangle is hit
 * within [tmin, tmax] (same boolean as the reference's closest-hit query, early exit). */
int lmb200_trace_any(lmb200_accel* a, const lmb200_ray* rays, uint8_t* occluded, uint64_t n);
int lmb200_trace_any_dev(lmb200_accel* a, const void* rays_dev, void* occluded_dev, uint64_t n, void* stream);

/* Traversal work counters for the roofline's algorithmic-byte figure: mean
 * 64-byte nodes and 48-byte triangle records fetched per ray over the given DEVICE ray batch
 * (instrumented copy of the closest-hit kernel, not timed). */
int lmb200_trace_count_dev(lmb200_accel* a, const void* rays_dev, uint64_t n, double* nodes_per_ray, double* tris_per_ray);

/* Number of kernels this library has launched so far in this process (bench.py gpu_launches). */
uint64_t lmb200_launch_count(void);

/* Host view of the flattened structure, for tests (host logic is checked without a GPU): ONE array of 64-byte units,
 * unit 0 = the root node (csrc/bvh.h: Node64 | 48-byte TriAccel record + 16 bytes of padding). A node's children are
 * contiguous from Node64::base: its internal children in slot order, then the triangles of its leaf slots. Node origins
 * are 16-bit coordinates on the scene grid: origin[a] = grid_lo[a] + k[a] * grid_step[a]. */
typedef struct lmb200_bvh_layout {
    const void* units;
    uint64_t num_units, num_nodes, num_triangles;
    float grid_lo[3], grid_step[3];
} lmb200_bvh_layout;
int lmb200_accel_host_layout(const lmb200_accel* a, lmb200_bvh_layout* out);
/* Build on the host only (no device needed); such an accel cannot trace. For tests. */
lmb200_accel* lmb200_accel_create_host_only(void);

/* ---------------------------------------------------------------------------------------------
 * Wavefront path tracer (replaces Renderer::Render of renderer::pt / renderer::ptdirect,
 * renderer_pt.cpp:64-231, renderer_ptdirect.cpp:72-282, and Scheduler_::Process,
 * scheduler.cpp:78-295).
 */

#define LMB200_BSDF_NULL          0   /* bsdf_null.cpp:38-56: Type()==None, path ends */
#define LMB200_BSDF_DIFFUSE       1   /* bsdf_diffuse.cpp:69-104 */
#define LMB200_BSDF_COOKTORRANCE  2   /* bsdf_cooktorrance.cpp:73-120,187-225,276-293 (GGX) */
#define LMB200_BSDF_REFLECT_ALL   3   /* bsdf_reflectall.cpp:57-105: perfect mirror (delta) */
#define LMB200_BSDF_REFRACT_ALL   4   /* bsdf_refractall.cpp:59-140: refraction, mirror on total internal reflection (delta) */
#define LMB200_BSDF_FLESNEL       5   /* bsdf_flesnel.cpp:59-160,222-240: Fresnel-weighted choice of the two (delta) */

typedef struct lmb200_bsdf {
    int32_t type;
    float R[3];
    float eta[3];
    float k[3];
    float roughness;
    float eta1, eta2;   /* refract_all / flesnel: indices of refraction outside / inside (defaults 1, 2) */
    int32_t texR;       /* diffuse / cook_torrance "TexR" (bsdf_diffuse.cpp:48-53,102, bsdf_cooktorrance.cpp:50-55):
                           0 = use R, k > 0 = R is textures[k-1] evaluated at the hit's interpolated uv */
} lmb200_bsdf;

/* A texture as texture::bitmap holds it (texture_bitmap.cpp:140-160): width*height RGB texels, row-major; evaluated with
 * x = clamp(int(fract(u)*width), 0, width-1), y likewise, texel y*width+x (:162-168). Other Texture implementations
 * are baked into this form by the host (renderer::lmb200pt param texture_resolution). */
typedef struct lmb200_texture {
    int32_t width, height;
    const float* rgb;   /* 3 floats / texel */
} lmb200_texture;

/* One entry per scene primitive that owns triangles (primitive.h:57-84). */
typedef struct lmb200_primitive {
    int32_t  bsdf;        /* index into bsdfs */
    int32_t  light;       /* index into lights or -1 */
    uint32_t first_tri;   /* its triangles are [first_tri, first_tri+num_tris) in build order */
    uint32_t num_tris;
    int32_t  has_normals; /* 0: shading normal = geometric normal (intersectionutils.h:96-100) */
} lmb200_primitive;

/* light::area (light_area.cpp:47-115). The area CDF over the primitive's triangles is built
 * by the library exactly as TriangleUtils::CreateTriangleAreaDist (triangleutils.h:47-68). */
#define LMB200_LIGHT_AREA   0
#define LMB200_LIGHT_POINT  1   /* light_point.cpp:47-105: delta position, emits Le in every direction */
#define LMB200_LIGHT_DIRECTIONAL 2   /* light_directional.cpp:92-215: delta direction; sampled on the virtual disk of the
                                        scene's bounding sphere; never hit by extend rays (no emitter shape registered) */
#define LMB200_LIGHT_ENV    3   /* light_env.cpp:96-275 with a constant Le (its envmap branch cannot load in the reference,
                                   :107-113): direction uniform on the sphere. Direct-light sampling only: accepted by
                                   LMB200_MODE_PTDIRECT; the reference's pt / ptmis dereference a null primitive when an
                                   extend ray escapes to the env emitter shape (scene3.cpp:463-475, renderer_pt.cpp:183),
                                   so lmb200_render* rejects those modes with LMB200_ERR_INVALID */
typedef struct lmb200_light {
    float   Le[3];
    int32_t primitive;     /* area: the primitive whose mesh is sampled; others: the light's own primitive (no mesh) */
    int32_t kind;
    float   position[3];   /* point: world-space position */
    float   direction[3];  /* directional: Mat3(transform) * normalize(direction), the direction light travels (light_directional.cpp:103) */
} lmb200_light;

/* sensor::pinhole (sensor_pinhole.cpp:47-61): position = column 3 of the primitive transform,
 * vx,vy,vz = columns 0..2, fov in radians (vertical), aspect = W/H.
 * sensor::thinlens (sensor_thinlens.cpp:44-68) adds the aperture radius and the focal distance; the lens point of a
 * sample is position + lens_radius * concentric_disk(u2) in the (vx,vy) plane (:87-106). */
#define LMB200_CAMERA_PINHOLE  0
#define LMB200_CAMERA_THINLENS 1
typedef struct lmb200_camera {
    float position[3];
    float vx[3], vy[3], vz[3];
    float fov;
    int32_t width, height;
    int32_t kind;
    float lens_radius, focal_distance;   /* thinlens only */
} lmb200_camera;

typedef struct lmb200_scene_desc {
    uint64_t num_tris;
    const float* verts;        /* 9 floats / triangle, world space (as lmb200_accel_build) */
    const float* normals;      /* 9 floats / triangle: normalTransform * n_i (intersectionutils.h:88-90), or NULL */
    const uint32_t* tri_prim;  /* primitive index (into prims) per triangle */
    uint32_t num_prims;
    const lmb200_primitive* prims;
    uint32_t num_bsdfs;
    const lmb200_bsdf* bsdfs;
    uint32_t num_lights;
    const lmb200_light* lights;
    lmb200_camera camera;
    float sphere_center[3];    /* Scene3::GetSphereBound() (scene3.cpp:56-78): bounding sphere of all mesh vertices and the */
    float sphere_radius;       /* sensor position, radius grown by 1 %. Read by directional / env lights only. */
    const float* uvs;          /* 6 floats / triangle: the three vertices' texture coordinates (TriangleMesh::Texcoords, */
                               /* intersectionutils.h:107-115), or NULL when no BSDF is textured (uv = 0 for meshes without) */
    uint32_t num_textures;
    const lmb200_texture* textures;
} lmb200_scene_desc;

typedef struct lmb200_scene lmb200_scene;

#define LMB200_MODE_PT        0   /* renderer::pt: emission on BSDF-sampled hits, no NEE */
#define LMB200_MODE_PTDIRECT  1   /* renderer::ptdirect: NEE at every vertex incl. the camera vertex */
#define LMB200_MODE_PTMIS     3   /* renderer::ptmis: NEE + BSDF-sampled emission, balance heuristic (renderer_ptmis.cpp:130-275) */
#define LMB200_MODE_NORMAL    2   /* primary rays at pixel centres, |sn| as RGB (renderer_raycast.cpp:72-105 ray gen,
                                     plugin/renderer_normal/renderer_normal.cpp:62-75 shading) */

typedef struct lmb200_render_params {
    int32_t  mode;
    int64_t  num_samples;        /* scheduler.cpp:55 num_samples: total over ALL ranks */
    int64_t  sample_begin;       /* this call renders global sample indices [sample_begin, sample_end) */
    int64_t  sample_end;
    int32_t  max_num_vertices;   /* renderer_pt.cpp:59, -1 = unbounded */
    int32_t  min_num_vertices;   /* renderer_pt.cpp:60 */
    uint64_t seed;               /* counter-based RNG key; same seed => same image at any GPU count */
    int32_t  pool_size;          /* wavefront size in paths, 0 = default */
    /* Optional tile partitioning (off = all zero). tile = raster sub-rectangle {x0, y0, x1, y1} in [0,1]^2: the camera
     * samples of THIS call are drawn uniformly inside it (raster sample u -> x0 + u.x (x1-x0), y0 + u.y (y1-y0)) instead
     * of over the whole image. Rendering tiles of area a_k with a_k * num_samples samples each gives the same expected
     * image as whole-image sampling (stratified over tiles: a different sampling pattern, statistically identical);
     * splats of camera-vertex light sampling still land anywhere, so films are summed as usual. */
    float    tile[4];
    int32_t  primary_tile;       /* coherent camera samples: the 32 samples of a group (sample index / 32) share one tile of
                                    primary_tile x primary_tile pixels and are uniform inside it; consecutive groups visit all tiles
                                    in a keyed pseudo-random order before any tile repeats. Every sample's raster position stays
                                    marginally uniform over the image (the estimator of renderer_pt.cpp:84 in expectation, for any
                                    sample count and sharding; complete rounds are stratified over the tiles), while the primary
                                    rays a warp traces together are coherent. 0 = automatic (lmb200_default_primary_tile), < 0 = off (independent) */
    int32_t  count_work;         /* 1 = run the instrumented traversal kernels and fill lmb200_render_stats::extend_nodes ...
                                    shadow_tris (for the roofline's algorithmic bytes per sample); such a run is never timed */
    int32_t  tile_partition;     /* lmb200_render_multi / _timed only: 1 = GPU g of n draws its raster positions in the
                                    horizontal strip [g/n, (g+1)/n) of the image (default 0: every GPU samples the whole image) */
} lmb200_render_params;

/* The tile edge primary_tile = 0 resolves to for a job of num_samples samples on a width x height film: the smallest of
 * 2, 4, ... 64 pixels whose rounds fit at least four times into the job; -1 (independent samples) for tiny jobs. */
int lmb200_default_primary_tile(int width, int height, int64_t num_samples);

typedef struct lmb200_render_stats {
    int64_t  samples;
    int64_t  extend_rays, shadow_rays;
    int64_t  iterations;
    uint64_t launches;
    double   seconds;            /* device time of the wavefront loop (CUDA events) */
    double   reduce_seconds;     /* lmb200_render_multi / _timed: device time of the NCCL film reductions; else 0 */
    int64_t  vertices;           /* path vertices processed (camera vertices included): one k_logic/k_nee/k_bsdf pass each */
    int64_t  extend_nodes, extend_tris;   /* count_work only: 64-byte nodes / 48-byte triangle records fetched by the extend rays */
    int64_t  shadow_nodes, shadow_tris;   /* ... and by the shadow rays */
} lmb200_render_stats;

/* Builds the accel (device) and uploads shading data. */
lmb200_scene* lmb200_scene_create(int device, const lmb200_scene_desc* desc);
/* Same with a choice of BVH builder (LMB200_BUILD_HOST_SAH / LMB200_BUILD_GPU_LBVH). */
lmb200_scene* lmb200_scene_create_ex(int device, const lmb200_scene_desc* desc, int builder);
/* Same, but traversal uses an accel that is already built on a device over the SAME triangle list (same
 * order); the scene borrows it and never frees it. Lets renderer::lmb200pt reuse the BVH of accel::lmb200
 * when the YAML selected both. */
lmb200_scene* lmb200_scene_create_shared(const lmb200_scene_desc* desc, lmb200_accel* accel);
void lmb200_scene_destroy(lmb200_scene* s);

/* Process-wide lookup table (owner address -> accel) through which the two plugins, loaded RTLD_LOCAL by the
 * host (component.cpp:127-164), find each other's objects. put(owner, NULL) removes the entry. */
void lmb200_registry_put(const void* owner, lmb200_accel* accel);
lmb200_accel* lmb200_registry_get(const void* owner);
lmb200_accel* lmb200_scene_accel(lmb200_scene* s);

/* Accumulates UNSCALED splats of the given sample range into film_dev: W*H float4 (rgb + pad,
 * the layout of film_hdr.cpp's Vec3 data_), row 0 = raster y in [0,1/H) (film_hdr.cpp:218-223).
 * The caller zeroes the film, sums films across GPUs (NCCL reduce) and applies
 * lmb200_film_rescale with W*H/num_samples (scheduler.cpp:280-288). Synchronises `stream` once at the end. */
int lmb200_render_dev(lmb200_scene* s, const lmb200_render_params* p, void* film_dev, void* stream, lmb200_render_stats* stats);
int lmb200_film_rescale_dev(void* film_dev, int64_t num_pixels, float scale, void* stream);

/* Host-buffer convenience: zero film, render [sample_begin,sample_end), rescale by
 * W*H/num_samples, copy W*H*4 floats back. The call the Renderer plugin makes on one GPU. */
int lmb200_render(lmb200_scene* s, const lmb200_render_params* p, float* film_rgba_host, lmb200_render_stats* stats);

/* Single-process multi-GPU render (one host thread + stream per device; per-GPU films summed to
 * device 0 with ncclReduce, then rescaled). scenes[g] must hold the same scene on device g. */
int lmb200_render_multi(lmb200_scene** scenes, int num_gpus, const lmb200_render_params* p, float* film_rgba_host, lmb200_render_stats* stats);

/* Time-budgeted / progressive rendering: Scheduler_'s `render_time` and
 * `progress_image_update_interval` (scheduler.cpp:54-57,108,191-255). Runs passes of pass_samples
 * samples (the reference uses grain_size*1000; <=0: 10^7) until render_time seconds have elapsed
 * (render_time <= 0: until [sample_begin,sample_end) is done). Every progress_interval seconds (> 0)
 * `progress` receives the image so far (W*H*4 floats, rescaled by W*H/processed), the number of
 * samples processed and a running tick count; a non-zero return stops the render. The final image
 * is rescaled by W*H/processed (scheduler.cpp:288) and stats->samples = samples processed. */
typedef int (*lmb200_progress_fn)(void* user, const float* film_rgba, int64_t samples_done, int64_t tick);
int lmb200_render_timed(lmb200_scene** scenes, int num_gpus, const lmb200_render_params* p, double render_time,
                        int64_t pass_samples, double progress_interval, lmb200_progress_fn progress, void* user,
                        float* film_rgba_host, lmb200_render_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* LMB200_H */
