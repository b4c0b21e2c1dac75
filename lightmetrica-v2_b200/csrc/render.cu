// Wavefront path tracer on sm_100a: the unidirectional sample loop of the reference
// (/root/reference/src/liblightmetrica/renderer/renderer_pt.cpp:68-231 and
// renderer_ptdirect.cpp:76-282, driven by Scheduler_::Process, scheduler.cpp:78-295)
// re-organised as a pool of path slots advanced by separate kernels per iteration:
//
//   k_logic   hit processing: miss / emission splat (pt) / Russian roulette / advance vertex,
//             regeneration of finished slots from the global sample counter (camera-ray generation),
//             compaction of live slots into the vertex queue (warp ballot + prefix sum)
//   k_nee     next-event estimation against area lights -> compacted shadow-ray queue
//   k_bsdf    BSDF sample + pdf + evaluate, throughput update -> compacted extend-ray queue
//   extend    closest-hit traversal over the extend queue   (accel.cu trace kernel)
//   k_shadow  any-hit traversal over the shadow queue fused with the film splat
//
// Random numbers are Philox4x32-10 keyed by (seed) with counter (sample index, block), so any
// partition of the sample range over GPUs gives the same image up to fp32 summation order.
#include "internal.h"
#include "traverse.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include <nccl.h>      // types and enum values only: libnccl is dlopen'ed at run time, the library has no link-time dependency on it

extern "C" int lmb200_default_primary_tile(int width, int height, int64_t num_samples);

namespace lmb200 {

#define LMB_PI 3.14159265358979323846f
#define LMB_INV_PI 0.31830988618379067154f
#define LMB_EPS 1e-4f
#define LMB_EPS_ISECT 1e-4f
#define LMB_FLT_MAX 3.402823466e+38f

struct DevScene {
    const float* verts;        // 9 floats / tri
    const float* normals;      // 9 floats / tri or nullptr
    const uint32_t* tri_prim;
    const lmb200_primitive* prims;
    const lmb200_bsdf* bsdfs;
    const lmb200_light* lights;
    const float* light_cdf;    // concatenated per-light CDFs
    const uint32_t* light_cdf_off;
    const float* light_inv_area;
    uint32_t num_lights;
    // camera
    float pos[3], vx[3], vy[3], vz[3];
    float tan_fov, aspect;
    int width, height;
    int cam_kind;              // LMB200_CAMERA_*
    float lens_radius, focal_distance;
    // Scene3::GetSphereBound (directional / env lights)
    float sph_c[3], sph_r;
    int has_env;
    // TexR textures (texture_bitmap.cpp layout): all texels concatenated, tex_info[k] = (width, height, first texel, 0)
    const float* uvs;          // 6 floats / tri or nullptr
    const float* tex_rgb;
    const int4* tex_info;
};

struct Pool {
    // per-slot path state
    unsigned long long* sample;   // global sample index
    int* nverts;                  // numVertices; 0 = idle slot
    float4* thr;                  // throughput rgb, w = raster pixel (int bits, -1 = unset)
    float4* ray_o;                // current extend ray of the slot: o.xyz,tmin
    float4* ray_d;                // d.xyz,tmax
    float4* hit;                  // result of the extend ray: t,u,v,tri
    uint8_t* traced;              // 1 if the slot has a pending hit record
    // current vertex (valid between k_logic and k_bsdf)
    float4* vtx_p;                // p.xyz, w = tri (uint bits) ; camera vertex: tri = 0xffffffff
    float4* vtx_wi;               // wi.xyz, w = u
    float*  vtx_v;                // barycentric v
    float4* prev;                 // ptmis: shading normal of the vertex the extend ray left from, w = pdf of the sampled direction
    // queues
    uint32_t* vq;                 // vertex queue (slot indices)
    uint32_t* eq;                 // extend queue (slot indices)
    float4* sq_o; float4* sq_d;   // shadow rays (compact)
    float4* sq_c;                 // contribution rgb, w = pixel (int bits)
    // counters: [0] vq size, [1] eq size, [2] sq size, [3] unused
    uint32_t* qcount;
    unsigned long long* next_sample;   // [0] next sample index to hand out, [1] extend rays, [2] shadow rays, [3] path vertices,
                                       // [4],[5] nodes / triangle records fetched by extend rays, [6],[7] by shadow rays (count_work runs only)
};

struct RenderCfg {
    int mode, max_verts, min_verts;
    unsigned long long seed;
    unsigned long long sample_end;
    uint32_t pool;
    float tile_x0, tile_y0, tile_sx, tile_sy;   // raster sample u -> (x0 + u.x sx, y0 + u.y sy); whole image = (0, 0, 1, 1)
    int gt_nx, gt_ny;                           // coherent camera samples: tiles per axis (0 = off), see camera_raster
};

// ------------------------------------------------------------------------------------------------
// small vector helpers

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 F3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 ld3(const float* p) { return F3(p[0], p[1], p[2]); }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 neg(f3 a) { return F3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// v * (1 / |v|): one IEEE square root, one IEEE division, three multiplications — the same sequence as the oracle's vnorm, so
// both sides produce identical bits (the reference itself normalises with the 12-bit SSE rsqrt, math.h:1872-1875)
__device__ __forceinline__ f3 normalize(f3 a) { const float inv = 1.0f / sqrtf(dot(a, a)); return F3(a.x * inv, a.y * inv, a.z * inv); }
__device__ __forceinline__ bool black(f3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

// Radiometric scalars (BSDF values, pdfs, geometry terms, MIS weights, throughput) use the hardware reciprocal / square root
// (2 ulp) instead of the IEEE sequences: they scale pixel values by 1 +- 2^-22 and never touch ray geometry. Everything that
// decides WHERE a path goes — camera rays, surface frames, sampled directions, light points, shadow rays, raster positions,
// the Fresnel coin of bsdf::flesnel — stays on the correctly rounded operations, so the device takes the same paths as the
// oracle, sample for sample (tests/test_gpu_render.py::test_same_samples_as_oracle).
__device__ __forceinline__ float qdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float qsqrt(float a) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ f3 qnormalize(f3 a) { float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(dot(a, a))); return a * r; }

// Philox4x32-10, counter (sample_lo, sample_hi, block, 0), key (seed_lo, seed_hi) -> 4 uniforms in [0,1)
__device__ __forceinline__ float4 rng_block(unsigned long long seed, unsigned long long sample, uint32_t block)
{
    uint32_t c0 = (uint32_t)sample, c1 = (uint32_t)(sample >> 32), c2 = block, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float s = 1.0f / 16777216.0f;
    return make_float4((float)(c0 >> 8) * s, (float)(c1 >> 8) * s, (float)(c2 >> 8) * s, (float)(c3 >> 8) * s);
}

// ------------------------------------------------------------------------------------------------
// surface geometry: the subset of IntersectionUtils::CreateTriangleIntersection
// (intersectionutils.h:59-135) the estimators read

struct Geom { f3 p, gn, sn, dpdu, dpdv; bool degenerated; };

__device__ __forceinline__ void basis(f3 a, f3& b, f3& c)   // Math::OrthonormalBasis, math.h:2355-2360
{
    c = fabsf(a.x) > fabsf(a.y) ? normalize(F3(a.z, 0.f, -a.x)) : normalize(F3(0.f, a.z, -a.y));
    b = cross(c, a);
}
__device__ __forceinline__ f3 to_local(const Geom& g, f3 w) { return F3(dot(g.dpdu, w), dot(g.dpdv, w), dot(g.sn, w)); }
__device__ __forceinline__ f3 to_world(const Geom& g, f3 l) { return (g.dpdu * l.x + g.dpdv * l.y) + g.sn * l.z; }

__device__ __forceinline__ void tri_geom(const DevScene& S, uint32_t tri, float b0, float b1, f3 p, Geom& g)
{
    const float* v = S.verts + 9 * (size_t)tri;
    const f3 p1 = ld3(v), p2 = ld3(v + 3), p3 = ld3(v + 6);
    g.p = p;
    g.degenerated = false;
    g.gn = normalize(cross(p2 - p1, p3 - p1));
    const lmb200_primitive& P = S.prims[S.tri_prim[tri]];
    if (P.has_normals && S.normals) {
        const float* n = S.normals + 9 * (size_t)tri;
        g.sn = normalize((ld3(n) * (1.0f - b0 - b1) + ld3(n + 3) * b0) + ld3(n + 6) * b1);
        if (isnan(g.sn.x) || isnan(g.sn.y) || isnan(g.sn.z)) g.sn = g.gn;
    } else g.sn = g.gn;
    basis(g.sn, g.dpdu, g.dpdv);
}

// ---- sensor::pinhole (sensor_pinhole.cpp:79-90, 137-154, 165-185) and sensor::thinlens
//      (sensor_thinlens.cpp:87-106, 150-176, 186-222); p = the sensor vertex (pinhole position / lens point) ----
__device__ __forceinline__ bool sensor_eye(const DevScene& S, f3 p, f3 wo, f3& e)
{
    if (S.cam_kind == LMB200_CAMERA_THINLENS) {
        const f3 nvz = neg(ld3(S.vz));
        const float c = dot(nvz, wo);
        if (c <= 0.f) return false;
        const float tf = S.focal_distance / c;
        const f3 Pf = p + wo * tf;                                  // intersection with the focal plane
        const f3 wo0 = normalize(Pf - ld3(S.pos));                  // direction before refraction
        e = F3(dot(ld3(S.vx), wo0), dot(ld3(S.vy), wo0), dot(ld3(S.vz), wo0));
    } else {
        e = F3(dot(ld3(S.vx), wo), dot(ld3(S.vy), wo), dot(ld3(S.vz), wo));
    }
    return e.z < 0.f;
}
__device__ __forceinline__ bool raster_from_eye(const DevScene& S, f3 e, float& rx, float& ry)
{
    rx = (-e.x / e.z / S.tan_fov / S.aspect + 1.0f) * 0.5f;
    ry = (-e.y / e.z / S.tan_fov + 1.0f) * 0.5f;
    return !(rx < 0.f || rx > 1.f || ry < 0.f || ry > 1.f);
}
__device__ __forceinline__ bool raster_position(const DevScene& S, f3 p, f3 wo, float& rx, float& ry)
{
    f3 e;
    if (!sensor_eye(S, p, wo, e)) return false;
    return raster_from_eye(S, e, rx, ry);
}
// Importance(wo) together with the raster position it is defined over (0 and inside = false outside the frustum): callers need
// both, and the raster position costs five IEEE divisions
__device__ __forceinline__ float importance_raster(const DevScene& S, f3 p, f3 wo, float& rx, float& ry, bool& inside)
{
    f3 e;
    inside = sensor_eye(S, p, wo, e) && raster_from_eye(S, e, rx, ry);
    if (!inside) return 0.f;
    const float cosT = -e.z, inv = qdiv(1.0f, cosT);
    const float A = S.tan_fov * S.tan_fov * S.aspect * 4.0f;
    return qdiv(inv * inv * inv, A);
}
__device__ __forceinline__ f3 camera_dir(const DevScene& S, float u0, float u1)
{
    const float x = 2.0f * u0 - 1.0f, y = 2.0f * u1 - 1.0f;
    const f3 e = normalize(F3(S.aspect * S.tan_fov * x, S.tan_fov * y, -1.0f));
    return (ld3(S.vx) * e.x + ld3(S.vy) * e.y) + ld3(S.vz) * e.z;
}
__device__ __forceinline__ void concentric_disk(float u0, float u1, float& sx, float& sy);
// Sensor::SamplePositionAndDirection split in two: the sensor vertex for lens sample (l0,l1) ...
__device__ __forceinline__ f3 camera_point(const DevScene& S, float l0, float l1)
{
    if (S.cam_kind != LMB200_CAMERA_THINLENS) return ld3(S.pos);
    float lx, ly;
    concentric_disk(l0, l1, lx, ly);
    lx *= S.lens_radius; ly *= S.lens_radius;
    return (ld3(S.pos) + ld3(S.vx) * lx) + ld3(S.vy) * ly;
}
// ... and the direction through raster sample (u0,u1) leaving from the sensor vertex p
__device__ __forceinline__ f3 camera_wo(const DevScene& S, float u0, float u1, f3 p)
{
    const f3 dir = camera_dir(S, u0, u1);
    if (S.cam_kind != LMB200_CAMERA_THINLENS) return dir;
    const float tf = S.focal_distance / dot(neg(ld3(S.vz)), dir);
    const f3 Pf = ld3(S.pos) + dir * tf;
    return normalize(Pf - p);
}
#define LMB_LENS_BLOCK 0xffffffffu   // Philox block of the lens sample (the second Next2D of renderer_pt.cpp:86)
// Raster position of camera sample `sidx` from its two uniforms (block 0). The reference draws it uniformly over the whole
// image, independently per sample (renderer_pt.cpp:84) - 32 neighbouring lanes then trace 32 unrelated primary rays. Here
// the 32 samples of a group (sidx / 32) share one tile of the image and are uniform inside it, and consecutive groups walk
// through ALL tiles in a pseudo-random order before any tile is visited again (round r = group / #tiles uses its own
// permutation of the tiles, a keyed 4-round Feistel network with cycle walking: integer arithmetic only, so the oracle
// port computes the same tiles). Every sample's raster position is still marginally uniform over the image, for any sample
// count and any split of the sample range (same expected image as the reference's estimator); complete rounds are
// stratified over the tiles (less noise than independent positions, never more); and the primary rays a warp generates and
// traces together are coherent.
__host__ __device__ __forceinline__ uint32_t lmb_hash32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t lmb_permute_tiles(uint32_t x, uint32_t n, uint32_t key)
{
    uint32_t hb = 1;
    while ((1u << (2u * hb)) < n) hb++;
    const uint32_t mask = (1u << hb) - 1u;
    do {
        uint32_t L = x >> hb, R = x & mask;
        for (uint32_t r = 0; r < 4u; r++) {
            const uint32_t F = lmb_hash32(R ^ key ^ (r * 0x9e3779b9u)) & mask;
            const uint32_t t = L ^ F; L = R; R = t;
        }
        x = (L << hb) | R;
    } while (x >= n);
    return x;
}
__device__ __forceinline__ void camera_raster(const RenderCfg& cfg, unsigned long long sidx, float ux, float uy, float& X, float& Y)
{
    float rx = ux, ry = uy;
    if (cfg.gt_nx > 0) {
        const uint32_t T = (uint32_t)cfg.gt_nx * (uint32_t)cfg.gt_ny;
        const unsigned long long g = sidx >> 5, round = g / T;
        const uint32_t key = lmb_hash32((uint32_t)cfg.seed ^ lmb_hash32((uint32_t)(cfg.seed >> 32) ^ lmb_hash32((uint32_t)round ^ lmb_hash32((uint32_t)(round >> 32)))));
        const uint32_t t = lmb_permute_tiles((uint32_t)(g - round * T), T, key);
        const uint32_t tx = t % (uint32_t)cfg.gt_nx, ty = t / (uint32_t)cfg.gt_nx;
        rx = ((float)tx + ux) / (float)cfg.gt_nx;
        ry = ((float)ty + uy) / (float)cfg.gt_ny;
    }
    X = cfg.tile_x0 + rx * cfg.tile_sx;
    Y = cfg.tile_y0 + ry * cfg.tile_sy;
}

__device__ __forceinline__ int pixel_index(const DevScene& S, float rx, float ry)   // film_hdr.cpp:218-223
{
    int px = (int)(rx * (float)S.width), py = (int)(ry * (float)S.height);
    px = min(max(px, 0), S.width - 1);
    py = min(max(py, 0), S.height - 1);
    return py * S.width + px;
}

// ---- BSDFs (bsdf_diffuse.cpp:69-104, bsdf_cooktorrance.cpp:73-120,187-225,276-293, bsdfutils.h:55-66) ----
__device__ __forceinline__ float snc(const Geom& g, f3 wi, f3 wo)
{
    const float wiNg = dot(wi, g.gn), woNg = dot(wo, g.gn);
    const float wiNs = to_local(g, wi).z, woNs = to_local(g, wo).z;
    return (wiNg * wiNs <= 0.f || woNg * woNs <= 0.f) ? 0.f : 1.f;
}
__device__ __forceinline__ void concentric_disk(float u0, float u1, float& sx, float& sy)   // sampler.h:44-60
{
    const float vx = 2.0f * u0 - 1.0f, vy = 2.0f * u1 - 1.0f;
    if (vx == 0.f && vy == 0.f) { sx = sy = 0.f; return; }
    float r, theta;
    if (vx > -vy) {
        if (vx > vy) { r = vx; theta = (LMB_PI * 0.25f) * vy / vx; }
        else { r = vy; theta = (LMB_PI * 0.25f) * (2.0f - vx / vy); }
    } else {
        if (vx < vy) { r = -vx; theta = (LMB_PI * 0.25f) * (4.0f + vy / vx); }
        else { r = -vy; theta = (LMB_PI * 0.25f) * (6.0f - vx / vy); }
    }
    float sn_, cs_;
    sincosf(theta, &sn_, &cs_);      // same values as sinf / cosf, one argument reduction
    sx = r * cs_; sy = r * sn_;
}
__device__ __forceinline__ float ggx_D(float alpha, f3 H)
{
    const float cosH = H.z;
    if (cosH <= 0.f) return 0.f;
    const float s2 = 1.0f - cosH * cosH;
    const float tanH = s2 <= 0.f ? 0.f : qdiv(qsqrt(s2), cosH);
    const float t1 = alpha * alpha;
    const float t = alpha * alpha + tanH * tanH;
    return qdiv(t1, LMB_PI * cosH * cosH * cosH * cosH * t * t);
}
// Fresnel term of bsdf::flesnel (bsdf_flesnel.cpp:224-240). EXACT = correctly rounded (the reflect / refract coin of the sampler),
// otherwise radiometric (pdf and value).
template <bool EXACT>
__device__ __forceinline__ float fresnel_term(f3 lwi, float etaI, float etaT)
{
    const float wiDotN = lwi.z, eta = etaI / etaT;
    const float c2 = 1.0f - eta * eta * (1.0f - wiDotN * wiDotN);
    if (c2 <= 0.f) return 1.0f;
    const float ci = fabsf(wiDotN), ct = EXACT ? sqrtf(c2) : qsqrt(c2);
    const float rhoS = EXACT ? (etaI * ci - etaT * ct) / (etaI * ci + etaT * ct) : qdiv(etaI * ci - etaT * ct, etaI * ci + etaT * ct);
    const float rhoT = EXACT ? (etaI * ct - etaT * ci) / (etaI * ct + etaT * ci) : qdiv(etaI * ct - etaT * ci, etaI * ct + etaT * ci);
    return (rhoS * rhoS + rhoT * rhoT) * 0.5f;
}
__device__ __forceinline__ bool is_specular(const lmb200_bsdf& B) { return B.type >= LMB200_BSDF_REFLECT_ALL && B.type <= LMB200_BSDF_FLESNEL; }

__device__ __forceinline__ bool bsdf_sample(const lmb200_bsdf& B, const Geom& g, f3 wi, float u0, float u1, float ucomp, f3& wo)
{
    const f3 lwi = to_local(g, wi);
    if (B.type == LMB200_BSDF_REFRACT_ALL || B.type == LMB200_BSDF_FLESNEL) {
        // bsdf_refractall.cpp:59-90, bsdf_flesnel.cpp:59-95: both sides of the surface
        float etaI = B.eta1, etaT = B.eta2;
        if (lwi.z < 0.f) { const float t = etaI; etaI = etaT; etaT = t; }
        const float eta = etaI / etaT;
        const float c2 = 1.0f - eta * eta * (1.0f - lwi.z * lwi.z);
        const bool reflect = B.type == LMB200_BSDF_REFRACT_ALL ? (c2 <= 0.f) : (ucomp <= fresnel_term<true>(lwi, etaI, etaT));
        if (reflect) wo = to_world(g, F3(-lwi.x, -lwi.y, lwi.z));                // BSDFUtils::LocalReflect
        else {
            const float ct = sqrtf(c2) * (lwi.z > 0.f ? -1.0f : 1.0f);
            wo = to_world(g, F3(-eta * lwi.x, -eta * lwi.y, ct));                 // BSDFUtils::LocalRefract
        }
        return true;
    }
    if (lwi.z <= 0.f) return false;
    if (B.type == LMB200_BSDF_REFLECT_ALL) { wo = to_world(g, F3(-lwi.x, -lwi.y, lwi.z)); return true; }   // bsdf_reflectall.cpp:57-68
    if (B.type == LMB200_BSDF_DIFFUSE) {
        float sx, sy;
        concentric_disk(u0, u1, sx, sy);
        wo = to_world(g, F3(sx, sy, sqrtf(fmaxf(0.f, 1.0f - sx * sx - sy * sy))));
        return true;
    }
    if (B.type == LMB200_BSDF_COOKTORRANCE) {
        const float a = B.roughness;
        const float v0 = (1.0f - LMB_EPS) * u0 + LMB_EPS;
        const float v1 = (1.0f - 2.0f * LMB_EPS) * u1 + LMB_EPS;
        const float den = sqrtf(1.0f - (1.0f - a * a) * v0);
        const float cosT = sqrtf(1.0f - v0) / den, sinT = a * (sqrtf(v0) / den);
        const float phi = LMB_PI * (2.0f * v1 - 1.0f);
        float sp, cp;
        sincosf(phi, &sp, &cp);
        const f3 H = F3(sinT * cp, sinT * sp, cosT);
        const f3 nwi = neg(lwi);
        const f3 lwo = nwi - H * (2.0f * dot(nwi, H));
        if (lwo.z <= 0.f) return false;
        wo = to_world(g, lwo);
        return true;
    }
    return false;
}
__device__ __forceinline__ float bsdf_pdf(const lmb200_bsdf& B, const Geom& g, f3 wi, f3 wo, bool eval_delta)
{
    const f3 lwi = to_local(g, wi), lwo = to_local(g, wo);
    if (is_specular(B)) {
        if (eval_delta) return 0.f;
        if (B.type == LMB200_BSDF_REFLECT_ALL) return (lwi.z <= 0.f || lwo.z <= 0.f) ? 0.f : 1.f;   // bsdf_reflectall.cpp:70-85
        if (B.type == LMB200_BSDF_REFRACT_ALL) return 1.f;                                            // bsdf_refractall.cpp:92-100
        float etaI = B.eta1, etaT = B.eta2;                                                           // bsdf_flesnel.cpp:97-128
        if (lwi.z < 0.f) { const float t = etaI; etaI = etaT; etaT = t; }
        const float Fr = fresnel_term<false>(lwi, etaI, etaT);
        return lwi.z * lwo.z >= 0.f ? Fr : 1.0f - Fr;
    }
    if (lwi.z <= 0.f || lwo.z <= 0.f) return 0.f;
    if (B.type == LMB200_BSDF_DIFFUSE) return LMB_INV_PI;
    if (B.type == LMB200_BSDF_COOKTORRANCE) {
        const f3 H = qnormalize(lwi + lwo);
        const float D = ggx_D(B.roughness, H);
        return qdiv(qdiv(D * H.z, 4.0f * dot(lwo, H)), lwo.z);
    }
    return 0.f;
}
// R of bsdf::diffuse / bsdf::cook_torrance: the constant, or TexR evaluated at the interpolated texture coordinates
// (intersectionutils.h:107-115; Texture_Bitmap::Evaluate, texture_bitmap.cpp:162-168; bsdf_diffuse.cpp:102)
__device__ __forceinline__ f3 bsdf_R(const DevScene& S, const lmb200_bsdf& B, uint32_t tri, float b0, float b1)
{
    if (B.texR <= 0) return ld3(B.R);
    float u = 0.f, v = 0.f;
    if (S.uvs) {
        const float* t = S.uvs + 6 * (size_t)tri;
        u = t[0] * (1.0f - b0 - b1) + t[2] * b0 + t[4] * b1;
        v = t[1] * (1.0f - b0 - b1) + t[3] * b0 + t[5] * b1;
    }
    const int4 ti = S.tex_info[B.texR - 1];
    const int x = min(max((int)((u - floorf(u)) * (float)ti.x), 0), ti.x - 1);
    const int y = min(max((int)((v - floorf(v)) * (float)ti.y), 0), ti.y - 1);
    return ld3(S.tex_rgb + 3 * ((size_t)ti.z + (size_t)ti.x * (size_t)y + (size_t)x));
}
// Rr = bsdf_R(...) of the vertex (only diffuse / cook_torrance read it)
__device__ __forceinline__ f3 bsdf_eval(const lmb200_bsdf& B, f3 Rr, const Geom& g, f3 wi, f3 wo, bool eval_delta)
{
    const f3 lwi = to_local(g, wi), lwo = to_local(g, wo);
    if (is_specular(B)) {
        if (eval_delta) return F3(0, 0, 0);
        if (B.type == LMB200_BSDF_REFLECT_ALL) {                                                      // bsdf_reflectall.cpp:87-103
            if (lwi.z <= 0.f || lwo.z <= 0.f) return F3(0, 0, 0);
            return ld3(B.R) * snc(g, wi, wo);
        }
        float etaI = B.eta1, etaT = B.eta2;
        if (lwi.z < 0.f) { const float t = etaI; etaI = etaT; etaT = t; }
        const float eta = etaI / etaT;
        const float Fr = B.type == LMB200_BSDF_FLESNEL ? fresnel_term<false>(lwi, etaI, etaT) : 0.f;
        if (lwi.z * lwo.z >= 0.f)        // reflection (total internal reflection for refract_all)
            return ld3(B.R) * ((B.type == LMB200_BSDF_FLESNEL ? Fr : 1.0f) * snc(g, wi, wo));
        // refraction, EL transport: eta^2 (bsdf_refractall.cpp:120-127, bsdf_flesnel.cpp:150-156)
        return ld3(B.R) * ((B.type == LMB200_BSDF_FLESNEL ? 1.0f - Fr : 1.0f) * snc(g, wi, wo) * eta * eta);
    }
    if (lwi.z <= 0.f || lwo.z <= 0.f) return F3(0, 0, 0);
    if (B.type == LMB200_BSDF_DIFFUSE) return (Rr * LMB_INV_PI) * snc(g, wi, wo);
    if (B.type == LMB200_BSDF_COOKTORRANCE) {
        const f3 H = qnormalize(lwi + lwo);
        const float D = ggx_D(B.roughness, H);
        const float woH = fabsf(dot(lwo, H));
        // sic: the reference uses wo.H for both masking terms (bsdf_cooktorrance.cpp:281-283)
        const float G = fminf(1.0f, fminf(qdiv(2.0f * H.z * lwo.z, woH), qdiv(2.0f * H.z * lwi.z, woH)));
        const float c = dot(lwi, H);
        float F[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float eta = B.eta[i], k = B.k[i];
            const float tmp = (eta * eta + k * k) * (c * c);
            const float rP = qdiv(tmp - eta * (2.0f * c) + 1.0f, tmp + eta * (2.0f * c) + 1.0f);
            const float tmpF = eta * eta + k * k;
            const float rS = qdiv(tmpF - eta * (2.0f * c) + c * c, tmpF + eta * (2.0f * c) + c * c);
            F[i] = (rP + rS) * 0.5f;
        }
        const float s = qdiv(qdiv(D * G, 4.0f * lwi.z), lwo.z) * snc(g, wi, wo);
        return F3(Rr.x * F[0] * s, Rr.y * F[1] * s, Rr.z * F[2] * s);
    }
    return F3(0, 0, 0);
}

// ---- emitter shape of light::directional / light::env (light_directional.cpp:52-73, light_env.cpp:56-77;
//      SphereBound::Intersect with minT=0, maxT=Inf, bound.h:125-167) ----
__device__ __forceinline__ bool emitter_shape_hit(const DevScene& S, f3 o, f3 d, Geom& g)
{
    const f3 center = ld3(S.sph_c);
    const f3 oo = o - center;
    const float a = dot(d, d), b = 2.0f * dot(oo, d), c = dot(oo, oo) - S.sph_r * S.sph_r;
    const float det = b * b - 4.0f * a * c;
    if (det < 0.f) return false;
    const float e = sqrtf(det), denom = 2.0f * a;
    const float t0 = (-b - e) / denom, t1 = (-b + e) / denom;
    if (t0 > LMB_FLT_MAX || t1 < 0.f) return false;
    float t = t0;
    if (t < 0.f) { t = t1; if (t > LMB_FLT_MAX) return false; }
    g.degenerated = false;
    g.gn = g.sn = neg(d);
    basis(g.sn, g.dpdu, g.dpdv);
    const f3 p = o + d * t;
    const f3 cc = center + d * S.sph_r;
    g.p = (cc + g.dpdu * dot(g.dpdu, p - cc)) + g.dpdv * dot(g.dpdv, p - cc);
    return true;
}

// ---- Light::SamplePositionGivenPreviousPosition + its area pdf (evalDelta=false) from the vertex at `from` ----
// light::area: triangleutils.h:71-122, dist.h:70-76, sampler.h:97-101, light_area.cpp:100-103
__device__ __forceinline__ bool light_sample(const DevScene& S, int li, f3 from, float u0, float u1, Geom& g, float& pdfPL)
{
    const int kind = S.lights[li].kind;
    if (kind == LMB200_LIGHT_POINT) {            // light_point.cpp:62-66, 88-91
        g.p = ld3(S.lights[li].position);
        g.degenerated = true;
        g.gn = g.sn = g.dpdu = g.dpdv = F3(0, 0, 0);
        pdfPL = 1.0f;
        return true;
    }
    if (kind == LMB200_LIGHT_DIRECTIONAL || kind == LMB200_LIGHT_ENV) {
        f3 d;
        float pdfSA = 1.0f;                      // light_directional.cpp:176-180
        if (kind == LMB200_LIGHT_DIRECTIONAL) d = neg(ld3(S.lights[li].direction));   // light_directional.cpp:131-145
        else {                                   // light_env.cpp:129-146, Sampler::UniformSampleSphere (sampler.h:79-85)
            const float z = 1.0f - 2.0f * u0, r = sqrtf(fmaxf(0.f, 1.0f - z * z)), phi = 2.0f * LMB_PI * u1;
            float sp, cp;
            sincosf(phi, &sp, &cp);
            d = F3(r * cp, r * sp, z);
            pdfSA = LMB_INV_PI * 0.25f;          // light_env.cpp:185-189
        }
        if (!emitter_shape_hit(S, from, d, g)) return false;
        // PDFVal(SolidAngle).ConvertToArea(geomPrev, geom), probability.h:59-71
        f3 w = g.p - from;
        const float d2 = dot(w, w);
        w = qnormalize(w);
        pdfPL = qdiv(pdfSA * fabsf(dot(g.sn, neg(w))), d2);
        return true;
    }
    const lmb200_primitive& P = S.prims[S.lights[li].primitive];
    const float* cdf = S.light_cdf + S.light_cdf_off[li];
    const int n = (int)P.num_tris;
    int lo = 0, hi = n + 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (u0 < cdf[mid]) hi = mid; else lo = mid + 1; }
    int i = min(max(lo - 1, 0), n - 1);
    const float u2x = (u0 - cdf[i]) / (cdf[i + 1] - cdf[i]);
    const float s = sqrtf(fmaxf(0.f, u2x)), bx = 1.0f - s, by = u1 * s;
    const float* v = S.verts + 9 * (size_t)(P.first_tri + (uint32_t)i);
    const f3 p1 = ld3(v), p2 = ld3(v + 3), p3 = ld3(v + 6);
    g.p = (p1 * (1.0f - bx - by) + p2 * bx) + p3 * by;
    g.degenerated = false;
    g.gn = normalize(cross(p2 - p1, p3 - p1));
    g.sn = g.gn;
    basis(g.sn, g.dpdu, g.dpdv);
    pdfPL = S.light_inv_area[li];
    return true;
}

// ------------------------------------------------------------------------------------------------
// queue append: warp-aggregated (ballot + popc prefix, one atomicAdd per warp)
__device__ __forceinline__ uint32_t queue_slot(uint32_t* counter, bool want)
{
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return 0;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void film_add(float4* film, int pixel, f3 c)
{
    float* f = reinterpret_cast<float*>(film + pixel);
    atomicAdd(f, c.x); atomicAdd(f + 1, c.y); atomicAdd(f + 2, c.z);
}

// ------------------------------------------------------------------------------------------------
// k_logic: consume the hit of every traced slot, then (re)generate camera paths into finished slots.
__global__ void __launch_bounds__(256) k_logic(DevScene S, Pool P, RenderCfg cfg, float4* film)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (cfg.pool + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t i = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool inb = i < cfg.pool;
        int nv = inb ? P.nverts[i] : 0;
        bool alive = false;
        if (inb && nv > 0 && P.traced[i]) {
            // --- the slot's extend ray came back ---
            const float4 h = P.hit[i];
            const uint32_t tri = __float_as_uint(h.w);
            P.traced[i] = 0;
            if (tri != LMB200_MISS) {
                const float4 ro = P.ray_o[i], rd = P.ray_d[i];
                const f3 o = F3(ro.x, ro.y, ro.z), d = F3(rd.x, rd.y, rd.z);
                float4 thr = P.thr[i];
                const lmb200_primitive& prim = S.prims[S.tri_prim[tri]];
                bool cont = true;
                if (cfg.mode != LMB200_MODE_PTDIRECT && prim.light >= 0 && nv + 1 >= cfg.min_verts) {
                    // emission on hit (renderer_pt.cpp:183-194; light_area.cpp:105-115)
                    Geom g;
                    tri_geom(S, tri, h.y, h.z, o + d * h.x, g);
                    if (to_local(g, neg(d)).z > 0.f) {
                        f3 C = F3(thr.x, thr.y, thr.z) * ld3(S.lights[prim.light].Le);
                        if (cfg.mode == LMB200_MODE_PTMIS) {
                            // balance heuristic against the light-sampling pdf (renderer_ptmis.cpp:247-260):
                            // pdfPL / G(hit, previous vertex) * pdfL, G as renderutils.h:46-56
                            const float4 pv = P.prev[i];
                            f3 dd = o - g.p;
                            const float d2 = dot(dd, dd);
                            dd = qnormalize(dd);
                            float G = fabsf(dot(g.sn, dd));
                            if (nv > 1) G *= fabsf(dot(F3(pv.x, pv.y, pv.z), neg(dd)));   // the camera vertex is degenerated
                            G = qdiv(G, d2);
                            const float pdfDL = pv.w < 0.f ? 0.f : qdiv(S.light_inv_area[prim.light], G) * qdiv(1.0f, (float)S.num_lights);
                            const float pdfBS = fabsf(pv.w);
                            C = C * qdiv(pdfBS, pdfBS + pdfDL);
                        }
                        film_add(film, __float_as_int(thr.w), C);
                    }
                }
                // Russian roulette with the 4th uniform of this iteration's first block (renderer_pt.cpp:207-215)
                const float4 ua = rng_block(cfg.seed, P.sample[i], (uint32_t)(2 * nv - 1));
                if (ua.w > 0.5f) cont = false;
                if (cont) {
                    thr.x = thr.x / 0.5f; thr.y = thr.y / 0.5f; thr.z = thr.z / 0.5f;
                    nv++;
                    // loop-top test of the next iteration (renderer_pt.cpp:120-123) and bsdf::null termination
                    if (cfg.max_verts != -1 && nv >= cfg.max_verts) cont = false;
                    if (S.bsdfs[prim.bsdf].type == LMB200_BSDF_NULL) cont = false;
                }
                if (cont) {
                    const f3 p = o + d * h.x;
                    P.thr[i] = thr;
                    P.nverts[i] = nv;
                    P.vtx_p[i] = make_float4(p.x, p.y, p.z, h.w);
                    P.vtx_wi[i] = make_float4(-d.x, -d.y, -d.z, h.y);
                    P.vtx_v[i] = h.z;
                    alive = true;
                }
            }
        }
        // --- regeneration: camera vertex (sensor_pinhole.cpp:79-90) ---
        bool need = inb && !alive;
        unsigned long long sidx = 0;
        {
            const unsigned mask = __ballot_sync(0xffffffffu, need);
            if (mask) {
                const unsigned lane = threadIdx.x & 31u;
                const int leader = __ffs(mask) - 1;
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(P.next_sample, (unsigned long long)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                sidx = base + __popc(mask & ((1u << lane) - 1u));
            }
        }
        if (need) {
            if (sidx < cfg.sample_end) {
                bool ok = true;
                int pixel = -1;
                f3 cp = ld3(S.pos);
                if (S.cam_kind == LMB200_CAMERA_THINLENS) {
                    const float4 ul = rng_block(cfg.seed, sidx, LMB_LENS_BLOCK);
                    cp = camera_point(S, ul.x, ul.y);
                }
                if (cfg.mode != LMB200_MODE_PTDIRECT) {
                    // renderer::pt / ptmis compute the raster position up front and drop the sample if it fails (renderer_pt.cpp:94-99)
                    const float4 u = rng_block(cfg.seed, sidx, 0u);
                    float rx, ry;
                    float sx, sy;
                    camera_raster(cfg, sidx, u.y, u.z, sx, sy);
                    ok = raster_position(S, cp, camera_wo(S, sx, sy, cp), rx, ry);
                    if (ok) pixel = pixel_index(S, rx, ry);
                }
                if (ok && !(cfg.max_verts != -1 && 1 >= cfg.max_verts)) {
                    P.sample[i] = sidx;
                    P.nverts[i] = 1;
                    P.thr[i] = make_float4(1.f, 1.f, 1.f, __int_as_float(pixel));
                    P.vtx_p[i] = make_float4(cp.x, cp.y, cp.z, __uint_as_float(LMB200_MISS));
                    alive = true;
                } else {
                    P.nverts[i] = 0;   // sample consumed without a path; the slot is refilled next iteration
                }
            } else {
                P.nverts[i] = 0;
            }
        }
        const uint32_t q = queue_slot(P.qcount + 0, alive);
        if (alive) P.vq[q] = i;
    }
}

// k_nee: direct light sampling at every live vertex, camera vertex included (renderer_ptdirect.cpp:123-177)
__global__ void __launch_bounds__(256) k_nee(DevScene S, Pool P, RenderCfg cfg)
{
    const uint32_t nq = P.qcount[0];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (nq + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t qi = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool emit = false;
        f3 C = F3(0, 0, 0), p = F3(0, 0, 0), pl = F3(0, 0, 0);
        int pixel = 0;
        if (qi < nq && S.num_lights > 0 && (cfg.mode == LMB200_MODE_PTDIRECT || P.nverts[P.vq[qi]] + 1 >= cfg.min_verts)) {
            const uint32_t i = P.vq[qi];
            const int nv = P.nverts[i];
            const float4 vp = P.vtx_p[i];
            const uint32_t tri = __float_as_uint(vp.w);
            const bool is_sensor = tri == LMB200_MISS;
            const float4 thr = P.thr[i];
            const float4 ua = rng_block(cfg.seed, P.sample[i], (uint32_t)(2 * nv - 1));
            const int nL = (int)S.num_lights;
            const int li = min(max((int)(ua.x * (float)nL), 0), nL - 1);      // scene3.cpp:508-513
            const float pdfL = 1.0f / (float)nL;                               // scene3.cpp:526-530
            Geom gL;
            float pdfPL = 1.0f;
            p = F3(vp.x, vp.y, vp.z);
            const bool sampled = light_sample(S, li, p, ua.y, ua.z, gL, pdfPL);
            pl = gL.p;
            const f3 ppL = normalize(gL.p - p);
            f3 fsE;
            Geom g;
            float pdfB;      // pdf of sampling ppL from this vertex (ptmis)
            float srx = 0.f, sry = 0.f;      // raster position of ppL (camera vertex)
            bool inside = false;
            if (is_sensor) { const float im = importance_raster(S, p, ppL, srx, sry, inside); fsE = F3(im, im, im); g.degenerated = true; pdfB = im; }
            else {
                const float4 vw = P.vtx_wi[i];
                tri_geom(S, tri, vw.w, P.vtx_v[i], p, g);
                const lmb200_bsdf& B = S.bsdfs[S.prims[S.tri_prim[tri]].bsdf];
                fsE = bsdf_eval(B, bsdf_R(S, B, tri, vw.w, P.vtx_v[i]), g, F3(vw.x, vw.y, vw.z), ppL, true);
                pdfB = cfg.mode == LMB200_MODE_PTMIS ? bsdf_pdf(B, g, F3(vw.x, vw.y, vw.z), ppL, true) : 0.f;
            }
            // light_point.cpp:95-98, light_directional.cpp:182-185, light_env.cpp:191-208 emit Le in every direction;
            // light::area only on its front side (light_area.cpp:105-110)
            const f3 fsL = S.lights[li].kind != LMB200_LIGHT_AREA ? ld3(S.lights[li].Le)
                                          : (to_local(gL, neg(ppL)).z <= 0.f ? F3(0, 0, 0) : ld3(S.lights[li].Le));
            f3 d = gL.p - p;                                                   // RenderUtils::GeometryTerm, renderutils.h:46-56
            const float d2 = dot(d, d);
            d = ppL;
            float G = 1.0f;
            if (!is_sensor) G *= fabsf(dot(g.sn, d));
            if (!gL.degenerated) G *= fabsf(dot(gL.sn, neg(d)));
            G = qdiv(G, d2);
            C = ((F3(thr.x, thr.y, thr.z) * fsE) * fsL) * G;
            if (sampled && !black(C)) {
                C = C * qdiv(qdiv(1.0f, pdfL), pdfPL);
                if (cfg.mode == LMB200_MODE_PTMIS) {          // renderer_ptmis.cpp:163-170
                    const float pdfDL = qdiv(pdfPL, G) * pdfL;
                    C = C * qdiv(pdfDL, pdfDL + pdfB);
                }
                pixel = __float_as_int(thr.w);
                if (is_sensor) pixel = pixel_index(S, srx, sry);             // renderer_ptdirect.cpp:165-170 (C != 0 => inside)
                emit = true;
            }
        }
        const uint32_t q = queue_slot(P.qcount + 2, emit);
        if (emit) {
            // Scene3::Visible's shadow ray (scene3.h:107-116)
            const f3 dd = pl - p;
            const float L = sqrtf(dot(dd, dd));
            P.sq_o[q] = make_float4(p.x, p.y, p.z, LMB_EPS_ISECT);
            P.sq_d[q] = make_float4(dd.x / L, dd.y / L, dd.z / L, L * (1.0f - LMB_EPS_ISECT));
            P.sq_c[q] = make_float4(C.x, C.y, C.z, __int_as_float(pixel));
        }
    }
}

// k_bsdf: sample the next direction, evaluate pdf and fs, update throughput, emit the extend ray
// (renderer_pt.cpp:128-176 / renderer_ptdirect.cpp:183-246)
__global__ void __launch_bounds__(256) k_bsdf(DevScene S, Pool P, RenderCfg cfg)
{
    const uint32_t nq = P.qcount[0];
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (nq + stride - 1) / stride;
    for (uint32_t r = 0; r < rounds; r++) {
        const uint32_t qi = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool emit = false, is_primary = false;
        uint32_t i = 0;
        if (qi < nq) {
            i = P.vq[qi];
            const int nv = P.nverts[i];
            const float4 vp = P.vtx_p[i];
            const uint32_t tri = __float_as_uint(vp.w);
            const bool is_sensor = tri == LMB200_MISS;
            is_primary = is_sensor;
            const f3 p = F3(vp.x, vp.y, vp.z);
            float4 thr = P.thr[i];
            f3 wo = F3(0, 0, 0), fs, sn_here = F3(0, 0, 0);
            float pdfD;
            bool ok = true, specular_here = false;
            if (is_sensor) {
                const float4 u = rng_block(cfg.seed, P.sample[i], 0u);
                float sx, sy;
                camera_raster(cfg, P.sample[i], u.y, u.z, sx, sy);
                wo = camera_wo(S, sx, sy, p);
                float rx = 0.f, ry = 0.f;
                bool inside;
                const float im = importance_raster(S, p, wo, rx, ry, inside);
                pdfD = im; fs = F3(im, im, im);
                if (cfg.mode == LMB200_MODE_PTDIRECT) {
                    ok = inside;                                               // renderer_ptdirect.cpp:200-208
                    if (ok) thr.w = __int_as_float(pixel_index(S, rx, ry));
                }
            } else {
                const float4 vw = P.vtx_wi[i];
                const f3 wi = F3(vw.x, vw.y, vw.z);
                Geom g;
                tri_geom(S, tri, vw.w, P.vtx_v[i], p, g);
                const lmb200_bsdf& B = S.bsdfs[S.prims[S.tri_prim[tri]].bsdf];
                const float4 ub = rng_block(cfg.seed, P.sample[i], (uint32_t)(2 * nv));
                bsdf_sample(B, g, wi, ub.x, ub.y, ub.z, wo);
                pdfD = bsdf_pdf(B, g, wi, wo, false);
                fs = bsdf_eval(B, bsdf_R(S, B, tri, vw.w, P.vtx_v[i]), g, wi, wo, false);
                sn_here = g.sn;
                specular_here = is_specular(B);
            }
            if (ok && !black(fs)) {
                const float ipdf = qdiv(1.0f, pdfD);
                thr.x *= fs.x * ipdf; thr.y *= fs.y * ipdf; thr.z *= fs.z * ipdf;
                P.thr[i] = thr;
                // ptmis: a negative pdf marks a specular vertex (light sampling cannot reach it, renderer_ptmis.cpp:251-254)
                if (cfg.mode == LMB200_MODE_PTMIS) P.prev[i] = make_float4(sn_here.x, sn_here.y, sn_here.z, specular_here ? -pdfD : pdfD);
                P.ray_o[i] = make_float4(p.x, p.y, p.z, LMB_EPS_ISECT);       // scene3.cpp:461
                P.ray_d[i] = make_float4(wo.x, wo.y, wo.z, LMB_FLT_MAX);
                P.traced[i] = 1;
                emit = true;
            } else {
                P.nverts[i] = 0;                                               // path ends; slot is refilled by k_logic
            }
        }
        // two-ended extend queue: primary rays (coherent in generation order, see camera_raster) from the front, all other
        // rays from the back, so that the warps of k_extend get runs of primary rays instead of a mix
        const bool prim = emit && is_primary;
        const uint32_t qp = queue_slot(P.qcount + 1, prim);
        const uint32_t qs = queue_slot(P.qcount + 3, emit && !prim);
        if (prim) P.eq[qp] = i;
        else if (emit) P.eq[cfg.pool - 1u - qs] = i;
    }
}

// extend: closest hit over the extend queue (slot indirection), persistent warps with lane refill
struct ExtendIo {
    Pool P; uint32_t n_primary, n, last;      // ray qi: eq[qi] for qi < n_primary (front of the queue), else eq[last - (qi - n_primary)] (back)
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ uint32_t slot(uint64_t qi) const { return P.eq[qi < n_primary ? (uint32_t)qi : last - ((uint32_t)qi - n_primary)]; }
    __device__ __forceinline__ void load(uint64_t qi, float4& ro, float4& rd) const { const uint32_t i = slot(qi); ro = P.ray_o[i]; rd = P.ray_d[i]; }
    __device__ __forceinline__ void store(uint64_t qi, const Trav& T) const
    {
        const bool hit = T.hid != LMB200_MISS;
        P.hit[slot(qi)] = make_float4(hit ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid));
    }
};
// COUNT variants (instrumented, never timed): nodes / triangle records fetched, summed into P.next_sample[4..7]
__device__ __forceinline__ void add_work(unsigned long long* dst, const TravCounters& cnt)
{
    unsigned long long a = cnt.nodes, b = cnt.tris;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31u) == 0) { atomicAdd(dst, a); atomicAdd(dst + 1, b); }
}
template <bool COUNT>
__global__ void __launch_bounds__(LMB_TRACE_BLOCK, LMB_TRACE_MIN_BLOCKS)
k_extend(const BvhDev bvh, Pool P, unsigned long long* __restrict__ counter, const uint32_t pool)
{
    __shared__ uint2 smem[LMB_TRAV_SMEM_UINT2(LMB_TRACE_BLOCK)];
    ExtendIo io{P, P.qcount[1], P.qcount[1] + P.qcount[3], pool - 1u};
    TravCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    persistent_trace<false, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
    if (COUNT) add_work(P.next_sample + 4, cnt);
}

// shadow: any hit over the shadow queue, unoccluded contributions are splatted (film_hdr.cpp:218-223)
struct ShadowIo {
    Pool P; uint32_t n; float4* film;
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ void load(uint64_t qi, float4& ro, float4& rd) const { ro = P.sq_o[qi]; rd = P.sq_d[qi]; }
    __device__ __forceinline__ void store(uint64_t qi, const Trav& T) const
    {
        if (T.hid == LMB200_MISS) {
            const float4 c = P.sq_c[qi];
            film_add(film, __float_as_int(c.w), F3(c.x, c.y, c.z));
        }
    }
};
template <bool COUNT>
__global__ void __launch_bounds__(LMB_TRACE_BLOCK, LMB_TRACE_MIN_BLOCKS)
k_shadow(const BvhDev bvh, Pool P, unsigned long long* __restrict__ counter, float4* film)
{
    __shared__ uint2 smem[LMB_TRAV_SMEM_UINT2(LMB_TRACE_BLOCK)];
    ShadowIo io{P, P.qcount[2], film};
    TravCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    persistent_trace<true, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
    if (COUNT) add_work(P.next_sample + 6, cnt);
}

__global__ void k_stats(Pool P)
{
    // fold the queue sizes of this iteration into the running ray counters
    P.next_sample[1] += P.qcount[1] + P.qcount[3];
    P.next_sample[2] += P.qcount[2];
    P.next_sample[3] += P.qcount[0];
}

__global__ void k_rescale(float4* film, long long n, float s)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = film[i]; v.x *= s; v.y *= s; v.z *= s; film[i] = v; }
}

// primary-ray normal renderer (config 1): pixel-centre rays (renderer_raycast.cpp:77-83), |sn| shading
__global__ void k_normal_raygen(DevScene S, float4* rays, int pix0, int npix)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // ray i of this call = pixel pix0 + i
    if (i >= npix) return;
    const int x = (pix0 + i) % S.width, y = (pix0 + i) / S.width;
    // thin lens: the lens centre (lens sample (.5,.5)), i.e. the in-focus pinhole image
    const f3 cp = camera_point(S, 0.5f, 0.5f);
    const f3 wo = camera_wo(S, ((float)x + 0.5f) / (float)S.width, ((float)y + 0.5f) / (float)S.height, cp);
    rays[2 * i] = make_float4(cp.x, cp.y, cp.z, LMB_EPS_ISECT);
    rays[2 * i + 1] = make_float4(wo.x, wo.y, wo.z, LMB_FLT_MAX);
}
__global__ void k_normal_shade(DevScene S, const float4* hits, float4* film, int pix0, int npix)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 h = hits[i];
    const uint32_t tri = __float_as_uint(h.w);
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tri != LMB200_MISS) {
        Geom g;
        tri_geom(S, tri, h.y, h.z, F3(0, 0, 0), g);
        c = make_float4(fabsf(g.sn.x), fabsf(g.sn.y), fabsf(g.sn.z), 0.f);
    }
    film[pix0 + i] = c;
}

// ------------------------------------------------------------------------------------------------
// host side

struct Scene {
    int device = 0;
    Accel own_accel;                 // used unless a built accel is shared in (lmb200_scene_create_shared)
    Accel* accel = &own_accel;
    unsigned long long* d_counter = nullptr;   // work counters of this scene's traversal launches ([0] extend, [1] shadow)
    DevScene dev{};
    std::vector<void*> allocs;
    // pool (lazily sized)
    Pool pool{};
    uint32_t pool_size = 0;      // slots ALLOCATED (a render may use fewer)
    std::vector<void*> pool_allocs;
    void* h_pinned = nullptr;   // 8 x u64 readback + 4 x u64 upload
    void* film = nullptr;       // device film of the host-buffer call lmb200_render (reused across calls)
    int64_t film_cap = 0;
    int num_sms = 148;
    // k_shadow runs on its own stream beside k_bsdf / k_extend of the same iteration (tail filling)
    cudaStream_t shadow_stream = nullptr;
    cudaEvent_t ev_nee = nullptr, ev_shadow = nullptr;
    std::recursive_mutex job_mu;      // a scene renders one job at a time (pool, film, counters and streams are per scene)

    ~Scene()
    {
        cudaSetDevice(device);
        for (void* p : pool_allocs) cudaFree(p);
        for (void* p : allocs) cudaFree(p);
        if (d_counter) cudaFree(d_counter);
        if (film) cudaFree(film);
        if (h_pinned) cudaFreeHost(h_pinned);
        if (shadow_stream) cudaStreamDestroy(shadow_stream);
        if (ev_nee) cudaEventDestroy(ev_nee);
        if (ev_shadow) cudaEventDestroy(ev_shadow);
    }
};

template <typename T>
static int dev_upload(Scene* s, const T* src, size_t n, const T** out)
{
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scene)");
    s->allocs.push_back(d);
    if (n && (e = cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy(scene)");
    *out = reinterpret_cast<const T*>(d);
    return LMB200_OK;
}

template <typename T>
static int pool_alloc(Scene* s, T** out, size_t n)
{
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(pool)");
    s->pool_allocs.push_back(d);
    *out = reinterpret_cast<T*>(d);
    return LMB200_OK;
}

// The pool only grows: a short render (a warm-up pass, the last pass of a time budget) followed by a long one must not pay
// for 18 cudaFree + cudaMalloc pairs (~30 ms at 8 Mi slots) inside the second call.
static int ensure_pool(Scene* s, uint32_t n)
{
    if (s->pool_size >= n) return LMB200_OK;
    for (void* p : s->pool_allocs) cudaFree(p);
    s->pool_allocs.clear();
    s->pool_size = 0;
    Pool& P = s->pool;
    int rc = 0;
    if ((rc = pool_alloc(s, &P.sample, n))) return rc;
    if ((rc = pool_alloc(s, &P.nverts, n))) return rc;
    if ((rc = pool_alloc(s, &P.thr, n))) return rc;
    if ((rc = pool_alloc(s, &P.ray_o, n))) return rc;
    if ((rc = pool_alloc(s, &P.ray_d, n))) return rc;
    if ((rc = pool_alloc(s, &P.hit, n))) return rc;
    if ((rc = pool_alloc(s, &P.traced, n))) return rc;
    if ((rc = pool_alloc(s, &P.vtx_p, n))) return rc;
    if ((rc = pool_alloc(s, &P.vtx_wi, n))) return rc;
    if ((rc = pool_alloc(s, &P.vtx_v, n))) return rc;
    if ((rc = pool_alloc(s, &P.prev, n))) return rc;
    if ((rc = pool_alloc(s, &P.vq, n))) return rc;
    if ((rc = pool_alloc(s, &P.eq, n))) return rc;
    if ((rc = pool_alloc(s, &P.sq_o, n))) return rc;
    if ((rc = pool_alloc(s, &P.sq_d, n))) return rc;
    if ((rc = pool_alloc(s, &P.sq_c, n))) return rc;
    if ((rc = pool_alloc(s, &P.qcount, 4))) return rc;
    if ((rc = pool_alloc(s, &P.next_sample, 8))) return rc;
    s->pool_size = n;
    return LMB200_OK;
}

// Pixels [pix0, pix0 + npx) of the primary-ray image (the whole image on one GPU; lmb200_render_multi gives each GPU a
// contiguous share of the pixels and leaves the rest of its film zero, so that the film sum is the image).
static int render_normal(Scene* s, float4* film, cudaStream_t st, int pix0, int npx)
{
    if (npx <= 0) return LMB200_OK;
    float4* rays = nullptr; float4* hits = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&rays, sizeof(float4) * 2 * (size_t)npx)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(rays)");
    if ((e = cudaMalloc(&hits, sizeof(float4) * (size_t)npx)) != cudaSuccess) { cudaFree(rays); return cuda_fail(e, "cudaMalloc(hits)"); }
    k_normal_raygen<<<(npx + 255) / 256, 256, 0, st>>>(s->dev, rays, pix0, npx); g_launch_count++;
    int rc = trace_closest_dev(s->accel, rays, hits, (uint64_t)npx, nullptr, st, s->accel->ring_slot());
    if (!rc) { k_normal_shade<<<(npx + 255) / 256, 256, 0, st>>>(s->dev, hits, film, pix0, npx); g_launch_count++; }
    e = cudaGetLastError();
    const cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(rays); cudaFree(hits);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "render_normal launch");
    if (e2 != cudaSuccess) return cuda_fail(e2, "render_normal");
    return LMB200_OK;
}

// pix0 / npix: MODE_NORMAL only, the pixel range this call renders (npix < 0: the whole image)
static int render_dev(Scene* s, const lmb200_render_params* p, void* film_dev, cudaStream_t st, lmb200_render_stats* stats, int pix0 = 0, int npix = -1)
{
    std::lock_guard<std::recursive_mutex> job(s->job_mu);
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    float4* film = reinterpret_cast<float4*>(film_dev);
    const uint64_t launches0 = g_launch_count.load();
    if (p->mode == LMB200_MODE_NORMAL) {
        const auto t0 = std::chrono::steady_clock::now();
        if (npix < 0) { pix0 = 0; npix = s->dev.width * s->dev.height; }
        const int rc = render_normal(s, film, st, pix0, npix);
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            stats->samples = (int64_t)npix;
            stats->extend_rays = stats->samples;
            stats->iterations = 1;
            stats->launches = g_launch_count.load() - launches0;
            stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        return rc;
    }
    if (p->mode != LMB200_MODE_PT && p->mode != LMB200_MODE_PTDIRECT && p->mode != LMB200_MODE_PTMIS) return set_error(LMB200_E_INVALID, "unknown render mode");
    const bool nee = p->mode != LMB200_MODE_PT;
    if (s->dev.has_env && p->mode != LMB200_MODE_PTDIRECT)
        return set_error(LMB200_E_INVALID, "light::env is direct-light sampled only: use mode ptdirect (the reference's pt / ptmis dereference a null "
                                           "primitive when a ray escapes to the env emitter shape, scene3.cpp:463-475 + renderer_pt.cpp:183)");
    if (p->sample_end < p->sample_begin) return set_error(LMB200_E_INVALID, "sample_end < sample_begin");
    const int64_t todo = p->sample_end - p->sample_begin;
    uint32_t pool = p->pool_size > 0 ? (uint32_t)p->pool_size : (1u << 23);   // 8 Mi slots: best on B200 (profiles/r02_sweep.md; 4 Mi in round 1)
    if ((int64_t)pool > todo) pool = (uint32_t)std::max<int64_t>(todo, 1);
    pool = (pool + 31u) & ~31u;
    int rc = ensure_pool(s, pool);
    if (rc) return rc;
    Pool& P = s->pool;
    if (!s->h_pinned && (e = cudaHostAlloc(&s->h_pinned, 16 * sizeof(unsigned long long), cudaHostAllocDefault)) != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
    volatile unsigned long long* hp = reinterpret_cast<volatile unsigned long long*>(s->h_pinned);
    if (!s->shadow_stream) {
        if ((e = cudaStreamCreateWithFlags(&s->shadow_stream, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&s->ev_nee, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&s->ev_shadow, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
    }

    RenderCfg cfg;
    cfg.mode = p->mode; cfg.max_verts = p->max_num_vertices; cfg.min_verts = p->min_num_vertices;
    cfg.seed = p->seed; cfg.sample_end = (unsigned long long)p->sample_end; cfg.pool = pool;
    cfg.tile_x0 = 0.f; cfg.tile_y0 = 0.f; cfg.tile_sx = 1.f; cfg.tile_sy = 1.f;
    {
        // coherent camera sample groups (camera_raster): tiles of primary_tile x primary_tile pixels; 0 = automatic, < 0 = off
        const int tp = p->primary_tile == 0 ? lmb200_default_primary_tile(s->dev.width, s->dev.height, p->num_samples) : p->primary_tile;
        cfg.gt_nx = tp > 0 ? (s->dev.width + tp - 1) / tp : 0;
        cfg.gt_ny = tp > 0 ? (s->dev.height + tp - 1) / tp : 0;
    }
    if (p->tile[0] != 0.f || p->tile[1] != 0.f || p->tile[2] != 0.f || p->tile[3] != 0.f) {
        if (!(p->tile[0] >= 0.f && p->tile[1] >= 0.f && p->tile[2] <= 1.f && p->tile[3] <= 1.f && p->tile[2] > p->tile[0] && p->tile[3] > p->tile[1]))
            return set_error(LMB200_E_INVALID, "tile must satisfy 0 <= x0 < x1 <= 1 and 0 <= y0 < y1 <= 1");
        cfg.tile_x0 = p->tile[0]; cfg.tile_y0 = p->tile[1];
        cfg.tile_sx = p->tile[2] - p->tile[0]; cfg.tile_sy = p->tile[3] - p->tile[1];
    }

    // every runtime call of the loop is checked: a failed memset / copy / event must not turn into a silently wrong image
    struct Events {
        cudaEvent_t a = nullptr, b = nullptr;
        ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } ev;
#define LMB_CK(call, what) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)
    LMB_CK(cudaMemsetAsync(P.nverts, 0, sizeof(int) * pool, st), "cudaMemsetAsync(nverts)");
    LMB_CK(cudaMemsetAsync(P.traced, 0, pool, st), "cudaMemsetAsync(traced)");
    unsigned long long* h_init = reinterpret_cast<unsigned long long*>(s->h_pinned) + 8;     // [8..16)     // pinned: the async copy really is async
    h_init[0] = (unsigned long long)p->sample_begin;
    for (int k = 1; k < 8; k++) h_init[k] = 0ull;
    LMB_CK(cudaMemcpyAsync(P.next_sample, h_init, 8 * sizeof(unsigned long long), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(next_sample)");

    LMB_CK(cudaEventCreate(&ev.a), "cudaEventCreate");
    LMB_CK(cudaEventCreate(&ev.b), "cudaEventCreate");
    LMB_CK(cudaEventRecord(ev.a, st), "cudaEventRecord");
    const int logic_blocks = s->num_sms * 8;
    const int trace_blocks = s->accel->num_sms * s->accel->trace_blocks_per_sm;
    const BvhDev bvh = bvh_dev(s->accel);
    const bool count = p->count_work != 0;
    int64_t iters = 0;
    for (;;) {
        LMB_CK(cudaMemsetAsync(P.qcount, 0, 4 * sizeof(uint32_t), st), "cudaMemsetAsync(qcount)");
        LMB_CK(cudaMemsetAsync(s->d_counter, 0, 2 * sizeof(unsigned long long), st), "cudaMemsetAsync(work counters)");
        k_logic<<<logic_blocks, 256, 0, st>>>(s->dev, P, cfg, film);
        if (nee) {
            // the shadow rays of this iteration are traced on a second stream while the main stream goes on with
            // k_bsdf and k_extend: the two persistent traversal kernels fill each other's tails
            k_nee<<<logic_blocks, 256, 0, st>>>(s->dev, P, cfg);
            LMB_CK(cudaEventRecord(s->ev_nee, st), "cudaEventRecord(nee)");
            LMB_CK(cudaStreamWaitEvent(s->shadow_stream, s->ev_nee, 0), "cudaStreamWaitEvent(nee)");
            if (count) k_shadow<true><<<trace_blocks, LMB_TRACE_BLOCK, 0, s->shadow_stream>>>(bvh, P, s->d_counter + 1, film);
            else k_shadow<false><<<trace_blocks, LMB_TRACE_BLOCK, 0, s->shadow_stream>>>(bvh, P, s->d_counter + 1, film);
            LMB_CK(cudaEventRecord(s->ev_shadow, s->shadow_stream), "cudaEventRecord(shadow)");
        }
        k_bsdf<<<logic_blocks, 256, 0, st>>>(s->dev, P, cfg);
        if (count) k_extend<true><<<trace_blocks, LMB_TRACE_BLOCK, 0, st>>>(bvh, P, s->d_counter, pool);
        else k_extend<false><<<trace_blocks, LMB_TRACE_BLOCK, 0, st>>>(bvh, P, s->d_counter, pool);
        if (nee) LMB_CK(cudaStreamWaitEvent(st, s->ev_shadow, 0), "cudaStreamWaitEvent(shadow)");
        k_stats<<<1, 1, 0, st>>>(P);
        LMB_CK(cudaGetLastError(), "wavefront kernel launch");
        g_launch_count += nee ? 6 : 4;
        LMB_CK(cudaMemcpyAsync(s->h_pinned, P.qcount, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(qcount)");
        LMB_CK(cudaMemcpyAsync(reinterpret_cast<unsigned long long*>(s->h_pinned) + 4, P.next_sample, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(sample counter)");
        LMB_CK(cudaStreamSynchronize(st), "wavefront iteration");
        iters++;
        const uint32_t live = reinterpret_cast<volatile uint32_t*>(s->h_pinned)[0];
        if (live == 0 && hp[4] >= cfg.sample_end) break;     // no live vertex and the sample counter is exhausted
    }
    LMB_CK(cudaEventRecord(ev.b, st), "cudaEventRecord");
    LMB_CK(cudaMemcpyAsync(s->h_pinned, P.next_sample, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(counters)");
    LMB_CK(cudaStreamSynchronize(st), "wavefront end");
    float ms = 0.f;
    LMB_CK(cudaEventElapsedTime(&ms, ev.a, ev.b), "cudaEventElapsedTime");
#undef LMB_CK
    if (stats) {
        stats->samples = todo;
        stats->extend_rays = (int64_t)hp[1];
        stats->shadow_rays = (int64_t)hp[2];
        stats->iterations = iters;
        stats->launches = g_launch_count.load() - launches0;
        stats->seconds = ms * 1e-3;
        stats->reduce_seconds = 0;
        stats->vertices = (int64_t)hp[3];
        stats->extend_nodes = (int64_t)hp[4]; stats->extend_tris = (int64_t)hp[5];
        stats->shadow_nodes = (int64_t)hp[6]; stats->shadow_tris = (int64_t)hp[7];
    }
    return LMB200_OK;
}

}  // namespace lmb200

using namespace lmb200;

extern "C" {

static lmb200_scene* scene_create(int device, const lmb200_scene_desc* d, int builder, Accel* shared);

lmb200_scene* lmb200_scene_create(int device, const lmb200_scene_desc* d) { return scene_create(device, d, LMB200_BUILD_DEFAULT, nullptr); }

lmb200_scene* lmb200_scene_create_ex(int device, const lmb200_scene_desc* d, int builder) { return scene_create(device, d, builder, nullptr); }

lmb200_scene* lmb200_scene_create_shared(const lmb200_scene_desc* d, lmb200_accel* accel)
{
    Accel* a = reinterpret_cast<Accel*>(accel);
    if (!a || !d) { set_error(LMB200_E_INVALID, "null argument"); return nullptr; }
    if (!a->built || a->host_only || !a->d_units) { set_error(LMB200_E_STATE, "shared accel is not built on a device"); return nullptr; }
    if (a->bvh.stats.num_triangles != d->num_tris) { set_error(LMB200_E_INVALID, "shared accel was built over a different triangle list"); return nullptr; }
    return scene_create(a->device, d, LMB200_BUILD_HOST_SAH, a);
}

static lmb200_scene* scene_create(int device, const lmb200_scene_desc* d, int builder, Accel* shared)
{
    if (builder < LMB200_BUILD_HOST_SAH || builder > LMB200_BUILD_GPU_LBVH_SAH) { set_error(LMB200_E_INVALID, "unknown builder"); return nullptr; }
    if (!d || (d->num_tris && (!d->verts || !d->tri_prim)) || !d->prims || !d->bsdfs) { set_error(LMB200_E_INVALID, "null scene field"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error(LMB200_E_CUDA, "no CUDA device available (lmb200 has no CPU fallback)"); return nullptr; }
    if (device < 0 || device >= ndev) { set_error(LMB200_E_INVALID, "bad device ordinal"); return nullptr; }
    for (uint64_t t = 0; t < d->num_tris; t++) if (d->tri_prim[t] >= d->num_prims) { set_error(LMB200_E_INVALID, "tri_prim out of range"); return nullptr; }
    for (uint32_t i = 0; i < d->num_prims; i++) {
        const lmb200_primitive& P = d->prims[i];
        if (P.bsdf < 0 || (uint32_t)P.bsdf >= d->num_bsdfs || P.light >= (int32_t)d->num_lights || (uint64_t)P.first_tri + P.num_tris > d->num_tris) { set_error(LMB200_E_INVALID, "primitive out of range"); return nullptr; }
    }
    for (uint32_t i = 0; i < d->num_bsdfs; i++)
        if (d->bsdfs[i].type < LMB200_BSDF_NULL || d->bsdfs[i].type > LMB200_BSDF_FLESNEL) { set_error(LMB200_E_INVALID, "unknown bsdf type"); return nullptr; }
    Scene* s = new Scene;
    s->device = device;
    cudaSetDevice(device);
    if (shared) s->accel = shared;
    else {
        s->own_accel.device = device;
        if (lmb200_accel_build_ex(reinterpret_cast<lmb200_accel*>(&s->own_accel), d->verts, d->num_tris, builder)) { delete s; return nullptr; }
    }
    if (cudaMalloc(&s->d_counter, 4 * sizeof(unsigned long long)) != cudaSuccess) { set_error(LMB200_E_CUDA, "cudaMalloc(scene counters)"); delete s; return nullptr; }
    s->num_sms = s->accel->num_sms;
    DevScene& D = s->dev;
    int rc = 0;
    rc |= dev_upload(s, d->verts, 9 * d->num_tris, &D.verts);
    if (d->normals) rc |= dev_upload(s, d->normals, 9 * d->num_tris, &D.normals); else D.normals = nullptr;
    rc |= dev_upload(s, d->tri_prim, d->num_tris, &D.tri_prim);
    rc |= dev_upload(s, d->prims, d->num_prims, &D.prims);
    rc |= dev_upload(s, d->bsdfs, d->num_bsdfs, &D.bsdfs);
    rc |= dev_upload(s, d->lights, d->num_lights, &D.lights);
    // TexR textures: concatenated texels + (width, height, first texel) per texture
    D.uvs = nullptr; D.tex_rgb = nullptr; D.tex_info = nullptr;
    {
        bool textured = false;
        for (uint32_t b = 0; b < d->num_bsdfs; b++) {
            if (d->bsdfs[b].texR < 0 || (uint32_t)d->bsdfs[b].texR > d->num_textures) { set_error(LMB200_E_INVALID, "bsdf texR out of range"); delete s; return nullptr; }
            textured = textured || d->bsdfs[b].texR > 0;
        }
        if (textured) {
            std::vector<float> texels; std::vector<int4> info;
            for (uint32_t t = 0; t < d->num_textures; t++) {
                const lmb200_texture& T = d->textures[t];
                if (T.width <= 0 || T.height <= 0 || !T.rgb) { set_error(LMB200_E_INVALID, "empty texture"); delete s; return nullptr; }
                info.push_back(make_int4(T.width, T.height, (int)(texels.size() / 3), 0));
                texels.insert(texels.end(), T.rgb, T.rgb + 3 * (size_t)T.width * (size_t)T.height);
            }
            rc |= dev_upload(s, texels.data(), texels.size(), &D.tex_rgb);
            rc |= dev_upload(s, info.data(), info.size(), &D.tex_info);
            if (d->uvs) rc |= dev_upload(s, d->uvs, 6 * d->num_tris, &D.uvs);
        }
    }
    // area CDFs exactly as TriangleUtils::CreateTriangleAreaDist + Distribution1D::Normalize
    // (triangleutils.h:47-68, dist.h:44-60); Length uses the _mm_dp_ps summation order
    std::vector<float> cdf; std::vector<uint32_t> off; std::vector<float> inv_area;
    for (uint32_t li = 0; li < d->num_lights; li++) {
        if (d->lights[li].primitive < 0 || (uint32_t)d->lights[li].primitive >= d->num_prims) { set_error(LMB200_E_INVALID, "light primitive out of range"); delete s; return nullptr; }
        const lmb200_primitive& P = d->prims[d->lights[li].primitive];
        off.push_back((uint32_t)cdf.size());
        const int kind = d->lights[li].kind;
        if (kind == LMB200_LIGHT_POINT || kind == LMB200_LIGHT_DIRECTIONAL || kind == LMB200_LIGHT_ENV) {   // no mesh to sample
            if (kind != LMB200_LIGHT_POINT && !(d->sphere_radius > 0.f)) { set_error(LMB200_E_INVALID, "directional / env lights need lmb200_scene_desc::sphere_radius > 0"); delete s; return nullptr; }
            cdf.push_back(0.f); cdf.push_back(1.f); inv_area.push_back(1.0f); continue;
        }
        if (kind != LMB200_LIGHT_AREA) { set_error(LMB200_E_INVALID, "unknown light kind"); delete s; return nullptr; }
        if (P.num_tris == 0) { set_error(LMB200_E_INVALID, "area light without triangles"); delete s; return nullptr; }
        const size_t base = cdf.size();
        cdf.push_back(0.f);
        float sum = 0.f;
        for (uint32_t i = 0; i < P.num_tris; i++) {
            const float* v = d->verts + 9 * (size_t)(P.first_tri + i);
            const float e1[3] = {v[3] - v[0], v[4] - v[1], v[5] - v[2]}, e2[3] = {v[6] - v[0], v[7] - v[1], v[8] - v[2]};
            volatile float a0 = e1[1] * e2[2], a1 = e1[2] * e2[1], b0 = e1[2] * e2[0], b1 = e1[0] * e2[2], c0 = e1[0] * e2[1], c1 = e1[1] * e2[0];
            const float cx = a0 - a1, cy = b0 - b1, cz = c0 - c1;
            volatile float xx = cx * cx, yy = cy * cy, zz = cz * cz;
            volatile float s01 = xx + yy, s23 = zz + 0.0f;
            const float area = std::sqrt(s01 + s23) * 0.5f;
            cdf.push_back(cdf.back() + area);
            sum += area;
        }
        const float inv = 1.0f / cdf.back();
        for (size_t k = base; k < cdf.size(); k++) cdf[k] *= inv;
        inv_area.push_back(1.0f / sum);
    }
    rc |= dev_upload(s, cdf.data(), cdf.size(), &D.light_cdf);
    rc |= dev_upload(s, off.data(), off.size(), &D.light_cdf_off);
    rc |= dev_upload(s, inv_area.data(), inv_area.size(), &D.light_inv_area);
    if (rc) { delete s; return nullptr; }
    D.num_lights = d->num_lights;
    for (int k = 0; k < 3; k++) { D.pos[k] = d->camera.position[k]; D.vx[k] = d->camera.vx[k]; D.vy[k] = d->camera.vy[k]; D.vz[k] = d->camera.vz[k]; }
    D.tan_fov = std::tan(d->camera.fov * 0.5f);
    D.width = d->camera.width; D.height = d->camera.height;
    D.aspect = (float)d->camera.width / (float)d->camera.height;
    D.cam_kind = d->camera.kind;
    D.lens_radius = d->camera.lens_radius; D.focal_distance = d->camera.focal_distance;
    for (int k = 0; k < 3; k++) D.sph_c[k] = d->sphere_center[k];
    D.sph_r = d->sphere_radius;
    D.has_env = 0;
    for (uint32_t li = 0; li < d->num_lights; li++) if (d->lights[li].kind == LMB200_LIGHT_ENV) D.has_env = 1;
    return reinterpret_cast<lmb200_scene*>(s);
}

void lmb200_scene_destroy(lmb200_scene* s) { delete reinterpret_cast<Scene*>(s); }

lmb200_accel* lmb200_scene_accel(lmb200_scene* s) { return s ? reinterpret_cast<lmb200_accel*>(reinterpret_cast<Scene*>(s)->accel) : nullptr; }

// Process-wide registry so that two plugins loaded RTLD_LOCAL can find each other's objects:
// accel::lmb200 publishes its device BVH under its component address, renderer::lmb200pt looks it up.
namespace { std::mutex g_reg_mu; std::map<const void*, lmb200_accel*> g_registry; }

void lmb200_registry_put(const void* owner, lmb200_accel* accel)
{
    std::lock_guard<std::mutex> lock(g_reg_mu);
    if (accel) g_registry[owner] = accel; else g_registry.erase(owner);
}

lmb200_accel* lmb200_registry_get(const void* owner)
{
    std::lock_guard<std::mutex> lock(g_reg_mu);
    auto it = g_registry.find(owner);
    return it == g_registry.end() ? nullptr : it->second;
}

int lmb200_default_primary_tile(int width, int height, int64_t num_samples)
{
    // the smallest tile whose rounds (every tile visited once: 32 * #tiles samples) fit at least four times into the job, so
    // that even a partial last round leaves no visible coverage pattern; small tiles make the primary rays of a group
    // more coherent (ptdirect on configs[2]: 986 Msamples/s ungrouped, 1046 / 1072 / 1093 / 1113 with 16 / 8 / 4 / 2 pixels)
    for (int tp = 2; tp <= 64; tp *= 2) {
        const int64_t tiles = (int64_t)((width + tp - 1) / tp) * ((height + tp - 1) / tp);
        if (num_samples >= 4 * 32 * tiles) return tp;
    }
    return -1;
}

int lmb200_render_dev(lmb200_scene* h, const lmb200_render_params* p, void* film_dev, void* stream, lmb200_render_stats* stats)
{
    if (!h || !p || !film_dev) return set_error(LMB200_E_INVALID, "null argument");
    return render_dev(reinterpret_cast<Scene*>(h), p, film_dev, reinterpret_cast<cudaStream_t>(stream), stats);
}

int lmb200_film_rescale_dev(void* film_dev, int64_t num_pixels, float scale, void* stream)
{
    if (!film_dev) return set_error(LMB200_E_INVALID, "null argument");
    k_rescale<<<(unsigned)((num_pixels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(film_dev), num_pixels, scale);
    g_launch_count++;
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? LMB200_OK : cuda_fail(e, "k_rescale");
}

int lmb200_render(lmb200_scene* h, const lmb200_render_params* p, float* film_host, lmb200_render_stats* stats)
{
    if (!h || !p || !film_host) return set_error(LMB200_E_INVALID, "null argument");
    Scene* s = reinterpret_cast<Scene*>(h);
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    std::lock_guard<std::recursive_mutex> job(s->job_mu);
    const int64_t npx = (int64_t)s->dev.width * s->dev.height;
    // the device film lives as long as the scene (a renderer calls this once per frame / progress image)
    if (s->film_cap < npx) {
        if (s->film) { cudaFree(s->film); s->film = nullptr; s->film_cap = 0; }
        if ((e = cudaMalloc(&s->film, sizeof(float4) * (size_t)npx)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(film)");
        s->film_cap = npx;
    }
    void* film = s->film;
    if ((e = cudaMemsetAsync(film, 0, sizeof(float4) * (size_t)npx, 0)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(film)");
    int rc = render_dev(s, p, film, 0, stats);
    if (!rc && p->mode != LMB200_MODE_NORMAL) rc = lmb200_film_rescale_dev(film, npx, (float)npx / (float)p->num_samples, 0);
    if (!rc && (e = cudaMemcpy(film_host, film, sizeof(float4) * (size_t)npx, cudaMemcpyDeviceToHost)) != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(film)");
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU / time-budgeted rendering in one process. A Session keeps one UNSCALED film per GPU that
// accumulates over passes; films are combined only when an image is needed (progress tick, end):
// one ncclReduce to device 0 over NVLink, replacing contexts.combine_each(film->Accumulate)
// (scheduler.cpp:280-285), followed by the W*H/processed rescale (scheduler.cpp:288).
namespace {

typedef ncclResult_t (*fn_init_all)(ncclComm_t*, int, const int*);
typedef ncclResult_t (*fn_void)(void);
typedef ncclResult_t (*fn_reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*fn_destroy)(ncclComm_t);
typedef ncclResult_t (*fn_version)(int*);
typedef const char* (*fn_errstr)(ncclResult_t);

// restores the calling thread's current device when a multi-GPU call returns (per-ray Accel3::Intersect and the
// caller's own CUDA code must not find another device current afterwards)
struct DeviceGuard {
    int dev = -1;
    DeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = -1; } }
    ~DeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};

struct Session {
    std::vector<Scene*> sc;
    std::vector<int> devs;
    std::vector<void*> films;
    std::vector<cudaStream_t> streams;
    std::vector<ncclComm_t> comms;
    void* scratch = nullptr;       // on device 0: reduced + rescaled copy that goes to the host
    int64_t npx = 0;
    fn_void gs = nullptr, ge = nullptr;
    fn_reduce reduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    int nccl_version = 0;
    double reduce_seconds = 0;     // device time of the film reductions so far (events on device 0's stream)
    lmb200_render_stats total{};

    int nccl_fail(ncclResult_t r, const char* what) { return set_error(LMB200_E_NCCL, std::string(what) + ": " + (errstr ? errstr(r) : "NCCL error")); }

    // sum over the GPUs of `src[g]` (count floats each) into `dst` on device 0, on the per-GPU streams
    int reduce_all(const std::vector<void*>& src, void* dst, size_t count)
    {
        const int n = (int)sc.size();
        ncclResult_t r = gs();
        if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
        ncclResult_t bad = ncclSuccess;
        for (int g = 0; g < n; g++) {
            cudaSetDevice(devs[g]);
            r = reduce(src[g], g == 0 ? dst : src[g], count, ncclFloat32, ncclSum, 0, comms[g], streams[g]);
            if (r != ncclSuccess) bad = r;
        }
        r = ge();
        if (bad != ncclSuccess) return nccl_fail(bad, "ncclReduce");
        if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
        return LMB200_OK;
    }

    // Known-answer reduce right after the communicators are created: GPU g contributes (g + 1) * (i % 7 + 1); small
    // integers, so the sum is exact in fp32 whatever the reduction order. Catches a libnccl whose datatype / operator
    // numbering or calling convention differs from the header this file was compiled against.
    int self_test()
    {
        const int n = (int)sc.size();
        const size_t cnt = 1024;
        std::vector<void*> buf(n, nullptr);
        std::vector<float> h(cnt);
        int rc = LMB200_OK;
        for (int g = 0; g < n && !rc; g++) {
            for (size_t i = 0; i < cnt; i++) h[i] = (float)((g + 1) * (int)(i % 7 + 1));
            cudaError_t e = cudaSetDevice(devs[g]);
            if (e == cudaSuccess) e = cudaMalloc(&buf[g], (g == 0 ? 2 : 1) * cnt * sizeof(float));
            if (e == cudaSuccess) e = cudaMemcpyAsync(buf[g], h.data(), cnt * sizeof(float), cudaMemcpyHostToDevice, streams[g]);
            if (e == cudaSuccess) e = cudaStreamSynchronize(streams[g]);
            if (e != cudaSuccess) rc = cuda_fail(e, "nccl self-test setup");
        }
        if (!rc) rc = reduce_all(buf, reinterpret_cast<float*>(buf[0]) + cnt, cnt);
        if (!rc) {
            for (int g = 0; g < n; g++) { cudaSetDevice(devs[g]); cudaStreamSynchronize(streams[g]); }
            cudaSetDevice(devs[0]);
            const cudaError_t e = cudaMemcpy(h.data(), reinterpret_cast<float*>(buf[0]) + cnt, cnt * sizeof(float), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = cuda_fail(e, "nccl self-test readback");
            const int tri = n * (n + 1) / 2;
            for (size_t i = 0; i < cnt && !rc; i++)
                if (h[i] != (float)(tri * (int)(i % 7 + 1)))
                    rc = set_error(LMB200_E_NCCL, "ncclReduce self-test returned a wrong sum (libnccl " + std::to_string(nccl_version) + " is not ABI-compatible with the NCCL 2 header)");
        }
        for (int g = 0; g < n; g++) if (buf[g]) { cudaSetDevice(devs[g]); cudaFree(buf[g]); }
        return rc;
    }

    int open(lmb200_scene** scenes, int n)
    {
        for (int g = 0; g < n; g++) {
            Scene* s = reinterpret_cast<Scene*>(scenes[g]);
            if (!s) return set_error(LMB200_E_INVALID, "null scene");
            sc.push_back(s); devs.push_back(s->device);
        }
        npx = (int64_t)sc[0]->dev.width * sc[0]->dev.height;
        films.assign(n, nullptr); streams.assign(n, nullptr); comms.assign(n, nullptr);
        for (int g = 0; g < n; g++) {
            cudaError_t e = cudaSetDevice(devs[g]);
            if (e == cudaSuccess) e = cudaMalloc(&films[g], sizeof(float4) * (size_t)npx);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&streams[g], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaMemsetAsync(films[g], 0, sizeof(float4) * (size_t)npx, streams[g]);
            if (e != cudaSuccess) return cuda_fail(e, "session film");
        }
        cudaSetDevice(devs[0]);
        cudaError_t e = cudaMalloc(&scratch, sizeof(float4) * (size_t)npx);
        if (e != cudaSuccess) return cuda_fail(e, "session scratch");
        if (n > 1) {
            // Which libnccl: (1) $LMB200_NCCL_LIB if set (lmb200py points it at the copy bundled with PyTorch, so that a
            // process that also imports torch ends up with ONE libnccl.so.2 — the dynamic linker shares objects by soname,
            // and torch's libtorch_cuda.so needs symbols newer than an older system copy has); (2) a libnccl.so.2 that is
            // already loaded in the process; (3) the system library. Never RTLD_GLOBAL: nobody else should bind to it.
            void* lib = nullptr;
            if (const char* path = getenv("LMB200_NCCL_LIB")) if (*path) lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
            if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
            if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
            if (!lib) return set_error(LMB200_E_NCCL, std::string("cannot load libnccl: ") + dlerror());
            fn_init_all init = (fn_init_all)dlsym(lib, "ncclCommInitAll");
            gs = (fn_void)dlsym(lib, "ncclGroupStart"); ge = (fn_void)dlsym(lib, "ncclGroupEnd");
            reduce = (fn_reduce)dlsym(lib, "ncclReduce");
            destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
            errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
            fn_version version = (fn_version)dlsym(lib, "ncclGetVersion");
            if (!init || !gs || !ge || !reduce || !destroy || !version) return set_error(LMB200_E_NCCL, "libnccl lacks required symbols");
            // the enum values passed to ncclReduce come from the NCCL 2 header: refuse any other major version
            if (version(&nccl_version) != ncclSuccess || nccl_version < 20000 || nccl_version >= 30000)
                return set_error(LMB200_E_NCCL, "unsupported libnccl version " + std::to_string(nccl_version) + " (need 2.x, built against " + std::to_string(NCCL_VERSION_CODE) + ")");
            for (int g = 0; g < n; g++) for (int h = 0; h < g; h++)
                if (devs[g] == devs[h]) return set_error(LMB200_E_INVALID, "lmb200_render_multi: two scenes live on the same device");
            const ncclResult_t r = init(comms.data(), n, devs.data());
            if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitAll");
            if (const int rc = self_test()) return rc;
        }
        memset(&total, 0, sizeof(total));
        return LMB200_OK;
    }

    // renders global samples [begin,end) of p, split contiguously over the GPUs, into the per-GPU films
    int pass(const lmb200_render_params* p, int64_t begin, int64_t end)
    {
        const int n = (int)sc.size();
        std::vector<int> rcs(n, 0);
        std::vector<std::string> errs(n);
        std::vector<lmb200_render_stats> st(n);
        auto work = [&](int g) {
            lmb200_render_params q = *p;
            q.sample_begin = begin + (end - begin) * g / n;
            q.sample_end = begin + (end - begin) * (g + 1) / n;
            if (p->tile_partition && n > 1) {
                // GPU g owns the horizontal strip [g/n, (g+1)/n) of the raster (of the caller's tile, if any): equal areas
                // and equal sample counts, so the summed film has the expected value of whole-image sampling
                const bool whole = p->tile[0] == 0.f && p->tile[1] == 0.f && p->tile[2] == 0.f && p->tile[3] == 0.f;
                const float x0 = whole ? 0.f : p->tile[0], x1 = whole ? 1.f : p->tile[2];
                const float y0 = whole ? 0.f : p->tile[1], y1 = whole ? 1.f : p->tile[3];
                q.tile[0] = x0; q.tile[2] = x1;
                q.tile[1] = y0 + (y1 - y0) * (float)g / (float)n;
                q.tile[3] = g + 1 == n ? y1 : y0 + (y1 - y0) * (float)(g + 1) / (float)n;
            }
            if (p->mode == LMB200_MODE_NORMAL) {
                // primary-ray image: GPU g renders pixels [npx g/n, npx (g+1)/n) and leaves the rest of its film zero
                const int pix0 = (int)(npx * g / n), pix1 = (int)(npx * (g + 1) / n);
                rcs[g] = render_dev(sc[g], &q, films[g], streams[g], &st[g], pix0, pix1 - pix0);
            } else {
                rcs[g] = render_dev(sc[g], &q, films[g], streams[g], &st[g]);
            }
            if (rcs[g]) errs[g] = g_last_error;
        };
        std::vector<std::thread> th;
        for (int g = 1; g < n; g++) th.emplace_back(work, g);
        work(0);
        for (auto& t : th) t.join();
        double secs = 0; int64_t iters = 0;
        for (int g = 0; g < n; g++) {
            if (rcs[g]) return set_error(rcs[g], errs[g]);
            total.samples += st[g].samples; total.extend_rays += st[g].extend_rays; total.shadow_rays += st[g].shadow_rays;
            total.launches += st[g].launches; total.vertices += st[g].vertices;
            total.extend_nodes += st[g].extend_nodes; total.extend_tris += st[g].extend_tris;
            total.shadow_nodes += st[g].shadow_nodes; total.shadow_tris += st[g].shadow_tris;
            secs = std::max(secs, st[g].seconds); iters = std::max(iters, st[g].iterations);
        }
        total.seconds += secs; total.iterations += iters;
        return LMB200_OK;
    }

    // sum of the films * scale -> host (the per-GPU films keep accumulating afterwards)
    int gather(float scale, float* film_host)
    {
        const int n = (int)sc.size();
        if (n > 1) {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            cudaSetDevice(devs[0]);
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, streams[0]);
            const int rrc = reduce_all(films, scratch, (size_t)npx * 4);
            cudaSetDevice(devs[0]);
            cudaEventRecord(e1, streams[0]);
            cudaError_t se = cudaSuccess;
            for (int g = 0; g < n; g++) { cudaSetDevice(devs[g]); const cudaError_t x = cudaStreamSynchronize(streams[g]); if (x != cudaSuccess) se = x; }
            cudaSetDevice(devs[0]);
            float ms = 0.f;
            if (!rrc && se == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) reduce_seconds += ms * 1e-3;
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            if (rrc) return rrc;
            if (se != cudaSuccess) return cuda_fail(se, "film reduce");
        } else {
            cudaSetDevice(devs[0]);
            cudaMemcpyAsync(scratch, films[0], sizeof(float4) * (size_t)npx, cudaMemcpyDeviceToDevice, streams[0]);
        }
        int rc = LMB200_OK;
        if (scale != 1.0f) rc = lmb200_film_rescale_dev(scratch, npx, scale, streams[0]);
        cudaError_t e = cudaMemcpyAsync(film_host, scratch, sizeof(float4) * (size_t)npx, cudaMemcpyDeviceToHost, streams[0]);
        if (e == cudaSuccess) e = cudaStreamSynchronize(streams[0]);
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "film readback");
        return rc;
    }

    ~Session()
    {
        for (size_t g = 0; g < sc.size(); g++) {
            cudaSetDevice(devs[g]);
            if (comms[g] && destroy) destroy(comms[g]);
            if (films[g]) cudaFree(films[g]);
            if (streams[g]) cudaStreamDestroy(streams[g]);
        }
        if (scratch) { cudaSetDevice(devs[0]); cudaFree(scratch); }
    }
};

}  // namespace

int lmb200_render_multi(lmb200_scene** scenes, int num_gpus, const lmb200_render_params* p, float* film_host, lmb200_render_stats* stats)
{
    if (!scenes || num_gpus < 1 || !p || !film_host) return set_error(LMB200_E_INVALID, "null argument");
    DeviceGuard guard;
    Session S;
    int rc = S.open(scenes, num_gpus);
    if (!rc) rc = S.pass(p, p->sample_begin, p->sample_end);
    if (!rc) rc = S.gather(p->mode == LMB200_MODE_NORMAL ? 1.0f : (float)S.npx / (float)p->num_samples, film_host);
    if (!rc && stats) { *stats = S.total; stats->reduce_seconds = S.reduce_seconds; }
    return rc;
}

// Time-budgeted / progressive rendering (Scheduler_'s render_time and progress_image_update_interval,
// scheduler.cpp:54-57,108,191-255): passes of `pass_samples` samples (the reference uses grain_size*1000)
// until `render_time` seconds have elapsed (render_time <= 0: until p->num_samples are done); every
// `progress_interval` seconds (> 0) the image so far, rescaled by W*H/processed, is handed to `progress`.
// The final image is rescaled by W*H/processed and stats->samples holds the number processed.
int lmb200_render_timed(lmb200_scene** scenes, int num_gpus, const lmb200_render_params* p, double render_time,
                        int64_t pass_samples, double progress_interval, lmb200_progress_fn progress, void* user,
                        float* film_host, lmb200_render_stats* stats)
{
    if (!scenes || num_gpus < 1 || !p || !film_host) return set_error(LMB200_E_INVALID, "null argument");
    if (p->mode == LMB200_MODE_NORMAL) return lmb200_render_multi(scenes, num_gpus, p, film_host, stats);
    if (pass_samples <= 0) pass_samples = 10000000;
    DeviceGuard guard;
    Session S;
    int rc = S.open(scenes, num_gpus);
    const auto t0 = std::chrono::steady_clock::now();
    auto last_tick = t0;
    int64_t done = 0, cursor = p->sample_begin;
    int64_t ticks = 0;
    // the camera-sample tiling (lmb200_render_params::primary_tile) is resolved ONCE for the job, not per pass: without a time
    // budget from the job's sample count (the passes then reproduce lmb200_render exactly), with one from the pass size
    int job_tile = p->primary_tile;
    if (job_tile == 0 && num_gpus >= 1 && scenes[0]) {
        const Scene* s0 = reinterpret_cast<const Scene*>(scenes[0]);
        job_tile = lmb200_default_primary_tile(s0->dev.width, s0->dev.height, render_time <= 0 ? p->num_samples : pass_samples);
    }
    while (!rc) {
        int64_t n = pass_samples;
        if (render_time <= 0) n = std::min<int64_t>(n, p->sample_end - cursor);
        if (n <= 0) break;
        lmb200_render_params q = *p;
        q.num_samples = n;
        q.primary_tile = job_tile;
        rc = S.pass(&q, cursor, cursor + n);
        if (rc) break;
        cursor += n; done += n;
        const auto now = std::chrono::steady_clock::now();
        if (progress && progress_interval > 0 && std::chrono::duration<double>(now - last_tick).count() > progress_interval) {
            rc = S.gather((float)S.npx / (float)done, film_host);
            if (!rc) { ticks++; if (progress(user, film_host, done, ticks) != 0) break; }
            last_tick = now;
        }
        if (render_time > 0 && std::chrono::duration<double>(now - t0).count() > render_time) break;
    }
    if (!rc) rc = S.gather(done > 0 ? (float)S.npx / (float)done : 1.0f, film_host);
    if (!rc && stats) { *stats = S.total; stats->samples = done; stats->reduce_seconds = S.reduce_seconds; }
    return rc;
}

}  // extern "C"
