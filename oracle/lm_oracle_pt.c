/*
 * TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's unidirectional path tracers
 * renderer::pt and renderer::ptdirect and of the primary-ray normal renderer. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 *
 * It follows the reference's per-sample loop line by line (one sample = one call of the lambda
 * handed to Scheduler::Process) but draws its random numbers from the counter-based generator
 * the CUDA renderer uses (Philox4x32-10 keyed by seed, counter = (sample, block)), so the same
 * sample index produces the same path on both sides up to libm-vs-CUDA ulp differences.
 *
 * Parity status: the estimator has no golden vectors in the reference (no renderer test exists,
 * SURVEY.md §4). It is pinned STATISTICALLY against the reference itself (oracle/_ref running
 * renderer::pt / renderer::ptdirect with dSFMT): tests/test_oracle_pt.py compares mean radiance
 * and per-pixel images within the Monte-Carlo noise floor measured between two reference seeds,
 * and tests/golden/pt_*.npz holds reference images rendered here for machines without
 * /root/reference. Differences by design: exact sqrt-normalisation instead of the reference's
 * 12-bit SSE rsqrt (math.h:1872-1875).
 *
 * Citations are paths under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* from lm_oracle.c */
typedef struct orc_scene orc_scene;
typedef struct { uint32_t k; float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv; uint32_t faceIndex, primIndex; } orc_tri;
orc_scene* orc_scene_create(const float* verts9, uint64_t ntris);
void orc_scene_destroy(orc_scene* s);
int orc_closest_one(const orc_scene* s, const float* ray, int use_bvh, float* tuv, int32_t* tri);
int orc_any_one(const orc_scene* s, const float* ray);

/* Same field layout as include/lmb200.h (restated here; the oracle does not include product headers). */
typedef struct { int32_t type; float R[3], eta[3], k[3], roughness, eta1, eta2; int32_t texR; } orc_bsdf;   /* texR: 0 none, k>0 = textures[k-1] */
typedef struct { int32_t width, height; const float* rgb; } orc_texture;
typedef struct { int32_t bsdf, light; uint32_t first_tri, num_tris; int32_t has_normals; } orc_prim;
typedef struct { float Le[3]; int32_t primitive; int32_t kind; float position[3]; float direction[3]; } orc_light;   /* kind: 0 area, 1 point, 2 directional, 3 env */
typedef struct { float position[3], vx[3], vy[3], vz[3], fov; int32_t width, height; int32_t kind; float lens_radius, focal_distance; } orc_camera;   /* kind: 0 pinhole, 1 thinlens */
typedef struct {
    uint64_t num_tris; const float* verts; const float* normals; const uint32_t* tri_prim;
    uint32_t num_prims; const orc_prim* prims;
    uint32_t num_bsdfs; const orc_bsdf* bsdfs;
    uint32_t num_lights; const orc_light* lights;
    orc_camera camera;
    float sphere_center[3], sphere_radius;   /* Scene3::GetSphereBound (scene3.cpp:56-78), used by directional / env lights */
    const float* uvs;                        /* 6 floats / triangle or NULL */
    uint32_t num_textures; const orc_texture* textures;
} orc_scene_desc;

typedef struct {
    orc_scene_desc d;
    orc_scene* accel;
    float** cdf;       /* per light: num_tris+1 entries (dist.h:37-60) */
    float* inv_area;   /* per light */
    float tan_fov, aspect;
} orc_pt_scene;

#define ORC_PI 3.14159265358979323846f
#define ORC_INV_PI 0.31830988618379067154f
#define ORC_EPS 1e-4f      /* Math::Eps(), math.h:1664 */
#define ORC_EPS_ISECT 1e-4f

typedef struct { float x, y, z; } v3;
static v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 vmul(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static v3 vmulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* v * (1 / |v|), every operation individually rounded: the exact sequence of the CUDA renderer's normalize() */
static v3 vnorm(v3 a) { const float inv = 1.0f / sqrtf(vdot(a, a)); return V(a.x * inv, a.y * inv, a.z * inv); }
static v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static int vblack(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }   /* SPD::Black, spectrum.h */
static v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }

/* ---- counter-based RNG: Philox4x32-10 (Salmon et al. 2011) ---- */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    int r;
    for (r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void rng_block(uint64_t seed, uint64_t sample, uint32_t block, float u[4])
{
    uint32_t o[4]; int i;
    philox4x32((uint32_t)sample, (uint32_t)(sample >> 32), block, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    for (i = 0; i < 4; i++) u[i] = (float)(o[i] >> 8) * (1.0f / 16777216.0f);
}

/* ---- surface geometry (include/lightmetrica/intersectionutils.h:59-135, subset used by the estimators) ---- */
typedef struct { v3 p, gn, sn, dpdu, dpdv; int degenerated; float uvx, uvy; } geom_t;

static void basis(v3 a, v3* b, v3* c)   /* Math::OrthonormalBasis, math.h:2355-2360 */
{
    *c = fabsf(a.x) > fabsf(a.y) ? vnorm(V(a.z, 0.0f, -a.x)) : vnorm(V(0.0f, a.z, -a.y));
    *b = vcross(*c, a);
}
static v3 to_local(const geom_t* g, v3 w) { return V(vdot(g->dpdu, w), vdot(g->dpdv, w), vdot(g->sn, w)); }
static v3 to_world(const geom_t* g, v3 l) { return vadd(vadd(vmul(g->dpdu, l.x), vmul(g->dpdv, l.y)), vmul(g->sn, l.z)); }

static void tri_geom(const orc_pt_scene* S, uint32_t tri, float b0, float b1, v3 p, geom_t* g)
{
    const float* v = S->d.verts + 9 * (size_t)tri;
    const v3 p1 = ld3(v), p2 = ld3(v + 3), p3 = ld3(v + 6);
    const orc_prim* P = &S->d.prims[S->d.tri_prim[tri]];
    g->p = p;
    g->degenerated = 0;
    g->gn = vnorm(vcross(vsub(p2, p1), vsub(p3, p1)));
    if (P->has_normals && S->d.normals) {
        const float* n = S->d.normals + 9 * (size_t)tri;
        g->sn = vnorm(vadd(vadd(vmul(ld3(n), 1.0f - b0 - b1), vmul(ld3(n + 3), b0)), vmul(ld3(n + 6), b1)));
        if (isnan(g->sn.x) || isnan(g->sn.y) || isnan(g->sn.z)) g->sn = g->gn;
    } else g->sn = g->gn;
    basis(g->sn, &g->dpdu, &g->dpdv);
    g->uvx = g->uvy = 0.0f;
    if (S->d.uvs) {   /* intersectionutils.h:107-115 */
        const float* t = S->d.uvs + 6 * (size_t)tri;
        g->uvx = t[0] * (1.0f - b0 - b1) + t[2] * b0 + t[4] * b1;
        g->uvy = t[1] * (1.0f - b0 - b1) + t[3] * b0 + t[5] * b1;
    }
}

/* Texture_Bitmap::Evaluate (src/liblightmetrica/asset/texture/texture_bitmap.cpp:162-168) */
static v3 tex_eval(const orc_texture* T, float u, float v)
{
    int x = (int)((u - floorf(u)) * (float)T->width), y = (int)((v - floorf(v)) * (float)T->height);
    if (x < 0) x = 0;
    if (x > T->width - 1) x = T->width - 1;
    if (y < 0) y = 0;
    if (y > T->height - 1) y = T->height - 1;
    return ld3(T->rgb + 3 * ((size_t)T->width * y + x));
}
/* R of bsdf::diffuse / bsdf::cook_torrance: constant or TexR at geom.uv (bsdf_diffuse.cpp:102, bsdf_cooktorrance.cpp:118) */
static v3 bsdf_R(const orc_pt_scene* S, const void* B_, const geom_t* g);

/* ---- sensor::pinhole (src/liblightmetrica/asset/sensor/sensor_pinhole.cpp) and
 *      sensor::thinlens (src/liblightmetrica/asset/sensor/sensor_thinlens.cpp); p = the sensor vertex (lens point) ---- */
static int sensor_eye(const orc_pt_scene* S, v3 p, v3 wo, v3* e)
{
    const orc_camera* c = &S->d.camera;
    if (c->kind == 1) {                      /* sensor_thinlens.cpp:186-209 */
        const v3 nvz = vneg(ld3(c->vz));
        v3 Pf, woOrig;
        float tf;
        if (vdot(nvz, wo) <= 0.0f) return 0;
        tf = c->focal_distance / vdot(nvz, wo);
        Pf = vadd(p, vmul(wo, tf));          /* intersection with the focal plane */
        woOrig = vnorm(vsub(Pf, ld3(c->position)));   /* direction before refraction */
        *e = V(vdot(ld3(c->vx), woOrig), vdot(ld3(c->vy), woOrig), vdot(ld3(c->vz), woOrig));
    } else {                                 /* sensor_pinhole.cpp:165-175 */
        *e = V(vdot(ld3(c->vx), wo), vdot(ld3(c->vy), wo), vdot(ld3(c->vz), wo));
    }
    return e->z < 0.0f;
}
static int raster_position(const orc_pt_scene* S, v3 p, v3 wo, float* rx, float* ry)   /* pinhole :165-185, thinlens :186-222 */
{
    v3 e;
    if (!sensor_eye(S, p, wo, &e)) return 0;
    *rx = (-e.x / e.z / S->tan_fov / S->aspect + 1.0f) * 0.5f;
    *ry = (-e.y / e.z / S->tan_fov + 1.0f) * 0.5f;
    if (*rx < 0.0f || *rx > 1.0f || *ry < 0.0f || *ry > 1.0f) return 0;
    return 1;
}
static float importance(const orc_pt_scene* S, v3 p, v3 wo)   /* pinhole :137-154, thinlens :150-176 */
{
    float rx, ry;
    v3 e;
    if (!raster_position(S, p, wo, &rx, &ry)) return 0.0f;
    sensor_eye(S, p, wo, &e);
    {
        const float cosT = -e.z, inv = 1.0f / cosT;
        const float A = S->tan_fov * S->tan_fov * S->aspect * 4.0f;
        return inv * inv * inv / A;
    }
}
static v3 camera_dir(const orc_pt_scene* S, float u0, float u1)   /* sensor_pinhole.cpp:79-90 */
{
    const orc_camera* c = &S->d.camera;
    const float x = 2.0f * u0 - 1.0f, y = 2.0f * u1 - 1.0f;
    const v3 e = vnorm(V(S->aspect * S->tan_fov * x, S->tan_fov * y, -1.0f));
    return vadd(vadd(vmul(ld3(c->vx), e.x), vmul(ld3(c->vy), e.y)), vmul(ld3(c->vz), e.z));
}
static void concentric_disk(float u0, float u1, float* sx, float* sy);
/* Sensor::SamplePositionAndDirection: u = raster sample, (l0,l1) = lens sample (ignored by the pinhole) */
static v3 camera_sample(const orc_pt_scene* S, float u0, float u1, float l0, float l1, v3* p)
{
    const orc_camera* c = &S->d.camera;
    const v3 dir = camera_dir(S, u0, u1);
    *p = ld3(c->position);
    if (c->kind == 1) {                      /* sensor_thinlens.cpp:87-106 */
        float lx, ly, tf;
        v3 Pf;
        concentric_disk(l0, l1, &lx, &ly);
        lx *= c->lens_radius; ly *= c->lens_radius;
        *p = vadd(vadd(ld3(c->position), vmul(ld3(c->vx), lx)), vmul(ld3(c->vy), ly));
        tf = c->focal_distance / vdot(vneg(ld3(c->vz)), dir);
        Pf = vadd(ld3(c->position), vmul(dir, tf));
        return vnorm(vsub(Pf, *p));
    }
    return dir;
}

/* ---- BSDFs ---- */
static float snc(const geom_t* g, v3 wi, v3 wo)   /* BSDFUtils::ShadingNormalCorrection, bsdfutils.h:55-66 (EL) */
{
    const float wiNg = vdot(wi, g->gn), woNg = vdot(wo, g->gn);
    const float wiNs = to_local(g, wi).z, woNs = to_local(g, wo).z;
    if (wiNg * wiNs <= 0.0f || woNg * woNs <= 0.0f) return 0.0f;
    return 1.0f;
}
static void concentric_disk(float u0, float u1, float* sx, float* sy)   /* sampler.h:44-60 */
{
    const float vx = 2.0f * u0 - 1.0f, vy = 2.0f * u1 - 1.0f;
    float r, theta;
    if (vx == 0.0f && vy == 0.0f) { *sx = *sy = 0.0f; return; }
    if (vx > -vy) {
        if (vx > vy) { r = vx; theta = (ORC_PI * 0.25f) * vy / vx; }
        else { r = vy; theta = (ORC_PI * 0.25f) * (2.0f - vx / vy); }
    } else {
        if (vx < vy) { r = -vx; theta = (ORC_PI * 0.25f) * (4.0f + vy / vx); }
        else { r = -vy; theta = (ORC_PI * 0.25f) * (6.0f - vx / vy); }
    }
    *sx = r * cosf(theta); *sy = r * sinf(theta);
}
static float ggx_D(float alpha, v3 H)   /* bsdf_cooktorrance.cpp:187-198 */
{
    const float cosH = H.z;
    float tanH, t1, t;
    if (cosH <= 0.0f) return 0.0f;
    {   /* Math::LocalTan (math.h): sqrt(1 - cos^2) / cos, with the sin^2 clamped at 0 */
        const float c2 = cosH * cosH, s2 = 1.0f - c2;
        tanH = s2 <= 0.0f ? 0.0f : sqrtf(s2) / cosH;
    }
    t1 = alpha * alpha;
    t = alpha * alpha + tanH * tanH;
    return t1 / (ORC_PI * cosH * cosH * cosH * cosH * t * t);
}
/* Sample wo (returns 0 if no direction is produced, in which case the reference leaves wo
 * untouched (zero) and the pdf / fs evaluate to 0). */
/* Fresnel term of bsdf::flesnel (bsdf_flesnel.cpp:224-240) */
static float fresnel_term(v3 lwi, float etaI, float etaT)
{
    const float wiDotN = lwi.z, eta = etaI / etaT;
    const float c2 = 1.0f - eta * eta * (1.0f - wiDotN * wiDotN);
    if (c2 <= 0.0f) return 1.0f;
    {
        const float ci = fabsf(wiDotN), ct = sqrtf(c2);
        const float rhoS = (etaI * ci - etaT * ct) / (etaI * ci + etaT * ct);
        const float rhoT = (etaI * ct - etaT * ci) / (etaI * ct + etaT * ci);
        return (rhoS * rhoS + rhoT * rhoT) * 0.5f;
    }
}
static int is_specular(const orc_bsdf* B) { return B->type >= 3 && B->type <= 5; }

static int bsdf_sample(const orc_bsdf* B, const geom_t* g, v3 wi, float u0, float u1, float ucomp, v3* wo)
{
    const v3 lwi = to_local(g, wi);
    if (B->type == 4 || B->type == 5) {   /* refract_all (bsdf_refractall.cpp:59-90), flesnel (bsdf_flesnel.cpp:59-95): both sides */
        float etaI = B->eta1, etaT = B->eta2, eta, c2;
        if (lwi.z < 0.0f) { const float t = etaI; etaI = etaT; etaT = t; }
        eta = etaI / etaT;
        c2 = 1.0f - eta * eta * (1.0f - lwi.z * lwi.z);
        if (B->type == 4 ? (c2 <= 0.0f) : (ucomp <= fresnel_term(lwi, etaI, etaT))) {
            *wo = to_world(g, V(-lwi.x, -lwi.y, lwi.z));            /* BSDFUtils::LocalReflect */
        } else {
            const float ct = sqrtf(c2) * (lwi.z > 0.0f ? -1.0f : 1.0f);
            *wo = to_world(g, V(-eta * lwi.x, -eta * lwi.y, ct));     /* BSDFUtils::LocalRefract */
        }
        return 1;
    }
    if (lwi.z <= 0.0f) return 0;
    if (B->type == 3) { *wo = to_world(g, V(-lwi.x, -lwi.y, lwi.z)); return 1; }   /* bsdf_reflectall.cpp:57-68 */
    if (B->type == 1) {            /* bsdf_diffuse.cpp:69-79 */
        float sx, sy;
        concentric_disk(u0, u1, &sx, &sy);
        *wo = to_world(g, V(sx, sy, sqrtf(fmaxf(0.0f, 1.0f - sx * sx - sy * sy))));
        return 1;
    }
    if (B->type == 2) {            /* bsdf_cooktorrance.cpp:73-89, 200-225 */
        const float a = B->roughness;
        const float v0 = (1.0f - ORC_EPS) * u0 + ORC_EPS;
        const float v1 = (1.0f - 2.0f * ORC_EPS) * u1 + ORC_EPS;
        const float den = sqrtf(1.0f - (1.0f - a * a) * v0);
        const float cosT = sqrtf(1.0f - v0) / den, sinT = a * (sqrtf(v0) / den);
        const float phi = ORC_PI * (2.0f * v1 - 1.0f);
        const v3 H = V(sinT * cosf(phi), sinT * sinf(phi), cosT);
        const v3 nwi = vneg(lwi);
        const v3 lwo = vsub(nwi, vmul(H, 2.0f * vdot(nwi, H)));
        if (lwo.z <= 0.0f) return 0;
        *wo = to_world(g, lwo);
        return 1;
    }
    return 0;
}
static float bsdf_pdf(const orc_bsdf* B, const geom_t* g, v3 wi, v3 wo, int eval_delta)   /* projected solid angle */
{
    const v3 lwi = to_local(g, wi), lwo = to_local(g, wo);
    if (is_specular(B)) {
        if (eval_delta) return 0.0f;
        if (B->type == 3) return (lwi.z <= 0.0f || lwo.z <= 0.0f) ? 0.0f : 1.0f;       /* bsdf_reflectall.cpp:70-85 */
        if (B->type == 4) return 1.0f;                                                  /* bsdf_refractall.cpp:92-100 */
        {                                                                               /* bsdf_flesnel.cpp:97-128 */
            float etaI = B->eta1, etaT = B->eta2, Fr;
            if (lwi.z < 0.0f) { const float t = etaI; etaI = etaT; etaT = t; }
            Fr = fresnel_term(lwi, etaI, etaT);
            return lwi.z * lwo.z >= 0.0f ? Fr : 1.0f - Fr;
        }
    }
    if (lwi.z <= 0.0f || lwo.z <= 0.0f) return 0.0f;
    if (B->type == 1) return ORC_INV_PI;   /* bsdf_diffuse.cpp:81-91 */
    if (B->type == 2) {                    /* bsdf_cooktorrance.cpp:91-103 */
        const v3 H = vnorm(vadd(lwi, lwo));
        const float D = ggx_D(B->roughness, H);
        return D * H.z / (4.0f * vdot(lwo, H)) / lwo.z;
    }
    return 0.0f;
}
static v3 bsdf_R(const orc_pt_scene* S, const void* B_, const geom_t* g)
{
    const orc_bsdf* B = (const orc_bsdf*)B_;
    if (B->texR > 0 && (uint32_t)B->texR <= S->d.num_textures) return tex_eval(&S->d.textures[B->texR - 1], g->uvx, g->uvy);
    return ld3(B->R);
}
static v3 bsdf_eval(const orc_pt_scene* S, const orc_bsdf* B, const geom_t* g, v3 wi, v3 wo, int eval_delta)
{
    const v3 lwi = to_local(g, wi), lwo = to_local(g, wo);
    if (is_specular(B)) {
        float etaI = B->eta1, etaT = B->eta2;
        if (eval_delta) return V(0, 0, 0);
        if (B->type == 3) {                                                            /* bsdf_reflectall.cpp:87-103 */
            if (lwi.z <= 0.0f || lwo.z <= 0.0f) return V(0, 0, 0);
            return vmul(ld3(B->R), snc(g, wi, wo));
        }
        if (lwi.z < 0.0f) { const float t = etaI; etaI = etaT; etaT = t; }
        {
            const float eta = etaI / etaT;
            const float Fr = B->type == 5 ? fresnel_term(lwi, etaI, etaT) : 0.0f;
            if (lwi.z * lwo.z >= 0.0f)      /* reflection (total internal reflection for refract_all) */
                return vmul(ld3(B->R), (B->type == 5 ? Fr : 1.0f) * snc(g, wi, wo));
            /* refraction, EL transport: radiance scaling eta^2 (bsdf_refractall.cpp:120-127, bsdf_flesnel.cpp:150-156) */
            return vmul(ld3(B->R), (B->type == 5 ? 1.0f - Fr : 1.0f) * snc(g, wi, wo) * eta * eta);
        }
    }
    if (lwi.z <= 0.0f || lwo.z <= 0.0f) return V(0, 0, 0);
    if (B->type == 1) {                    /* bsdf_diffuse.cpp:93-104 */
        return vmul(vmul(bsdf_R(S, B, g), ORC_INV_PI), snc(g, wi, wo));
    }
    if (B->type == 2) {                    /* bsdf_cooktorrance.cpp:105-120, 276-293 */
        const v3 H = vnorm(vadd(lwi, lwo));
        const float D = ggx_D(B->roughness, H);
        const float woH = fabsf(vdot(lwo, H));
        /* sic: the reference uses wo.H for both terms (bsdf_cooktorrance.cpp:281-283) */
        const float G = fminf(1.0f, fminf(2.0f * H.z * lwo.z / woH, 2.0f * H.z * lwi.z / woH));
        const float c = vdot(lwi, H);
        float F[3]; int i;
        for (i = 0; i < 3; i++) {
            const float eta = B->eta[i], k = B->k[i];
            const float tmp = (eta * eta + k * k) * (c * c);
            const float rP = (tmp - eta * (2.0f * c) + 1.0f) / (tmp + eta * (2.0f * c) + 1.0f);
            const float tmpF = eta * eta + k * k;
            const float rS = (tmpF - eta * (2.0f * c) + c * c) / (tmpF + eta * (2.0f * c) + c * c);
            F[i] = (rP + rS) * 0.5f;
        }
        {
            const float s = D * G / (4.0f * lwi.z) / lwo.z * snc(g, wi, wo);
            const v3 R = bsdf_R(S, B, g);
            return V(R.x * F[0] * s, R.y * F[1] * s, R.z * F[2] * s);
        }
    }
    return V(0, 0, 0);
}

/* EmitterShape_{DirectionalLight,EnvLight}::Intersect with minT=0, maxT=Inf (light_directional.cpp:52-73,
 * light_env.cpp:56-77; SphereBound::Intersect, bound.h:125-167): the point where the ray leaves the scene's
 * bounding sphere, moved onto the virtual disk perpendicular to d. */
static int emitter_shape_hit(const orc_pt_scene* S, v3 o, v3 d, geom_t* g)
{
    const v3 center = ld3(S->d.sphere_center);
    const float radius = S->d.sphere_radius;
    const v3 oo = vsub(o, center);
    const float a = vdot(d, d), b = 2.0f * vdot(oo, d), c = vdot(oo, oo) - radius * radius;
    const float det = b * b - 4.0f * a * c;
    float e, denom, t0, t1, t;
    v3 p, cc;
    if (det < 0.0f) return 0;
    e = sqrtf(det); denom = 2.0f * a;
    t0 = (-b - e) / denom; t1 = (-b + e) / denom;
    if (t0 > FLT_MAX || t1 < 0.0f) return 0;
    t = t0;
    if (t < 0.0f) { t = t1; if (t > FLT_MAX) return 0; }
    g->degenerated = 0;
    g->gn = vneg(d); g->sn = g->gn;
    basis(g->sn, &g->dpdu, &g->dpdv);
    p = vadd(o, vmul(d, t));
    cc = vadd(center, vmul(d, radius));
    g->p = vadd(vadd(cc, vmul(g->dpdu, vdot(g->dpdu, vsub(p, cc)))), vmul(g->dpdv, vdot(g->dpdv, vsub(p, cc))));
    return 1;
}

/* Light::SamplePositionGivenPreviousPosition + EvaluatePositionGivenPreviousPositionPDF(evalDelta=false) from the
 * vertex at `from`. Returns 0 when no position is produced (the reference's LM_UNREACHABLE branches).
 * light::area: light_area.cpp:67-103, triangleutils.h:71-122; light::point: light_point.cpp:62-66,88-91;
 * light::directional: light_directional.cpp:131-145,176-180; light::env: light_env.cpp:129-146,185-189. */
static int light_sample(const orc_pt_scene* S, int li, v3 from, float u0, float u1, geom_t* g, float* pdfPL)
{
    const orc_light* L = &S->d.lights[li];
    const orc_prim* P = &S->d.prims[L->primitive];
    const float* cdf = S->cdf[li];
    if (L->kind == 1) {            /* light::point */
        memset(g, 0, sizeof(*g));
        g->degenerated = 1;
        g->p = ld3(L->position);
        *pdfPL = 1.0f;
        return 1;
    }
    if (L->kind == 2 || L->kind == 3) {
        v3 d, w;
        float d2, pdfSA = 1.0f;
        if (L->kind == 2) d = vneg(ld3(L->direction));
        else {                     /* Sampler::UniformSampleSphere, sampler.h:79-85 */
            const float z = 1.0f - 2.0f * u0, r = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = 2.0f * ORC_PI * u1;
            d = V(r * cosf(phi), r * sinf(phi), z);
            pdfSA = ORC_INV_PI * 0.25f;
        }
        if (!emitter_shape_hit(S, from, d, g)) return 0;
        /* PDFVal(SolidAngle, pdfSA).ConvertToArea(geomPrev, geom), probability.h:59-71 */
        w = vsub(g->p, from); d2 = vdot(w, w);
        { const float dl = sqrtf(d2); w = V(w.x / dl, w.y / dl, w.z / dl); }
        *pdfPL = pdfSA * fabsf(vdot(g->sn, vneg(w))) / d2;
        return 1;
    }
    {
    const int n = (int)P->num_tris;
    int lo = 0, hi = n + 1, i;
    float u2x, s, bx, by;
    /* upper_bound(cdf, cdf+n+1, u0) - 1, clamped to [0, n-1] (dist.h:70-76) */
    while (lo < hi) { const int mid = (lo + hi) / 2; if (u0 < cdf[mid]) hi = mid; else lo = mid + 1; }
    i = lo - 1; if (i < 0) i = 0; if (i > n - 1) i = n - 1;
    u2x = (u0 - cdf[i]) / (cdf[i + 1] - cdf[i]);
    s = sqrtf(fmaxf(0.0f, u2x)); bx = 1.0f - s; by = u1 * s;    /* sampler.h:97-101 */
    {
        const float* v = S->d.verts + 9 * (size_t)(P->first_tri + (uint32_t)i);
        const v3 p1 = ld3(v), p2 = ld3(v + 3), p3 = ld3(v + 6);
        g->p = vadd(vadd(vmul(p1, 1.0f - bx - by), vmul(p2, bx)), vmul(p3, by));
        g->degenerated = 0;
        g->gn = vnorm(vcross(vsub(p2, p1), vsub(p3, p1)));
        g->sn = g->gn;
        basis(g->sn, &g->dpdu, &g->dpdv);
    }
    *pdfPL = S->inv_area[li];                                    /* light_area.cpp:100-103 */
    return 1;
    }
}

static int visible(const orc_pt_scene* S, v3 p1, v3 p2)   /* Scene3::Visible, scene3.h:107-116 */
{
    const v3 d = vsub(p2, p1);
    const float L = sqrtf(vdot(d, d));
    float ray[8];
    ray[0] = p1.x; ray[1] = p1.y; ray[2] = p1.z; ray[3] = ORC_EPS_ISECT;
    ray[4] = d.x / L; ray[5] = d.y / L; ray[6] = d.z / L; ray[7] = L * (1.0f - ORC_EPS_ISECT);
    return !orc_any_one(S->accel, ray);
}

static void splat(const orc_pt_scene* S, float* film, float rx, float ry, v3 c)   /* film_hdr.cpp:218-223 */
{
    const int W = S->d.camera.width, H = S->d.camera.height;
    int px = (int)(rx * (float)W), py = (int)(ry * (float)H);
    float* f;
    if (px < 0) px = 0;
    if (px > W - 1) px = W - 1;
    if (py < 0) py = 0;
    if (py > H - 1) py = H - 1;
    f = film + 4 * ((size_t)py * W + px);
    f[0] += c.x; f[1] += c.y; f[2] += c.z;
}

/* RenderUtils::GeometryTerm (renderutils.h:46-56) */
static float geometry_term(const geom_t* g1, const geom_t* g2)
{
    v3 d = vsub(g2->p, g1->p);
    const float d2 = vdot(d, d), dl = sqrtf(d2);
    float t = 1.0f;
    d = V(d.x / dl, d.y / dl, d.z / dl);
    if (!g1->degenerated) t *= fabsf(vdot(g1->sn, d));
    if (!g2->degenerated) t *= fabsf(vdot(g2->sn, vneg(d)));
    return t / d2;
}

/* keyed bijection of [0, n): 4-round Feistel network on 2 hb bits with cycle walking (render.cu lmb_permute_tiles) */
static uint32_t hash32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
static uint32_t permute_tiles(uint32_t x, uint32_t n, uint32_t key)
{
    uint32_t hb = 1, mask;
    while ((1u << (2u * hb)) < n) hb++;
    mask = (1u << hb) - 1u;
    do {
        uint32_t L = x >> hb, R = x & mask, r;
        for (r = 0; r < 4u; r++) {
            const uint32_t F = hash32(R ^ key ^ (r * 0x9e3779b9u)) & mask;
            const uint32_t t = L ^ F; L = R; R = t;
        }
        x = (L << hb) | R;
    } while (x >= n);
    return x;
}

/* One sample of renderer::pt (mode 0, renderer_pt.cpp:68-231), renderer::ptdirect (mode 1,
 * renderer_ptdirect.cpp:76-282) or renderer::ptmis (mode 3, renderer_ptmis.cpp:79-290). RNG blocks: 0 = camera (x1,x2 = raster sample); iteration `it`
 * (= numVertices at loop top) uses block 2it-1 = (light pick, light u0, light u1, RR) and block
 * 2it = (bsdf u0, bsdf u1, component, -). */
static void sample_path(const orc_pt_scene* S, int mode, int max_verts, int min_verts, uint64_t seed, uint64_t sample,
                        const float* tile, int gt_nx, int gt_ny, float* film, int64_t* n_extend, int64_t* n_shadow)
{
    float u[4], rx = 0.0f, ry = 0.0f;
    v3 init_wo, thr = V(1, 1, 1), wi = V(0, 0, 0);
    geom_t geom;
    int is_sensor = 1, num_verts = 1;
    const orc_bsdf* bsdf = NULL;
    float ul[4] = {0.5f, 0.5f, 0.0f, 0.0f};
    rng_block(seed, sample, 0u, u);
    if (S->d.camera.kind == 1) rng_block(seed, sample, 0xffffffffu, ul);   /* lens sample (the second Next2D of renderer_pt.cpp:86) */
    memset(&geom, 0, sizeof(geom));
    geom.degenerated = 1;
    /* coherent camera sample groups (render.cu camera_raster): the 32 samples of group sample / 32 share one of the
     * gt_nx x gt_ny tiles; consecutive groups walk through all tiles in a keyed pseudo-random order per round */
    if (gt_nx > 0) {
        const uint32_t T = (uint32_t)gt_nx * (uint32_t)gt_ny;
        const uint64_t g = sample >> 5, round = g / T;
        const uint32_t key = hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) ^ hash32((uint32_t)round ^ hash32((uint32_t)(round >> 32)))));
        const uint32_t t = permute_tiles((uint32_t)(g - round * T), T, key);
        const uint32_t tx = t % (uint32_t)gt_nx, ty = t / (uint32_t)gt_nx;
        u[1] = ((float)tx + u[1]) / (float)gt_nx;
        u[2] = ((float)ty + u[2]) / (float)gt_ny;
    }
    /* optional tile partitioning: the raster sample is drawn inside tile = {x0, y0, x1, y1} (whole image: 0,0,1,1) */
    init_wo = camera_sample(S, tile[0] + u[1] * (tile[2] - tile[0]), tile[1] + u[2] * (tile[3] - tile[1]), ul[0], ul[1], &geom.p);
    if (mode != 1 && !raster_position(S, geom.p, init_wo, &rx, &ry)) return;      /* renderer_pt.cpp:94-99, renderer_ptmis.cpp:100-105 */

    for (;;) {
        float ua[4], ub[4];
        v3 wo, fs;
        float pdfD;
        if (max_verts != -1 && num_verts >= max_verts) break;
        rng_block(seed, sample, (uint32_t)(2 * num_verts - 1), ua);
        rng_block(seed, sample, (uint32_t)(2 * num_verts), ub);

        if ((mode == 1 || (mode == 3 && num_verts + 1 >= min_verts)) && S->d.num_lights > 0) {   /* direct light sampling, renderer_ptdirect.cpp:123-177, renderer_ptmis.cpp:130-205 */
            const int nL = (int)S->d.num_lights;
            int li = (int)(ua[0] * (float)nL);
            geom_t gL;
            v3 ppL, fsE, fsL, C, d;
            float pdfL, pdfPL, G, d2, dl;
            if (li < 0) li = 0;                                                      /* scene3.cpp:508-513 */
            if (li > nL - 1) li = nL - 1;
            pdfL = 1.0f / (float)nL;                                                 /* scene3.cpp:526-530 */
            if (!light_sample(S, li, geom.p, ua[1], ua[2], &gL, &pdfPL)) goto nee_done;
            ppL = vnorm(vsub(gL.p, geom.p));
            if (is_sensor) { const float im = importance(S, geom.p, ppL); fsE = V(im, im, im); }
            else fsE = bsdf_eval(S, bsdf, &geom, wi, ppL, 1);
            if (S->d.lights[li].kind != 0) fsL = ld3(S->d.lights[li].Le);   /* light_point.cpp:95-98, light_directional.cpp:182-185, light_env.cpp:191-208 (constant Le) */
            else fsL = to_local(&gL, vneg(ppL)).z <= 0.0f ? V(0, 0, 0) : ld3(S->d.lights[li].Le);   /* light_area.cpp:105-110 */
            d = vsub(gL.p, geom.p); d2 = vdot(d, d); dl = sqrtf(d2); d = V(d.x / dl, d.y / dl, d.z / dl);   /* renderutils.h:46-56 */
            G = 1.0f;
            if (!geom.degenerated) G *= fabsf(vdot(geom.sn, d));
            if (!gL.degenerated) G *= fabsf(vdot(gL.sn, vneg(d)));
            G = G / d2;
            C = vmul(vmulv(vmulv(thr, fsE), fsL), G);
            if (!vblack(C)) {
                /* V is evaluated unconditionally by the reference; skipping it for black C does not change the film */
                (*n_shadow)++;
                if (visible(S, geom.p, gL.p)) {
                    float prx = rx, pry = ry;
                    C = vmul(C, 1.0f / pdfL / pdfPL);
                    if (mode == 3) {   /* MIS weight, renderer_ptmis.cpp:163-170 */
                        const float pdfDL = pdfPL / geometry_term(&geom, &gL) * pdfL;
                        const float pdfB = is_sensor ? importance(S, geom.p, ppL) : bsdf_pdf(bsdf, &geom, wi, ppL, 1);
                        C = vmul(C, pdfDL / (pdfDL + pdfB));
                    }
                    if (is_sensor) raster_position(S, geom.p, ppL, &prx, &pry);      /* renderer_ptdirect.cpp:165-170 */
                    splat(S, film, prx, pry, C);
                }
            }
        }
nee_done:

        if (is_sensor) wo = init_wo;
        else { wo = V(0, 0, 0); bsdf_sample(bsdf, &geom, wi, ub[0], ub[1], ub[2], &wo); }
        pdfD = is_sensor ? importance(S, geom.p, wo) : bsdf_pdf(bsdf, &geom, wi, wo, 0);
        if (mode == 1 && is_sensor) { if (!raster_position(S, geom.p, wo, &rx, &ry)) break; }   /* renderer_ptdirect.cpp:200-208 */
        if (is_sensor) { const float im = importance(S, geom.p, wo); fs = V(im, im, im); }
        else fs = bsdf_eval(S, bsdf, &geom, wi, wo, 0);
        if (vblack(fs)) break;
        thr = vmulv(thr, V(fs.x / pdfD, fs.y / pdfD, fs.z / pdfD));

        {   /* intersection, scene3.cpp:458-478 */
            float ray[8], tuv[3]; int32_t tri;
            const orc_prim* P;
            geom_t prev = geom;
            v3 hp;
            ray[0] = geom.p.x; ray[1] = geom.p.y; ray[2] = geom.p.z; ray[3] = ORC_EPS_ISECT;
            ray[4] = wo.x; ray[5] = wo.y; ray[6] = wo.z; ray[7] = FLT_MAX;
            (*n_extend)++;
            orc_closest_one(S->accel, ray, 1, tuv, &tri);
            if (tri < 0) break;
            hp = vadd(geom.p, vmul(wo, tuv[0]));
            P = &S->d.prims[S->d.tri_prim[tri]];
            tri_geom(S, (uint32_t)tri, tuv[1], tuv[2], hp, &geom);
            if (mode != 1 && P->light >= 0 && num_verts + 1 >= min_verts) {          /* renderer_pt.cpp:183-194 */
                if (to_local(&geom, vneg(wo)).z > 0.0f) {
                    v3 C = vmulv(thr, ld3(S->d.lights[P->light].Le));
                    if (mode == 3) {   /* renderer_ptmis.cpp:247-260: balance heuristic against the light-sampling pdf */
                        /* (type & S) > 0 ? 0 : ... : a specular previous vertex cannot be reached by light sampling */
                        const float pdfDL = (!is_sensor && is_specular(bsdf)) ? 0.0f
                            : S->inv_area[P->light] / geometry_term(&geom, &prev) * (1.0f / (float)S->d.num_lights);
                        C = vmul(C, pdfD / (pdfD + pdfDL));
                    }
                    splat(S, film, rx, ry, C);
                }
            }
            if (ua[3] > 0.5f) break;                                                 /* renderer_pt.cpp:207-215 */
            thr = V(thr.x / 0.5f, thr.y / 0.5f, thr.z / 0.5f);
            bsdf = &S->d.bsdfs[P->bsdf];
            is_sensor = 0;
            wi = vneg(wo);
            num_verts++;
        }
    }
}

orc_pt_scene* orc_pt_scene_create(const orc_scene_desc* d)
{
    orc_pt_scene* S = (orc_pt_scene*)calloc(1, sizeof(orc_pt_scene));
    uint32_t li;
    S->d = *d;
    S->accel = orc_scene_create(d->verts, d->num_tris);
    S->tan_fov = tanf(d->camera.fov * 0.5f);
    S->aspect = (float)d->camera.width / (float)d->camera.height;
    S->cdf = (float**)calloc(d->num_lights ? d->num_lights : 1, sizeof(float*));
    S->inv_area = (float*)calloc(d->num_lights ? d->num_lights : 1, sizeof(float));
    for (li = 0; li < d->num_lights; li++) {     /* TriangleUtils::CreateTriangleAreaDist, triangleutils.h:47-68 */
        const orc_prim* P = &d->prims[d->lights[li].primitive];
        float* cdf;
        if (d->lights[li].kind != 0) { S->cdf[li] = NULL; S->inv_area[li] = 1.0f; continue; }   /* only light::area samples a mesh */
        cdf = (float*)malloc(sizeof(float) * (P->num_tris + 1));
        float sum = 0.0f, inv;
        uint32_t i;
        cdf[0] = 0.0f;
        for (i = 0; i < P->num_tris; i++) {
            const float* v = d->verts + 9 * (size_t)(P->first_tri + i);
            const v3 c = vcross(vsub(ld3(v + 3), ld3(v)), vsub(ld3(v + 6), ld3(v)));
            const float area = sqrtf((c.x * c.x + c.y * c.y) + (c.z * c.z + 0.0f)) * 0.5f;   /* _mm_dp_ps order */
            cdf[i + 1] = cdf[i] + area;
            sum += area;
        }
        inv = 1.0f / cdf[P->num_tris];
        for (i = 0; i <= P->num_tris; i++) cdf[i] *= inv;
        S->cdf[li] = cdf;
        S->inv_area[li] = 1.0f / sum;
    }
    return S;
}

void orc_pt_scene_destroy(orc_pt_scene* S)
{
    uint32_t li;
    if (!S) return;
    for (li = 0; li < S->d.num_lights; li++) free(S->cdf[li]);
    free(S->cdf); free(S->inv_area);
    orc_scene_destroy(S->accel);
    free(S);
}

/* Renders global sample indices [begin,end) of a num_samples job into film (W*H*4 floats, zeroed by
 * the caller), UNSCALED; threads accumulate into private films that are summed at the end, as
 * Scheduler_::Process does (scheduler.cpp:157-164, 280-285). counts[0]=extend rays, [1]=shadow rays. */
void orc_pt_render_ex(const orc_pt_scene* S, int mode, int max_verts, int min_verts, uint64_t seed,
                      int64_t begin, int64_t end, const float* tile, int primary_tile, float* film, int64_t* counts);
void orc_pt_render(const orc_pt_scene* S, int mode, int max_verts, int min_verts, uint64_t seed,
                   int64_t begin, int64_t end, float* film, int64_t* counts)
{
    const float whole[4] = {0.0f, 0.0f, 1.0f, 1.0f};
    orc_pt_render_ex(S, mode, max_verts, min_verts, seed, begin, end, whole, 0, film, counts);
}
/* Same with the camera samples drawn inside the raster rectangle tile = {x0, y0, x1, y1} (include/lmb200.h, tile partitioning). */
void orc_pt_render_tile(const orc_pt_scene* S, int mode, int max_verts, int min_verts, uint64_t seed,
                        int64_t begin, int64_t end, const float* tile, float* film, int64_t* counts)
{
    orc_pt_render_ex(S, mode, max_verts, min_verts, seed, begin, end, tile, 0, film, counts);
}
/* primary_tile: the RESOLVED lmb200_render_params::primary_tile (tile edge in pixels; <= 0 = off; 0 is resolved by the caller) */
void orc_pt_render_ex(const orc_pt_scene* S, int mode, int max_verts, int min_verts, uint64_t seed,
                      int64_t begin, int64_t end, const float* tile, int primary_tile, float* film, int64_t* counts)
{
    const size_t npx = (size_t)S->d.camera.width * S->d.camera.height;
    const int tp = primary_tile;
    const int gt_nx = tp > 0 ? (S->d.camera.width + tp - 1) / tp : 0, gt_ny = tp > 0 ? (S->d.camera.height + tp - 1) / tp : 0;
    int64_t ne = 0, ns = 0;
#pragma omp parallel reduction(+ : ne, ns)
    {
        float* local = (float*)calloc(npx * 4, sizeof(float));
        int64_t i;
        size_t k;
#pragma omp for schedule(dynamic, 4096)
        for (i = begin; i < end; i++) sample_path(S, mode, max_verts, min_verts, seed, (uint64_t)i, tile, gt_nx, gt_ny, local, &ne, &ns);
#pragma omp critical
        for (k = 0; k < npx * 4; k++) film[k] += local[k];
        free(local);
    }
    if (counts) { counts[0] = ne; counts[1] = ns; }
}

/* Primary-ray normal renderer: pixel-centre rays as renderer_raycast.cpp:72-105 generates them,
 * shaded abs(sn) as plugin/renderer_normal/renderer_normal.cpp:62-75 intends. film: W*H*4. */
void orc_render_normal(const orc_pt_scene* S, float* film, int32_t* tri_out)
{
    const int W = S->d.camera.width, H = S->d.camera.height;
    int y;
#pragma omp parallel for schedule(dynamic, 4)
    for (y = 0; y < H; y++) {
        int x;
        for (x = 0; x < W; x++) {
            v3 cp;   /* thin lens: the lens centre (u2 = (.5,.5)), i.e. the in-focus pinhole image */
            const v3 wo = camera_sample(S, ((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H, 0.5f, 0.5f, &cp);
            float ray[8], tuv[3]; int32_t tri;
            float* f = film + 4 * ((size_t)y * W + x);
            ray[0] = cp.x; ray[1] = cp.y; ray[2] = cp.z; ray[3] = ORC_EPS_ISECT;
            ray[4] = wo.x; ray[5] = wo.y; ray[6] = wo.z; ray[7] = FLT_MAX;
            orc_closest_one(S->accel, ray, 1, tuv, &tri);
            if (tri_out) tri_out[(size_t)y * W + x] = tri;
            if (tri < 0) { f[0] = f[1] = f[2] = 0.0f; continue; }
            {
                geom_t g;
                tri_geom(S, (uint32_t)tri, tuv[1], tuv[2], V(0, 0, 0), &g);
                f[0] = fabsf(g.sn.x); f[1] = fabsf(g.sn.y); f[2] = fabsf(g.sn.z);
            }
        }
    }
}
