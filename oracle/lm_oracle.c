/*
 * TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's traversal hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; it is never
 * shipped and never on the product path.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function below bit-for-bit against
 * the reference itself compiled from /root/reference (oracle/_ref, see oracle/ref/build_ref.sh)
 * and against the reference's own known-answer tests Accel3Test.Simple/Simple2
 * (src/lightmetrica-test/test_accel3.cpp:272-345); the resulting vectors are committed under
 * tests/golden/ so the pin also holds on machines without /root/reference.
 *
 * Build: gcc -O2 -msse4.2 -ffp-contract=off -fopenmp -fPIC -shared (no FMA contraction: every
 * operation is an individually rounded IEEE single op, exactly like the oracle/_ref build).
 * Citations are paths under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef struct {            /* include/lightmetrica/triaccel.h:32-48 (48 bytes) */
    uint32_t k;
    float n_u, n_v, n_d;
    float a_u, a_v, b_nu, b_nv;
    float c_nu, c_nv;
    uint32_t faceIndex, primIndex;
} orc_tri;

/* include/lightmetrica/triaccel.h:50-91. Cross and Dot follow the SSE specialisations
 * (math.h:1791-1802: separately rounded products; math.h:1728-1731: _mm_dp_ps mask 0x71 sums
 * (x*x' + y*y') + (z*z' + 0)). */
int orc_triaccel_load(orc_tri* r, const float* A, const float* B, const float* C)
{
    static const int waldModulo[4] = {1, 2, 0, 1};
    float b[3], c[3], N[3];
    int j;
    uint32_t k = 0;
    for (j = 0; j < 3; j++) { b[j] = C[j] - A[j]; c[j] = B[j] - A[j]; }
    N[0] = c[1] * b[2] - c[2] * b[1];
    N[1] = c[2] * b[0] - c[0] * b[2];
    N[2] = c[0] * b[1] - c[1] * b[0];
    for (j = 0; j < 3; j++) if (fabsf(N[j]) > fabsf(N[k])) k = (uint32_t)j;
    {
        const int u = waldModulo[k], v = waldModulo[k + 1];
        const float n_k = N[k];
        const float denom = b[u] * c[v] - b[v] * c[u];
        if (denom == 0) { r->k = 3; return 1; }
        r->k = k;
        r->n_u = N[u] / n_k;
        r->n_v = N[v] / n_k;
        r->n_d = ((A[0] * N[0] + A[1] * N[1]) + (A[2] * N[2] + 0.0f)) / n_k;
        r->b_nu = b[u] / denom;
        r->b_nv = -b[v] / denom;
        r->a_u = A[u];
        r->a_v = A[v];
        r->c_nu = c[v] / denom;
        r->c_nv = -c[u] / denom;
    }
    return 0;
}

/* include/lightmetrica/triaccel.h:93-151 */
int orc_triaccel_intersect(const orc_tri* r, const float* o, const float* d, float mint, float maxt, float* u, float* v, float* t)
{
    float o_u, o_v, o_k, d_u, d_v, d_k, demon, hu, hv;
    switch (r->k) {
        case 0: o_u = o[1]; o_v = o[2]; o_k = o[0]; d_u = d[1]; d_v = d[2]; d_k = d[0]; break;
        case 1: o_u = o[2]; o_v = o[0]; o_k = o[1]; d_u = d[2]; d_v = d[0]; d_k = d[1]; break;
        case 2: o_u = o[0]; o_v = o[1]; o_k = o[2]; d_u = d[0]; d_v = d[1]; d_k = d[2]; break;
        default: return 0;
    }
    demon = d_u * r->n_u + d_v * r->n_v + d_k;
    if (demon == 0) return 0;
    *t = (r->n_d - o_u * r->n_u - o_v * r->n_v - o_k) / demon;
    if (*t < mint || *t > maxt) return 0;
    hu = o_u + *t * d_u - r->a_u;
    hv = o_v + *t * d_v - r->a_v;
    *u = hv * r->b_nu + hu * r->b_nv;
    *v = hu * r->c_nu + hv * r->c_nv;
    return *u >= 0.0f && *v >= 0.0f && *u + *v <= 1.0f;
}

/* ---------------------------------------------------------------------------------------------
 * Scene = flat triangle list in the reference accels' order (primitive-major, face-minor,
 * src/liblightmetrica/accel/accel_qbvh.cpp:161-194) + a binary BVH used only to make the scan
 * fast; node culling follows Bound::Intersect (include/lightmetrica/bound.h:66-91) on boxes padded
 * by Math::Eps()=1e-4 (accel_qbvh.cpp:189-190). The closest hit is defined by the triangle test
 * and the tie rule alone (a later triangle with t == maxT replaces the earlier, triaccel.h:137),
 * so any conservative tree gives accel::naive's answer up to exact ties; exact ties are resolved
 * like accel::naive's linear scan (accel_naive.cpp:92-124): the larger index wins. */

typedef struct { float lo[3], hi[3]; int32_t left, right, first, count; } orc_node;

typedef struct {
    uint64_t n;
    orc_tri* tris;        /* input order */
    float* lo; float* hi; /* padded per-triangle boxes */
    uint32_t* order;      /* leaf order -> input index */
    orc_node* nodes; int32_t num_nodes, cap_nodes;
} orc_scene;

static int32_t orc_build_rec(orc_scene* s, int32_t begin, int32_t end)
{
    int32_t ni = s->num_nodes++, i, a, axis = 0;
    orc_node* nd = &s->nodes[ni];
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (a = 0; a < 3; a++) { nd->lo[a] = FLT_MAX; nd->hi[a] = -FLT_MAX; }
    for (i = begin; i < end; i++) {
        const uint32_t id = s->order[i];
        for (a = 0; a < 3; a++) {
            const float l = s->lo[3 * id + a], h = s->hi[3 * id + a], c = 0.5f * (l + h);
            if (l < nd->lo[a]) nd->lo[a] = l;
            if (h > nd->hi[a]) nd->hi[a] = h;
            if (c < clo[a]) clo[a] = c;
            if (c > chi[a]) chi[a] = c;
        }
    }
    nd->left = nd->right = -1; nd->first = begin; nd->count = end - begin;
    if (end - begin <= 4) return ni;
    for (a = 1; a < 3; a++) if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
    {
        const float split = 0.5f * (clo[axis] + chi[axis]);
        int32_t l = begin, r = end - 1, mid;
        while (l <= r) {
            const uint32_t id = s->order[l];
            const float c = 0.5f * (s->lo[3 * id + axis] + s->hi[3 * id + axis]);
            if (c < split) l++;
            else { const uint32_t tmp = s->order[l]; s->order[l] = s->order[r]; s->order[r] = tmp; r--; }
        }
        mid = l;
        if (mid == begin || mid == end) mid = (begin + end) / 2;
        nd->count = 0;
        {
            const int32_t lc = orc_build_rec(s, begin, mid);
            const int32_t rc = orc_build_rec(s, mid, end);
            s->nodes[ni].left = lc; s->nodes[ni].right = rc;
        }
    }
    return ni;
}

orc_scene* orc_scene_create(const float* verts9, uint64_t ntris)
{
    orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
    uint64_t i; int a, k;
    s->n = ntris;
    s->tris = (orc_tri*)calloc(ntris ? ntris : 1, sizeof(orc_tri));
    s->lo = (float*)malloc(sizeof(float) * 3 * (ntris ? ntris : 1));
    s->hi = (float*)malloc(sizeof(float) * 3 * (ntris ? ntris : 1));
    s->order = (uint32_t*)malloc(sizeof(uint32_t) * (ntris ? ntris : 1));
    s->cap_nodes = (int32_t)(2 * ntris + 2);
    s->nodes = (orc_node*)malloc(sizeof(orc_node) * (size_t)s->cap_nodes);
    for (i = 0; i < ntris; i++) {
        const float* v = verts9 + 9 * i;
        orc_triaccel_load(&s->tris[i], v, v + 3, v + 6);
        s->tris[i].faceIndex = (uint32_t)i; s->tris[i].primIndex = 0;
        for (a = 0; a < 3; a++) {
            float l = v[a], h = v[a];
            for (k = 1; k < 3; k++) { if (v[3 * k + a] < l) l = v[3 * k + a]; if (v[3 * k + a] > h) h = v[3 * k + a]; }
            s->lo[3 * i + a] = l - 1e-4f; s->hi[3 * i + a] = h + 1e-4f;
        }
        s->order[i] = (uint32_t)i;
    }
    s->num_nodes = 0;
    if (ntris) orc_build_rec(s, 0, (int32_t)ntris);
    return s;
}

void orc_scene_destroy(orc_scene* s)
{
    if (!s) return;
    free(s->tris); free(s->lo); free(s->hi); free(s->order); free(s->nodes); free(s);
}

const orc_tri* orc_scene_tris(const orc_scene* s) { return s->tris; }

/* include/lightmetrica/bound.h:66-91 */
static int orc_bound_intersect(const orc_node* b, const float* o, const float* d, float tMin, float tMax)
{
    float tmin, tmax;
    int a;
    {
        const int neg = d[0] < 0.0f;
        const float inv_max = d[0] == 0 ? FLT_MAX : 1.0f / d[0], inv_min = d[0] == 0 ? 0.0f : 1.0f / d[0];
        tmax = ((neg ? b->lo[0] : b->hi[0]) - o[0]) * inv_max;
        tmin = ((neg ? b->hi[0] : b->lo[0]) - o[0]) * inv_min;
    }
    for (a = 1; a < 3; a++) {
        const int neg = d[a] < 0.0f;
        const float inv_max = d[a] == 0 ? FLT_MAX : 1.0f / d[a], inv_min = d[a] == 0 ? 0.0f : 1.0f / d[a];
        const float tamax = ((neg ? b->lo[a] : b->hi[a]) - o[a]) * inv_max;
        const float tamin = ((neg ? b->hi[a] : b->lo[a]) - o[a]) * inv_min;
        if (tmin > tamax || tamin > tmax) return 0;
        if (tamin > tmin) tmin = tamin;
        if (tamax < tmax) tmax = tamax;
    }
    return tmin < tMax && tmax > tMin;
}

/* One ray: 8 floats (ox,oy,oz,tmin,dx,dy,dz,tmax) -> t,u,v and the triangle index (-1 = miss).
 * use_bvh = 0: accel::naive's linear scan (accel_naive.cpp:92-124). */
int orc_closest_one(const orc_scene* s, const float* r, int use_bvh, float* tuv, int32_t* tri)
{
    const float* o = r; const float* d = r + 4;
    const float mint = r[3];
    float maxt = r[7], bu = 0, bv = 0;
    int32_t best = -1;
    if (!use_bvh) {
        uint64_t j;
        for (j = 0; j < s->n; j++) {
            float t, u, v;
            if (orc_triaccel_intersect(&s->tris[j], o, d, mint, maxt, &u, &v, &t)) { maxt = t; bu = u; bv = v; best = (int32_t)j; }
        }
    } else if (s->n) {
        int32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const orc_node* nd = &s->nodes[stack[--sp]];
            /* Bound::Intersect compares strictly against [tMin,tMax]; the boxes are padded by 1e-4,
             * so a triangle hit at exactly maxt still lies strictly inside its box's range. */
            if (!orc_bound_intersect(nd, o, d, mint, maxt)) continue;
            if (nd->count) {
                int32_t k;
                for (k = nd->first; k < nd->first + nd->count; k++) {
                    const uint32_t id = s->order[k];
                    float t, u, v;
                    if (orc_triaccel_intersect(&s->tris[id], o, d, mint, maxt, &u, &v, &t)) {
                        if (t < maxt || best < 0 || (int32_t)id > best) { maxt = t; bu = u; bv = v; best = (int32_t)id; }
                    }
                }
            } else { stack[sp++] = nd->left; stack[sp++] = nd->right; }
        }
    }
    if (best >= 0) { tuv[0] = maxt; tuv[1] = bu; tuv[2] = bv; }
    else { tuv[0] = tuv[1] = tuv[2] = 0.0f; }
    *tri = best;
    return best >= 0;
}

/* Batch form; returns the hit count. */
long long orc_closest(const orc_scene* s, const float* rays, long long n, int use_bvh, float* tuv, int32_t* tri)
{
    long long hits = 0, i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : hits)
    for (i = 0; i < n; i++) hits += orc_closest_one(s, rays + 8 * i, use_bvh, tuv + 3 * i, tri + i);
    return hits;
}

/* Scene3::Visible's query (include/lightmetrica/scene3.h:107-116): any triangle within [tmin,tmax]. */
int orc_any_one(const orc_scene* s, const float* r)
{
    int found = 0;
    if (s->n) {
        int32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp && !found) {
            const orc_node* nd = &s->nodes[stack[--sp]];
            if (!orc_bound_intersect(nd, r, r + 4, r[3], r[7])) continue;
            if (nd->count) {
                int32_t k;
                for (k = nd->first; k < nd->first + nd->count && !found; k++) {
                    float t, u, v;
                    if (orc_triaccel_intersect(&s->tris[s->order[k]], r, r + 4, r[3], r[7], &u, &v, &t)) found = 1;
                }
            } else { stack[sp++] = nd->left; stack[sp++] = nd->right; }
        }
    }
    return found;
}

long long orc_any(const orc_scene* s, const float* rays, long long n, uint8_t* occluded)
{
    long long hits = 0, i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : hits)
    for (i = 0; i < n; i++) { const int f = orc_any_one(s, rays + 8 * i); occluded[i] = (uint8_t)f; hits += f; }
    return hits;
}

/* ---------------------------------------------------------------------------------------------
 * Checker for the PRODUCT's flattened structure (lightmetrica-v2_b200/csrc/bvh.h): a scalar walk
 * of the 64-byte units with exact box decoding in double precision. It lets the CPU test-suite
 * verify the builders (every triangle reachable, boxes conservative, slot order encoding)
 * without a GPU. It visits ALL children whose decoded box the ray touches, in arbitrary order. */
typedef struct {
    uint16_t k[3]; uint16_t counts; uint8_t e[3]; uint8_t imask; uint32_t base;
    uint8_t qlo[3][8]; uint8_t qhi[3][8];
} orc_node64;
typedef struct { orc_tri rec; uint32_t pad[4]; } orc_triunit;

static uint32_t orc_slot_tri_offset(const orc_node64* nd, int s)
{
    uint32_t off = 0; int j;
    for (j = 0; j < s; j++) off += (nd->counts >> (2 * j)) & 3u;
    return off;
}
static uint32_t orc_popc8(uint32_t x) { uint32_t c = 0; for (; x; x &= x - 1) c++; return c; }

static int orc_slot_hit(const orc_node64* nd, const float* grid, int s, const float* o, const float* d, double mint, double maxt, double* tnear)
{
    double tn = mint, tf = maxt;
    int a;
    for (a = 0; a < 3; a++) {
        const double org = (double)grid[a] + (double)grid[3 + a] * nd->k[a];
        const double sc = ldexp(1.0, (int)nd->e[a] - 127);
        const double lo = org + sc * nd->qlo[a][s], hi = org + sc * nd->qhi[a][s];
        if (nd->qlo[a][s] > nd->qhi[a][s]) return 0;       /* empty slot */
        if (d[a] == 0.0f) { if (o[a] < lo || o[a] > hi) return 0; }
        else {
            double t0 = (lo - o[a]) / d[a], t1 = (hi - o[a]) / d[a];
            if (t0 > t1) { const double tt = t0; t0 = t1; t1 = tt; }
            if (t0 > tn) tn = t0;
            if (t1 < tf) tf = t1;
        }
    }
    if (tnear) *tnear = tn;
    return tn <= tf;
}

/* grid: lo[3], step[3] of the scene grid (lmb200_bvh_layout) */
long long orc_wide_closest(const void* units64, const float* grid, const float* rays, long long n, float* tuv, int32_t* tri)
{
    const orc_node64* units = (const orc_node64*)units64;
    long long hits = 0, i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : hits)
    for (i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        const float* o = r; const float* d = r + 4;
        const float mint = r[3];
        float maxt = r[7], bu = 0, bv = 0;
        int32_t best = -1;
        uint32_t stack[512]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const orc_node64* nd = &units[stack[--sp]];
            const uint32_t nint = orc_popc8(nd->imask);
            int s;
            uint32_t rel = 0;
            for (s = 0; s < 8; s++) {
                const int internal = (nd->imask >> s) & 1;
                const uint32_t cnt = (nd->counts >> (2 * s)) & 3u;
                const int hit = orc_slot_hit(nd, grid, s, o, d, mint, maxt, 0);
                if (internal) { const uint32_t child = nd->base + rel; rel++; if (hit && sp < 512) stack[sp++] = child; }
                else if (cnt && hit) {
                    const uint32_t off = orc_slot_tri_offset(nd, s);
                    uint32_t k;
                    for (k = 0; k < cnt; k++) {
                        const orc_tri* T = &((const orc_triunit*)&units[nd->base + nint + off + k])->rec;
                        float t, u, v;
                        if (orc_triaccel_intersect(T, o, d, mint, maxt, &u, &v, &t)) {
                            const int32_t id = (int32_t)T->faceIndex;
                            if (t < maxt || best < 0 || id > best) { maxt = t; bu = u; bv = v; best = id; }
                        }
                    }
                }
            }
        }
        if (best >= 0) { hits++; tuv[3 * i] = maxt; tuv[3 * i + 1] = bu; tuv[3 * i + 2] = bv; }
        else { tuv[3 * i] = tuv[3 * i + 1] = tuv[3 * i + 2] = 0.0f; }
        tri[i] = best;
    }
    return hits;
}

/* Work estimator for tree-quality experiments on the CPU (no GPU needed): walks the flattened structure in the
 * PRODUCT's traversal order — children of a node in octant priority (slot ^ (7 - ray octant), highest first), the
 * remaining siblings pushed as one stack entry, triangles tested as soon as their leaf slot is hit — and counts the
 * 64-byte nodes and triangle records fetched. The device kernel defers triangle tests by a few steps, so its counts
 * are slightly higher; ratios between trees carry over. out[0] = nodes, out[1] = triangle records (totals). */
void orc_wide_count(const void* units64, const float* grid, const float* rays, long long n, double* out)
{
    const orc_node64* units = (const orc_node64*)units64;
    double tn_nodes = 0, tn_tris = 0;
    long long i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : tn_nodes, tn_tris)
    for (i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        const float* o = r; const float* d = r + 4;
        const float mint = r[3];
        float maxt = r[7];
        int32_t best = -1;
        const uint32_t oct = (d[0] < 0 ? 1u : 0u) | (d[1] < 0 ? 2u : 0u) | (d[2] < 0 ? 4u : 0u);
        const uint32_t oi = 7u - oct;
        /* stack entry: base unit of a sibling group, pending priority bits, imask */
        struct { uint32_t base, pend, imask; } stack[256];
        int sp = 0;
        uint32_t g_base = 0, g_pend = 0x80u, g_imask = 0;       /* root: relative index 0 */
        for (;;) {
            if (!g_pend) { if (!sp) break; sp--; g_base = stack[sp].base; g_pend = stack[sp].pend; g_imask = stack[sp].imask; continue; }
            {
                uint32_t bit = 7; const orc_node64* nd; uint32_t slot, rel, nint, hits8 = 0, pr = 0; int s;
                while (!((g_pend >> bit) & 1u)) bit--;
                g_pend &= ~(1u << bit);
                if (g_pend && sp < 256) { stack[sp].base = g_base; stack[sp].pend = g_pend; stack[sp].imask = g_imask; sp++; }
                slot = bit ^ oi;
                rel = orc_popc8(g_imask & ((1u << slot) - 1u));
                nd = &units[g_base + rel];
                tn_nodes += 1;
                nint = orc_popc8(nd->imask);
                for (s = 0; s < 8; s++) if (orc_slot_hit(nd, grid, s, o, d, mint, maxt, 0)) hits8 |= 1u << s;
                for (s = 0; s < 8; s++) {
                    const uint32_t cnt = (nd->counts >> (2 * s)) & 3u;
                    if (!((hits8 >> s) & 1u) || ((nd->imask >> s) & 1u) || !cnt) continue;
                    {
                        const uint32_t off = orc_slot_tri_offset(nd, s);
                        uint32_t k;
                        for (k = 0; k < cnt; k++) {
                            const orc_tri* T = &((const orc_triunit*)&units[nd->base + nint + off + k])->rec;
                            float t, u, v;
                            tn_tris += 1;
                            if (orc_triaccel_intersect(T, o, d, mint, maxt, &u, &v, &t)) {
                                const int32_t id = (int32_t)T->faceIndex;
                                if (t < maxt || best < 0 || id > best) { maxt = t; best = id; }
                            }
                        }
                    }
                }
                for (s = 0; s < 8; s++) if (((hits8 & nd->imask) >> s) & 1u) pr |= 1u << (s ^ oi);
                g_base = nd->base; g_pend = pr; g_imask = nd->imask;
            }
        }
    }
    out[0] = tn_nodes; out[1] = tn_tris;
}

/* Analysis only (scripts/cpu_tree_quality.py): the same tree walked in EXACT front-to-back order - hit internal children
 * pushed individually with their entry distance, farthest first, and dropped at pop time when the entry distance has
 * fallen behind the closest hit. Counts what an ideally ordered traversal would fetch; the distance between this and
 * orc_wide_count is what the octant slot order and the missing entry-distance cull leave on the table.
 * out[0] = nodes, out[1] = triangle records, out[2] = entries dropped at pop time. */
struct orc_sorted_entry { uint32_t unit; double tn; };
void orc_wide_count_sorted(const void* units64, const float* grid, const float* rays, long long n, double* out)
{
    const orc_node64* units = (const orc_node64*)units64;
    double tn_nodes = 0, tn_tris = 0, tn_drop = 0;
    long long i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : tn_nodes, tn_tris, tn_drop)
    for (i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        const float* o = r; const float* d = r + 4;
        const float mint = r[3];
        float maxt = r[7];
        int32_t best = -1;
        struct orc_sorted_entry stack[512];
        int sp = 0;
        stack[sp].unit = 0; stack[sp].tn = mint; sp++;
        while (sp) {
            const orc_node64* nd;
            uint32_t nint, rel = 0;
            int s, first, a, b;
            sp--;
            if (stack[sp].tn > maxt) { tn_drop += 1; continue; }
            nd = &units[stack[sp].unit];
            tn_nodes += 1;
            nint = orc_popc8(nd->imask);
            first = sp;
            for (s = 0; s < 8; s++) {
                const int internal = (nd->imask >> s) & 1;
                const uint32_t cnt = (nd->counts >> (2 * s)) & 3u;
                double tnear = 0;
                const int hit = orc_slot_hit(nd, grid, s, o, d, mint, maxt, &tnear);
                if (internal) { const uint32_t child = nd->base + rel; rel++; if (hit && sp < 512) { stack[sp].unit = child; stack[sp].tn = tnear; sp++; } }
                else if (cnt && hit) {
                    const uint32_t off = orc_slot_tri_offset(nd, s);
                    uint32_t k;
                    for (k = 0; k < cnt; k++) {
                        const orc_tri* T = &((const orc_triunit*)&units[nd->base + nint + off + k])->rec;
                        float t, u, v;
                        tn_tris += 1;
                        if (orc_triaccel_intersect(T, o, d, mint, maxt, &u, &v, &t)) {
                            const int32_t id = (int32_t)T->faceIndex;
                            if (t < maxt || best < 0 || id > best) { maxt = t; best = id; }
                        }
                    }
                }
            }
            /* the children just pushed: farthest at the bottom, nearest on top */
            for (a = first + 1; a < sp; a++) {
                const struct orc_sorted_entry key = stack[a];
                for (b = a; b > first && stack[b - 1].tn < key.tn; b--) stack[b] = stack[b - 1];
                stack[b] = key;
            }
        }
    }
    out[0] = tn_nodes; out[1] = tn_tris; out[2] = tn_drop;
}

/* Analysis only: orc_wide_count's octant-ordered walk with an entry-distance cull.
 * mode 1: one lower bound per stacked sibling group = the smallest entry distance among the children still pending when the
 *         group is pushed; the whole group is dropped at pop time when that bound lies behind the closest hit;
 * mode 2: every pending child keeps its own entry distance and is dropped individually (the most a cull can do in this order).
 * out[0] = nodes, out[1] = triangle records, out[2] = children dropped unfetched. */
void orc_wide_count_cull(const void* units64, const float* grid, const float* rays, long long n, int mode, double* out)
{
    const orc_node64* units = (const orc_node64*)units64;
    double tn_nodes = 0, tn_tris = 0, tn_drop = 0;
    long long i;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : tn_nodes, tn_tris, tn_drop)
    for (i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        const float* o = r; const float* d = r + 4;
        const float mint = r[3];
        float maxt = r[7];
        int32_t best = -1;
        const uint32_t oct = (d[0] < 0 ? 1u : 0u) | (d[1] < 0 ? 2u : 0u) | (d[2] < 0 ? 4u : 0u);
        const uint32_t oi = 7u - oct;
        struct { uint32_t base, pend, imask; double tn[8]; } stack[256], g;      /* tn indexed by priority bit */
        int sp = 0, k;
        g.base = 0; g.pend = 0x80u; g.imask = 0;
        for (k = 0; k < 8; k++) g.tn[k] = mint;
        for (;;) {
            if (!g.pend) { if (!sp) break; g = stack[--sp]; continue; }
            {
                uint32_t bit = 7; const orc_node64* nd; uint32_t slot, rel, nint, hits8 = 0, pr = 0; int s;
                double ctn[8];
                if (mode == 1) {
                    /* group bound: all pending children behind the hit -> drop them together */
                    double m = 1e300; int cnt = 0;
                    for (k = 0; k < 8; k++) if ((g.pend >> k) & 1u) { if (g.tn[k] < m) m = g.tn[k]; cnt++; }
                    if (m > maxt) { tn_drop += cnt; g.pend = 0; continue; }
                }
                while (!((g.pend >> bit) & 1u)) bit--;
                g.pend &= ~(1u << bit);
                if (mode == 2 && g.tn[bit] > maxt) { tn_drop += 1; continue; }
                if (g.pend && sp < 256) stack[sp++] = g;
                slot = bit ^ oi;
                rel = orc_popc8(g.imask & ((1u << slot) - 1u));
                nd = &units[g.base + rel];
                tn_nodes += 1;
                nint = orc_popc8(nd->imask);
                for (s = 0; s < 8; s++) { ctn[s] = 0; if (orc_slot_hit(nd, grid, s, o, d, mint, maxt, &ctn[s])) hits8 |= 1u << s; }
                for (s = 0; s < 8; s++) {
                    const uint32_t cnt = (nd->counts >> (2 * s)) & 3u;
                    if (!((hits8 >> s) & 1u) || ((nd->imask >> s) & 1u) || !cnt) continue;
                    {
                        const uint32_t off = orc_slot_tri_offset(nd, s);
                        uint32_t q;
                        for (q = 0; q < cnt; q++) {
                            const orc_tri* T = &((const orc_triunit*)&units[nd->base + nint + off + q])->rec;
                            float t, u, v;
                            tn_tris += 1;
                            if (orc_triaccel_intersect(T, o, d, mint, maxt, &u, &v, &t)) {
                                const int32_t id = (int32_t)T->faceIndex;
                                if (t < maxt || best < 0 || id > best) { maxt = t; best = id; }
                            }
                        }
                    }
                }
                for (s = 0; s < 8; s++) if (((hits8 & nd->imask) >> s) & 1u) { pr |= 1u << (s ^ oi); g.tn[s ^ oi] = ctn[s]; }
                g.base = nd->base; g.pend = pr; g.imask = nd->imask;
            }
        }
    }
    out[0] = tn_nodes; out[1] = tn_tris; out[2] = tn_drop;
}
