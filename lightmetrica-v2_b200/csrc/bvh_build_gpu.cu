// GPU builders for the same flattened structure as bvh_build.cpp (one array of 64-byte units, bvh.h), all on the device:
//   LMB200_BUILD_GPU_LBVH      Morton-ordered binary radix tree (Karras 2012) -> bottom-up box fit -> level-by-level greedy
//                              collapse to 8-wide nodes (open the largest child first, single-triangle leaves)
//   LMB200_BUILD_GPU_LBVH_SAH  the same tree with the host builder's SAH-optimal collapse (dynamic programme)
//   LMB200_BUILD_GPU_PLOC      parallel locally-ordered clustering instead of the radix tree, SAH-optimal collapse
// followed by octant slot assignment + quantisation.
// Replaces the serial recursive build of the reference (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:202-383: ~2 s per 100 k triangles) when time-to-first-ray matters more than tree
// quality: LBVH trees are not SAH-optimised, so traversal is slower than on the host-built tree
// (DESIGN.md §6). The triangle records are bit-identical to the host path (same operation order,
// __f*_rn intrinsics), hence closest-hit results are identical whichever builder is used.
#include "internal.h"

#include <cub/cub.cuh>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cfloat>

namespace lmb200 {

namespace {

constexpr double kQSlackDev = 1.0 / 256.0;   // as kQSlack in bvh_build.cpp
#ifndef LMB_GPU_LEAF
#define LMB_GPU_LEAF 1u      // subtrees with at most this many triangles become one leaf slot (<= 3); 1 measured best (profiles/r01_sweep.md)
#endif

// ---- bit-exact TriAccel precompute (triaccel.h host version; /root/reference/include/lightmetrica/triaccel.h:50-91) ----
__device__ __forceinline__ bool triaccel_load_dev(TriRecord& r, const float* A, const float* B, const float* C, uint32_t tri)
{
    const int waldModulo[4] = {1, 2, 0, 1};
    r.tri = tri; r.pad = 0;
    const float b[3] = {__fsub_rn(C[0], A[0]), __fsub_rn(C[1], A[1]), __fsub_rn(C[2], A[2])};
    const float c[3] = {__fsub_rn(B[0], A[0]), __fsub_rn(B[1], A[1]), __fsub_rn(B[2], A[2])};
    const float N[3] = {__fsub_rn(__fmul_rn(c[1], b[2]), __fmul_rn(c[2], b[1])),
                        __fsub_rn(__fmul_rn(c[2], b[0]), __fmul_rn(c[0], b[2])),
                        __fsub_rn(__fmul_rn(c[0], b[1]), __fmul_rn(c[1], b[0]))};
    uint32_t k = 0;
    for (uint32_t j = 0; j < 3; j++) if (fabsf(N[j]) > fabsf(N[k])) k = j;
    const int u = waldModulo[k], v = waldModulo[k + 1];
    const float n_k = N[k];
    const float denom = __fsub_rn(__fmul_rn(b[u], c[v]), __fmul_rn(b[v], c[u]));
    if (denom == 0.f) {
        r.k = 3; r.n_u = r.n_v = r.n_d = r.a_u = r.a_v = r.b_nu = r.b_nv = r.c_nu = r.c_nv = 0.f;
        return false;
    }
    r.k = k;
    r.n_u = __fdiv_rn(N[u], n_k);
    r.n_v = __fdiv_rn(N[v], n_k);
    // Dot3 with the _mm_dp_ps summation order: (x*x' + y*y') + (z*z' + 0)
    r.n_d = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(A[0], N[0]), __fmul_rn(A[1], N[1])), __fadd_rn(__fmul_rn(A[2], N[2]), 0.0f)), n_k);
    r.b_nu = __fdiv_rn(b[u], denom);
    r.b_nv = __fdiv_rn(-b[v], denom);
    r.a_u = A[u];
    r.a_v = A[v];
    r.c_nu = __fdiv_rn(c[v], denom);
    r.c_nv = __fdiv_rn(-c[u], denom);
    return true;
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v)
{
    // valid for any finite floats: order-preserving integer view
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v)
{
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct Box6 { float lo[3], hi[3]; };

// records, validity, raw boxes, scene bounds
__global__ void k_prep(const float* __restrict__ verts, uint32_t n, TriRecord* recs, Box6* boxes, uint8_t* valid, float* scene /*[6]*/)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        float v[9];
        for (int k = 0; k < 9; k++) v[k] = verts[9 * (size_t)i + k];
        TriRecord r;
        bool ok = triaccel_load_dev(r, v, v + 3, v + 6, i);
        for (int k = 0; k < 9; k++) ok = ok && isfinite(v[k]);
        recs[i] = r;
        valid[i] = ok ? 1 : 0;
        Box6 b;
        for (int a = 0; a < 3; a++) {
            b.lo[a] = fminf(v[a], fminf(v[3 + a], v[6 + a]));
            b.hi[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
        }
        boxes[i] = b;
        if (ok) for (int a = 0; a < 3; a++) { lo[a] = b.lo[a]; hi[a] = b.hi[a]; }
    }
    // warp reduce, then one atomic per warp and axis
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31u) == 0 && lo[a] <= hi[a]) { atomic_min_float(scene + a, lo[a]); atomic_max_float(scene + 3 + a, hi[a]); }
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 63-bit Morton keys of the box centres; invalid triangles sort to the end
__global__ void k_morton(const Box6* __restrict__ boxes, const uint8_t* __restrict__ valid, uint32_t n, const float* __restrict__ scene,
                         unsigned long long* keys, uint32_t* idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = i;
    if (!valid[i]) { keys[i] = ~0ull; return; }
    unsigned long long q[3];
    for (int a = 0; a < 3; a++) {
        const float ext = scene[3 + a] - scene[a];
        const float c = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
        float t = ext > 0.f ? (c - scene[a]) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        q[a] = (unsigned long long)fminf(t * 2097152.0f, 2097151.0f);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
}

// binary radix tree over the sorted keys (Karras 2012). child refs: bit 31 set = leaf (sorted position)
struct RadixTree {
    uint32_t* left; uint32_t* right; uint32_t* parent_internal; uint32_t* parent_leaf;
    uint32_t* first; uint32_t* last;   // sorted-position range of each internal node
};

__device__ __forceinline__ int lcp(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_radix_tree(const unsigned long long* __restrict__ keys, int n, RadixTree T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lcp(keys, n, i, i + 1) - lcp(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lcp(keys, n, i, i - d);
    int lmax = 2;
    while (lcp(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (lcp(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lcp(keys, n, i, j);
    int s = 0, t = l;
    do { t = (t + 1) / 2; if (lcp(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t L = lo == gamma ? (0x80000000u | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t R = hi == gamma + 1 ? (0x80000000u | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    T.left[i] = L; T.right[i] = R; T.first[i] = (uint32_t)lo; T.last[i] = (uint32_t)hi;
    if (L & 0x80000000u) T.parent_leaf[gamma] = (uint32_t)i; else T.parent_internal[gamma] = (uint32_t)i;
    if (R & 0x80000000u) T.parent_leaf[gamma + 1] = (uint32_t)i; else T.parent_internal[gamma + 1] = (uint32_t)i;
    if (i == 0) T.parent_internal[0] = 0xffffffffu;
}

__device__ __forceinline__ Box6 load_box_cg(const Box6* p)
{
    // boxes written by other SMs in the same launch: bypass the (non-coherent) L1
    Box6 b;
    const float* f = reinterpret_cast<const float*>(p);
    for (int a = 0; a < 3; a++) { b.lo[a] = __ldcg(f + a); b.hi[a] = __ldcg(f + 3 + a); }
    return b;
}

// bottom-up box fit: the second thread to reach an internal node merges its children
__global__ void k_fit(const Box6* __restrict__ boxes, const uint32_t* __restrict__ idx, int n, float pad, RadixTree T,
                      Box6* leaf_box, Box6* node_box, int* flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Box6 b = boxes[idx[k]];
    for (int a = 0; a < 3; a++) { b.lo[a] -= pad; b.hi[a] += pad; }
    leaf_box[k] = b;
    if (n == 1) return;
    uint32_t p = T.parent_leaf[k];
    for (;;) {
        __threadfence();
        if (atomicAdd(&flags[p], 1) == 0) return;       // first arrival: the sibling is not ready yet
        __threadfence();
        const uint32_t L = T.left[p], R = T.right[p];
        const Box6 bl = load_box_cg((L & 0x80000000u) ? &leaf_box[L & 0x7fffffffu] : &node_box[L]);
        const Box6 br = load_box_cg((R & 0x80000000u) ? &leaf_box[R & 0x7fffffffu] : &node_box[R]);
        Box6 m;
        for (int a = 0; a < 3; a++) { m.lo[a] = fminf(bl.lo[a], br.lo[a]); m.hi[a] = fmaxf(bl.hi[a], br.hi[a]); }
        node_box[p] = m;
        if (p == 0) return;
        p = T.parent_internal[p];
    }
}

__device__ __forceinline__ float half_area(const Box6& b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

struct WorkItem { uint32_t bin; uint32_t wide; };

// One thread per wide node: gather <= 8 slots by opening the largest expandable child, assign octant
// slots, quantise, allocate children / triangles, emit.
__global__ void k_collapse(const WorkItem* __restrict__ in, uint32_t n_in, WorkItem* out, uint32_t* out_count,
                           RadixTree T, const Box6* __restrict__ leaf_box, const Box6* __restrict__ node_box,
                           const uint32_t* __restrict__ idx, const TriRecord* __restrict__ recs,
                           Unit64* units, uint32_t* node_count, uint32_t* unit_count, SceneGrid grid, int n_prims)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_in) return;
    const WorkItem it = in[w];
    uint32_t ch[8]; int n = 0;
    auto count_of = [&](uint32_t ref) -> uint32_t { return (ref & 0x80000000u) ? 1u : T.last[ref] - T.first[ref] + 1u; };
    auto box_of = [&](uint32_t ref) -> Box6 { return (ref & 0x80000000u) ? leaf_box[ref & 0x7fffffffu] : node_box[ref]; };
    if (n_prims == 1) { ch[n++] = 0x80000000u; }
    else if (count_of(it.bin) <= LMB_GPU_LEAF) { ch[n++] = it.bin; }       // the whole (sub)tree is one leaf slot
    else { ch[n++] = T.left[it.bin]; ch[n++] = T.right[it.bin]; }
    for (;;) {
        int best = -1; float best_area = -1.f;
        for (int i = 0; i < n; i++) {
            if ((ch[i] & 0x80000000u) || count_of(ch[i]) <= LMB_GPU_LEAF) continue;   // leaf slot
            const float a = half_area(box_of(ch[i]));
            if (a > best_area) { best_area = a; best = i; }
        }
        if (best < 0 || n == 8) break;
        const uint32_t b = ch[best];
        ch[best] = T.left[b]; ch[n++] = T.right[b];
    }
    Box6 cb[8];
    Box6 nb; for (int a = 0; a < 3; a++) { nb.lo[a] = FLT_MAX; nb.hi[a] = -FLT_MAX; }
    for (int i = 0; i < n; i++) {
        cb[i] = box_of(ch[i]);
        for (int a = 0; a < 3; a++) { nb.lo[a] = fminf(nb.lo[a], cb[i].lo[a]); nb.hi[a] = fmaxf(nb.hi[a], cb[i].hi[a]); }
    }
    Node64 node;
    memset(&node, 0, sizeof(node));
    double scale[3];
    float origin[3];
    for (int a = 0; a < 3; a++) {
        // node origin = the box minimum snapped down to the 16-bit scene grid (bvh.h); decoded exactly as the traversal does
        double kd = floor(((double)nb.lo[a] - (double)grid.lo[a]) / (double)grid.step[a]);
        kd = fmax(0.0, fmin(65535.0, kd));
        uint32_t ki = (uint32_t)kd;
        while (ki > 0 && fmaf((float)ki, grid.step[a], grid.lo[a]) > nb.lo[a]) ki--;
        node.k[a] = (uint16_t)ki;
        origin[a] = fmaf((float)ki, grid.step[a], grid.lo[a]);
        const double ext = fmax(0.0, (double)nb.hi[a] - (double)origin[a]);
        int e = ext > 0 ? (int)ceil(log2(ext / 254.0)) : -126;
        e = max(-126, min(110, e));
        while (e < 110 && ceil(ext / ldexp(1.0, e) + 2 * kQSlackDev) > 255.0) e++;
        node.e[a] = (uint8_t)(e + 127);
        scale[a] = ldexp(1.0, e);
    }
    // octant slot assignment (greedy on the centroid-offset score), as Emitter::emit_node
    int slot_of[8]; bool slot_used[8], child_done[8];
    for (int i = 0; i < 8; i++) { slot_used[i] = false; child_done[i] = false; slot_of[i] = -1; }
    for (int round = 0; round < n; round++) {
        int bc = -1, bs = -1; float bscore = -FLT_MAX;
        for (int i = 0; i < n; i++) {
            if (child_done[i]) continue;
            float cen[3];
            for (int a = 0; a < 3; a++) cen[a] = 0.5f * (cb[i].lo[a] + cb[i].hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
            for (int s = 0; s < 8; s++) {
                if (slot_used[s]) continue;
                const float score = ((s & 1) ? cen[0] : -cen[0]) + ((s & 2) ? cen[1] : -cen[1]) + ((s & 4) ? cen[2] : -cen[2]);
                if (score > bscore) { bscore = score; bc = i; bs = s; }
            }
        }
        slot_of[bc] = bs; slot_used[bs] = true; child_done[bc] = true;
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int i = 0; i < n; i++) child_in_slot[slot_of[i]] = i;
    // allocation
    uint32_t n_internal = 0, n_tris = 0;
    for (int i = 0; i < n; i++) {
        const bool leaf = (ch[i] & 0x80000000u) || count_of(ch[i]) <= LMB_GPU_LEAF;
        if (leaf) n_tris += count_of(ch[i]); else n_internal++;
    }
    // children: internal nodes first (slot order), then the triangles of the leaf slots (slot order), one allocation
    const uint32_t base = atomicAdd(unit_count, n_internal + n_tris);
    if (n_internal) atomicAdd(node_count, n_internal);
    const uint32_t qbase = n_internal ? atomicAdd(out_count, n_internal) : 0u;
    node.base = base;
    uint32_t rel = 0, toff = 0;
    for (int s = 0; s < 8; s++) {
        const int i = child_in_slot[s];
        if (i < 0) { for (int a = 0; a < 3; a++) { node.qlo[a][s] = 255; node.qhi[a][s] = 0; } continue; }
        for (int a = 0; a < 3; a++) {
            double ql = floor(((double)cb[i].lo[a] - (double)origin[a]) / scale[a] - kQSlackDev);
            double qh = ceil(((double)cb[i].hi[a] - (double)origin[a]) / scale[a] + kQSlackDev);
            ql = fmax(0.0, fmin(255.0, ql)); qh = fmax(0.0, fmin(255.0, qh));
            node.qlo[a][s] = (uint8_t)ql; node.qhi[a][s] = (uint8_t)qh;
        }
        const bool leaf = (ch[i] & 0x80000000u) || count_of(ch[i]) <= LMB_GPU_LEAF;
        if (leaf) {
            const uint32_t cnt = count_of(ch[i]);
            const uint32_t first = (ch[i] & 0x80000000u) ? (ch[i] & 0x7fffffffu) : T.first[ch[i]];
            node.counts |= (uint16_t)(cnt << (2 * s));
            for (uint32_t k = 0; k < cnt; k++) {
                TriUnit tu;
                tu.rec = recs[idx[first + k]];
                tu.pad[0] = tu.pad[1] = tu.pad[2] = tu.pad[3] = 0u;
                units[base + n_internal + toff + k].tri = tu;
            }
            toff += cnt;
        } else {
            node.imask |= (uint8_t)(1u << s);
            out[qbase + rel] = WorkItem{ch[i], base + rel};
            rel++;
        }
    }
    units[it.wide].node = node;
}


// ------------------------------------------------------------------------------------------------
// PLOC builder (parallel locally-ordered clustering, Meister & Bittner 2018) + SAH-optimal wide collapse on the device.
// The Morton-ordered radix tree above splits by key bits only; PLOC builds the binary tree bottom-up by repeatedly
// merging mutual nearest neighbours (smallest merged surface area within +-LMB_PLOC_RADIUS positions of the Morton-ordered
// cluster array), which gives trees close to a top-down SAH build. The same dynamic programme as the host builder
// (bvh_build.cpp Emitter::plan, Ylitie et al. 2017 sec. 4.1) then chooses the 8-wide nodes: its tables are filled for
// every new binary node right when the node is created (children are always older than their parent).
#ifndef LMB_PLOC_RADIUS
#define LMB_PLOC_RADIUS 12
#endif
#ifndef LMB_WIDE_COST_NODE
#define LMB_WIDE_COST_NODE 1.0f
#endif
#ifndef LMB_GPU_MAX_LEAF
#define LMB_GPU_MAX_LEAF 3      // triangles per leaf slot the collapse may choose (<= 3: two count bits per slot)
#endif
#ifndef LMB_GPU_COST_TRI
#define LMB_GPU_COST_TRI 1.2f
#endif
#undef LMB_WIDE_COST_TRI
#define LMB_WIDE_COST_TRI LMB_GPU_COST_TRI

struct PlocTree {
    uint32_t* left; uint32_t* right;      // children of internal node k (refs: bit 31 set = leaf, sorted position)
    uint32_t* count;                      // triangles below internal node k
    Box6* box;                            // box of internal node k
    float* cost;                          // 7 per internal node: cheapest representation with <= i+1 slots
    unsigned long long* choice;           // per internal node: take[7] (2 bits each) | dist_k[7] (3 bits each, at bit 14) | int_k (3 bits, at bit 35)
};

__device__ __forceinline__ Box6 box_union(const Box6& a, const Box6& b)
{
    Box6 m;
    for (int k = 0; k < 3; k++) { m.lo[k] = fminf(a.lo[k], b.lo[k]); m.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
    return m;
}

// nearest neighbour of every cluster within the search radius (ties: the lower position)
__global__ void k_ploc_nn(const Box6* __restrict__ cb, uint32_t n, uint32_t* __restrict__ nn)
{
    constexpr int R = LMB_PLOC_RADIUS, TB = 256;
    __shared__ Box6 tile[TB + 2 * R];
    const int base = (int)(blockIdx.x * TB) - R;
    for (int t = threadIdx.x; t < TB + 2 * R; t += TB) {
        const int g = base + t;
        if (g >= 0 && g < (int)n) tile[t] = cb[g];
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * TB + threadIdx.x;
    if (i >= n) return;
    const Box6 me = tile[threadIdx.x + R];
    float best = FLT_MAX; uint32_t bj = i;
    for (int d = -R; d <= R; d++) {
        const int j = (int)i + d;
        if (d == 0 || j < 0 || j >= (int)n) continue;
        const float a = half_area(box_union(me, tile[threadIdx.x + R + d]));
        if (a < best) { best = a; bj = (uint32_t)j; }
    }
    nn[i] = bj;
}

// flags for the scan: low word = the cluster survives (1) or is absorbed by its partner (0); high word = it creates a node
__global__ void k_ploc_flags(const uint32_t* __restrict__ nn, uint32_t n, int force_pairs, unsigned long long* __restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t j = nn[i];
    bool mutual = j != i && nn[j] == i;
    if (force_pairs) { j = i ^ 1u; mutual = j < n; }      // guaranteed progress: neighbours (2k, 2k+1) merge
    unsigned long long f = 1ull;
    if (mutual) f = i < j ? (1ull | (1ull << 32)) : 0ull;
    flags[i] = f;
}

__device__ __forceinline__ void dp_children(const PlocTree& T, const Box6* __restrict__ leaf_box, uint32_t ref, float* c /*[7]*/, uint32_t& total)
{
    if (ref & 0x80000000u) {
        const float v = half_area(leaf_box[ref & 0x7fffffffu]) * LMB_WIDE_COST_TRI;
        for (int i = 0; i < 7; i++) c[i] = v;
        total = 1;
    } else {
        // L2 loads: in the radix-tree pass these were written by another SM earlier in the same launch
        for (int i = 0; i < 7; i++) c[i] = __ldcg(T.cost + (size_t)ref * 7 + i);
        total = __ldcg(T.count + ref);
    }
}

// Collapse tables of the new binary node k with children L, R and box b (bvh_build.cpp Emitter::plan for one node): cost[i-1] =
// cheapest way to represent the subtree with at most i wide-node slots; take / dist_k / int_k remember how.
__device__ __forceinline__ void dp_node(const PlocTree& T, const Box6* __restrict__ leaf_box, uint32_t k, uint32_t L, uint32_t R, const Box6& b)
{
    float cl7[7], cr7[7];
    uint32_t tl, tr;
    dp_children(T, leaf_box, L, cl7, tl);
    dp_children(T, leaf_box, R, cr7, tr);
    const uint32_t total = tl + tr;
    const float area = half_area(b);
    const float leaf = total <= (uint32_t)LMB_GPU_MAX_LEAF ? area * (float)total * LMB_WIDE_COST_TRI : FLT_MAX;
    unsigned long long ch = 0;
    float c[7];
    // LMB_GPU_DP_FULL: a subtree that has at least i triangles must use exactly i slots (the budget is only left unused when
    // there are not enough triangles), i.e. the collapse picks the SAH-best cut among the FULLEST wide nodes; without it the
    // classic "at most i slots" programme of the host builder.
    auto distribute = [&](int jn, uint32_t& kbest) {
        float best = FLT_MAX; kbest = 1;
        for (int kk = 1; kk < jn; kk++) {
            if (kk > 7 || jn - kk > 7) continue;
#ifdef LMB_GPU_DP_FULL
            if ((uint32_t)kk > tl || (uint32_t)(jn - kk) > tr) continue;
#endif
            const float v = cl7[kk - 1] + cr7[jn - kk - 1];
            if (v < best) { best = v; kbest = (uint32_t)kk; }
        }
        return best;
    };
    uint32_t ik;
#ifdef LMB_GPU_DP_FULL
    const float internal = distribute((int)min(total, 8u), ik) + area * LMB_WIDE_COST_NODE;
#else
    const float internal = distribute(8, ik) + area * LMB_WIDE_COST_NODE;
#endif
    ch |= (unsigned long long)ik << 35;
    if (leaf <= internal) c[0] = leaf; else { c[0] = internal; ch |= 1ull; }
    for (int i2 = 2; i2 <= 7; i2++) {
        uint32_t dk = 1;
#ifdef LMB_GPU_DP_FULL
        const float d = (uint32_t)i2 <= total ? distribute(i2, dk) : FLT_MAX;
        ch |= (unsigned long long)dk << (14 + 3 * (i2 - 1));
        if (d < FLT_MAX) { c[i2 - 1] = d; ch |= 2ull << (2 * (i2 - 1)); } else { c[i2 - 1] = c[i2 - 2]; ch |= 3ull << (2 * (i2 - 1)); }
#else
        const float d = distribute(i2, dk);
        ch |= (unsigned long long)dk << (14 + 3 * (i2 - 1));
        if (d < c[i2 - 2]) { c[i2 - 1] = d; ch |= 2ull << (2 * (i2 - 1)); } else { c[i2 - 1] = c[i2 - 2]; ch |= 3ull << (2 * (i2 - 1)); }
#endif
    }
    T.left[k] = L; T.right[k] = R; T.count[k] = total; T.box[k] = b; T.choice[k] = ch;
    for (int i2 = 0; i2 < 7; i2++) T.cost[(size_t)k * 7 + i2] = c[i2];
}

// applies the merges of one round: new internal nodes (with their collapse tables) and the compacted cluster array
__global__ void k_ploc_apply(const uint32_t* __restrict__ cl, const Box6* __restrict__ cb, const uint32_t* __restrict__ nn, uint32_t n, int force_pairs,
                             const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan, uint32_t node_base,
                             uint32_t* __restrict__ cl_out, Box6* __restrict__ cb_out, PlocTree T, const Box6* __restrict__ leaf_box)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long f = flags[i];
    if (!(f & 1ull)) return;                               // absorbed
    const uint32_t pos = (uint32_t)(scan[i] & 0xffffffffull);
    if (!(f >> 32)) { cl_out[pos] = cl[i]; cb_out[pos] = cb[i]; return; }
    const uint32_t j = force_pairs ? (i ^ 1u) : nn[i];
    const uint32_t k = node_base + (uint32_t)(scan[i] >> 32);
    const uint32_t L = cl[i], R = cl[j];
    const Box6 b = box_union(cb[i], cb[j]);
    dp_node(T, leaf_box, k, L, R, b);
    cl_out[pos] = k; cb_out[pos] = b;
}

// Collapse tables for the Morton radix tree: a second bottom-up pass in the pattern of k_fit (the second thread to reach a
// node handles it), after the boxes are known.
__global__ void k_dp_radix(int n, RadixTree R, PlocTree T, const Box6* __restrict__ leaf_box, int* flags)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || n == 1) return;
    uint32_t p = R.parent_leaf[k];
    for (;;) {
        __threadfence();
        if (atomicAdd(&flags[p], 1) == 0) return;
        __threadfence();
        dp_node(T, leaf_box, p, R.left[p], R.right[p], load_box_cg(&T.box[p]));
        if (p == 0) return;
        p = R.parent_internal[p];
    }
}

__global__ void k_ploc_init(uint32_t n, const Box6* __restrict__ leaf_box, uint32_t* cl, Box6* cb)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cl[i] = 0x80000000u | i; cb[i] = leaf_box[i];
}

__global__ void k_leaf_boxes(const Box6* __restrict__ boxes, const uint32_t* __restrict__ idx, uint32_t n, float pad, Box6* leaf_box)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Box6 b = boxes[idx[k]];
    for (int a = 0; a < 3; a++) { b.lo[a] -= pad; b.hi[a] += pad; }
    leaf_box[k] = b;
}

// One thread per wide node: its slots are what the collapse tables planned (bvh_build.cpp Emitter::collect), then octant
// slot assignment, quantisation and allocation exactly as k_collapse / Emitter::emit_node.
__global__ void k_emit_dp(const WorkItem* __restrict__ in, uint32_t n_in, WorkItem* out, uint32_t* out_count,
                          PlocTree T, const Box6* __restrict__ leaf_box, const uint32_t* __restrict__ idx, const TriRecord* __restrict__ recs,
                          Unit64* units, uint32_t* node_count, uint32_t* unit_count, SceneGrid grid, int n_prims)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_in) return;
    const WorkItem it = in[w];
    uint32_t ch[8]; bool is_leaf[8]; int n = 0;
    auto box_of = [&](uint32_t ref) -> Box6 { return (ref & 0x80000000u) ? leaf_box[ref & 0x7fffffffu] : T.box[ref]; };
    auto count_of = [&](uint32_t ref) -> uint32_t { return (ref & 0x80000000u) ? 1u : T.count[ref]; };
    if (n_prims == 1 || (it.bin & 0x80000000u)) { ch[n] = n_prims == 1 ? 0x80000000u : it.bin; is_leaf[n++] = true; }
    else if (it.wide == 0 && T.count[it.bin] <= 3u && !(T.choice[it.bin] & 3ull)) { ch[n] = it.bin; is_leaf[n++] = true; }   // the whole tree is one leaf slot
    else {
        // collect(left, int_k) + collect(right, 8 - int_k), iteratively
        uint32_t sref[16]; int sbud[16]; int sp = 0;
        const int ik = (int)((T.choice[it.bin] >> 35) & 7ull);
        sref[sp] = T.right[it.bin]; sbud[sp++] = 8 - ik;
        sref[sp] = T.left[it.bin]; sbud[sp++] = ik;
        while (sp) {
            uint32_t ref = sref[--sp]; int bud = sbud[sp];
            if (ref & 0x80000000u) { ch[n] = ref; is_leaf[n++] = true; continue; }
            const unsigned long long c = T.choice[ref];
            int take = (int)((c >> (2 * (bud - 1))) & 3ull);
            while (take == 3) { bud--; take = (int)((c >> (2 * (bud - 1))) & 3ull); }
            if (take == 0) { ch[n] = ref; is_leaf[n++] = true; }
            else if (take == 1) { ch[n] = ref; is_leaf[n++] = false; }
            else {
                const int k = (int)((c >> (14 + 3 * (bud - 1))) & 7ull);
                sref[sp] = T.right[ref]; sbud[sp++] = bud - k;
                sref[sp] = T.left[ref]; sbud[sp++] = k;
            }
        }
    }
    Box6 cb[8];
    Box6 nb; for (int a = 0; a < 3; a++) { nb.lo[a] = FLT_MAX; nb.hi[a] = -FLT_MAX; }
    for (int i = 0; i < n; i++) { cb[i] = box_of(ch[i]); nb = box_union(nb, cb[i]); }
    Node64 node;
    memset(&node, 0, sizeof(node));
    double scale[3];
    float origin[3];
    for (int a = 0; a < 3; a++) {
        double kd = floor(((double)nb.lo[a] - (double)grid.lo[a]) / (double)grid.step[a]);
        kd = fmax(0.0, fmin(65535.0, kd));
        uint32_t ki = (uint32_t)kd;
        while (ki > 0 && fmaf((float)ki, grid.step[a], grid.lo[a]) > nb.lo[a]) ki--;
        node.k[a] = (uint16_t)ki;
        origin[a] = fmaf((float)ki, grid.step[a], grid.lo[a]);
        const double ext = fmax(0.0, (double)nb.hi[a] - (double)origin[a]);
        int e = ext > 0 ? (int)ceil(log2(ext / 254.0)) : -126;
        e = max(-126, min(110, e));
        while (e < 110 && ceil(ext / ldexp(1.0, e) + 2 * kQSlackDev) > 255.0) e++;
        node.e[a] = (uint8_t)(e + 127);
        scale[a] = ldexp(1.0, e);
    }
    int slot_of[8]; bool slot_used[8], child_done[8];
    for (int i = 0; i < 8; i++) { slot_used[i] = false; child_done[i] = false; slot_of[i] = -1; }
    for (int round = 0; round < n; round++) {
        int bc = -1, bs = -1; float bscore = -FLT_MAX;
        for (int i = 0; i < n; i++) {
            if (child_done[i]) continue;
            float cen[3];
            for (int a = 0; a < 3; a++) cen[a] = 0.5f * (cb[i].lo[a] + cb[i].hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
            for (int s = 0; s < 8; s++) {
                if (slot_used[s]) continue;
                const float score = ((s & 1) ? cen[0] : -cen[0]) + ((s & 2) ? cen[1] : -cen[1]) + ((s & 4) ? cen[2] : -cen[2]);
                if (score > bscore) { bscore = score; bc = i; bs = s; }
            }
        }
        slot_of[bc] = bs; slot_used[bs] = true; child_done[bc] = true;
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int i = 0; i < n; i++) child_in_slot[slot_of[i]] = i;
    uint32_t n_internal = 0, n_tris = 0;
    for (int i = 0; i < n; i++) { if (is_leaf[i]) n_tris += count_of(ch[i]); else n_internal++; }
    const uint32_t base = atomicAdd(unit_count, n_internal + n_tris);
    if (n_internal) atomicAdd(node_count, n_internal);
    const uint32_t qbase = n_internal ? atomicAdd(out_count, n_internal) : 0u;
    node.base = base;
    uint32_t rel = 0, toff = 0;
    for (int s = 0; s < 8; s++) {
        const int i = child_in_slot[s];
        if (i < 0) { for (int a = 0; a < 3; a++) { node.qlo[a][s] = 255; node.qhi[a][s] = 0; } continue; }
        for (int a = 0; a < 3; a++) {
            double ql = floor(((double)cb[i].lo[a] - (double)origin[a]) / scale[a] - kQSlackDev);
            double qh = ceil(((double)cb[i].hi[a] - (double)origin[a]) / scale[a] + kQSlackDev);
            ql = fmax(0.0, fmin(255.0, ql)); qh = fmax(0.0, fmin(255.0, qh));
            node.qlo[a][s] = (uint8_t)ql; node.qhi[a][s] = (uint8_t)qh;
        }
        if (is_leaf[i]) {
            // the (<= 3) triangles below this binary subtree
            uint32_t st[4]; int sp = 0; uint32_t cnt = 0;
            st[sp++] = ch[i];
            while (sp) {
                const uint32_t ref = st[--sp];
                if (ref & 0x80000000u) {
                    TriUnit tu;
                    tu.rec = recs[idx[ref & 0x7fffffffu]];
                    tu.pad[0] = tu.pad[1] = tu.pad[2] = tu.pad[3] = 0u;
                    units[base + n_internal + toff + cnt].tri = tu;
                    cnt++;
                } else { st[sp++] = T.right[ref]; st[sp++] = T.left[ref]; }
            }
            node.counts |= (uint16_t)(cnt << (2 * s));
            toff += cnt;
        } else {
            node.imask |= (uint8_t)(1u << s);
            out[qbase + rel] = WorkItem{ch[i], base + rel};
            rel++;
        }
    }
    units[it.wide].node = node;
}

// Scratch memory of one build: ONE device allocation carved up by a bump pointer (a cudaMalloc / cudaFree pair per
// temporary array - about 25 of them - cost more than all the kernels of the build together), with a plain cudaMalloc
// as the fallback should the estimate ever be too small.
struct DevBuf {
    uint8_t* arena = nullptr;
    size_t cap = 0, used = 0;
    std::vector<void*> extra;
    bool reserve(size_t bytes)
    {
        if (cudaMalloc(reinterpret_cast<void**>(&arena), bytes) != cudaSuccess) { cudaGetLastError(); arena = nullptr; return false; }
        cap = bytes; used = 0;
        return true;
    }
    template <typename T> T* alloc(size_t n)
    {
        const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
        if (arena && used + bytes <= cap) { T* p = reinterpret_cast<T*>(arena + used); used += bytes; return p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        extra.push_back(p);
        return reinterpret_cast<T*>(p);
    }
    ~DevBuf() { if (arena) cudaFree(arena); for (void* p : extra) cudaFree(p); }
};

}  // namespace

// Builds on the accel's device from HOST vertices; fills a->d_units and the stats.
int build_bvh_gpu(Accel* a, const float* verts_host, uint64_t ntris, int builder)
{
    const bool ploc = builder == LMB200_BUILD_GPU_PLOC;
    const bool greedy = builder == LMB200_BUILD_GPU_LBVH;      // radix tree + greedy collapse; LMB200_BUILD_GPU_LBVH_SAH: + collapse tables
    const auto t0 = std::chrono::steady_clock::now();
    const bool timing = getenv("LMB200_BUILD_TIMING") != nullptr;
    auto tick = [&](const char* what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        fprintf(stderr, "[lmb200 gpu build] %-28s %8.3f ms\n", what, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    service_destroy(a);
    if (a->d_units) { cudaFree(a->d_units); a->d_units = nullptr; a->num_units = 0; }
    if (!a->d_counter && (e = cudaMalloc(&a->d_counter, LMB_NUM_COUNTERS * sizeof(unsigned long long))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(counter)");
    const uint32_t n = (uint32_t)ntris;
    DevBuf D;
    // per triangle: vertices 36, record 48, box 24, keys 2 x 8, indices 2 x 4, validity 1, radix tree 24, leaf / node boxes 48, flags 4,
    // work queues 16, collapse tables 40, sort scratch ~20; the clustering builder adds cluster arrays, neighbours and scan buffers (~110)
    D.reserve((size_t)std::max<uint32_t>(n, 1024) * (ploc ? 420 : 300) + (64u << 20));
    float* d_verts = D.alloc<float>(9 * (size_t)n);
    TriRecord* recs = D.alloc<TriRecord>(n);
    Box6* boxes = D.alloc<Box6>(n);
    uint8_t* valid = D.alloc<uint8_t>(n);
    float* scene = D.alloc<float>(6);
    unsigned long long* keys = D.alloc<unsigned long long>(n);
    unsigned long long* keys2 = D.alloc<unsigned long long>(n);
    uint32_t* idx = D.alloc<uint32_t>(n);
    uint32_t* idx2 = D.alloc<uint32_t>(n);
    if (!d_verts || !recs || !boxes || !valid || !scene || !keys || !keys2 || !idx || !idx2) return set_error(LMB200_E_CUDA, "out of device memory (gpu build)");
    tick("alloc 1");
    if (n && (e = cudaMemcpy(d_verts, verts_host, sizeof(float) * 9 * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "H2D verts");
    tick("H2D verts");
    const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    cudaMemcpy(scene, init, sizeof(init), cudaMemcpyHostToDevice);
    const int TB = 256;
    if (n) {
        k_prep<<<(n + TB - 1) / TB, TB>>>(d_verts, n, recs, boxes, valid, scene); g_launch_count++;
        k_morton<<<(n + TB - 1) / TB, TB>>>(boxes, valid, n, scene, keys, idx); g_launch_count++;
    }
    float h_scene[6];
    cudaMemcpy(h_scene, scene, sizeof(h_scene), cudaMemcpyDeviceToHost);
    // number of valid triangles = count of keys != ~0 after sorting
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)n);
    void* tmp = D.alloc<uint8_t>(tmp_bytes);
    if (!tmp) return set_error(LMB200_E_CUDA, "out of device memory (sort)");
    if (n) { cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, idx2, (int)n); g_launch_count += 8; }
    tick("prep + morton + sort");
    std::vector<uint8_t> h_valid(n);
    if (n) cudaMemcpy(h_valid.data(), valid, n, cudaMemcpyDeviceToHost);
    uint32_t nv = 0;
    for (uint32_t i = 0; i < n; i++) nv += h_valid[i];
    float extent = 0.f;
    if (nv) for (int k = 0; k < 6; k++) extent = std::max(extent, std::fabs(h_scene[k]));
    const float pad = 1e-4f + 4e-6f * extent;     // as bvh_build.cpp

    tick("count valid");
    // tree
    RadixTree T;
    T.left = D.alloc<uint32_t>(nv); T.right = D.alloc<uint32_t>(nv); T.parent_internal = D.alloc<uint32_t>(nv); T.parent_leaf = D.alloc<uint32_t>(nv);
    T.first = D.alloc<uint32_t>(nv); T.last = D.alloc<uint32_t>(nv);
    Box6* leaf_box = D.alloc<Box6>(nv);
    Box6* node_box = D.alloc<Box6>(nv);
    int* flags = D.alloc<int>(nv);
    WorkItem* q0 = D.alloc<WorkItem>(nv);
    WorkItem* q1 = D.alloc<WorkItem>(nv);
    uint32_t* counters = D.alloc<uint32_t>(4);   // [0] node count, [1] tri count, [2] out queue count
    if (!T.left || !T.right || !T.parent_internal || !T.parent_leaf || !T.first || !T.last || !leaf_box || !node_box || !flags || !q0 || !q1 || !counters)
        return set_error(LMB200_E_CUDA, "out of device memory (gpu build)");
    // scene grid for the 16-bit node origins (as build_bvh): the padded bounds of the valid triangles
    {
        float glo[3], ghi[3];
        for (int k = 0; k < 3; k++) { glo[k] = nv ? h_scene[k] - pad : 0.f; ghi[k] = nv ? h_scene[3 + k] + pad : 0.f; }
        make_scene_grid(glo, ghi, a->bvh.grid);
    }
    tick("alloc 2");
    // every wide node has >= 2 children or is the root, so there are at most nv nodes; plus nv triangle units
    const size_t unit_cap = 2 * (size_t)std::max<uint32_t>(nv, 1) + 1;
    if ((e = cudaMalloc(&a->d_units, unit_cap * sizeof(Unit64))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(units)");
    uint32_t num_nodes = 1, num_units = 1;
    int depth = 1;
    if (nv == 0) {
        Unit64 root; memset(&root, 0, sizeof(root));
        root.node.e[0] = root.node.e[1] = root.node.e[2] = 127;
        for (int s = 0; s < 8; s++) for (int ax = 0; ax < 3; ax++) { root.node.qlo[ax][s] = 255; root.node.qhi[ax][s] = 0; }
        cudaMemcpy(a->d_units, &root, sizeof(root), cudaMemcpyHostToDevice);
    } else if (ploc) {
        // ---- PLOC: bottom-up clustering of the Morton-ordered triangles, collapse tables filled on the way ----
        PlocTree P;
        P.left = D.alloc<uint32_t>(nv); P.right = D.alloc<uint32_t>(nv); P.count = D.alloc<uint32_t>(nv);
        P.box = node_box; P.cost = D.alloc<float>(7 * (size_t)nv); P.choice = D.alloc<unsigned long long>(nv);
        uint32_t* cl[2] = {D.alloc<uint32_t>(nv), D.alloc<uint32_t>(nv)};
        Box6* cbx[2] = {D.alloc<Box6>(nv), D.alloc<Box6>(nv)};
        uint32_t* nn = D.alloc<uint32_t>(nv);
        unsigned long long* fl = D.alloc<unsigned long long>(nv + 1);
        unsigned long long* sc = D.alloc<unsigned long long>(nv + 1);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, fl, sc, (int)nv + 1);
        void* scan_tmp = D.alloc<uint8_t>(scan_bytes);
        if (!P.left || !P.right || !P.count || !P.cost || !P.choice || !cl[0] || !cl[1] || !cbx[0] || !cbx[1] || !nn || !fl || !sc || !scan_tmp)
            return set_error(LMB200_E_CUDA, "out of device memory (gpu build, clustering)");
        k_leaf_boxes<<<(nv + TB - 1) / TB, TB>>>(boxes, idx2, nv, pad, leaf_box); g_launch_count++;
        k_ploc_init<<<(nv + TB - 1) / TB, TB>>>(nv, leaf_box, cl[0], cbx[0]); g_launch_count++;
        uint32_t nc = nv, node_base = 0;
        int cur = 0, rounds = 0;
        while (nc > 1) {
            int force = 0;
            for (;;) {
                k_ploc_nn<<<(nc + 255) / 256, 256>>>(cbx[cur], nc, nn);
                k_ploc_flags<<<(nc + TB - 1) / TB, TB>>>(nn, nc, force, fl);
                cudaMemsetAsync(fl + nc, 0, sizeof(unsigned long long));
                cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, fl, sc, (int)nc + 1);
                g_launch_count += 4;
                unsigned long long tot = 0;
                if ((e = cudaMemcpy(&tot, sc + nc, sizeof(tot), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "gpu clustering");
                const uint32_t merges = (uint32_t)(tot >> 32), survivors = (uint32_t)(tot & 0xffffffffull);
                // mutual nearest neighbours normally pair up a third or more of the clusters; if a round pairs fewer than
                // 1/16, neighbours (2k, 2k+1) are merged instead so that the number of rounds stays logarithmic
                if (!force && merges < std::max<uint32_t>(1u, nc / 16u)) { force = 1; continue; }
                k_ploc_apply<<<(nc + TB - 1) / TB, TB>>>(cl[cur], cbx[cur], nn, nc, force, fl, sc, node_base, cl[cur ^ 1], cbx[cur ^ 1], P, leaf_box);
                g_launch_count++;
                node_base += merges; nc = survivors; cur ^= 1;
                break;
            }
            if (++rounds > 4096) return set_error(LMB200_E_STATE, "gpu build: clustering did not converge");
        }
        uint32_t root_ref = 0;
        if ((e = cudaMemcpy(&root_ref, cl[cur], sizeof(root_ref), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "gpu clustering");
        const uint32_t hc[4] = {1u, 1u, 0u, 0u};
        cudaMemcpy(counters, hc, sizeof(hc), cudaMemcpyHostToDevice);
        const WorkItem rootw{root_ref, 0u};
        cudaMemcpy(q0, &rootw, sizeof(rootw), cudaMemcpyHostToDevice);
        uint32_t n_in = 1;
        WorkItem* qin = q0; WorkItem* qout = q1;
        while (n_in) {
            cudaMemset(counters + 2, 0, sizeof(uint32_t));
            k_emit_dp<<<(n_in + 127) / 128, 128>>>(qin, n_in, qout, counters + 2, P, leaf_box, idx2, recs,
                                                   reinterpret_cast<Unit64*>(a->d_units), counters, counters + 1, a->bvh.grid, (int)nv);
            g_launch_count++;
            uint32_t h[3];
            if ((e = cudaMemcpy(h, counters, sizeof(h), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "gpu collapse");
            num_nodes = h[0]; num_units = h[1]; n_in = h[2];
            std::swap(qin, qout);
            if (n_in) depth++;
            if (depth > 64) return set_error(LMB200_E_STATE, "gpu build: tree too deep");
        }
    } else {
        // ---- Morton radix tree (Karras 2012), boxes and collapse tables bottom-up, SAH-optimal collapse top-down ----
        PlocTree P;
        P.left = T.left; P.right = T.right; P.box = node_box; P.count = nullptr; P.cost = nullptr; P.choice = nullptr;
        if (!greedy) {
            P.count = D.alloc<uint32_t>(nv); P.cost = D.alloc<float>(7 * (size_t)nv); P.choice = D.alloc<unsigned long long>(nv);
            if (!P.count || !P.cost || !P.choice) return set_error(LMB200_E_CUDA, "out of device memory (gpu build, collapse tables)");
        }
        cudaMemset(flags, 0, sizeof(int) * nv);
        if (nv > 1) { k_radix_tree<<<(nv - 1 + TB - 1) / TB, TB>>>(keys2, (int)nv, T); g_launch_count++; }
        k_fit<<<(nv + TB - 1) / TB, TB>>>(boxes, idx2, (int)nv, pad, T, leaf_box, node_box, flags); g_launch_count++;
        if (!greedy) {
            cudaMemset(flags, 0, sizeof(int) * nv);
            k_dp_radix<<<(nv + TB - 1) / TB, TB>>>((int)nv, T, P, leaf_box, flags); g_launch_count++;
        }
        const uint32_t hc[4] = {1u, 1u, 0u, 0u};      // [0] nodes, [1] units (the root is unit 0), [2] out queue count
        cudaMemcpy(counters, hc, sizeof(hc), cudaMemcpyHostToDevice);
        const WorkItem rootw{0u, 0u};
        cudaMemcpy(q0, &rootw, sizeof(rootw), cudaMemcpyHostToDevice);
        uint32_t n_in = 1;
        WorkItem* qin = q0; WorkItem* qout = q1;
        while (n_in) {
            cudaMemset(counters + 2, 0, sizeof(uint32_t));
            if (greedy)      // open the largest child first, single-triangle leaves
                k_collapse<<<(n_in + 127) / 128, 128>>>(qin, n_in, qout, counters + 2, T, leaf_box, node_box, idx2, recs,
                                                        reinterpret_cast<Unit64*>(a->d_units), counters, counters + 1, a->bvh.grid, (int)nv);
            else
            k_emit_dp<<<(n_in + 127) / 128, 128>>>(qin, n_in, qout, counters + 2, P, leaf_box, idx2, recs,
                                                   reinterpret_cast<Unit64*>(a->d_units), counters, counters + 1, a->bvh.grid, (int)nv);
            g_launch_count++;
            uint32_t h[3];
            if ((e = cudaMemcpy(h, counters, sizeof(h), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "gpu collapse");
            num_nodes = h[0];
            num_units = h[1];
            n_in = h[2];
            std::swap(qin, qout);
            if (n_in) depth++;
            if (depth > 64) return set_error(LMB200_E_STATE, "gpu build: tree too deep");
        }
    }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return cuda_fail(e, "gpu build");
    tick("tree + collapse");
    // the host mirror (lmb200_accel_host_arrays) is filled lazily from the device arrays
    a->bvh.units.clear();
    a->num_units = num_units;
    a->gpu_built = true;
    a->bvh.stats.num_triangles = ntris;
    a->bvh.stats.num_valid = nv;
    a->bvh.stats.num_nodes = num_nodes;
    a->bvh.stats.sah_cost = 0.f;
    a->bvh.stats.max_depth = depth;
    a->bvh.stats.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return LMB200_OK;
}

}  // namespace lmb200
