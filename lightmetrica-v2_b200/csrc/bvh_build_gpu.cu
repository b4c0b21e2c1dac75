// GPU builder for the same flattened structure as bvh_build.cpp (one array of 64-byte units, bvh.h):
// Morton-ordered binary radix tree (Karras 2012) -> bottom-up box fit -> level-by-level collapse to
// 8-wide nodes with <=3-triangle leaves -> octant slot assignment + quantisation, all on the device.
// Replaces the serial recursive build of the reference (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:202-383: ~2 s per 100 k triangles) when time-to-first-ray matters more than tree
// quality: LBVH trees are not SAH-optimised, so traversal is slower than on the host-built tree
// (DESIGN.md §6). The triangle records are bit-identical to the host path (same operation order,
// __f*_rn intrinsics), hence closest-hit results are identical whichever builder is used.
#include "internal.h"

#include <cub/cub.cuh>
#include <chrono>
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

struct DevBuf {
    std::vector<void*> ptrs;
    template <typename T> T* alloc(size_t n) { void* p = nullptr; if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr; ptrs.push_back(p); return reinterpret_cast<T*>(p); }
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
};

}  // namespace

// Builds on the accel's device from HOST vertices; fills a->d_units and the stats.
int build_bvh_gpu(Accel* a, const float* verts_host, uint64_t ntris)
{
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    service_destroy(a);
    if (a->d_units) { cudaFree(a->d_units); a->d_units = nullptr; a->num_units = 0; }
    if (!a->d_counter && (e = cudaMalloc(&a->d_counter, LMB_NUM_COUNTERS * sizeof(unsigned long long))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(counter)");
    const uint32_t n = (uint32_t)ntris;
    DevBuf D;
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
    if (n && (e = cudaMemcpy(d_verts, verts_host, sizeof(float) * 9 * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "H2D verts");
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
    std::vector<uint8_t> h_valid(n);
    if (n) cudaMemcpy(h_valid.data(), valid, n, cudaMemcpyDeviceToHost);
    uint32_t nv = 0;
    for (uint32_t i = 0; i < n; i++) nv += h_valid[i];
    float extent = 0.f;
    if (nv) for (int k = 0; k < 6; k++) extent = std::max(extent, std::fabs(h_scene[k]));
    const float pad = 1e-4f + 4e-6f * extent;     // as bvh_build.cpp

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
    } else {
        cudaMemset(flags, 0, sizeof(int) * nv);
        if (nv > 1) { k_radix_tree<<<(nv - 1 + TB - 1) / TB, TB>>>(keys2, (int)nv, T); g_launch_count++; }
        k_fit<<<(nv + TB - 1) / TB, TB>>>(boxes, idx2, (int)nv, pad, T, leaf_box, node_box, flags); g_launch_count++;
        const uint32_t hc[4] = {1u, 1u, 0u, 0u};      // [0] nodes, [1] units (the root is unit 0), [2] out queue count
        cudaMemcpy(counters, hc, sizeof(hc), cudaMemcpyHostToDevice);
        const WorkItem rootw{0u, 0u};
        cudaMemcpy(q0, &rootw, sizeof(rootw), cudaMemcpyHostToDevice);
        uint32_t n_in = 1;
        WorkItem* qin = q0; WorkItem* qout = q1;
        while (n_in) {
            cudaMemset(counters + 2, 0, sizeof(uint32_t));
            k_collapse<<<(n_in + 127) / 128, 128>>>(qin, n_in, qout, counters + 2, T, leaf_box, node_box, idx2, recs,
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
