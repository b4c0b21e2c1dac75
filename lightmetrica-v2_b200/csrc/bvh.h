// Flattened wide BVH shared by the host builder, the device builder and the device traversal.
// Replaces the pointer-chasing QBVH of the reference (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:54-137 node, :202-383 builder) with ONE array of 64-byte units:
//   * an 8-wide node in the spirit of Ylitie/Karras/Laine 2017 — child boxes quantised to 8 bits on a
//     per-node power-of-two grid, children stored in octant order so traversal needs no distance sort —
//     squeezed into 64 bytes = two 32-byte sectors of one cache line, fetched with two 256-bit loads
//     (round 1's 80-byte node touched three sectors and took five 128-bit loads; the random-record fetch
//     microbenchmark scripts/micro/l1_wavefront.cu puts the ceiling at 113 G records/s for this shape against
//     81 G/s for the 80-byte one);
//   * a triangle unit: the reference's 48-byte TriAccel record + 16 bytes of padding.
// A node's children are contiguous: first its internal children (slot order), then the triangles of its
// leaf slots (slot order), so a single 32-bit `base` addresses both.
#pragma once
#include <stdint.h>
#include <vector>
#include "triaccel.h"

namespace lmb200 {

// The node origin is stored as three 16-bit coordinates on a scene-wide grid: origin[a] = grid_lo[a] + k[a] * grid_step[a]
// (grid_step a power of two, grid_lo a multiple of it, so the product is exact). The origin is only the zero of the
// node's quantisation grid — the builder quantises child boxes against the decoded origin — so snapping it DOWN
// to the scene grid costs no tightness beyond at most one more grid step of node extent.
struct alignas(64) Node64 {
    uint16_t k[3];         // origin on the scene grid
    uint16_t counts;       // 2 bits per slot: triangles of a leaf slot (1..3); 0 for internal and empty slots
    uint8_t  e[3];         // biased exponents: child-grid step on axis i = 2^(e[i]-127)
    uint8_t  imask;        // bit s set <=> slot s holds an internal child
    uint32_t base;         // unit index of the first child: internal children first, then the leaf slots' triangles
    uint8_t  qlo[3][8];    // quantised child box minima  [axis][slot]   (empty slot: qlo = 255 > qhi = 0)
    uint8_t  qhi[3][8];    // quantised child box maxima  [axis][slot]
};
static_assert(sizeof(Node64) == 64, "Node64 must be 64 bytes");

struct alignas(64) TriUnit {
    TriRecord rec;
    uint32_t pad[4];
};
static_assert(sizeof(TriUnit) == 64, "TriUnit must be 64 bytes");

union alignas(64) Unit64 {
    Node64 node;
    TriUnit tri;
    uint32_t w[16];
};
static_assert(sizeof(Unit64) == 64, "Unit64 must be 64 bytes");

struct SceneGrid {
    float lo[3]   = {0.f, 0.f, 0.f};
    float step[3] = {1.f, 1.f, 1.f};
};

// Chooses the scene grid for the bounds [lo, hi] (already padded): per axis the smallest power-of-two step with
// 65535 steps covering the extent, lo snapped down to a multiple of the step.
void make_scene_grid(const float lo[3], const float hi[3], SceneGrid& g);

// Decoded origin exactly as the device computes it: fma(2^23 + k, step, lo - 2^23 step) == lo + k step.
inline float grid_origin(const SceneGrid& g, int axis, uint32_t k) { return __builtin_fmaf((float)k, g.step[axis], g.lo[axis]); }

// Largest k with grid_origin(k) <= x (clamped to [0, 65535]).
uint16_t grid_floor(const SceneGrid& g, int axis, float x);

struct BuildStats {
    uint64_t num_triangles = 0, num_valid = 0, num_nodes = 0;
    double build_seconds = 0;
    float sah_cost = 0;
    int max_depth = 0;
};

struct HostBVH {
    std::vector<Unit64> units;       // units[0] is the root node
    SceneGrid grid;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    BuildStats stats;
};

// verts: 9 floats per triangle (world space). num_threads <= 0: hardware concurrency.
void build_bvh(const float* verts, uint64_t ntris, HostBVH& out, int num_threads = 0);

}  // namespace lmb200
