// Flattened wide BVH shared by the host builder and the device traversal.
// Replaces the pointer-chasing QBVH of the reference (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:54-137 node, :202-383 builder) with an 8-wide, 80-byte compressed node in the
// spirit of Ylitie/Karras/Laine 2017: child boxes quantised to 8 bits on a per-node power-of-two
// grid, children stored in octant order so traversal order needs no distance sort.
#pragma once
#include <stdint.h>
#include <vector>
#include "triaccel.h"

namespace lmb200 {

// 80 bytes = 5 x 16-byte rows (fetched as 5 x LDG.128).
struct alignas(16) Node80 {
    float   p[3];          // node box minimum (origin of the quantisation grid)
    uint8_t e[3];          // biased exponents: grid step on axis i = 2^(e[i]-127)
    uint8_t imask;         // bit s set <=> slot s holds an internal child
    uint32_t child_base;   // index of the first internal child (children are contiguous, slot order)
    uint32_t tri_base;     // index of the first triangle referenced by this node's leaf slots
    uint8_t meta[8];       // per slot: 0 empty | internal: 0x20 | (24+s) | leaf: unary count<<5 | tri offset
    uint8_t qlo[3][8];     // quantised child box minima  [axis][slot]
    uint8_t qhi[3][8];     // quantised child box maxima  [axis][slot]
};
static_assert(sizeof(Node80) == 80, "Node80 must be 80 bytes");

struct BuildStats {
    uint64_t num_triangles = 0, num_valid = 0, num_nodes = 0;
    double build_seconds = 0;
    float sah_cost = 0;
    int max_depth = 0;
};

struct HostBVH {
    std::vector<Node80> nodes;       // nodes[0] is the root
    std::vector<TriRecord> tris;     // in leaf order; TriRecord::tri = input index
    std::vector<uint32_t> tri_index; // leaf order -> input index (same as tris[i].tri)
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
    BuildStats stats;
};

// verts: 9 floats per triangle (world space). num_threads <= 0: hardware concurrency.
void build_bvh(const float* verts, uint64_t ntris, HostBVH& out, int num_threads = 0);

}  // namespace lmb200
