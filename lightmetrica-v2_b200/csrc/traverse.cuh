// Device traversal of the 80-byte wide BVH (bvh.h) with the bit-exact TriAccel leaf test.
// Replaces the reference's stack(64) QBVH loop (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:398-497: unordered child push, SSE 4-box slab test) with an octant-ordered
// 8-wide traversal: one thread per ray, a (node-group, triangle-group) pair in registers and a
// short stack of 8-byte entries.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "triaccel.h"

namespace lmb200 {

#define LMB_STACK_SIZE 32

struct TravCounters { uint32_t nodes, tris; };

__device__ __forceinline__ float lmb_safe_inv(float d)
{
    // 1/d with |d| clamped away from zero so that 0 * inf never appears in the slab test. The
    // reference substitutes EpsLarge/Inf for zero components instead (accel_qbvh.cpp:411-416);
    // both make the slab of a zero component span everything iff the origin lies inside it.
    const float tiny = 1.0e-30f;
    const float c = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    return 1.0f / c;
}

__device__ __forceinline__ float lmb_q2f(uint32_t word, uint32_t sel)
{
    // byte `sel&3` of word -> float, via the 2^23 mantissa trick (PRMT + FADD instead of I2F)
    return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f;
}

__device__ __forceinline__ uint32_t lmb_sign_extend_s8x4(uint32_t x)
{
    // 0xff for every byte whose top bit is set, else 0x00. PRMT's sign-replicate mode (selector
    // nibble bit 3) is only reachable through PTX: the __byte_perm intrinsic masks it off.
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

// Intersects the 8 quantised child boxes of one node. Returns the hit mask: bits 31..24 =
// internal children in traversal priority order, bits 23..0 = triangles of hit leaf slots.
__device__ __forceinline__ uint32_t lmb_intersect_node(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4,
                                                       const float ox, const float oy, const float oz,
                                                       const float idx, const float idy, const float idz,
                                                       const bool negx, const bool negy, const bool negz,
                                                       const uint32_t oct_inv4, const float tmin, const float tmax)
{
    const uint32_t ew = __float_as_uint(n0.w);
    const float sx = __uint_as_float((ew & 0xffu) << 23) * idx;
    const float sy = __uint_as_float(((ew >> 8) & 0xffu) << 23) * idy;
    const float sz = __uint_as_float(((ew >> 16) & 0xffu) << 23) * idz;
    const float bx = (n0.x - ox) * idx;
    const float by = (n0.y - oy) * idy;
    const float bz = (n0.z - oz) * idz;

    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
        const uint32_t qlox = __float_as_uint(half == 0 ? n2.x : n2.y);
        const uint32_t qloy = __float_as_uint(half == 0 ? n2.z : n2.w);
        const uint32_t qloz = __float_as_uint(half == 0 ? n3.x : n3.y);
        const uint32_t qhix = __float_as_uint(half == 0 ? n3.z : n3.w);
        const uint32_t qhiy = __float_as_uint(half == 0 ? n4.x : n4.y);
        const uint32_t qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
        const uint32_t nearx = negx ? qhix : qlox, farx = negx ? qlox : qhix;
        const uint32_t neary = negy ? qhiy : qloy, fary = negy ? qloy : qhiy;
        const uint32_t nearz = negz ? qhiz : qloz, farz = negz ? qloz : qhiz;
        // internal slots carry 24+s in their low 5 bits (both bits 3 and 4 set): flip the slot
        // number by the ray octant so that the highest set bit is the nearest child
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = lmb_sign_extend_s8x4(is_inner4 << 3);    // 0xff per inner byte
        const uint32_t bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t sel = 0x7650u | (uint32_t)j;
            const float tnx = fmaf(lmb_q2f(nearx, sel), sx, bx);
            const float tny = fmaf(lmb_q2f(neary, sel), sy, by);
            const float tnz = fmaf(lmb_q2f(nearz, sel), sz, bz);
            const float tfx = fmaf(lmb_q2f(farx, sel), sx, bx);
            const float tfy = fmaf(lmb_q2f(fary, sel), sy, by);
            const float tfz = fmaf(lmb_q2f(farz, sel), sz, bz);
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            if (tn <= tf) {
                const uint32_t bits = (child_bits4 >> (8 * j)) & 0xffu;
                const uint32_t idx5 = (bit_index4 >> (8 * j)) & 0xffu;
                hitmask |= bits << idx5;
            }
        }
    }
    return hitmask;
}

// Closest hit (ANY=false) or any hit (ANY=true). On return for closest: tmax/hu/hv/hid hold the
// winner (hid = 0xffffffff if none). Tie rule: equal t -> larger triangle index wins, which is
// what a linear scan with the reference's "reject t > maxT" rule yields (accel_naive.cpp:92-124).
template <bool ANY, bool COUNT>
__device__ __forceinline__ bool lmb_traverse(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                             const float ox, const float oy, const float oz,
                                             const float dx, const float dy, const float dz,
                                             const float tmin, float& tmax, float& hu, float& hv, uint32_t& hid,
                                             TravCounters* cnt)
{
    const float idx = lmb_safe_inv(dx), idy = lmb_safe_inv(dy), idz = lmb_safe_inv(dz);
    // signs taken from the clamped reciprocal so that -0.0 picks the same near/far planes it scales
    const bool negx = idx < 0.f, negy = idy < 0.f, negz = idz < 0.f;
    const uint32_t oct = (negx ? 1u : 0u) | (negy ? 2u : 0u) | (negz ? 4u : 0u);
    const uint32_t oct_inv4 = (7u - oct) * 0x01010101u;

    uint2 stack[LMB_STACK_SIZE];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);   // root: node 0, "slot 7 ^ oct_inv" resolved below via imask=0
    uint2 tgroup = make_uint2(0u, 0u);
    hid = 0xffffffffu;
    bool found = false;

    for (;;) {
        if (ngroup.y & 0xff000000u) {
            const uint32_t hits_imask = ngroup.y;
            const uint32_t bit = 31u - __clz(hits_imask);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) { stack[sp++] = ngroup; }
            const uint32_t slot = (bit - 24u) ^ (oct_inv4 & 7u);
            const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot) & 0xffu);
            const uint32_t ni = ngroup.x + rel;
            const float4* np = nodes + (size_t)ni * 5u;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) cnt->nodes++;
            const uint32_t hitmask = lmb_intersect_node(n0, n1, n2, n3, n4, ox, oy, oz, idx, idy, idz, negx, negy, negz, oct_inv4, tmin, tmax);
            ngroup.x = __float_as_uint(n1.x);
            ngroup.y = (hitmask & 0xff000000u) | (__float_as_uint(n0.w) >> 24);
            tgroup.x = __float_as_uint(n1.y);
            tgroup.y = hitmask & 0x00ffffffu;
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y) {
            const uint32_t i = __ffs(tgroup.y) - 1;
            tgroup.y &= tgroup.y - 1;
            const float4* tp = tris + (size_t)(tgroup.x + i) * 3u;
            const float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
            if (COUNT) cnt->tris++;
            float t, u, v;
            if (triaccel_intersect(r0, r1, r2, ox, oy, oz, dx, dy, dz, tmin, tmax, t, u, v)) {
                if (ANY) return true;
                const uint32_t id = __float_as_uint(r2.z);
                if (t < tmax || !found || id > hid) { tmax = t; hu = u; hv = v; hid = id; found = true; }
            }
        }

        if ((ngroup.y & 0xff000000u) == 0u) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    return found;
}

}  // namespace lmb200
