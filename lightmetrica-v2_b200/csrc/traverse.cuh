// Device traversal of the 64-byte-unit wide BVH (bvh.h) with the bit-exact TriAccel leaf test.
// Replaces the reference's stack(64) QBVH loop (/root/reference/src/liblightmetrica/accel/
// accel_qbvh.cpp:398-497: unordered child push, SSE 4-box slab test) with an octant-ordered
// 8-wide traversal written for SIMT efficiency:
//   * persistent warps; a lane whose ray finishes is refilled from the global work counter
//     (warp-level fetch: one atomicAdd per refill for all idle lanes) instead of idling until the
//     slowest ray of its warp is done,
//   * one thread per ray, (node-group, triangle-group) pairs in registers and a short stack of
//     8-byte entries in shared memory (16 levels, a sentinel at the bottom; deeper trees are refused by the host),
//   * a node is one 64-byte-aligned unit fetched with two 256-bit loads (two sectors of one line),
//   * triangle tests are batched across the warp (parked per lane until some lane has to flush: then all lanes that have
//     some), so the triangle code runs with many lanes instead of ~2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "triaccel.h"
#include "bvh_dev.h"

namespace lmb200 {

#ifndef LMB_SM_STACK
#define LMB_SM_STACK 16         // entries per thread in shared memory (one of them the sentinel): trees up to 15 levels
#endif
#define LMB_LOCAL_STACK 1       // (no overflow area any more; the array argument is kept for the call sites)
#ifndef LMB_REFILL_BELOW
#define LMB_REFILL_BELOW 26     // refill the warp when fewer lanes than this are active
#endif

struct TravCounters { uint32_t nodes, tris; };

// 32 bytes with one 256-bit load through the read-only path (sm_100: ld.global.nc.v8)
__device__ __forceinline__ void lmb_ld256(const float4* __restrict__ p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// 16-bit grid coordinate (low or high half of `word`) -> the float 2^23 + k, exactly (bytes of k under the 0x4B exponent)
template <int HIGH>
__device__ __forceinline__ float lmb_k2f(uint32_t word)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(0x4B000000u), "n"(HIGH ? 0x7632 : 0x7610));
    return __uint_as_float(r);
}

__device__ __forceinline__ float lmb_safe_inv(float d)
{
    // 1/d with |d| clamped away from zero so that 0 * inf never appears in the slab test. The
    // reference substitutes EpsLarge/Inf for zero components instead (accel_qbvh.cpp:411-416);
    // both make the slab of a zero component span everything iff the origin lies inside it.
    const float tiny = 1.0e-30f;
    const float c = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    return __fdiv_rn(1.0f, c);      // explicitly IEEE: the traversal must not depend on the unit's -prec-div setting
}

// 0x3F800000 read from constant memory: ptxas cannot fold it, so PRMT takes it as its register /
// constant-bank operand and the byte selectors stay immediates (an immediate here forces every
// selector into a register, one extra move per PRMT).
static __constant__ uint32_t lmb_c_one = 0x3F800000u;
__device__ __forceinline__ uint32_t lmb_one_bits() { return lmb_c_one; }

// Byte J of `word` placed into mantissa bits 15..8 of 1.0f: the float 1 + q * 2^-15, exactly.
// With s' = 2^15 * s and b' = b - s' (per node and axis) the slab distance q*s + b becomes a
// single fma(m, s', b'): PRMT + FFMA per plane instead of PRMT + FADD + FFMA. The rounding of b'
// costs at most 2^-9 grid steps, which the builder covers by widening child boxes by 2^-8 step.
template <int J>
__device__ __forceinline__ float lmb_q2m(uint32_t word, uint32_t one)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word), "r"(one), "n"(0x7604 | (J << 4)));
    return __uint_as_float(r);
}

__device__ __forceinline__ uint32_t lmb_sign_extend_s8x4(uint32_t x)
{
    // 0xff for every byte whose top bit is set, else 0x00. PRMT's sign-replicate mode (selector
    // nibble bit 3) is only reachable through PTX: the __byte_perm intrinsic masks it off.
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}

template <int SLOT>
__device__ __forceinline__ void lmb_child(const uint32_t nearx, const uint32_t neary, const uint32_t nearz,
                                          const uint32_t farx, const uint32_t fary, const uint32_t farz,
                                          const float sx, const float sy, const float sz, const float bx, const float by, const float bz,
                                          const float tmin, const float tmax, const uint32_t one, uint32_t& hits8)
{
    constexpr int J = SLOT & 3;
    const float tnx = fmaf(lmb_q2m<J>(nearx, one), sx, bx);
    const float tny = fmaf(lmb_q2m<J>(neary, one), sy, by);
    const float tnz = fmaf(lmb_q2m<J>(nearz, one), sz, bz);
    const float tfx = fmaf(lmb_q2m<J>(farx, one), sx, bx);
    const float tfy = fmaf(lmb_q2m<J>(fary, one), sy, by);
    const float tfz = fmaf(lmb_q2m<J>(farz, one), sz, bz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
    if (tn <= tf) hits8 |= (1u << SLOT);
}

// Intersects the 8 quantised child boxes of one node; bit s of the result = slot s was hit. Empty slots
// carry an inverted box (qlo = 255 > qhi = 0), so they never contribute. px,py,pz: the node's grid origin; ew: the
// exponent bytes; q[0..11]: the six plane rows (qlo x,y,z then qhi x,y,z), two words (slots 0-3, 4-7) each.
__device__ __forceinline__ uint32_t lmb_intersect_node(const float px, const float py, const float pz, const uint32_t ew,
                                                       const uint32_t* __restrict__ q,
                                                       const float ox, const float oy, const float oz,
                                                       const float idx, const float idy, const float idz,
                                                       const float tmin, const float tmax, const uint32_t one)
{
    // grid step * 2^15 (exponent bytes are biased; the builder keeps e + 15 <= 254)
    const float sx = __uint_as_float(((ew & 0xffu) + 15u) << 23) * idx;
    const float sy = __uint_as_float((((ew >> 8) & 0xffu) + 15u) << 23) * idy;
    const float sz = __uint_as_float((((ew >> 16) & 0xffu) + 15u) << 23) * idz;
    const float bx = fmaf(px - ox, idx, -sx);
    const float by = fmaf(py - oy, idy, -sy);
    const float bz = fmaf(pz - oz, idz, -sz);
    const bool negx = idx < 0.f, negy = idy < 0.f, negz = idz < 0.f;
    uint32_t hits8 = 0;
    {
        const uint32_t qlox = q[0], qloy = q[2], qloz = q[4];
        const uint32_t qhix = q[6], qhiy = q[8], qhiz = q[10];
        const uint32_t nearx = negx ? qhix : qlox, farx = negx ? qlox : qhix;
        const uint32_t neary = negy ? qhiy : qloy, fary = negy ? qloy : qhiy;
        const uint32_t nearz = negz ? qhiz : qloz, farz = negz ? qloz : qhiz;
        lmb_child<0>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<1>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<2>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<3>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
    }
    {
        const uint32_t qlox = q[1], qloy = q[3], qloz = q[5];
        const uint32_t qhix = q[7], qhiy = q[9], qhiz = q[11];
        const uint32_t nearx = negx ? qhix : qlox, farx = negx ? qlox : qhix;
        const uint32_t neary = negy ? qhiy : qloy, fary = negy ? qloy : qhiy;
        const uint32_t nearz = negz ? qhiz : qloz, farz = negz ? qloz : qhiz;
        lmb_child<4>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<5>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<6>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
        lmb_child<7>(nearx, neary, nearz, farx, fary, farz, sx, sy, sz, bx, by, bz, tmin, tmax, one, hits8);
    }
    return hits8;
}

// Traversal-priority order of a slot-order hit mask: bit s moves to bit s ^ o (o = 7 - ray octant; highest bit = nearest child).
// From a table in shared memory: row o of 256 bytes, entry x = bits of x moved from s to s ^ o. One LDS per node step (round 1
// permuted with three masked delta swaps, ~15 ALU-pipe operations). The table lives behind the block's stack area.
#define LMB_LUT_UINT2 256
template <int STRIDE>
__device__ __forceinline__ void trav_lut_init(const uint32_t sm_base)      // all threads of the block, then __syncthreads()
{
    const uint32_t lut = sm_base + 8u * LMB_SM_STACK * STRIDE;
    for (uint32_t i = threadIdx.x; i < 2048u; i += blockDim.x) {
        const uint32_t o = i >> 8, x = i & 255u;
        uint32_t r = 0;
        for (uint32_t b = 0; b < 8u; b++) if ((x >> b) & 1u) r |= 1u << (b ^ o);
        asm volatile("st.shared.u8 [%0], %1;" :: "r"(lut + i), "r"(r) : "memory");
    }
}

// Per-lane traversal state. hid == 0xffffffff <=> no hit yet (triangle ids are < 2^27).
struct Trav {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, tmin, tmax, hu, hv;
    uint32_t hid, oct_inv4;   // oct_inv4: bits 2..0 = 7 - octant, bits 31..8 = shared address of the ray's row of the permutation table (trav_set_lut)
    uint32_t one;      // lmb_one_bits()
    uint2 ngroup;      // x: child_base; y: bits 31..24 pending internal children (priority order) | imask
    uint2 pend;        // parked leaf slots of one node: x = its first triangle unit, y = hit leaf-slot mask | 2-bit counts << 8 (0 = none)
    int sp;
};

__device__ __forceinline__ void trav_init(Trav& T, const float4 ro, const float4 rd)
{
    T.one = lmb_one_bits();
    T.ox = ro.x; T.oy = ro.y; T.oz = ro.z; T.tmin = ro.w;
    T.dx = rd.x; T.dy = rd.y; T.dz = rd.z; T.tmax = rd.w;
    T.idx = lmb_safe_inv(rd.x); T.idy = lmb_safe_inv(rd.y); T.idz = lmb_safe_inv(rd.z);
    // signs taken from the clamped reciprocal so that -0.0 picks the same near/far planes it scales
    const uint32_t oct = (T.idx < 0.f ? 1u : 0u) | (T.idy < 0.f ? 2u : 0u) | (T.idz < 0.f ? 4u : 0u);
    const uint32_t oi = 7u - oct;
    T.oct_inv4 = oi;      // the table row is attached by trav_set_lut
    T.hu = 0.f; T.hv = 0.f; T.hid = 0xffffffffu;
    // A ray with a NaN or infinite origin / direction component never hits anything (every comparison of the triangle test
    // fails) but its slab tests would not cull anything either - NaNs drop out of the min / max - and the walk would
    // visit the whole tree. Such a ray starts in the finished state and reports a miss.
    const bool finite = (ro.x * 0.f + ro.y * 0.f + ro.z * 0.f + rd.x * 0.f + rd.y * 0.f + rd.z * 0.f) == 0.f;
    T.ngroup = make_uint2(0u, finite ? 0x80000000u : 0u);   // root: node 0 (imask 0 resolves to relative index 0)
    T.pend = make_uint2(0u, 0u);
    T.sp = 0;
}

// Shared-memory stack: entry e of thread t lives at shared address sm_base + 8 * (e * STRIDE + t) (conflict-free), LMB_SM_STACK
// levels, no overflow area (builds deeper than LMB_SM_STACK - 1 are refused by the host). T.sp holds the SHARED-SPACE BYTE
// ADDRESS of the next free entry of the thread's column, and a sentinel entry ("no pending child") sits at the bottom of
// every column: a pop of an empty stack returns the sentinel and leaves it in place. A push is one STS and an add, a pop one
// LDS, a test and an add - neither needs the level or the column base (round 1 kept a level index, re-derived the column
// from S2R on every push and pop and spilled levels >= 8 to local memory: ~16 more instructions per node step).
template <int STRIDE>
__device__ __forceinline__ void trav_push(Trav& T, const uint32_t, uint2* __restrict__, const uint2 v)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"((uint32_t)T.sp), "r"(v.x), "r"(v.y) : "memory");
    T.sp += 8 * STRIDE;
}
template <int STRIDE>
__device__ __forceinline__ uint2 trav_pop(Trav& T, const uint32_t, const uint2* __restrict__)
{
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((uint32_t)T.sp - 8u * STRIDE) : "memory");
    if (v.y & 0xff000000u) T.sp -= 8 * STRIDE;
    return v;
}
template <int STRIDE>
__device__ __forceinline__ bool trav_stack_empty(const Trav&, const uint32_t) { return false; }      // a pop is always safe
template <int STRIDE>
__device__ __forceinline__ void trav_stack_reset_t(Trav& T, const uint32_t sm_base, const uint32_t column)
{
    const uint32_t a = sm_base + 8u * column;
    asm volatile("st.shared.v2.b32 [%0], {%1, %1};" :: "r"(a), "r"(0u) : "memory");
    T.sp = (int)(a + 8u * STRIDE);
}

template <int STRIDE>
__device__ __forceinline__ void trav_set_lut(Trav& T, const uint32_t sm_base)
{
    const uint32_t oi = T.oct_inv4 & 7u;
    T.oct_inv4 = oi | ((sm_base + 8u * LMB_SM_STACK * STRIDE + (oi << 8)) << 8);
}
// shared-space address of a block's stack area (call it on the __shared__ array itself so that it folds to a constant)
#define LMB_SM_BASE(arr) ((uint32_t)__cvta_generic_to_shared(arr))

// One traversal step: open the nearest pending node, then maybe test triangles. Every lane of the warp takes every step
// (`lanes` = the full mask; a lane without a ray is in the finished state and its step does nothing), the per-ray service
// calls it from a single lane.
// Triangle work is batched across the warp: the leaf slots hit by a node test are parked in a per-lane pending group
// (two registers) and tested only when some lane must (it hit a second group, or it has nothing else left to do) - then
// by every lane that has parked work. (A second trigger, "at least N lanes have parked work", made no difference for N
// between 8 and 24 and was dropped.)
// Incoherent rays reach a leaf on ~5 % of their node visits, so testing at once would run the
// triangle code on ~80 % of the warp's steps with ~2 of 32 lanes active (ncu, profiles/).
// Returns true when the ray is finished. Closest hit: tie on t -> larger triangle index wins, which
// is what a linear scan with the reference's "reject t > maxT" rule yields (accel_naive.cpp:92-124).
template <bool ANY, bool COUNT, int STRIDE>
__device__ __forceinline__ bool trav_step(Trav& T, const BvhDev& bvh,
                                          const uint32_t smem, uint2* __restrict__ lstack, TravCounters& cnt, const unsigned lanes)
{
    // `lanes`: the lanes of this warp that execute this step together (the caller's ballot of active lanes)
    uint2 fresh = make_uint2(0u, 0u);
    if (T.ngroup.y & 0xff000000u) {
        const uint32_t hits_imask = T.ngroup.y;
        const uint32_t bit = 31u - __clz(hits_imask);
        T.ngroup.y &= ~(1u << bit);
        if (T.ngroup.y & 0xff000000u) trav_push<STRIDE>(T, smem, lstack, T.ngroup);
        const uint32_t slot = (bit - 24u) ^ (T.oct_inv4 & 7u);
        const uint32_t rel = __popc(hits_imask & ~(0xffffffffu << slot) & 0xffu);
        const float4* np = bvh.units + (size_t)(T.ngroup.x + rel) * 4u;
        uint32_t h[8], q[8];
        lmb_ld256(np, h);            // header (4 words) + qlo x, y
        lmb_ld256(np + 2, q);        // qlo z + qhi x, y, z
        if (COUNT) cnt.nodes++;
        uint32_t planes[12];
        planes[0] = h[4]; planes[1] = h[5]; planes[2] = h[6]; planes[3] = h[7];
#pragma unroll
        for (int k = 0; k < 8; k++) planes[4 + k] = q[k];
        const float px = fmaf(lmb_k2f<0>(h[0]), bvh.gstep[0], bvh.glo2[0]);
        const float py = fmaf(lmb_k2f<1>(h[0]), bvh.gstep[1], bvh.glo2[1]);
        const float pz = fmaf(lmb_k2f<0>(h[1]), bvh.gstep[2], bvh.glo2[2]);
        const uint32_t hits8 = lmb_intersect_node(px, py, pz, h[2], planes, T.ox, T.oy, T.oz, T.idx, T.idy, T.idz, T.tmin, T.tmax, T.one);
        const uint32_t imask = h[2] >> 24;
        T.ngroup.x = h[3];
        uint32_t pr;      // hit internal children in traversal-priority order: one table lookup (trav_lut_init)
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(pr) : "r"((T.oct_inv4 >> 8) + (hits8 & imask)));
        T.ngroup.y = (pr << 24) | imask;
        // hit leaf slots are parked as (first triangle unit of the node, slot mask | 2-bit counts << 8): the expansion into
        // triangle offsets is left to the flush, which runs once per parked group instead of once per node step
        const uint32_t leaf = hits8 & ~imask;
        fresh.x = h[3] + __popc(imask);
        fresh.y = leaf ? (leaf | (h[1] >> 16 << 8)) : 0u;
    }
    // next node group
    if ((T.ngroup.y & 0xff000000u) == 0u && !trav_stack_empty<STRIDE>(T, smem)) T.ngroup = trav_pop<STRIDE>(T, smem, lstack);
    const bool no_nodes = (T.ngroup.y & 0xff000000u) == 0u;

    // park the fresh triangles if the pending slot is free
    const bool collide = T.pend.y != 0u && fresh.y != 0u;
    if (T.pend.y == 0u) { T.pend = fresh; fresh.y = 0u; }
    {
        const unsigned must = __ballot_sync(lanes, collide || (no_nodes && T.pend.y != 0u));
        if (must != 0u) {
            // expand the parked slots into a mask of triangle offsets (the triangle units follow the node's internal
            // children in slot order; two count bits per slot), then test them
            uint32_t tmask = 0;
            {
                const uint32_t counts = T.pend.y >> 8;
                for (uint32_t leaf = T.pend.y & 0xffu; leaf; leaf &= leaf - 1) {
                    const uint32_t sl = __ffs(leaf) - 1;
                    const uint32_t c = (counts >> (2u * sl)) & 3u;
                    const uint32_t below = counts & ~(0xffffffffu << (2u * sl));
                    tmask |= ((1u << c) - 1u) << (__popc(below & 0x5555u) + 2u * __popc(below & 0xaaaau));
                }
            }
            T.pend.y = 0u;
            while (tmask) {
                const uint32_t i = __ffs(tmask) - 1;
                tmask &= tmask - 1;
                const float4* tp = bvh.units + (size_t)(T.pend.x + i) * 4u;
                const float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
                if (COUNT) cnt.tris++;
                float t, u, v;
                if (triaccel_intersect(r0, r1, r2, T.ox, T.oy, T.oz, T.dx, T.dy, T.dz, T.tmin, T.tmax, t, u, v)) {
                    if (ANY) { T.hid = 0u; return true; }
                    const uint32_t id = __float_as_uint(r2.z);
                    if (t < T.tmax || T.hid == 0xffffffffu || id > T.hid) { T.tmax = t; T.hu = u; T.hv = v; T.hid = id; }
                }
            }
            if (fresh.y) { T.pend = fresh; }      // the second group of a collision waits for the next batch
        }
    }
    return no_nodes && T.pend.y == 0u;
}

// Shared memory a block needs (uint2 entries): the per-thread short stacks.
#define LMB_TRAV_SMEM_UINT2(block) (LMB_SM_STACK * (block) + LMB_LUT_UINT2)

// ------------------------------------------------------------------------------------------------
// Streaming gate (host-buffer calls, accel.cu trace_host_stream): ONE persistent launch serves a whole host-buffer call whose
// rays arrive chunk by chunk while it runs. The gate lives entirely in the refill path of persistent_trace:
//   * a warp that fetched ray indices of a chunk whose upload has not landed yet (ready[c], written by a 4-byte copy enqueued
//     behind the chunk's upload) keeps the indices, goes on traversing the rays it has, and looks again at its next refill;
//     only a warp with nothing else to do spins on the flag;
//   * it counts, per warp and chunk, the rays fetched and finished, and adds a chunk's finished rays to done[c] once the warp
//     has no ray of that chunk left in flight (one fence + one atomic per warp and chunk); the warp that completes a chunk
//     publishes h_done[c] in mapped host memory, on which the host starts that chunk's download.
// The per-warp bookkeeping is 16 words of shared memory; the common refill touches nothing else. NoGate compiles all of it
// out (every other kernel).
struct NoGate { static constexpr bool enabled = false; };
struct StreamGate {
    static constexpr bool enabled = true;
    const unsigned long long* first;      // [C + 1] first global ray index of every chunk, first[C] = n
    uint32_t C;
    const uint32_t* ready;                // [C] device memory: chunk uploaded
    unsigned int* done;                   // [C] device memory: finished rays per chunk
    volatile uint32_t* h_done;            // [C] mapped host memory: chunk complete (all its hits are in device memory)
    volatile uint32_t* h_ctl;             // mapped host memory: [0] error raised by the kernel, [1] abort requested by the host
    uint32_t* ws;                         // shared memory, 16 words per warp (LMB_GW_*)
};
enum { LMB_GW_CUR = 0, LMB_GW_INFL_PREV, LMB_GW_INFL_CUR, LMB_GW_PEND_PREV, LMB_GW_PEND_CUR, LMB_GW_READY_UPTO /* chunks < this are uploaded */,
       LMB_GW_LO_PREV = 6 /* u64 first[cur-1] */, LMB_GW_LO_CUR = 8 /* u64 first[cur] */, LMB_GW_HI_CUR = 10 /* u64 first[cur+1] */, LMB_GW_WORDS = 16 };
#define LMB_GATE_NONE 0xffffffffffffffffull
#ifndef LMB_GATE_TIMEOUT_CYCLES
#define LMB_GATE_TIMEOUT_CYCLES 20000000000ll      // ~10 s at 1.9 GHz: an upload that never lands ends the kernel with an error
#endif

__device__ __forceinline__ unsigned long long gate_ws64(const StreamGate& g, const int k) { return (unsigned long long)g.ws[k] | ((unsigned long long)g.ws[k + 1] << 32); }
__device__ __forceinline__ void gate_ws64_set(const StreamGate& g, const int k, const unsigned long long v) { g.ws[k] = (uint32_t)v; g.ws[k + 1] = (uint32_t)(v >> 32); }
__device__ __forceinline__ void gate_init(const StreamGate& g, const unsigned lane)      // every lane of the warp
{
    if (lane < (unsigned)LMB_GW_WORDS) g.ws[lane] = 0u;
    __syncwarp();
    if (lane == 0u) gate_ws64_set(g, LMB_GW_HI_CUR, g.first[1]);      // cur = 0: [first[0] = 0, first[1])
    __syncwarp();
}
__device__ __forceinline__ uint32_t gate_chunk_of(const StreamGate& g, const uint64_t i, uint32_t c)
{
    while (i < g.first[c]) c--;                      // first[0] = 0
    while (i >= g.first[c + 1]) c++;                 // i < first[C]
    return c;
}
// every lane of the warp calls it with the same (c, k)
__device__ __forceinline__ void gate_add_done(const StreamGate& g, const uint32_t c, const uint32_t k, const unsigned lane)
{
    if (k == 0u) return;
    __threadfence();                                 // this lane's hit stores before the count that announces them
    __syncwarp();
    if (lane == 0u) {
        const unsigned old = atomicAdd(g.done + c, k);
        if ((unsigned long long)old + k == g.first[c + 1] - g.first[c]) { __threadfence_system(); g.h_done[c] = 1u; }
    }
    __syncwarp();
}
// the finished rays of the lanes in `fin` (ray_index still holds them) are credited to their chunks
__device__ __forceinline__ void gate_account(const StreamGate& g, const bool fin, uint64_t& ray_index, const unsigned lane)
{
    const unsigned m_fin = __ballot_sync(0xffffffffu, fin);
    if (m_fin == 0u) return;
    const uint32_t cur = g.ws[LMB_GW_CUR];
    const bool in_cur = fin && ray_index >= gate_ws64(g, LMB_GW_LO_CUR);
    const bool in_prev = fin && !in_cur && cur > 0u && ray_index >= gate_ws64(g, LMB_GW_LO_PREV);
    const unsigned m_cur = __ballot_sync(0xffffffffu, in_cur), m_prev = __ballot_sync(0xffffffffu, in_prev);
    __syncwarp();
    if (lane == 0u) {
        g.ws[LMB_GW_INFL_CUR] -= __popc(m_cur); g.ws[LMB_GW_PEND_CUR] += __popc(m_cur);
        g.ws[LMB_GW_INFL_PREV] -= __popc(m_prev); g.ws[LMB_GW_PEND_PREV] += __popc(m_prev);
    }
    const unsigned m_other = m_fin & ~(m_cur | m_prev);
    if (m_other) {
        // rays of chunks the warp's two-chunk window has already left (chunks shorter than a ray lives): one by one
        __threadfence();
        if ((m_other >> lane) & 1u) {
            const uint32_t c_old = gate_chunk_of(g, ray_index, cur);
            const unsigned old = atomicAdd(g.done + c_old, 1u);
            if ((unsigned long long)old + 1u == g.first[c_old + 1] - g.first[c_old]) { __threadfence_system(); g.h_done[c_old] = 1u; }
        }
    }
    if (fin) ray_index = LMB_GATE_NONE;
    __syncwarp();
}
// The lanes in `got` hold ray indices they have not loaded yet (`newly`: fetched just now, not counted yet). Moves the
// window to the newest chunk among them, counts the new ones, and reports whether that chunk has been uploaded:
// 0 = yes, load the rays; 1 = not yet (only if `busy`: the warp has other rays to traverse meanwhile); 2 = given up.
__device__ __forceinline__ int gate_acquire(const StreamGate& g, const bool got, const bool newly, const uint64_t i, const bool busy, const unsigned lane)
{
    const unsigned m_got = __ballot_sync(0xffffffffu, got);
    uint32_t cur = g.ws[LMB_GW_CUR];
    int state = 0;
    if (m_got) {
        const uint64_t i_max = __shfl_sync(0xffffffffu, i, 31 - __clz(m_got));      // indices ascend with the lane
        while (i_max >= gate_ws64(g, LMB_GW_HI_CUR)) {                              // move the window one chunk on
            if (cur > 0u) gate_add_done(g, cur - 1u, g.ws[LMB_GW_PEND_PREV], lane); // what is still in flight of cur-1 finishes one by one
            __syncwarp();
            if (lane == 0u) {
                g.ws[LMB_GW_INFL_PREV] = g.ws[LMB_GW_INFL_CUR]; g.ws[LMB_GW_PEND_PREV] = g.ws[LMB_GW_PEND_CUR];
                g.ws[LMB_GW_INFL_CUR] = 0u; g.ws[LMB_GW_PEND_CUR] = 0u; g.ws[LMB_GW_CUR] = cur + 1u;
                gate_ws64_set(g, LMB_GW_LO_PREV, gate_ws64(g, LMB_GW_LO_CUR));
                gate_ws64_set(g, LMB_GW_LO_CUR, gate_ws64(g, LMB_GW_HI_CUR));
                gate_ws64_set(g, LMB_GW_HI_CUR, g.first[cur + 2u]);
            }
            cur++;
            __syncwarp();
        }
        const bool n_cur = newly && i >= gate_ws64(g, LMB_GW_LO_CUR);
        const bool n_prev = newly && !n_cur && cur > 0u && i >= gate_ws64(g, LMB_GW_LO_PREV);
        const unsigned f_cur = __ballot_sync(0xffffffffu, n_cur), f_prev = __ballot_sync(0xffffffffu, n_prev);
        __syncwarp();
        if (lane == 0u) { g.ws[LMB_GW_INFL_CUR] += __popc(f_cur); g.ws[LMB_GW_INFL_PREV] += __popc(f_prev); }
        __syncwarp();
        const uint32_t upto = g.ws[LMB_GW_READY_UPTO];
        __syncwarp();                                // every lane has read it before lane 0 may update it below
        if (upto <= cur) {                           // not known to be uploaded yet: look at the flag (uploads land in chunk order)
            if (lane == 0u) {
                uint32_t r;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(g.ready + cur) : "memory");
                if (r == 0u && !busy) {
                    // nothing else to do: poll the flag, backing off to 32 us (thousands of warps may be waiting here, all on
                    // one address), and look at the host's abort word - a read across PCIe - only every 64th time
                    const long long t0 = clock64();
                    unsigned ns = 500u;
                    for (unsigned it = 1u;; it++) {
                        __nanosleep(ns);
                        if (ns < 32000u) ns *= 2u;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(g.ready + cur) : "memory");
                        if (r != 0u) break;
                        if ((it & 63u) == 0u && (g.h_ctl[1] != 0u || clock64() - t0 > LMB_GATE_TIMEOUT_CYCLES)) { g.h_ctl[0] = 1u; state = 2; break; }
                    }
                }
                if (r != 0u) g.ws[LMB_GW_READY_UPTO] = cur + 1u; else if (state == 0) state = 1;
            }
            state = __shfl_sync(0xffffffffu, state, 0);
        }
    }
    __syncwarp();
    if (cur > 0u && g.ws[LMB_GW_INFL_PREV] == 0u && g.ws[LMB_GW_PEND_PREV] != 0u) {        // the warp's last ray of cur-1 is done: hand the count in
        gate_add_done(g, cur - 1u, g.ws[LMB_GW_PEND_PREV], lane);
        if (lane == 0u) g.ws[LMB_GW_PEND_PREV] = 0u;
        __syncwarp();
    }
    return state;
}
// at the end of the warp's life: everything it fetched is finished
__device__ __forceinline__ void gate_finish(const StreamGate& g, uint64_t& ray_index, const unsigned lane)
{
    gate_account(g, ray_index != LMB_GATE_NONE, ray_index, lane);
    const uint32_t cur = g.ws[LMB_GW_CUR];
    if (cur > 0u) gate_add_done(g, cur - 1u, g.ws[LMB_GW_PEND_PREV], lane);
    gate_add_done(g, cur, g.ws[LMB_GW_PEND_CUR], lane);
}

// Persistent-warp driver. `Io` supplies rays and consumes results:
//   uint64_t count() const;                          number of rays
//   void load(uint64_t i, float4& ro, float4& rd);   ray i
//   void store(uint64_t i, const Trav& T);           result of ray i (T.hid == 0xffffffff: miss)
template <bool ANY, bool COUNT, int STRIDE, typename Io, typename Gate = NoGate>
__device__ __forceinline__ void persistent_trace(const BvhDev& bvh, Io& io,
                                                 unsigned long long* __restrict__ counter, const uint32_t smem, TravCounters& cnt,
                                                 const Gate& gate = Gate())
{
    trav_lut_init<STRIDE>(smem);
    __syncthreads();
    const uint64_t n = io.count();
    const unsigned lane = threadIdx.x & 31u;
    uint2 lstack[LMB_LOCAL_STACK];
    Trav T;
    bool active = false;
    bool exhausted = false;      // warp-uniform: the work counter has run past n
    uint64_t ray_index = Gate::enabled ? LMB_GATE_NONE : 0ull;
    bool waiting = false, any_waiting = false;      // streaming gate only: the lane holds an index whose chunk has not been uploaded yet
    if constexpr (Gate::enabled) gate_init(gate, lane);
    // every lane takes every step (a lane without a ray is in the "finished" state: nothing pending, empty stack, nothing
    // parked - its step does nothing), so the warp votes of a step use the full mask
    T.ngroup = make_uint2(0u, 0u); T.pend = make_uint2(0u, 0u); T.oct_inv4 = 0u; T.hid = 0xffffffffu;
    T.ox = T.oy = T.oz = T.dx = T.dy = T.dz = T.idx = T.idy = T.idz = T.tmin = T.tmax = T.hu = T.hv = 0.f; T.one = lmb_one_bits();
    trav_stack_reset_t<STRIDE>(T, smem, threadIdx.x);
    trav_set_lut<STRIDE>(T, smem);

    for (;;) {
        // ---- refill idle lanes (all 32 lanes converge here) ----
        if constexpr (Gate::enabled) {
            if (!exhausted || any_waiting) {
                const unsigned idle = __ballot_sync(0xffffffffu, !active);
                if (idle) {
                    gate_account(gate, !active && !waiting && ray_index != LMB_GATE_NONE, ray_index, lane);
                    uint64_t i = ray_index;      // a waiting lane keeps the index it fetched earlier
                    bool got = waiting, newly = false;
                    if (!any_waiting && !exhausted) {
                        unsigned long long base = 0;
                        const int leader = __ffs(idle) - 1;
                        if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(idle));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (!active) { i = base + __popc(idle & ((1u << lane) - 1u)); got = i < n; newly = got; }
                        if (base + __popc(idle) >= n) exhausted = true;
                    }
                    const int state = gate_acquire(gate, got, newly, i, idle != 0xffffffffu, lane);
                    if (state == 0) {
                        if (got) {
                            float4 ro, rd;
                            io.load(i, ro, rd);
                            trav_init(T, ro, rd);
                            trav_stack_reset_t<STRIDE>(T, smem, threadIdx.x);
                            trav_set_lut<STRIDE>(T, smem);
                            ray_index = i;
                            active = true;
                        }
                        waiting = false;
                    } else if (state == 1) {
                        if (got) { ray_index = i; waiting = true; }
                    } else {
                        if (got) ray_index = LMB_GATE_NONE;      // an upload never came: these rays are dropped, the call fails
                        waiting = false; exhausted = true;
                    }
                    any_waiting = __any_sync(0xffffffffu, waiting);
                }
            }
        } else if (!exhausted) {
            const unsigned idle = __ballot_sync(0xffffffffu, !active);
            if (idle) {
                unsigned long long base = 0;
                const int leader = __ffs(idle) - 1;
                if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!active) {
                    const uint64_t i = base + __popc(idle & ((1u << lane) - 1u));
                    if (i < n) {
                        float4 ro, rd;
                        io.load(i, ro, rd);
                        trav_init(T, ro, rd);
                        trav_stack_reset_t<STRIDE>(T, smem, threadIdx.x);
                        trav_set_lut<STRIDE>(T, smem);
                        ray_index = i;
                        active = true;
                    }
                }
                if (base + __popc(idle) >= n) exhausted = true;
            }
        }
        if (!__any_sync(0xffffffffu, active)) break;

        // ---- traverse until enough lanes have finished to make a refill worthwhile ----
        unsigned live = __ballot_sync(0xffffffffu, active);
        for (;;) {
            if (trav_step<ANY, COUNT, STRIDE>(T, bvh, smem, lstack, cnt, 0xffffffffu) && active) {
                io.store(ray_index, T);
                active = false;
                T.ngroup.y = 0u; T.pend.y = 0u;      // (an any-hit ray leaves early: put the lane into the finished state)
                trav_stack_reset_t<STRIDE>(T, smem, threadIdx.x);
            }
            live = __ballot_sync(0xffffffffu, active);
            if (live == 0u) break;
            if ((!exhausted || (Gate::enabled && any_waiting)) && __popc(live) < LMB_REFILL_BELOW) break;
        }
    }
    if constexpr (Gate::enabled) gate_finish(gate, ray_index, lane);
}

// Whole-ray helper for single-ray callers (per-ray Accel3::Intersect path).
template <bool ANY, bool COUNT, int STRIDE>
__device__ __forceinline__ bool lmb_traverse(const BvhDev& bvh,
                                             const float4 ro, const float4 rd, Trav& T, const uint32_t smem, TravCounters& cnt)
{
    uint2 lstack[LMB_LOCAL_STACK];
    trav_init(T, ro, rd);
    trav_stack_reset_t<STRIDE>(T, smem, threadIdx.x);
    trav_set_lut<STRIDE>(T, smem);
    while (!trav_step<ANY, COUNT, STRIDE>(T, bvh, smem, lstack, cnt, __activemask())) {}
    return T.hid != 0xffffffffu;
}

}  // namespace lmb200
