// Wald "TriAccel" projected triangle test, restated for host + device with the exact operation
// order of the reference so results are bit-identical to it (no FMA contraction anywhere: the
// device side uses __f*_rn intrinsics, the host side is compiled with -ffp-contract=off).
//   record layout + precompute : /root/reference/include/lightmetrica/triaccel.h:32-91
//   intersection               : /root/reference/include/lightmetrica/triaccel.h:93-151
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LMB_HD __host__ __device__ __forceinline__
#else
#define LMB_HD inline
#endif

namespace lmb200 {

// 48 bytes, three 16-byte rows so the device fetches it as 3 x LDG.128.
struct alignas(16) TriRecord {
    uint32_t k;      // projection axis 0..2, 3 = degenerate (never hit)
    float n_u, n_v, n_d;
    float a_u, a_v, b_nu, b_nv;
    float c_nu, c_nv;
    uint32_t tri;    // index in build (input) order; the reference keeps (faceIndex, primIndex) here
    uint32_t pad;
};
static_assert(sizeof(TriRecord) == 48, "TriRecord must be 48 bytes");

#if !defined(__CUDA_ARCH__)
// Host-only precompute. The reference evaluates Cross and Dot with SSE (math.h:1791-1802,
// math.h:1728-1731): Cross = (y1*z2 - z1*y2, z1*x2 - x1*z2, x1*y2 - y1*x2) with separately rounded
// products; Dot3 = _mm_dp_ps(.,.,0x71) = (x1*x2 + y1*y2) + (z1*z2 + 0).
inline int triaccel_load(TriRecord& r, const float* A, const float* B, const float* C, uint32_t tri)
{
    static const int waldModulo[4] = {1, 2, 0, 1};
    r.tri = tri;
    r.pad = 0;
    const float b[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    const float c[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    // N = Cross(c, b)
    volatile float m0 = c[1] * b[2], m1 = c[2] * b[0], m2 = c[0] * b[1];
    volatile float s0 = c[2] * b[1], s1 = c[0] * b[2], s2 = c[1] * b[0];
    const float N[3] = {m0 - s0, m1 - s1, m2 - s2};
    uint32_t k = 0;
    for (uint32_t j = 0; j < 3; j++) {
        if (__builtin_fabsf(N[j]) > __builtin_fabsf(N[k])) k = j;
    }
    const int u = waldModulo[k], v = waldModulo[k + 1];
    const float n_k = N[k];
    volatile float d0 = b[u] * c[v], d1 = b[v] * c[u];
    const float denom = d0 - d1;
    if (denom == 0) {
        r.k = 3;
        r.n_u = r.n_v = r.n_d = r.a_u = r.a_v = r.b_nu = r.b_nv = r.c_nu = r.c_nv = 0.f;
        return 1;
    }
    r.k = k;
    r.n_u = N[u] / n_k;
    r.n_v = N[v] / n_k;
    volatile float p0 = A[0] * N[0], p1 = A[1] * N[1], p2 = A[2] * N[2];
    volatile float q0 = p0 + p1, q1 = p2 + 0.0f;
    r.n_d = (q0 + q1) / n_k;
    r.b_nu = b[u] / denom;
    r.b_nv = -b[v] / denom;
    r.a_u = A[u];
    r.a_v = A[v];
    r.c_nu = c[v] / denom;
    r.c_nv = -c[u] / denom;
    return 0;
}
#endif

#if defined(__CUDACC__)
#define LMB_MUL(a, b) __fmul_rn((a), (b))
#define LMB_ADD(a, b) __fadd_rn((a), (b))
#define LMB_SUB(a, b) __fsub_rn((a), (b))
#define LMB_DIV(a, b) __fdiv_rn((a), (b))

// Device test on the three rows of a record. Returns true iff the reference's Intersect would:
// same expression tree, every operation individually rounded to nearest.
__device__ __forceinline__ bool triaccel_intersect(const float4 r0, const float4 r1, const float4 r2,
                                                   const float ox, const float oy, const float oz,
                                                   const float dx, const float dy, const float dz,
                                                   const float mint, const float maxt,
                                                   float& t, float& u, float& v)
{
    const uint32_t k = __float_as_uint(r0.x);
    if (k > 2u) return false;
    float o_u, o_v, o_k, d_u, d_v, d_k;
    if (k == 0u)      { o_u = oy; o_v = oz; o_k = ox; d_u = dy; d_v = dz; d_k = dx; }
    else if (k == 1u) { o_u = oz; o_v = ox; o_k = oy; d_u = dz; d_v = dx; d_k = dy; }
    else              { o_u = ox; o_v = oy; o_k = oz; d_u = dx; d_v = dy; d_k = dz; }
    const float n_u = r0.y, n_v = r0.z, n_d = r0.w;
    const float demon = LMB_ADD(LMB_ADD(LMB_MUL(d_u, n_u), LMB_MUL(d_v, n_v)), d_k);
    if (demon == 0.f) return false;
    t = LMB_DIV(LMB_SUB(LMB_SUB(LMB_SUB(n_d, LMB_MUL(o_u, n_u)), LMB_MUL(o_v, n_v)), o_k), demon);
    if (t < mint || t > maxt) return false;   // NaN t falls through exactly as on the CPU
    const float hu = LMB_SUB(LMB_ADD(o_u, LMB_MUL(t, d_u)), r1.x);
    const float hv = LMB_SUB(LMB_ADD(o_v, LMB_MUL(t, d_v)), r1.y);
    u = LMB_ADD(LMB_MUL(hv, r1.z), LMB_MUL(hu, r1.w));
    v = LMB_ADD(LMB_MUL(hu, r2.x), LMB_MUL(hv, r2.y));
    return u >= 0.f && v >= 0.f && LMB_ADD(u, v) <= 1.f;
}
#endif

}  // namespace lmb200
