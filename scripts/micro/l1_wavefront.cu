// Microbenchmark: how does the L1 data pipe charge divergent 16-byte loads?
//  A: every lane reads the five 16-byte pieces of ITS OWN random 80-byte record (5 loads, each with 32 distinct lines)  [= the traversal]
//  B: the same bytes, but consecutive lanes read consecutive pieces of the same record (5 loads, ~6-7 records = 6-12 lines each)
//  C: like A with three 32-byte loads of the 96 aligned bytes
// Records are picked from a table small enough to live in L2 (so DRAM is not the limiter).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <random>
__global__ void kA(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = idx[(tid + it * 7919) % n];
        const float4* p = tab + (size_t)r * 5;
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
        acc += a.x + b.y + c.z + d.w + e.x;
    }
    out[tid] = acc;
}
__global__ void kB(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wbase = tid - lane;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        // the warp still consumes 32 records (160 pieces): piece q = 32 j + lane of the warp's piece list, record q / 5, piece q % 5
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int q = 32 * j + lane;
            const uint32_t r = idx[(wbase + q / 5 + it * 7919) % n];
            const float4 a = __ldg(tab + (size_t)r * 5 + q % 5);
            acc += a.x + a.w;
        }
    }
    out[tid] = acc;
}
__global__ void kC(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = idx[(tid + it * 7919) % n];
        const char* p = reinterpret_cast<const char*>(tab + (size_t)r * 5);
        const char* b = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)31);
        float v[24];
#pragma unroll
        for (int k = 0; k < 3; k++)
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[8*k]), "=f"(v[8*k+1]), "=f"(v[8*k+2]), "=f"(v[8*k+3]), "=f"(v[8*k+4]), "=f"(v[8*k+5]), "=f"(v[8*k+6]), "=f"(v[8*k+7]) : "l"(b + 32 * k));
        acc += v[0] + v[5] + v[10] + v[15] + v[20];
    }
    out[tid] = acc;
}
// D/E: 64-byte records aligned to 64 bytes (two 32-byte sectors, never straddling a line): four 16-byte / two 32-byte loads
__global__ void kD(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = idx[(tid + it * 7919) % n];
        const float4* p = tab + (size_t)r * 4;
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
        acc += a.x + b.y + c.z + d.w;
    }
    out[tid] = acc;
}
__global__ void kE(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = idx[(tid + it * 7919) % n];
        const char* b = reinterpret_cast<const char*>(tab + (size_t)r * 4);
        float v[16];
#pragma unroll
        for (int k = 0; k < 2; k++)
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[8*k]), "=f"(v[8*k+1]), "=f"(v[8*k+2]), "=f"(v[8*k+3]), "=f"(v[8*k+4]), "=f"(v[8*k+5]), "=f"(v[8*k+6]), "=f"(v[8*k+7]) : "l"(b + 32 * k));
        acc += v[0] + v[5] + v[10] + v[15];
    }
    out[tid] = acc;
}
// F: 48-byte records (triangle records), three 16-byte loads
__global__ void kF(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const uint32_t r = idx[(tid + it * 7919) % n];
        const float4* p = tab + (size_t)r * 3;
        const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        acc += a.x + b.y + c.z;
    }
    out[tid] = acc;
}
int main()
{
    const int nrec = 1 << 20;                 // 80 MB of records: L2-resident on B200 (126 MB)
    const int n = 1 << 22;
    std::vector<uint32_t> h(n);
    std::mt19937 g(1);
    for (auto& x : h) x = g() % nrec;
    float4* tab; uint32_t* idx; float* out;
    cudaMalloc(&tab, (size_t)nrec * 80 + 64); cudaMemset(tab, 0, (size_t)nrec * 80 + 64);
    cudaMalloc(&idx, n * 4); cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice);
    const int blocks = 148 * 9, threads = 128, iters = 64;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 6; which++) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (which == 0) kA<<<blocks, threads>>>(tab, idx, n, iters, out);
            if (which == 1) kB<<<blocks, threads>>>(tab, idx, n, iters, out);
            if (which == 2) kC<<<blocks, threads>>>(tab, idx, n, iters, out);
            if (which == 3) kD<<<blocks, threads>>>(tab, idx, n, iters, out);
            if (which == 4) kE<<<blocks, threads>>>(tab, idx, n, iters, out);
            if (which == 5) kF<<<blocks, threads>>>(tab, idx, n, iters, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double recs = (double)blocks * threads * iters;
        static const char* names[6] = {"A own 80 B record, 5 x 16 B", "B 80 B pieces dealt across lanes", "C own 80 B record, 3 x 32 B", "D own 64 B aligned record, 4 x 16 B", "E own 64 B aligned record, 2 x 32 B", "F own 48 B record, 3 x 16 B"};
        static const int bytes[6] = {80, 80, 80, 64, 64, 48};
        printf("%s: %.3f ms  %.1f G records/s  %.0f GB/s useful\n", names[which], best, recs / best / 1e6, recs * bytes[which] / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
