// Accel object, persistent trace kernels and the C ABI for the traversal path (include/lmb200.h).
// Replaces Accel::Build / Accel3::Intersect of the reference's in-tree accels
// (/root/reference/src/liblightmetrica/accel/accel_qbvh.cpp:152-497) for ray BATCHES.
#include "internal.h"
#include <cstdlib>
#include <cstdio>
#include "traverse.cuh"

#include <chrono>
#include <immintrin.h>
#include <cstring>
#include <vector>

namespace lmb200 {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};

int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }

int cuda_fail(cudaError_t e, const char* what)
{
    return set_error(LMB200_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------------
// Kernels. Persistent grid of warps; lanes are refilled individually from a global work counter
// (traverse.cuh persistent_trace), so long rays do not strand the rest of their warp.

struct BatchIoClosest {
    const float4* __restrict__ rays; float4* __restrict__ out; uint64_t n;
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ void load(uint64_t i, float4& ro, float4& rd) const { ro = __ldg(rays + 2 * i); rd = __ldg(rays + 2 * i + 1); }
    __device__ __forceinline__ void store(uint64_t i, const Trav& T) const
    {
        const bool hit = T.hid != LMB200_MISS;
        out[i] = make_float4(hit ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid));
    }
};
struct BatchIoAny {
    const float4* __restrict__ rays; uint8_t* __restrict__ out; uint64_t n;
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ void load(uint64_t i, float4& ro, float4& rd) const { ro = __ldg(rays + 2 * i); rd = __ldg(rays + 2 * i + 1); }
    __device__ __forceinline__ void store(uint64_t i, const Trav& T) const { out[i] = T.hid != LMB200_MISS ? 1 : 0; }
};

// compact wire form: 24 bytes per ray (origin, direction), one [tmin, tmax] for the whole batch (lmb200_trace_*_compact)
template <typename Out>
struct BatchIoCompact {
    const float2* __restrict__ rays; Out* __restrict__ out; uint64_t n; float tmin, tmax;
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ void load(uint64_t i, float4& ro, float4& rd) const
    {
        const float2 a = __ldg(rays + 3 * i), b = __ldg(rays + 3 * i + 1), c = __ldg(rays + 3 * i + 2);
        ro = make_float4(a.x, a.y, b.x, tmin); rd = make_float4(b.y, c.x, c.y, tmax);
    }
    __device__ __forceinline__ void store(uint64_t i, const Trav& T) const;
};
template <> __device__ __forceinline__ void BatchIoCompact<float4>::store(uint64_t i, const Trav& T) const
{
    const bool hit = T.hid != LMB200_MISS;
    out[i] = make_float4(hit ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid));
}
template <> __device__ __forceinline__ void BatchIoCompact<uint8_t>::store(uint64_t i, const Trav& T) const { out[i] = T.hid != LMB200_MISS ? 1 : 0; }

// COMPACT: `rays` holds 24-byte rays and (tmin, tmax) apply to all of them
template <bool ANY, bool COUNT, bool COMPACT = false>
__global__ void __launch_bounds__(LMB_TRACE_BLOCK, LMB_TRACE_MIN_BLOCKS)
trace_kernel(const BvhDev bvh,
             const float4* __restrict__ rays, void* __restrict__ out,
             const uint64_t n_host, const uint32_t* __restrict__ n_dev,
             unsigned long long* __restrict__ counter, unsigned long long* __restrict__ work_counters, const float tmin = 0.f, const float tmax = 0.f)
{
    __shared__ uint2 smem[LMB_TRAV_SMEM_UINT2(LMB_TRACE_BLOCK)];
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_host;
    TravCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    if (COMPACT) {
        if (ANY) {
            BatchIoCompact<uint8_t> io{reinterpret_cast<const float2*>(rays), reinterpret_cast<uint8_t*>(out), n, tmin, tmax};
            persistent_trace<true, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
        } else {
            BatchIoCompact<float4> io{reinterpret_cast<const float2*>(rays), reinterpret_cast<float4*>(out), n, tmin, tmax};
            persistent_trace<false, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
        }
    } else if (ANY) {
        BatchIoAny io{rays, reinterpret_cast<uint8_t*>(out), n};
        persistent_trace<true, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
    } else {
        BatchIoClosest io{rays, reinterpret_cast<float4*>(out), n};
        persistent_trace<false, COUNT, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt);
    }
    if (COUNT) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned long long a = cnt.nodes, b = cnt.tris;
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
        if (lane == 0) { atomicAdd(work_counters, a); atomicAdd(work_counters + 1, b); }
    }
}

// Streaming form (trace_host_stream): rays and results live in RINGS over the global ray index (slot = index & mask), filled
// and emptied by the host while the kernel runs. A ring slot is reused within one launch, so the rays are read past L1 (ld.cg).
template <typename Out, bool COMPACT>
struct StreamIo {
    const void* __restrict__ rays; Out* __restrict__ out; uint64_t n, mask; float tmin, tmax;
    __device__ __forceinline__ uint64_t count() const { return n; }
    __device__ __forceinline__ void load(uint64_t i, float4& ro, float4& rd) const
    {
        const uint64_t r = i & mask;
        if (COMPACT) {
            const float2* p = reinterpret_cast<const float2*>(rays) + 3 * r;
            const float2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
            ro = make_float4(a.x, a.y, b.x, tmin); rd = make_float4(b.y, c.x, c.y, tmax);
        } else {
            const float4* p = reinterpret_cast<const float4*>(rays) + 2 * r;
            ro = __ldcg(p); rd = __ldcg(p + 1);
        }
    }
    __device__ __forceinline__ void store(uint64_t i, const Trav& T) const;
};
template <> __device__ __forceinline__ void StreamIo<float4, false>::store(uint64_t i, const Trav& T) const
{ out[i & mask] = make_float4(T.hid != LMB200_MISS ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid)); }
template <> __device__ __forceinline__ void StreamIo<float4, true>::store(uint64_t i, const Trav& T) const
{ out[i & mask] = make_float4(T.hid != LMB200_MISS ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid)); }
template <> __device__ __forceinline__ void StreamIo<uint8_t, false>::store(uint64_t i, const Trav& T) const { out[i & mask] = T.hid != LMB200_MISS ? 1 : 0; }
template <> __device__ __forceinline__ void StreamIo<uint8_t, true>::store(uint64_t i, const Trav& T) const { out[i & mask] = T.hid != LMB200_MISS ? 1 : 0; }

template <bool ANY, bool COMPACT>
__global__ void __launch_bounds__(LMB_TRACE_BLOCK, LMB_TRACE_MIN_BLOCKS)
trace_stream_kernel(const BvhDev bvh, const void* __restrict__ rays, void* __restrict__ out, const uint64_t n, const uint64_t mask,
                    unsigned long long* __restrict__ counter, StreamGate gate, const float tmin, const float tmax)
{
    __shared__ uint2 smem[LMB_TRAV_SMEM_UINT2(LMB_TRACE_BLOCK)];
    __shared__ uint32_t ws[(LMB_TRACE_BLOCK / 32) * LMB_GW_WORDS];
    gate.ws = ws + (threadIdx.x >> 5) * LMB_GW_WORDS;
    TravCounters cnt; cnt.nodes = 0; cnt.tris = 0;
    if (ANY) {
        StreamIo<uint8_t, COMPACT> io{rays, reinterpret_cast<uint8_t*>(out), n, mask, tmin, tmax};
        persistent_trace<true, false, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt, gate);
    } else {
        StreamIo<float4, COMPACT> io{rays, reinterpret_cast<float4*>(out), n, mask, tmin, tmax};
        persistent_trace<false, false, LMB_TRACE_BLOCK>(bvh, io, counter, LMB_SM_BASE(smem), cnt, gate);
    }
}

BvhDev bvh_dev(const Accel* a)
{
    BvhDev b;
    b.units = reinterpret_cast<const float4*>(a->d_units);
    for (int k = 0; k < 3; k++) { b.gstep[k] = a->bvh.grid.step[k]; b.glo2[k] = a->bvh.grid.lo[k] - 8388608.0f * a->bvh.grid.step[k]; }
    return b;
}

template <bool ANY, bool COUNT, bool COMPACT = false>
static int launch_trace(Accel* a, const void* rays, void* out, uint64_t n, const uint32_t* n_dev, cudaStream_t st, unsigned long long* work, int slot,
                        float tmin = 0.f, float tmax = 0.f)
{
    unsigned long long* counter = a->d_counter + slot;
    if (!a->d_units) return set_error(LMB200_E_STATE, "accel not built on a device");
    if (n == 0 && !n_dev) return LMB200_OK;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(counter)");
    const uint64_t warps_needed = (n + 31) / 32;
    uint64_t blocks = (warps_needed + (LMB_TRACE_BLOCK / 32) - 1) / (LMB_TRACE_BLOCK / 32);
    const uint64_t persistent = (uint64_t)a->num_sms * a->trace_blocks_per_sm;
    if (n_dev || blocks > persistent) blocks = persistent;
    if (blocks == 0) blocks = 1;
    trace_kernel<ANY, COUNT, COMPACT><<<(unsigned)blocks, LMB_TRACE_BLOCK, 0, st>>>(
        bvh_dev(a), reinterpret_cast<const float4*>(rays), out, n, n_dev, counter, work, tmin, tmax);
    g_launch_count++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "trace_kernel launch");
    return LMB200_OK;
}

int trace_closest_dev(Accel* a, const void* rays, void* hits, uint64_t n, const uint32_t* n_dev, cudaStream_t st, int slot)
{
    return launch_trace<false, false>(a, rays, hits, n, n_dev, st, nullptr, slot);
}

int trace_any_dev(Accel* a, const void* rays, void* occ, uint64_t n, const uint32_t* n_dev, cudaStream_t st, int slot)
{
    return launch_trace<true, false>(a, rays, occ, n, n_dev, st, nullptr, slot);
}

// ------------------------------------------------------------------------------------------------

void Accel::free_device()
{
    service_destroy(this);       // the service kernel reads d_units: it must be gone first
    if (device >= 0) cudaSetDevice(device);
    if (d_units) cudaFree(d_units);
    if (d_counter) cudaFree(d_counter);
    for (int i = 0; i < LMB_NBUF; i++) {
        if (stage_rays[i]) cudaFree(stage_rays[i]);
        if (stage_out[i]) cudaFree(stage_out[i]);
        stage_rays[i] = stage_out[i] = nullptr;
    }
    if (ring_rays) cudaFree(ring_rays);
    if (ring_out) cudaFree(ring_out);
    if (d_gate) cudaFree(d_gate);
    if (h_gate) cudaFreeHost(h_gate);
    for (cudaEvent_t ev : out_events) cudaEventDestroy(ev);
    out_events.clear();
    ring_rays = ring_out = d_gate = h_gate = nullptr; ring_cap = 0; gate_cap = 0;
    for (int i = 0; i < 4; i++) { if (streams[i]) cudaStreamDestroy(streams[i]); streams[i] = nullptr; }
    for (int i = 0; i < 3 * LMB_NBUF; i++) { if (events[i]) cudaEventDestroy(events[i]); events[i] = nullptr; }
    stage_cap = 0;
    d_units = nullptr; num_units = 0; d_counter = nullptr;
}

Accel::~Accel() { free_device(); }

// The traversal stack holds one entry per level of the wide tree below the root (the not yet visited siblings of the node a
// ray descended into): LMB_SM_STACK - 1 entries in shared memory, pushed without a bounds check. A deeper tree is refused
// here instead of corrupting a neighbouring thread's stack on the device.
static int depth_limit()
{
    // LMB200_DEPTH_LIMIT lowers the limit (tests exercise the refusal / fallback paths with ordinary scenes)
    static const int limit = [] { const char* e = getenv("LMB200_DEPTH_LIMIT"); const int v = e ? atoi(e) : 0; return v >= 1 && v < LMB_SM_STACK ? v : LMB_SM_STACK; }();
    return limit;
}
static bool depth_fits(int max_depth) { return max_depth <= depth_limit(); }
static int check_depth(int max_depth)
{
    if (!depth_fits(max_depth))
        return set_error(LMB200_E_STATE, "BVH depth " + std::to_string(max_depth) + " exceeds the traversal stack capacity of " +
                                             std::to_string(depth_limit()) + " levels");
    return LMB200_OK;
}

int mirror_to_host(Accel* a)
{
    if (!a->gpu_built || !a->bvh.units.empty()) return LMB200_OK;
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    a->bvh.units.resize(a->num_units);
    if ((e = cudaMemcpy(a->bvh.units.data(), a->d_units, a->num_units * sizeof(Unit64), cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "D2H units");
    return LMB200_OK;
}

int Accel::finish_device_setup()
{
    cudaDeviceProp prop;
    const cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    num_sms = prop.multiProcessorCount;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, trace_kernel<false, false>, LMB_TRACE_BLOCK, 0);
    trace_blocks_per_sm = occ > 0 ? occ : 4;
    return LMB200_OK;
}

int Accel::upload()
{
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    service_destroy(this);       // a rebuild invalidates what a running service kernel traverses
    if (d_units) { cudaFree(d_units); d_units = nullptr; }
    num_units = bvh.units.size();
    const size_t nb = num_units * sizeof(Unit64);
    if ((e = cudaMalloc(&d_units, nb)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(units)");
    if (!d_counter && (e = cudaMalloc(&d_counter, LMB_NUM_COUNTERS * sizeof(unsigned long long))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(counter)");
    if ((e = cudaMemcpy(d_units, bvh.units.data(), nb, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "cudaMemcpy(units)");
    if (const int rc = finish_device_setup()) return rc;
    upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return LMB200_OK;
}

// Streaming host-buffer trace (the default for calls of three chunks and more): ONE persistent launch for the whole call.
// Every boundary between the per-chunk launches of the pipeline below costs ~0.2 ms (the ending launch's warps stop refilling
// and run on with fewer and fewer live lanes; profiles/r02_sweep.md) - 14 of them in a 64 Mi-ray call. Here the rays and the
// results live in rings over the ray index (4 chunks deep), the kernel takes the rays of a chunk as soon as its upload has
// landed (traverse.cuh StreamGate) and publishes each chunk's completion in mapped host memory; this thread waits for that
// flag, starts the chunk's download and enqueues the uploads the freed ring space admits.
#ifndef LMB_STREAM_RING_CHUNKS
#define LMB_STREAM_RING_CHUNKS 4
#endif
template <bool ANY, bool COMPACT>
static int trace_host_stream(Accel* a, const uint8_t* rays, void* out, const uint64_t n, const float tmin, const float tmax,
                             const std::vector<uint64_t>& sched, const uint64_t chunk)
{
    const size_t ray_elem = COMPACT ? 24 : sizeof(lmb200_ray), out_elem = ANY ? 1 : sizeof(lmb200_hit);
    const uint32_t C = (uint32_t)sched.size();
    const uint64_t R = (uint64_t)LMB_STREAM_RING_CHUNKS * chunk;      // chunk is a power of two here (the call has several chunks)
    cudaError_t e;
    if (a->ring_cap < R) {
        if (a->ring_rays) cudaFree(a->ring_rays);
        if (a->ring_out) cudaFree(a->ring_out);
        a->ring_rays = a->ring_out = nullptr; a->ring_cap = 0;
        if ((e = cudaMalloc(&a->ring_rays, R * sizeof(lmb200_ray))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(ray ring)");
        if ((e = cudaMalloc(&a->ring_out, R * sizeof(lmb200_hit))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(hit ring)");
        a->ring_cap = R;
    }
    if (a->gate_cap < C) {
        uint32_t cap = 64;
        while (cap < C) cap *= 2;
        if (a->d_gate) cudaFree(a->d_gate);
        if (a->h_gate) cudaFreeHost(a->h_gate);
        a->d_gate = a->h_gate = nullptr; a->gate_cap = 0;
        if ((e = cudaMalloc(&a->d_gate, (cap + 1) * sizeof(unsigned long long) + 2 * cap * sizeof(uint32_t))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(gate)");
        if ((e = cudaHostAlloc(&a->h_gate, (cap + 4) * sizeof(uint32_t), cudaHostAllocMapped)) != cudaSuccess) return cuda_fail(e, "cudaHostAlloc(gate)");
        a->gate_cap = cap;
    }
    while (a->out_events.size() < C) {
        cudaEvent_t ev;
        if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        a->out_events.push_back(ev);
    }
    const uint32_t cap = a->gate_cap;
    unsigned long long* d_first = reinterpret_cast<unsigned long long*>(a->d_gate);
    uint32_t* d_ready = reinterpret_cast<uint32_t*>(d_first + cap + 1);
    uint32_t* d_done = d_ready + cap;
    volatile uint32_t* h_done = reinterpret_cast<volatile uint32_t*>(a->h_gate);
    volatile uint32_t* h_ctl = h_done + cap;
    uint32_t* h_one = const_cast<uint32_t*>(h_done) + cap + 2;
    void* dh = nullptr;
    if ((e = cudaHostGetDevicePointer(&dh, a->h_gate, 0)) != cudaSuccess) return cuda_fail(e, "cudaHostGetDevicePointer");
    std::vector<unsigned long long> first(C + 1, 0ull);
    for (uint32_t c = 0; c < C; c++) first[c + 1] = first[c] + sched[c];
    for (uint32_t c = 0; c < C; c++) h_done[c] = 0u;
    h_ctl[0] = 0u; h_ctl[1] = 0u; *h_one = 1u;

    cudaStream_t s_in = a->streams[0], s_k = a->streams[1], s_out = a->streams[2];
    cudaEvent_t ev_init = a->events[0];
    unsigned long long* counter = a->d_counter + 2;
    int rc = LMB200_OK;
#define LMB_ST(call, what) do { if (!rc) { const cudaError_t e_ = (call); if (e_ != cudaSuccess) rc = cuda_fail(e_, what); } } while (0)
    LMB_ST(cudaMemcpyAsync(d_first, first.data(), (C + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, s_in), "H2D chunk table");
    LMB_ST(cudaMemsetAsync(d_ready, 0, 2 * (size_t)cap * sizeof(uint32_t), s_in), "cudaMemsetAsync(gate flags)");
    LMB_ST(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s_in), "cudaMemsetAsync(counter)");
    LMB_ST(cudaEventRecord(ev_init, s_in), "cudaEventRecord");
    LMB_ST(cudaStreamWaitEvent(s_k, ev_init, 0), "cudaStreamWaitEvent");
    bool launched = false;
    if (!rc) {
        StreamGate g;
        g.first = d_first; g.C = C; g.ready = d_ready; g.done = d_done;
        g.h_done = reinterpret_cast<volatile uint32_t*>(dh); g.h_ctl = g.h_done + cap; g.ws = nullptr;
        const unsigned blocks = (unsigned)((uint64_t)a->num_sms * a->trace_blocks_per_sm);
        trace_stream_kernel<ANY, COMPACT><<<blocks, LMB_TRACE_BLOCK, 0, s_k>>>(bvh_dev(a), a->ring_rays, a->ring_out, n, R - 1, counter, g, tmin, tmax);
        g_launch_count++;
        LMB_ST(cudaGetLastError(), "trace_stream_kernel launch");
        launched = !rc;
    }
    // copies between the caller's buffers and the rings: a chunk may wrap around the end of the ring
    auto ring_copy = [&](const bool up, const uint32_t c, cudaStream_t st) {
        const uint64_t g0 = first[c], m = sched[c], o = g0 & (R - 1), m0 = std::min(m, R - o);
        for (int part = 0; part < 2 && !rc; part++) {
            const uint64_t cnt = part == 0 ? m0 : m - m0, src = part == 0 ? g0 : g0 + m0, slot = part == 0 ? o : 0;
            if (cnt == 0) continue;
            if (up) LMB_ST(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(a->ring_rays) + slot * ray_elem, rays + src * ray_elem, cnt * ray_elem, cudaMemcpyHostToDevice, st), "H2D rays");
            else LMB_ST(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + src * out_elem, reinterpret_cast<uint8_t*>(a->ring_out) + slot * out_elem, cnt * out_elem, cudaMemcpyDeviceToHost, st), "D2H hits");
        }
    };
    uint32_t up = 0, down = 0;      // chunks whose upload / download has been enqueued
    auto pump = [&]() {
        // chunk `up` overwrites the ring slots of indices [first[up] - R, first[up + 1] - R): they must belong to downloaded chunks
        while (!rc && up < C && first[up + 1] <= R + first[down]) {
            if (down > 0) LMB_ST(cudaStreamWaitEvent(s_in, a->out_events[down - 1], 0), "cudaStreamWaitEvent");
            ring_copy(true, up, s_in);
            LMB_ST(cudaMemcpyAsync(d_ready + up, h_one, sizeof(uint32_t), cudaMemcpyHostToDevice, s_in), "H2D ready flag");
            up++;
        }
    };
    pump();
    const auto t_start = std::chrono::steady_clock::now();
    static const bool dbg = getenv("LMB200_STREAM_DEBUG") != nullptr;
    auto since = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() * 1e3; };
    for (uint32_t c = 0; c < C && !rc; c++) {
        const double t_w0 = dbg ? since() : 0;
        // the kernel's flag: every ray of chunk c is finished and its hits are in device memory
        for (uint64_t spins = 0; h_done[c] == 0u; spins++) {
            if (h_ctl[0] != 0u) { rc = set_error(LMB200_E_CUDA, "streaming trace: the kernel gave up waiting for an upload"); break; }
            if ((spins & 0xfffu) == 0xfffu) {
                const cudaError_t q = cudaStreamQuery(s_k);
                if (q == cudaSuccess && h_done[c] == 0u) { rc = set_error(LMB200_E_CUDA, "streaming trace: the kernel ended before all chunks were complete"); break; }
                if (q != cudaSuccess && q != cudaErrorNotReady) { rc = cuda_fail(q, "trace_stream_kernel"); break; }
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 120.0) { rc = set_error(LMB200_E_CUDA, "streaming trace: timed out"); break; }
            }
            _mm_pause();
        }
        if (rc) break;
        const double t_w1 = dbg ? since() : 0;
        ring_copy(false, c, s_out);
        LMB_ST(cudaEventRecord(a->out_events[c], s_out), "cudaEventRecord");
        down = c + 1;
        const double t_w2 = dbg ? since() : 0;
        pump();
        if (dbg) fprintf(stderr, "[stream] chunk %u (%llu rays): waited %.3f ms (from %.3f), download enqueue %.3f ms, pump %.3f ms (uploaded %u)\n", c, (unsigned long long)sched[c], t_w1 - t_w0, t_w0, t_w2 - t_w1, since() - t_w2, up);
    }
#undef LMB_ST
    if (rc && launched) h_ctl[1] = 1u;      // a kernel still waiting for uploads that will not come must leave
    for (int i = 0; i < 4; i++) {
        e = cudaStreamSynchronize(a->streams[i]);
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    if (!rc && h_ctl[0] != 0u) rc = set_error(LMB200_E_CUDA, "streaming trace: the kernel reported an error");
    return rc;
}

// Host-buffer trace: a three-stage pipeline (H2D copy | kernel | D2H copy) over three streams and
#ifndef LMB_E2E_CHUNK_LOG2
#define LMB_E2E_CHUNK_LOG2 23      // rays per full pipeline chunk (tuning override: env LMB200_E2E_CHUNK_LOG2); with the graded schedule below 2^21: 1050, 2^22: 1121, 2^23: 1149 Mrays/s
#endif
// LMB_NBUF staging buffers, so that with pinned host memory the PCIe traffic of chunk k+1 and k-1
// overlaps the kernel of chunk k.
// COMPACT: rays = 24 bytes each (o.xyz, d.xyz), tmin / tmax shared by the batch
template <bool ANY, bool COMPACT = false>
static int trace_host(Accel* a, const void* rays_v, void* out, uint64_t n, float tmin = 0.f, float tmax = 0.f)
{
    const size_t ray_elem = COMPACT ? 24 : sizeof(lmb200_ray);
    const uint8_t* rays = reinterpret_cast<const uint8_t*>(rays_v);
    if (!a || (!rays && n) || (!out && n)) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only) return set_error(LMB200_E_STATE, "host-only accel cannot trace");
    if (!a->d_units) return set_error(LMB200_E_STATE, "accel not built");
    if (n == 0) return LMB200_OK;
    std::lock_guard<std::mutex> lock(a->stage_mu);
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    const size_t out_elem = ANY ? 1 : sizeof(lmb200_hit);
    static const int chunk_log2 = [] { const char* e = getenv("LMB200_E2E_CHUNK_LOG2"); const int v = e ? atoi(e) : 0; return v >= 16 && v <= 26 ? v : LMB_E2E_CHUNK_LOG2; }();
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << chunk_log2);
    // Chunk schedule: full-size chunks in the middle, geometrically smaller ones at both ends (1/8, 1/4, 1/2 of a chunk). The
    // pipeline's fill (first H2D copy, nothing else running) and drain (last D2H copy) then cost an eighth of what a
    // full chunk costs: with 64 Mi rays and 4 Mi-ray chunks that is ~2.5 ms of 61 ms.
    // LMB200_E2E_STREAM=0 selects the per-chunk launches below for every call (they also serve calls of one or two chunks)
    static const bool stream_on = [] { const char* e = getenv("LMB200_E2E_STREAM"); return !(e && atoi(e) == 0); }();
    std::vector<uint64_t> sched;
    static const uint64_t grade_env = [] { const char* e = getenv("LMB200_E2E_GRADE"); const int v = e ? atoi(e) : 0; return (uint64_t)(v >= 1 && v <= 1024 ? v : 0); }();
    // mid-size calls stream too, with a chunk scaled down to a quarter of the call (a call of one or two full chunks would run
    // upload, kernel and download one after the other)
    uint64_t chunk_s = chunk;
    while (chunk_s > (1u << 17) && n < 4 * chunk_s) chunk_s /= 2;      // below 512 Ki rays the plain three-step call is as fast
    const bool streaming = stream_on && n >= 4 * chunk_s && (chunk_s & (chunk_s - 1)) == 0 && chunk_s >= 4096;
    if (streaming) {
        const uint64_t chunk = chunk_s;      // (shadows the pipeline's chunk inside this block)
        // The streaming launch pays nothing per chunk but a few small copies. Its schedule starts with chunk / 64 and grows by
        // 5/4 per chunk: the kernel can start a chunk only when ALL of it has been uploaded, and uploads run only ~1.3x faster
        // than the kernel eats rays, so a chunk twice the size of everything before it (the schedule of the per-chunk launches
        // below) makes the kernel wait for its upload (~2 ms in a 64 Mi-ray call). The tail halves down to chunk / 64: the last
        // download is all that is not overlapped.
        const uint64_t small = chunk / (grade_env ? grade_env : 64);
        std::vector<uint64_t> tail;
        uint64_t tail_sum = 0;
        for (uint64_t c = chunk / 2; c >= std::max<uint64_t>(small, 1); c /= 2) { tail.push_back(c); tail_sum += c; }
        uint64_t left = n - tail_sum;
        for (uint64_t c = std::max<uint64_t>(small, 1); c < chunk && left > chunk; c = (c * 5 / 4 + 1023) & ~1023ull) { sched.push_back(c); left -= c; }
        while (left > 0) { const uint64_t m = std::min(chunk, left); sched.push_back(m); left -= m; }
        for (uint64_t c : tail) sched.push_back(c);
    } else {
        uint64_t left = n;
        std::vector<uint64_t> tail;
        const uint64_t grade = grade_env ? grade_env : 8;
        for (uint64_t c = std::max<uint64_t>(chunk / grade, 1); c < chunk && left > 2 * chunk; c *= 2) {
            sched.push_back(c); tail.push_back(c); left -= 2 * c;
        }
        while (left > 0) { const uint64_t m = std::min(chunk, left); sched.push_back(m); left -= m; }
        for (size_t k = tail.size(); k-- > 0;) sched.push_back(tail[k]);
    }
    for (int i = 0; i < 4; i++) {
        if (!a->streams[i] && (e = cudaStreamCreateWithFlags(&a->streams[i], cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
    }
    for (int i = 0; i < 3 * LMB_NBUF; i++) {
        if (!a->events[i] && (e = cudaEventCreateWithFlags(&a->events[i], cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
    }
    if (streaming) return trace_host_stream<ANY, COMPACT>(a, rays, out, n, tmin, tmax, sched, chunk_s);
    if (a->stage_cap < chunk) {
        for (int i = 0; i < LMB_NBUF; i++) {
            if (a->stage_rays[i]) cudaFree(a->stage_rays[i]);
            if (a->stage_out[i]) cudaFree(a->stage_out[i]);
            a->stage_rays[i] = a->stage_out[i] = nullptr;
            if ((e = cudaMalloc(&a->stage_rays[i], chunk * sizeof(lmb200_ray))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage rays)");
            if ((e = cudaMalloc(&a->stage_out[i], chunk * sizeof(lmb200_hit))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(stage out)");
        }
        a->stage_cap = chunk;
    }
    // kernels of consecutive chunks go to two alternating streams (own work counters: slots 2 and 3), so that the next
    // chunk's persistent blocks move in while the previous chunk's last rays drain
    cudaStream_t s_in = a->streams[0], s_out = a->streams[2];
    // a failure in the middle of the pipeline must not return while earlier chunks are still in flight on the
    // caller's buffers: every exit path drains the four streams first
    int rc = LMB200_OK;
    uint64_t c = 0, off = 0;
    for (; c < sched.size() && !rc; off += sched[c], c++) {
        const int b = (int)(c % LMB_NBUF);
        cudaEvent_t ev_in = a->events[3 * b], ev_k = a->events[3 * b + 1], ev_out = a->events[3 * b + 2];
        const uint64_t m = sched[c];
        if (c >= LMB_NBUF && (e = cudaStreamWaitEvent(s_in, ev_out, 0)) != cudaSuccess) { rc = cuda_fail(e, "cudaStreamWaitEvent"); break; }   // buffer b is free again
        if ((e = cudaMemcpyAsync(a->stage_rays[b], rays + off * ray_elem, m * ray_elem, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) { rc = cuda_fail(e, "H2D rays"); break; }
        if ((e = cudaEventRecord(ev_in, s_in)) != cudaSuccess) { rc = cuda_fail(e, "cudaEventRecord"); break; }
        cudaStream_t s_k = a->streams[(c & 1) ? 3 : 1];
        const int slot = (c & 1) ? 3 : 2;
        if ((e = cudaStreamWaitEvent(s_k, ev_in, 0)) != cudaSuccess) { rc = cuda_fail(e, "cudaStreamWaitEvent"); break; }
        rc = launch_trace<ANY, false, COMPACT>(a, a->stage_rays[b], a->stage_out[b], m, nullptr, s_k, nullptr, slot, tmin, tmax);
        if (rc) break;
        if ((e = cudaEventRecord(ev_k, s_k)) != cudaSuccess) { rc = cuda_fail(e, "cudaEventRecord"); break; }
        if ((e = cudaStreamWaitEvent(s_out, ev_k, 0)) != cudaSuccess) { rc = cuda_fail(e, "cudaStreamWaitEvent"); break; }
        if ((e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(out) + off * out_elem, a->stage_out[b], m * out_elem, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) { rc = cuda_fail(e, "D2H hits"); break; }
        if ((e = cudaEventRecord(ev_out, s_out)) != cudaSuccess) { rc = cuda_fail(e, "cudaEventRecord"); break; }
    }
    for (int i = 0; i < 4; i++) {
        e = cudaStreamSynchronize(a->streams[i]);
        if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
    return rc;
}

}  // namespace lmb200

using namespace lmb200;

extern "C" {

const char* lmb200_last_error(void) { return g_last_error.c_str(); }

int lmb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t lmb200_launch_count(void) { return g_launch_count.load(); }

lmb200_accel* lmb200_accel_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error(LMB200_E_CUDA, "no CUDA device available (lmb200 has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= n) { set_error(LMB200_E_INVALID, "bad device ordinal"); return nullptr; }
    Accel* a = new Accel;
    a->device = device;
    return reinterpret_cast<lmb200_accel*>(a);
}

lmb200_accel* lmb200_accel_create_host_only(void)
{
    Accel* a = new Accel;
    a->host_only = true;
    return reinterpret_cast<lmb200_accel*>(a);
}

void lmb200_accel_destroy(lmb200_accel* a) { delete reinterpret_cast<Accel*>(a); }

static int build_host_sah(Accel* a, const float* verts, uint64_t ntris)
{
    build_bvh(verts, ntris, a->bvh, 0);
    a->built = false;
    a->gpu_built = false;
    if (const int rc = check_depth(a->bvh.stats.max_depth)) return rc;
    a->built = true;
    if (a->host_only) return LMB200_OK;
    return a->upload();
}

int lmb200_accel_build(lmb200_accel* h, const float* verts, uint64_t ntris)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || (!verts && ntris)) return set_error(LMB200_E_INVALID, "null argument");
    if (ntris >= (1ull << 27)) return set_error(LMB200_E_INVALID, "too many triangles (limit 2^27, as the reference's leaf encoding accel_qbvh.cpp:62-72)");
    // the default builder of a device accel is the device builder (milliseconds instead of seconds, 99-101 % of the SAH
    // tree's traversal rate); a host-only accel can only be built on the host
    return lmb200_accel_build_ex(h, verts, ntris, a->host_only ? LMB200_BUILD_HOST_SAH : LMB200_BUILD_DEFAULT);
}

int lmb200_accel_build_ex(lmb200_accel* h, const float* verts, uint64_t ntris, int builder)
{
    if (builder == LMB200_BUILD_HOST_SAH) {
        Accel* a0 = reinterpret_cast<Accel*>(h);
        if (!a0 || (!verts && ntris)) return set_error(LMB200_E_INVALID, "null argument");
        if (ntris >= (1ull << 27)) return set_error(LMB200_E_INVALID, "too many triangles (limit 2^27, as the reference's leaf encoding accel_qbvh.cpp:62-72)");
        return build_host_sah(a0, verts, ntris);
    }
    if (builder != LMB200_BUILD_GPU_LBVH && builder != LMB200_BUILD_GPU_PLOC && builder != LMB200_BUILD_GPU_LBVH_SAH) return set_error(LMB200_E_INVALID, "unknown builder");
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || (!verts && ntris)) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only) return set_error(LMB200_E_STATE, "the GPU builder needs a device accel");
    if (ntris >= (1ull << 27)) return set_error(LMB200_E_INVALID, "too many triangles (limit 2^27)");
    a->built = false;
    int rc = build_bvh_gpu(a, verts, ntris, builder);
    if (rc) return rc;
    // a Morton / clustering tree over adversarial input (long chains) can be deeper than the traversal stack: the binned SAH
    // builder, which falls back to median splits, takes over in that case
    if (!depth_fits(a->bvh.stats.max_depth)) return build_host_sah(a, verts, ntris);
    a->built = true;
    a->upload_seconds = 0;
    return a->finish_device_setup();
}

int lmb200_accel_device(const lmb200_accel* h)
{
    const Accel* a = reinterpret_cast<const Accel*>(h);
    return a && !a->host_only ? a->device : -1;
}

// Replica of a built device accel on another GPU: the node and record arrays are copied device to device (over
// NVLink where the GPUs are peers), so an N-GPU render builds the BVH once instead of N times.
lmb200_accel* lmb200_accel_replicate(const lmb200_accel* h, int device)
{
    const Accel* src = reinterpret_cast<const Accel*>(h);
    if (!src || !src->built || src->host_only || !src->d_units) { set_error(LMB200_E_STATE, "replicate: source accel is not built on a device"); return nullptr; }
    Accel* a = reinterpret_cast<Accel*>(lmb200_accel_create(device));
    if (!a) return nullptr;
    const size_t nb = src->num_units * sizeof(Unit64);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&a->d_units, nb);
    if (e == cudaSuccess) e = cudaMalloc(&a->d_counter, LMB_NUM_COUNTERS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemcpyPeer(a->d_units, device, src->d_units, src->device, nb);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cuda_fail(e, "lmb200_accel_replicate"); delete a; return nullptr; }
    a->num_units = src->num_units;
    a->bvh.stats = src->bvh.stats;
    a->bvh.grid = src->bvh.grid;
    for (int k = 0; k < 3; k++) { a->bvh.scene_lo[k] = src->bvh.scene_lo[k]; a->bvh.scene_hi[k] = src->bvh.scene_hi[k]; }
    a->gpu_built = true;          // host mirror is filled on demand from the device array
    a->built = true;
    if (a->finish_device_setup()) { delete a; return nullptr; }
    return reinterpret_cast<lmb200_accel*>(a);
}

int lmb200_accel_get_stats(const lmb200_accel* h, lmb200_accel_stats* out)
{
    const Accel* a = reinterpret_cast<const Accel*>(h);
    if (!a || !out) return set_error(LMB200_E_INVALID, "null argument");
    if (!a->built) return set_error(LMB200_E_STATE, "accel not built");
    out->num_triangles = a->bvh.stats.num_triangles;
    out->num_valid_triangles = a->bvh.stats.num_valid;
    out->num_nodes = a->bvh.stats.num_nodes;
    out->node_bytes = out->num_nodes * sizeof(Node64);
    out->tri_bytes = a->bvh.stats.num_valid * sizeof(TriUnit);
    out->build_seconds = a->bvh.stats.build_seconds;
    out->upload_seconds = a->upload_seconds;
    out->sah_cost = a->bvh.stats.sah_cost;
    out->max_depth = a->bvh.stats.max_depth;
    return LMB200_OK;
}

int lmb200_accel_host_layout(const lmb200_accel* h, lmb200_bvh_layout* out)
{
    Accel* a = const_cast<Accel*>(reinterpret_cast<const Accel*>(h));
    if (!a || !out) return set_error(LMB200_E_INVALID, "null argument");
    if (!a->built) return set_error(LMB200_E_STATE, "accel not built");
    if (const int rc = mirror_to_host(a)) return rc;
    out->units = a->bvh.units.data();
    out->num_units = a->bvh.units.size();
    out->num_nodes = a->bvh.stats.num_nodes;
    out->num_triangles = a->bvh.stats.num_valid;
    for (int k = 0; k < 3; k++) { out->grid_lo[k] = a->bvh.grid.lo[k]; out->grid_step[k] = a->bvh.grid.step[k]; }
    return LMB200_OK;
}

int lmb200_trace_closest(lmb200_accel* h, const lmb200_ray* rays, lmb200_hit* hits, uint64_t n)
{
    return trace_host<false>(reinterpret_cast<Accel*>(h), rays, hits, n);
}

int lmb200_trace_any(lmb200_accel* h, const lmb200_ray* rays, uint8_t* occluded, uint64_t n)
{
    return trace_host<true>(reinterpret_cast<Accel*>(h), rays, occluded, n);
}

int lmb200_trace_closest_compact(lmb200_accel* h, const float* rays24, float tmin, float tmax, lmb200_hit* hits, uint64_t n)
{
    return trace_host<false, true>(reinterpret_cast<Accel*>(h), rays24, hits, n, tmin, tmax);
}

int lmb200_trace_any_compact(lmb200_accel* h, const float* rays24, float tmin, float tmax, uint8_t* occluded, uint64_t n)
{
    return trace_host<true, true>(reinterpret_cast<Accel*>(h), rays24, occluded, n, tmin, tmax);
}

int lmb200_trace_closest_one(lmb200_accel* h, const lmb200_ray* ray, lmb200_hit* hit)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || !ray || !hit) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only || !a->d_units) return set_error(LMB200_E_STATE, "accel not built on a device");
    return service_trace_one(a, ray, hit);
}

int lmb200_trace_closest_dev(lmb200_accel* h, const void* rays_dev, void* hits_dev, uint64_t n, void* stream)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || (n && (!rays_dev || !hits_dev))) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only) return set_error(LMB200_E_STATE, "host-only accel cannot trace");
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return trace_closest_dev(a, rays_dev, hits_dev, n, nullptr, reinterpret_cast<cudaStream_t>(stream), a->ring_slot());
}

int lmb200_trace_any_dev(lmb200_accel* h, const void* rays_dev, void* occ_dev, uint64_t n, void* stream)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || (n && (!rays_dev || !occ_dev))) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only) return set_error(LMB200_E_STATE, "host-only accel cannot trace");
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return trace_any_dev(a, rays_dev, occ_dev, n, nullptr, reinterpret_cast<cudaStream_t>(stream), a->ring_slot());
}

int lmb200_trace_count_dev(lmb200_accel* h, const void* rays_dev, uint64_t n, double* nodes_per_ray, double* tris_per_ray)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || !rays_dev || !n) return set_error(LMB200_E_INVALID, "null argument");
    if (a->host_only || !a->d_units) return set_error(LMB200_E_STATE, "accel not built on a device");
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    void* scratch = nullptr;
    unsigned long long* work = nullptr;
    if ((e = cudaMalloc(&scratch, n * sizeof(lmb200_hit))) != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
    if ((e = cudaMalloc(&work, 2 * sizeof(unsigned long long))) != cudaSuccess) { cudaFree(scratch); return cuda_fail(e, "cudaMalloc(work)"); }
    cudaMemset(work, 0, 2 * sizeof(unsigned long long));
    int rc = launch_trace<false, true>(a, rays_dev, scratch, n, nullptr, 0, work, a->ring_slot());
    unsigned long long hw[2] = {0, 0};
    if (!rc) {
        e = cudaMemcpy(hw, work, sizeof(hw), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy(work)");
    }
    cudaFree(scratch); cudaFree(work);
    if (rc) return rc;
    if (nodes_per_ray) *nodes_per_ray = (double)hw[0] / (double)n;
    if (tris_per_ray) *tris_per_ray = (double)hw[1] / (double)n;
    return LMB200_OK;
}

}  // extern "C"

