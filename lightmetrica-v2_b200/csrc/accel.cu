// Accel object, persistent trace kernels and the C ABI for the traversal path (include/lmb200.h).
// Replaces Accel::Build / Accel3::Intersect of the reference's in-tree accels
// (/root/reference/src/liblightmetrica/accel/accel_qbvh.cpp:152-497) for ray BATCHES.
#include "internal.h"
#include <cstdlib>
#include "traverse.cuh"

#include <chrono>
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
    std::vector<uint64_t> sched;
    {
        uint64_t left = n;
        std::vector<uint64_t> tail;
        for (uint64_t c = std::max<uint64_t>(chunk / 8, 1); c < chunk && left > 2 * chunk; c *= 2) {
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

