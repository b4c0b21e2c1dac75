// Internal declarations shared by the .cu translation units of liblmb200.so.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <string>
#include "../../include/lmb200.h"
#include "bvh.h"
#include "bvh_dev.h"

#ifndef LMB_TRACE_BLOCK
#define LMB_TRACE_BLOCK 128
#endif
#ifndef LMB_TRACE_MIN_BLOCKS
#define LMB_TRACE_MIN_BLOCKS 9     // resident blocks per SM the traversal kernels are compiled for (register cap)
#endif

#define LMB_NBUF 3
// work counters of an accel's persistent trace launches: [0],[1] reserved, [2],[3] the staging streams of the host-buffer
// calls, [4 .. 4+LMB_COUNTER_RING) handed out round-robin to the *_dev entry points, so that calls in flight on
// different streams (or on several scenes sharing one BVH) never share a counter
#define LMB_COUNTER_RING 60
#define LMB_NUM_COUNTERS (4 + LMB_COUNTER_RING)

namespace lmb200 {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;
int set_error(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

struct Service;      // service.cu: the persistent per-ray kernel and its mailboxes

struct Accel {
    int device = -1;
    bool host_only = false;
    bool built = false;
    bool gpu_built = false;     // built by build_bvh_gpu: bvh.nodes/tris mirror is filled on demand
    HostBVH bvh;
    void* d_units = nullptr;     // the flattened BVH: bvh.h Unit64 array (nodes + triangle units), root = unit 0
    uint64_t num_units = 0;
    unsigned long long* d_counter = nullptr;   // LMB_NUM_COUNTERS work counters (see above)
    std::atomic<unsigned> ring{0};
    int ring_slot() { return 4 + (int)(ring.fetch_add(1) % LMB_COUNTER_RING); }
    int num_sms = 148;
    int trace_blocks_per_sm = 4;
    double upload_seconds = 0;
    // staging for the host-pointer entry points
    cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};      // copy-in, kernel (even chunks), copy-out, kernel (odd chunks)
    void* stage_rays[LMB_NBUF] = {};
    void* stage_out[LMB_NBUF] = {};
    cudaEvent_t events[3 * LMB_NBUF] = {};
    uint64_t stage_cap = 0;
    // streaming host-buffer path (accel.cu trace_host_stream): ring staging over the ray index space + per-chunk flags
    void* ring_rays = nullptr; void* ring_out = nullptr;
    uint64_t ring_cap = 0;              // rays the ring holds (a power of two)
    void* d_gate = nullptr;             // device: first[cap + 1] (u64), ready[cap], done[cap] (u32)
    void* h_gate = nullptr;             // mapped pinned: h_done[cap], ctl[2], one[1] (u32)
    uint32_t gate_cap = 0;              // chunks the gate arrays hold
    std::vector<cudaEvent_t> out_events;
    std::mutex stage_mu;          // host-buffer calls on one accel take turns: they share the staging buffers and streams

    Service* service = nullptr;   // per-ray Accel3::Intersect service, created on first use
    std::mutex service_mu;

    ~Accel();
    int upload();
    void free_device();
    int finish_device_setup();   // occupancy / SM count of the device the units live on
};

int build_bvh_gpu(Accel* a, const float* verts_host, uint64_t ntris, int builder);   // bvh_build_gpu.cu
int mirror_to_host(Accel* a);                                            // accel.cu
BvhDev bvh_dev(const Accel* a);                                          // accel.cu
int service_trace_one(Accel* a, const lmb200_ray* ray, lmb200_hit* hit);  // service.cu
void service_destroy(Accel* a);                                          // service.cu: stops the kernel, frees the mailboxes

// n_dev != nullptr: the ray count is read from device memory (wavefront queues).
int trace_closest_dev(Accel* a, const void* rays, void* hits, uint64_t n, const uint32_t* n_dev, cudaStream_t st, int slot);
int trace_any_dev(Accel* a, const void* rays, void* occ, uint64_t n, const uint32_t* n_dev, cudaStream_t st, int slot);

}  // namespace lmb200
