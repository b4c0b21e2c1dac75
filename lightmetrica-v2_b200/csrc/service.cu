// Per-ray Accel3::Intersect on the GPU without a kernel launch per ray.
// Replaces the reference's synchronous, per-thread call shape (/root/reference/include/lightmetrica/accel3.h:68,
// called concurrently from every worker thread of Scheduler_::Process, scheduler.cpp:146-175) with a persistent
// SERVICE KERNEL: one block that stays resident while rays keep coming. Host threads post rays into mailboxes in
// mapped pinned host memory; a poller warp watches the mailboxes over PCIe and hands new rays to worker lanes through
// shared memory; a worker lane traverses its ray (the same trav_step as the batch kernels, bit-exact results) and
// writes the hit and a completion stamp straight back into the mailbox, where the host thread is spinning.
// Round 1 launched one kernel and synchronised one stream per ray.
//
// Life cycle: the kernel is started by the first call that finds it not running and exits on its own after
// LMB_SERVICE_IDLE_US without a request, so that a device-wide synchronisation elsewhere in the process (cudaFree,
// cudaDeviceSynchronize) is never blocked for longer than that. Any waiting caller restarts it.
#include "internal.h"
#include "traverse.cuh"

#include <chrono>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace lmb200 {

#define LMB_SERVICE_SLOTS 64            // mailboxes
#define LMB_SERVICE_BLOCKS 8            // service blocks; each owns LMB_SERVICE_SLOTS / LMB_SERVICE_BLOCKS mailboxes
#define LMB_SERVICE_BSLOTS (LMB_SERVICE_SLOTS / LMB_SERVICE_BLOCKS)
// one poller warp + one worker WARP per mailbox: rays of different callers never share a warp, and the 32 lanes of a
// worker warp traverse ONE ray together (wide_traverse below)
#define LMB_SERVICE_THREADS (32 + 32 * LMB_SERVICE_BSLOTS)
#ifndef LMB_SERVICE_IDLE_US
#define LMB_SERVICE_IDLE_US 2000
#endif
#define LMB_WIDE_STACK 512              // pending nodes per ray (shared memory, per worker warp)
#define LMB_WIDE_TRIS 512               // pending triangle tests per ray

// One mailbox = two 64-byte lines in mapped pinned host memory.
// Request line: three 16-byte chunks, EACH carrying the request stamp in its last word. A 16-byte chunk is read with one
// naturally aligned PCIe read, so it is consistent in itself; the poller fetches the three chunks of a mailbox with three
// concurrent loads (ONE round trip) and accepts the ray when all three show the same new stamp - no second, dependent read
// of the ray after seeing a flag. Answer line: hit (16 bytes) and completion stamp written with ONE 32-byte store, so the
// host never sees the stamp without the hit and no system-wide fence is needed between them.
struct alignas(128) ServiceSlot {
    float4 req[3];                 // host -> device: (o.xyz, stamp) (d.xyz, stamp) (tmin, tmax, 0, stamp); stamp as uint bits
    uint32_t pad0[4];
    float4 hit;                    // device -> host
    volatile uint32_t done;        // completion stamp, same 32-byte store as the hit
    uint32_t pad1[11];
};
static_assert(sizeof(ServiceSlot) == 128, "ServiceSlot layout");

struct ServiceShared {
    ServiceSlot slot[LMB_SERVICE_SLOTS];
    volatile uint32_t alive;       // 1 while a service kernel instance may pick up requests (host sets, device clears)
    volatile uint32_t stop;        // host asks the kernel to leave now
};

// device-memory control block of one kernel instance (zeroed by the host before each launch)
struct ServiceCtl {
    unsigned long long last_ns;    // %globaltimer of the last request any block picked up
    uint32_t exiting;              // set by the first block that decides to leave: all blocks leave together, so that a
                                   // mailbox is never left without a poller while `alive` still says 1
    uint32_t leaving;              // blocks that have left
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------
// Latency-oriented traversal of ONE ray by a whole warp. A single lane walking the tree pays one dependent memory round
// trip plus ~300 dependent instructions per visited node (~0.9 us each with nothing else on the SM to hide it: 37 us for an
// incoherent ray in the 4 M-triangle soup). Here every round pops up to 32 pending nodes (nearest on top of the stack) and
// up to 32 pending triangles, one per lane, fetches and tests them all at once and pushes the surviving children: the
// number of sequential rounds is a small multiple of the tree depth instead of the number of visited nodes. The price is
// speculative work (nodes a strictly ordered traversal would have culled), which a service with idle lanes can afford.
// The result is the same closest hit as everywhere else: the triangle test is triaccel_intersect, and the best hit is the
// minimum over (t, larger triangle index first), which does not depend on the order of the tests.
struct WideShared {
    uint32_t stk[LMB_WIDE_STACK + 8];
    uint32_t tl[LMB_WIDE_TRIS + 24];
    unsigned long long key;            // ordered(t) << 32 | (0xfffffffe - triangle id); low word 0xffffffff = no hit yet
    float u, v;
    uint32_t hid;
    uint32_t pad;
};

__device__ __forceinline__ uint32_t ordered_bits(float t) { const uint32_t b = __float_as_uint(t); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float from_ordered(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// returns false if the pending-node stack would overflow (never seen; the caller reports an error)
__device__ __forceinline__ bool wide_traverse(const BvhDev& bvh, const float4 ro, const float4 rd, WideShared* W, float4& result)
{
    const unsigned lane = threadIdx.x & 31u;
    const float ox = ro.x, oy = ro.y, oz = ro.z, tmin = ro.w, dx = rd.x, dy = rd.y, dz = rd.z;
    // a ray with NaN or infinite components is a miss (trav_init does the same for the lane-ordered walk)
    if (!((ox * 0.f + oy * 0.f + oz * 0.f + dx * 0.f + dy * 0.f + dz * 0.f) == 0.f)) { result = make_float4(0.f, 0.f, 0.f, __uint_as_float(LMB200_MISS)); return true; }
    const float idx = lmb_safe_inv(dx), idy = lmb_safe_inv(dy), idz = lmb_safe_inv(dz);
    const uint32_t oct = (idx < 0.f ? 1u : 0u) | (idy < 0.f ? 2u : 0u) | (idz < 0.f ? 4u : 0u);
    const uint32_t oi = 7u - oct;
    const uint32_t one = lmb_one_bits();
    if (lane == 0) { W->key = ((unsigned long long)ordered_bits(rd.w) << 32) | 0xffffffffull; W->hid = LMB200_MISS; W->u = 0.f; W->v = 0.f; W->stk[0] = 0u; }
    __syncwarp();
    uint32_t top = 1, ntri = 0;          // warp-uniform
    while (top | ntri) {
        uint32_t P = min(min(32u, top), min((LMB_WIDE_STACK - min(top, (uint32_t)LMB_WIDE_STACK)) / 7u, (LMB_WIDE_TRIS - min(ntri, (uint32_t)LMB_WIDE_TRIS)) / 24u));
        if (P == 0u && ntri == 0u) return false;
        const uint32_t Tn = min(32u, ntri);
        const bool has_node = lane < P, has_tri = lane < Tn;
        const uint32_t node = has_node ? W->stk[top - 1u - lane] : 0u;
        const uint32_t tri = has_tri ? W->tl[ntri - 1u - lane] : 0u;
        top -= P; ntri -= Tn;
        uint32_t h[8], q[8];
        float4 r0, r1, r2;
        if (has_node) { const float4* np = bvh.units + (size_t)node * 4u; lmb_ld256(np, h); lmb_ld256(np + 2, q); }
        if (has_tri) { const float4* tp = bvh.units + (size_t)tri * 4u; r0 = __ldg(tp); r1 = __ldg(tp + 1); r2 = __ldg(tp + 2); }
        const float tmax = from_ordered((uint32_t)(*reinterpret_cast<volatile unsigned long long*>(&W->key) >> 32));
        uint32_t internal = 0, imask = 0, base = 0, trimask = 0, tribase = 0;
        if (has_node) {
            uint32_t planes[12];
            planes[0] = h[4]; planes[1] = h[5]; planes[2] = h[6]; planes[3] = h[7];
#pragma unroll
            for (int k = 0; k < 8; k++) planes[4 + k] = q[k];
            const float px = fmaf(lmb_k2f<0>(h[0]), bvh.gstep[0], bvh.glo2[0]);
            const float py = fmaf(lmb_k2f<1>(h[0]), bvh.gstep[1], bvh.glo2[1]);
            const float pz = fmaf(lmb_k2f<0>(h[1]), bvh.gstep[2], bvh.glo2[2]);
            const uint32_t hits8 = lmb_intersect_node(px, py, pz, h[2], planes, ox, oy, oz, idx, idy, idz, tmin, tmax, one);
            imask = h[2] >> 24; base = h[3];
            internal = hits8 & imask;
            uint32_t leaf = hits8 & ~imask;
            const uint32_t counts = h[1] >> 16;
            tribase = base + __popc(imask);
            while (leaf) {
                const uint32_t sl = __ffs(leaf) - 1; leaf &= leaf - 1;
                const uint32_t c = (counts >> (2u * sl)) & 3u;
                const uint32_t below = counts & ~(0xffffffffu << (2u * sl));
                trimask |= ((1u << c) - 1u) << (__popc(below & 0x5555u) + 2u * __popc(below & 0xaaaau));
            }
        }
        unsigned long long mykey = ~0ull;
        float mu = 0.f, mv = 0.f;
        if (has_tri) {
            float t, u, v;
            if (triaccel_intersect(r0, r1, r2, ox, oy, oz, dx, dy, dz, tmin, tmax, t, u, v)) {
                mykey = ((unsigned long long)ordered_bits(t) << 32) | (unsigned long long)(0xfffffffeu - __float_as_uint(r2.z));
                mu = u; mv = v;
                atomicMin(&W->key, mykey);
            }
        }
        // children of lane 0's node must end up on top: offsets are suffix sums over the lanes (lane 31 deepest)
        const uint32_t cn = __popc(internal), ct = __popc(trimask);
        uint32_t sn = cn, st = ct;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_down_sync(0xffffffffu, sn, o), b = __shfl_down_sync(0xffffffffu, st, o);
            if (lane + o < 32u) { sn += a; st += b; }
        }
        const uint32_t totn = __shfl_sync(0xffffffffu, sn, 0), tott = __shfl_sync(0xffffffffu, st, 0);
        {
            // far children first: ascending traversal priority (slot ^ (7 - octant)), so the nearest child is written last
            uint32_t pr = 0;
            for (uint32_t m = internal; m; m &= m - 1) pr |= 1u << ((__ffs(m) - 1u) ^ oi);
            uint32_t w = top + (sn - cn);
            while (pr) {
                const uint32_t b = __ffs(pr) - 1u; pr &= pr - 1;
                const uint32_t slot = b ^ oi;
                W->stk[w++] = base + __popc(imask & ~(0xffffffffu << slot));
            }
            uint32_t wt = ntri + (st - ct);
            for (uint32_t m = trimask; m; m &= m - 1) W->tl[wt++] = tribase + (__ffs(m) - 1u);
        }
        top += totn; ntri += tott;
        __syncwarp();
        if (mykey != ~0ull && *reinterpret_cast<volatile unsigned long long*>(&W->key) == mykey) { W->u = mu; W->v = mv; W->hid = __float_as_uint(r2.z); }
        __syncwarp();
    }
    const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&W->key);
    const bool hit = W->hid != LMB200_MISS;
    result = make_float4(hit ? from_ordered((uint32_t)(k >> 32)) : 0.f, W->u, W->v, __uint_as_float(W->hid));
    return true;
}

__global__ void __launch_bounds__(LMB_SERVICE_THREADS, 1)
service_kernel(const BvhDev bvh, ServiceShared* S, ServiceCtl* ctl, const unsigned long long idle_ns)
{
    __shared__ WideShared s_wide[LMB_SERVICE_BSLOTS];
    __shared__ uint2 s_stack[LMB_TRAV_SMEM_UINT2(LMB_SERVICE_BSLOTS)];      // short stacks of the single-lane walks (one column per worker warp)
    __shared__ float4 s_ray[LMB_SERVICE_BSLOTS][2];
    __shared__ volatile uint32_t s_req[LMB_SERVICE_BSLOTS];     // stamp of the ray waiting in s_ray
    __shared__ volatile uint32_t s_exit;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned slot0 = blockIdx.x * LMB_SERVICE_BSLOTS;      // this block's mailboxes: [slot0, slot0 + BSLOTS)
    if (threadIdx.x < LMB_SERVICE_BSLOTS) s_req[threadIdx.x] = S->slot[slot0 + threadIdx.x].done;     // nothing pending below this stamp
    if (threadIdx.x == 0) s_exit = 0;
    trav_lut_init<LMB_SERVICE_BSLOTS>(LMB_SM_BASE(s_stack));
    __syncthreads();

    if (warp == 0) {
        // ---- poller: lane l < BSLOTS watches mailbox slot0 + l over PCIe ----
        const bool mine = lane < LMB_SERVICE_BSLOTS;
        uint32_t seen = mine ? s_req[lane] : 0u;
        if (lane == 0) atomicMax(&ctl->last_ns, global_ns());
        for (;;) {
            bool got = false;
            if (mine) {
                ServiceSlot* m = &S->slot[slot0 + lane];
                const float4 c0 = ld_sys_f4(&m->req[0]), c1 = ld_sys_f4(&m->req[1]), c2 = ld_sys_f4(&m->req[2]);
                const uint32_t r = __float_as_uint(c0.w);
                if (r != seen && __float_as_uint(c1.w) == r && __float_as_uint(c2.w) == r) {
                    s_ray[lane][0] = make_float4(c0.x, c0.y, c0.z, c2.x); s_ray[lane][1] = make_float4(c1.x, c1.y, c1.z, c2.y);
                    __threadfence_block();
                    s_req[lane] = r; seen = r; got = true;
                }
            }
            const unsigned long long now = global_ns();
            if (__any_sync(0xffffffffu, got) && lane == 0) atomicMax(&ctl->last_ns, now);
            bool leave = false;
            if (lane == 31) leave = ld_sys_u32(&S->stop) != 0u;
            if (lane == 30) leave = *reinterpret_cast<volatile uint32_t*>(&ctl->exiting) != 0u;
            if (lane == 29) { const unsigned long long last = *reinterpret_cast<volatile unsigned long long*>(&ctl->last_ns); leave = now > last && now - last > idle_ns; }
            if (__any_sync(0xffffffffu, leave)) break;
        }
        if (lane == 0) { atomicExch(&ctl->exiting, 1u); s_exit = 1; }
    } else {
        // ---- workers: warp w serves mailbox slot0 + w - 1; its 32 lanes traverse the ray together ----
        const unsigned ls = warp - 1u;
        ServiceSlot* m = &S->slot[slot0 + ls];
        uint32_t cur = s_req[ls], served = 0;
        long long cyc_wide = 0, cyc_lane = 0;
        const uint32_t sm_base = LMB_SM_BASE(s_stack);
        for (;;) {
            const uint32_t v = s_req[ls];
            if (v == cur) {
                if (s_exit) break;
                __nanosleep(100);
                continue;
            }
            // Two ways to answer, both bit-identical: the whole warp speculatively in parallel (few sequential rounds, wins
            // when a ray visits far more nodes than the tree is deep) or the first lane alone in strict front-to-back order
            // (fewest nodes, wins on shallow walks). Which one is faster depends on the scene and the rays, so the warp
            // times both on alternating rays for a while and then keeps the faster one, re-sampling every 4096 rays.
            cur = v;
            const bool sampling = (served & 4095u) < 64u;
            if ((served & 4095u) == 0u) { cyc_wide = 0; cyc_lane = 0; }
            const bool use_wide = __shfl_sync(0xffffffffu, (int)(sampling ? (served & 1u) == 0u : cyc_wide <= cyc_lane), 0) != 0;
            const long long c0 = clock64();
            float4 res;
            bool ok = true;
            if (use_wide) ok = wide_traverse(bvh, s_ray[ls][0], s_ray[ls][1], &s_wide[ls], res);
            else if (lane == 0) {
                Trav T;
                TravCounters cnt;
                uint2 lstack[LMB_LOCAL_STACK];
                trav_init(T, s_ray[ls][0], s_ray[ls][1]);
                trav_stack_reset_t<LMB_SERVICE_BSLOTS>(T, sm_base, ls);
                trav_set_lut<LMB_SERVICE_BSLOTS>(T, sm_base);
                while (!trav_step<false, false, LMB_SERVICE_BSLOTS>(T, bvh, sm_base, lstack, cnt, 1u)) {}
                res = make_float4(T.hid != LMB200_MISS ? T.tmax : 0.f, T.hu, T.hv, __uint_as_float(T.hid));
            }
            __syncwarp();
            if (sampling) { if (use_wide) cyc_wide += clock64() - c0; else cyc_lane += clock64() - c0; }
            served++;
            if (lane == 0) {
                // an overflow of the pending-node stack is reported as a hit record the host recognises as an error
                if (!ok) res = make_float4(__int_as_float(0x7fc00001), 0.f, 0.f, __uint_as_float(0xfffffffeu));
                asm volatile("st.volatile.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(&m->hit), "f"(res.x), "f"(res.y), "f"(res.z), "f"(res.w),
                             "f"(__uint_as_float(cur)), "f"(0.f), "f"(0.f), "f"(0.f) : "memory");
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // the LAST block to leave clears `alive` (the host restarts the service only when no instance can pick up requests)
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&ctl->leaving, 1u) == gridDim.x - 1u) { __threadfence_system(); S->alive = 0u; }
    }
}

struct Service {
    int device = -1;
    ServiceShared* host = nullptr;       // mapped pinned
    ServiceShared* dev = nullptr;        // device view of the same memory
    cudaStream_t stream = nullptr;
    std::mutex launch_mu;
    std::mutex slot_mu[LMB_SERVICE_SLOTS];
    std::atomic<uint32_t> next_slot{0};
    ServiceCtl* ctl = nullptr;           // device memory
    uint64_t generation = 0;             // distinguishes services of accels that reuse an address
};

static std::atomic<uint64_t> g_service_generation{1};

static inline void cpu_relax()
{
#if defined(__x86_64__)
    _mm_pause();
#else
    std::this_thread::yield();
#endif
}

int service_create(Accel* a)
{
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    Service* s = new Service;
    s->device = a->device;
    s->generation = g_service_generation.fetch_add(1);
    if ((e = cudaHostAlloc(reinterpret_cast<void**>(&s->host), sizeof(ServiceShared), cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess) { delete s; return cuda_fail(e, "cudaHostAlloc(service mailboxes)"); }
    memset(s->host, 0, sizeof(ServiceShared));
    if ((e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->dev), s->host, 0)) != cudaSuccess) { cudaFreeHost(s->host); delete s; return cuda_fail(e, "cudaHostGetDevicePointer"); }
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) { cudaFreeHost(s->host); delete s; return cuda_fail(e, "cudaStreamCreate"); }
    if ((e = cudaMalloc(reinterpret_cast<void**>(&s->ctl), sizeof(ServiceCtl))) != cudaSuccess) { cudaStreamDestroy(s->stream); cudaFreeHost(s->host); delete s; return cuda_fail(e, "cudaMalloc(service control)"); }
    a->service = s;
    return LMB200_OK;
}

void service_destroy(Accel* a)
{
    Service* s = a->service;
    if (!s) return;
    cudaSetDevice(s->device);
    s->host->stop = 1u;
    cudaStreamSynchronize(s->stream);       // the kernel (if any) leaves at its next poll
    cudaStreamDestroy(s->stream);
    cudaFree(s->ctl);
    cudaFreeHost(s->host);
    delete s;
    a->service = nullptr;
}

// starts a kernel instance unless one is (still) accepting requests
static int service_ensure_running(Accel* a, Service* s)
{
    std::lock_guard<std::mutex> lock(s->launch_mu);
    if (s->host->alive) return LMB200_OK;
    cudaError_t e = cudaSetDevice(a->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    if ((e = cudaMemsetAsync(s->ctl, 0, sizeof(ServiceCtl), s->stream)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(service control)");
    s->host->alive = 1u;
    service_kernel<<<LMB_SERVICE_BLOCKS, LMB_SERVICE_THREADS, 0, s->stream>>>(bvh_dev(a), s->dev, s->ctl, (unsigned long long)LMB_SERVICE_IDLE_US * 1000ull);
    g_launch_count++;
    if ((e = cudaGetLastError()) != cudaSuccess) { s->host->alive = 0u; return cuda_fail(e, "service_kernel launch"); }
    return LMB200_OK;
}

int service_trace_one(Accel* a, const lmb200_ray* ray, lmb200_hit* hit)
{
    if (!a->service) {
        std::lock_guard<std::mutex> lock(a->service_mu);
        if (!a->service) { if (const int rc = service_create(a)) return rc; }
    }
    Service* s = a->service;
    // a thread keeps the mailbox it was given for this service; more threads than mailboxes share them under a lock
    static thread_local uint64_t t_gen = 0;
    static thread_local uint32_t t_slot = 0;
    if (t_gen != s->generation) { t_slot = s->next_slot.fetch_add(1) % LMB_SERVICE_SLOTS; t_gen = s->generation; }
    std::lock_guard<std::mutex> slot_lock(s->slot_mu[t_slot]);
    ServiceSlot& m = s->host->slot[t_slot];
    // the request stamp continues this mailbox's sequence (stamps only need to differ from the previous one)
    uint32_t stamp;
    memcpy(&stamp, reinterpret_cast<const char*>(&m.req[0]) + 12, 4);
    stamp += 1u;
    float sf;
    memcpy(&sf, &stamp, 4);
    const float c[3][4] = {{ray->ox, ray->oy, ray->oz, sf}, {ray->dx, ray->dy, ray->dz, sf}, {ray->tmin, ray->tmax, 0.f, sf}};
#if defined(__x86_64__)
    for (int k = 0; k < 3; k++) _mm_store_ps(reinterpret_cast<float*>(&m.req[k]), _mm_loadu_ps(c[k]));      // one 16-byte store per chunk
#else
    memcpy(m.req, c, sizeof(c));
#endif
    std::atomic_thread_fence(std::memory_order_seq_cst);
    if (!s->host->alive) { if (const int rc = service_ensure_running(a, s)) return rc; }
    uint32_t spins = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (m.done != stamp) {
        cpu_relax();
        if ((++spins & 0x3ffu) == 0u) {
            if (!s->host->alive) { if (const int rc = service_ensure_running(a, s)) return rc; }
            if ((spins & 0xfffffu) == 0u && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0) {
                const cudaError_t e = cudaStreamQuery(s->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "service kernel");
                return set_error(LMB200_E_CUDA, "per-ray service did not answer within 20 s");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(hit, const_cast<float4*>(&m.hit), sizeof(lmb200_hit));
    if (hit->tri == 0xfffffffeu) return set_error(LMB200_E_STATE, "per-ray service: pending-node stack overflow");
    return LMB200_OK;
}

}  // namespace lmb200

using namespace lmb200;

// n rays through lmb200_trace_closest_one from `threads` host threads at once (each thread takes a contiguous share and
// calls the per-ray entry point in a loop): the call pattern of the reference's renderers on Accel3::Intersect
// (scheduler.cpp:146-175), as one C call so that it can be timed without a foreign-function layer in the loop.
extern "C" int lmb200_trace_closest_one_mt(lmb200_accel* h, const lmb200_ray* rays, lmb200_hit* hits, uint64_t n, int threads, double* seconds)
{
    Accel* a = reinterpret_cast<Accel*>(h);
    if (!a || (n && (!rays || !hits)) || threads < 1) return set_error(LMB200_E_INVALID, "bad argument");
    if (a->host_only || !a->d_units) return set_error(LMB200_E_STATE, "accel not built on a device");
    std::vector<int> rcs(threads, 0);
    std::vector<std::string> errs(threads);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        for (uint64_t i = n * t / threads; i < n * (t + 1) / threads; i++) {
            const int rc = service_trace_one(a, rays + i, hits + i);
            if (rc) { rcs[t] = rc; errs[t] = g_last_error; return; }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < threads; t++) if (rcs[t]) return set_error(rcs[t], errs[t]);
    return LMB200_OK;
}
