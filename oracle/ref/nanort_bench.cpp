// TEST INFRASTRUCTURE. Throughput baseline named by the task's north star: the reference's *vendored* third-party
// ray tracer nanort (/root/reference/include/nanort/nanort.h, compiled from where it lies), driven with the same call
// sequence as the reference's accel::nanort wrapper (src/liblightmetrica/accel/accel_nanort.cpp:132-136 Build,
// :141-156 Traverse). The wrapper itself is broken in the reference (ignores minT/maxT, swaps prim/face, marked
// "TODO: Falling tests", :66) and nanort's triangle test differs from TriAccel, so this is a SPEED baseline only —
// never a results oracle.
#include <vector>
#include <thread>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#define NANORT_IMPLEMENTATION
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wunused-but-set-variable"
#pragma GCC diagnostic ignored "-Wsign-compare"
#include <nanort/nanort.h>
#pragma GCC diagnostic pop

namespace {
struct NanoScene {
    nanort::BVHAccel accel;
    std::vector<float> ps;
    std::vector<unsigned int> fs;
};
}

extern "C" {

void* ref_nanort_create(const float* verts9, unsigned long long ntris, double* build_seconds)
{
    auto* s = new NanoScene;
    s->ps.assign(verts9, verts9 + 9 * ntris);
    s->fs.resize(3 * ntris);
    for (unsigned long long i = 0; i < 3 * ntris; i++) s->fs[i] = (unsigned int)i;
    const auto t0 = std::chrono::steady_clock::now();
    nanort::BVHBuildOptions options;
    const bool ok = s->accel.Build(s->ps.data(), s->fs.data(), (unsigned int)ntris, options);
    if (build_seconds) *build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) { delete s; return nullptr; }
    return s;
}

void ref_nanort_destroy(void* h) { delete static_cast<NanoScene*>(h); }

// rays: 8 floats (o, tmin, d, tmax) as everywhere else; tmin is ignored by nanort, tmax seeds isect.t as the wrapper does.
// Returns seconds (wall clock, `threads` std::threads over contiguous ray ranges); face_out/t_out optional.
double ref_nanort_trace(void* h, const float* rays, unsigned long long n, int threads, int* face_out, float* t_out)
{
    auto* s = static_cast<NanoScene*>(h);
    const int T = std::max(1, threads);
    auto work = [&](int t)
    {
        const unsigned long long b = n * t / T, e = n * (t + 1) / T;
        for (unsigned long long i = b; i < e; i++)
        {
            const float* r = rays + 8 * i;
            nanort::Ray ray;
            ray.org[0] = r[0]; ray.org[1] = r[1]; ray.org[2] = r[2];
            ray.dir[0] = r[4]; ray.dir[1] = r[5]; ray.dir[2] = r[6];
            nanort::Intersection isect;
            isect.t = r[7];
            nanort::BVHTraceOptions opt;
            const bool hit = s->accel.Traverse(isect, s->ps.data(), s->fs.data(), ray, opt);
            if (face_out) face_out[i] = hit ? (int)isect.faceID : -1;
            if (t_out) t_out[i] = hit ? isect.t : 0.f;
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}
