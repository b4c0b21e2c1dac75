// TEST INFRASTRUCTURE — see host_shims.cpp.
#pragma once
#include <lightmetrica/lightmetrica.h>

LM_NAMESPACE_BEGIN
struct RefHost
{
    static int numThreads;   // worker threads of the Scheduler_ shim
    static auto FilmData(const Film* film, int& w, int& h) -> const float*;   // 4 floats per pixel (Vec3 is 16 B)
    static auto RegisterMesh(const float* ps, int nv, const float* ns, const float* ts, const unsigned int* fs, int nf) -> int;
    static auto ClearMeshes() -> void;
};
LM_NAMESPACE_END
