// TEST INFRASTRUCTURE — part of oracle/_ref/liblightmetrica.so (never shipped, never on the
// product path). A C API over the UNMODIFIED reference hot path compiled from /root/reference:
// it mirrors the CLI's flow (/root/reference/src/lightmetrica/main.cpp:533-684): load the YAML
// tree, create assets -> accel -> scene -> renderer through ComponentFactory, Render.
// Used by tests/ (as the checker) and by bench.py's cpu_baseline / --impl reference arm only.
#include <pch.h>
#include <lightmetrica/lightmetrica.h>
#include <lightmetrica/triaccel.h>
#include <lightmetrica/intersectionutils.h>
#include "host_shims.h"

using namespace lightmetrica_v2;

namespace {

struct Session
{
    PropertyTree::UniquePtr tree{nullptr, nullptr};
    Assets::UniquePtr assets{nullptr, nullptr};
    Accel::UniquePtr accel{nullptr, nullptr};
    Scene::UniquePtr scene{nullptr, nullptr};
    const PropertyNode* root = nullptr;   // the "lightmetrica" node
    std::string error;
};

std::string g_lastError;

template <typename T>
typename T::UniquePtr CreateConfigurable(const PropertyNode* root, const char* name, const char* typeOverride, const char* defType, const PropertyNode*& params)
{
    // main.cpp:697-775 InitializeConfigurable: "<name>::<type>" with params = child "params" (may be nullptr)
    const auto* n = root->Child(name);
    std::string type = defType ? defType : "";
    params = nullptr;
    if (n)
    {
        const auto* tn = n->Child("type");
        if (tn) type = tn->RawScalar();
        params = n->Child("params");
    }
    if (typeOverride && *typeOverride) type = typeOverride;
    if (type.empty()) return typename T::UniquePtr(nullptr, nullptr);
    return ComponentFactory::Create<T>(std::string(name) + "::" + type);
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_lastError.c_str(); }

void ref_set_verbose(int level) { Logger_SetVerboseLevel(level); }

int ref_load_plugin(const char* pathWithoutExt) { return ComponentFactory::LoadPlugin(pathWithoutExt) ? 1 : 0; }

int ref_register_mesh(const float* ps, int nv, const float* ns, const float* ts, const unsigned int* fs, int nf)
{
    return RefHost::RegisterMesh(ps, nv, ns, ts, fs, nf);
}

void ref_clear_meshes() { RefHost::ClearMeshes(); }

// accelType: nullptr/"" = take "accel: {type: ...}" from the YAML (default qbvh instead of embree,
// which cannot be built here).
void* ref_session_create(const char* yaml, const char* accelType)
{
    std::unique_ptr<Session> s(new Session);
    s->tree = ComponentFactory::Create<PropertyTree>();
    if (!s->tree || !s->tree->LoadFromString(yaml)) { g_lastError = "failed to parse scene"; return nullptr; }
    s->root = s->tree->Root()->Child("lightmetrica");
    if (!s->root) { g_lastError = "missing 'lightmetrica' node"; return nullptr; }

    s->assets = ComponentFactory::Create<Assets>();
    // main.cpp:604-607 hands "assets/params" to Assets::Initialize; accept the bare node too.
    const auto* assetsNode = s->root->Child("assets");
    if (assetsNode && assetsNode->Child("params") && assetsNode->Child("type")) assetsNode = assetsNode->Child("params");
    if (!s->assets->Initialize(assetsNode)) { g_lastError = "assets init failed"; return nullptr; }

    const PropertyNode* params = nullptr;
    s->accel = CreateConfigurable<Accel>(s->root, "accel", accelType, "qbvh", params);
    if (!s->accel) { g_lastError = "failed to create accel"; return nullptr; }
    if (!s->accel->Initialize(params)) { g_lastError = "accel init failed"; return nullptr; }

    s->scene = CreateConfigurable<Scene>(s->root, "scene", nullptr, "scene3", params);
    if (!s->scene) { g_lastError = "failed to create scene"; return nullptr; }
    // main.cpp:630-633,767 hands "scene/params" to Scene3_::Initialize (scene3.cpp:128 reads
    // "nodes"/"sensor" from it); accept a bare "scene: {sensor, nodes}" node too.
    const auto* sceneNode = s->root->Child("scene");
    if (!sceneNode) { g_lastError = "missing 'scene' node"; return nullptr; }
    if (sceneNode->Child("params")) sceneNode = sceneNode->Child("params");
    if (!s->scene->Initialize(sceneNode, s->assets.get(), s->accel.get())) { g_lastError = "scene init failed"; return nullptr; }
    return s.release();
}

void ref_session_destroy(void* h) { delete static_cast<Session*>(h); }

int ref_num_primitives(void* h)
{
    return static_cast<const Scene3*>(static_cast<Session*>(h)->scene.get())->NumPrimitives();
}

// rays: 8 floats each (ox,oy,oz,tmin, dx,dy,dz,tmax). Outputs (any may be nullptr):
//   prim[n], face[n] : -1 on miss;  tuv[3n]: t,u,v recomputed with the reference's own
//   TriAccelTriangle on the winning face (bit-identical to what the accel computed);
//   geom[11n]: p(3) gn(3) sn(3) uv(2) from the Intersection the accel filled.
// Returns the number of hits; *seconds = wall time of the Intersect loop over `threads` threads.
long long ref_intersect_batch(void* h, long long n, const float* rays, int threads,
                              int* prim, int* face, float* tuv, float* geom, double* seconds)
{
    auto* s = static_cast<Session*>(h);
    const auto* scene = static_cast<const Scene3*>(s->scene.get());
    const auto* accel = static_cast<const Accel3*>(s->accel.get());
    const int T = std::max(1, threads);
    std::vector<long long> hits(T, 0);
    auto work = [&](int t)
    {
        const long long b = n * t / T, e = n * (t + 1) / T;
        for (long long i = b; i < e; i++)
        {
            const float* r = rays + 8 * i;
            Ray ray;
            ray.o = Vec3(r[0], r[1], r[2]);
            ray.d = Vec3(r[4], r[5], r[6]);
            Intersection isect;
            const bool hit = accel->Intersect(scene, ray, isect, r[3], r[7]);
            if (!hit)
            {
                if (prim) prim[i] = -1;
                if (face) face[i] = -1;
                if (tuv) { tuv[3 * i] = tuv[3 * i + 1] = tuv[3 * i + 2] = 0.f; }
                if (geom) for (int k = 0; k < 11; k++) geom[11 * i + k] = 0.f;
                continue;
            }
            hits[t]++;
            const auto* P = isect.primitive;
            if (prim) prim[i] = P->index;
            if (face) face[i] = isect.geom.faceindex;
            if (tuv)
            {
                const auto* ps = P->mesh->Positions();
                const auto* fs = P->mesh->Faces();
                const int f = isect.geom.faceindex;
                const unsigned i1 = fs[3 * f], i2 = fs[3 * f + 1], i3 = fs[3 * f + 2];
                Vec3 p1(P->transform * Vec4(ps[3 * i1], ps[3 * i1 + 1], ps[3 * i1 + 2], 1_f));
                Vec3 p2(P->transform * Vec4(ps[3 * i2], ps[3 * i2 + 1], ps[3 * i2 + 2], 1_f));
                Vec3 p3(P->transform * Vec4(ps[3 * i3], ps[3 * i3 + 1], ps[3 * i3 + 2], 1_f));
                TriAccelTriangle tri;
                tri.Load(p1, p2, p3);
                Float u = 0, v = 0, tt = 0;
                tri.Intersect(ray, r[3], r[7], u, v, tt);
                tuv[3 * i] = tt; tuv[3 * i + 1] = u; tuv[3 * i + 2] = v;
            }
            if (geom)
            {
                float* g = geom + 11 * i;
                g[0] = isect.geom.p.x;  g[1] = isect.geom.p.y;  g[2] = isect.geom.p.z;
                g[3] = isect.geom.gn.x; g[4] = isect.geom.gn.y; g[5] = isect.geom.gn.z;
                g[6] = isect.geom.sn.x; g[7] = isect.geom.sn.y; g[8] = isect.geom.sn.z;
                g[9] = isect.geom.uv.x; g[10] = isect.geom.uv.y;
            }
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long total = 0;
    for (auto v : hits) total += v;
    return total;
}

// The reference's own Wald precompute on three world-space vertices -> 48-byte record.
int ref_triaccel_load(const float* A, const float* B, const float* C, void* out48)
{
    TriAccelTriangle tri;
    memset(&tri, 0, sizeof(tri));
    const int r = tri.Load(Vec3(A[0], A[1], A[2]), Vec3(B[0], B[1], B[2]), Vec3(C[0], C[1], C[2]));
    static_assert(sizeof(TriAccelTriangle) == 48, "TriAccelTriangle must be 48 bytes in float mode");
    memcpy(out48, &tri, 48);
    return r;
}

// World-space vertex of a primitive's mesh exactly as the reference accels compute it
// (accel_qbvh.cpp:182-184): Vec3(prim->transform * Vec4(p, 1)).
int ref_world_triangles(void* h, int primIndex, float* out9PerFace, int maxFaces)
{
    auto* s = static_cast<Session*>(h);
    const auto* scene = static_cast<const Scene3*>(s->scene.get());
    const auto* P = scene->PrimitiveAt(primIndex);
    if (!P->mesh) return 0;
    const auto* ps = P->mesh->Positions();
    const auto* fs = P->mesh->Faces();
    const int nf = std::min(P->mesh->NumFaces(), maxFaces);
    for (int f = 0; f < nf; f++)
        for (int k = 0; k < 3; k++)
        {
            const unsigned i = fs[3 * f + k];
            Vec3 p(P->transform * Vec4(ps[3 * i], ps[3 * i + 1], ps[3 * i + 2], 1_f));
            out9PerFace[9 * f + 3 * k] = p.x; out9PerFace[9 * f + 3 * k + 1] = p.y; out9PerFace[9 * f + 3 * k + 2] = p.z;
        }
    return P->mesh->NumFaces();
}

// Renders with "renderer::<rendererType>" (nullptr = from YAML). paramsYaml (may be nullptr)
// replaces the YAML's renderer params, e.g. "num_samples: 1000\nmax_num_vertices: -1".
// out: W*H*3 floats, row 0 = raster y in [0, 1/H) (bottom scanline, film_hdr.cpp:218-223).
int ref_render(void* h, const char* rendererType, const char* paramsYaml, unsigned int seed, int threads,
               float* out, int* outW, int* outH, double* seconds)
{
    auto* s = static_cast<Session*>(h);
    const PropertyNode* params = nullptr;
    auto renderer = CreateConfigurable<Renderer>(s->root, "renderer", rendererType, nullptr, params);
    if (!renderer) { g_lastError = "failed to create renderer"; return 0; }
    PropertyTree::UniquePtr ptree(nullptr, nullptr);
    if (paramsYaml && *paramsYaml)
    {
        ptree = ComponentFactory::Create<PropertyTree>();
        if (!ptree->LoadFromString(paramsYaml)) { g_lastError = "failed to parse renderer params"; return 0; }
        params = ptree->Root();
    }
    if (!renderer->Initialize(params)) { g_lastError = "renderer init failed"; return 0; }
    RefHost::numThreads = std::max(1, threads);
    Random initRng;
    initRng.SetSeed(seed);
    const auto* scene3 = static_cast<const Scene3*>(s->scene.get());
    auto* film = static_cast<const Sensor*>(scene3->GetSensor()->emitter)->GetFilm();
    film->Clear();
    const auto t0 = std::chrono::steady_clock::now();
    renderer->Render(s->scene.get(), &initRng, "/tmp/lmb200_ref_render");
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int w = 0, hh = 0;
    const float* d = RefHost::FilmData(film, w, hh);
    if (outW) *outW = w;
    if (outH) *outH = hh;
    if (out)
        for (size_t i = 0; i < (size_t)w * hh; i++) { out[3 * i] = d[4 * i]; out[3 * i + 1] = d[4 * i + 1]; out[3 * i + 2] = d[4 * i + 2]; }
    return 1;
}

int ref_film_size(void* h, int* w, int* hgt)
{
    auto* s = static_cast<Session*>(h);
    const auto* scene3 = static_cast<const Scene3*>(s->scene.get());
    auto* film = static_cast<const Sensor*>(scene3->GetSensor()->emitter)->GetFilm();
    *w = film->Width(); *hgt = film->Height();
    return 1;
}

}  // extern "C"
