// renderer::lmb200pt — Lightmetrica v2 plugin that drops in for renderer::pt / renderer::ptdirect
// (/root/reference/src/liblightmetrica/renderer/renderer_pt.cpp, renderer_ptdirect.cpp) and, with
// mode "normal", for the primary-ray renderers (renderer_raycast.cpp, plugin/renderer_normal).
// Selected from the scene YAML with
//     renderer: {type: lmb200pt, params: {mode: ptdirect, num_samples: ..., max_num_vertices: -1,
//                                         min_num_vertices: 0, num_gpus: 1, device: 0, pool_size: 0,
//                                         render_time: -1, progress_image_update_interval: -1, grain_size: 10000, builder: gpu,
//                                         texture_resolution: 1024, tile_partitioning: 0, primary_tile: 0}}
// It reads the scene through the reference's interfaces (Scene3::PrimitiveAt, TriangleMesh::*,
// BSDF::Reflectance/Glossiness, Light::Emittance, Sensor::GetFilm/GetProjectionMatrix), flattens
// it into the POD arrays of include/lmb200.h and calls liblmb200.so; the film comes back through
// Film::SetPixel and is written with Film::Save, exactly where the reference renderers save
// (renderer_pt.cpp:236-241). Scenes using assets outside the supported set (textured BSDFs, light::env
// outside mode ptdirect, unknown plugins) are rejected with an error, never mis-rendered.
#include <lightmetrica/lightmetrica.h>
#include <chrono>
#include <vector>
#include <map>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "lmb200.h"
#include "flatten.h"

LM_NAMESPACE_BEGIN

class Renderer_LMB200PT final : public Renderer
{
public:

    LM_IMPL_CLASS(Renderer_LMB200PT, Renderer);

public:

    // prop may be nullptr (main.cpp:767)
    LM_IMPL_F(Initialize) = [this](const PropertyNode* prop) -> bool
    {
        prop_ = prop;
        std::string mode = "ptdirect";
        if (prop)
        {
            if (prop->Child("mode")) mode = prop->ChildAs<std::string>("mode", "ptdirect");
            // same keys and defaults as Scheduler_::Load (scheduler.cpp:44-58) and renderer_pt.cpp:56-62
            numSamples_ = prop->ChildAs<long long>("num_samples", 10000000L);
            maxNumVertices_ = prop->ChildAs<int>("max_num_vertices", -1);
            minNumVertices_ = prop->ChildAs<int>("min_num_vertices", 0);
            // scheduler.cpp:44-58: time budget and progressive output
            if (prop->Child("render_time")) renderTime_ = prop->ChildAs<double>("render_time", -1.0);
            if (prop->Child("progress_image_update_interval")) progressImageInterval_ = prop->ChildAs<double>("progress_image_update_interval", -1.0);
            if (prop->Child("grain_size")) grainSize_ = prop->ChildAs<long long>("grain_size", 10000);
            if (prop->Child("num_gpus")) numGpus_ = prop->ChildAs<int>("num_gpus", 1);
            if (prop->Child("device")) device_ = prop->ChildAs<int>("device", 0);
            if (prop->Child("pool_size")) poolSize_ = prop->ChildAs<int>("pool_size", 0);
            if (prop->Child("texture_resolution")) textureResolution_ = prop->ChildAs<int>("texture_resolution", 1024);
            if (prop->Child("tile_partitioning")) tilePartitioning_ = prop->ChildAs<int>("tile_partitioning", 0) != 0;
            if (prop->Child("primary_tile")) primaryTile_ = prop->ChildAs<int>("primary_tile", 0);
            if (prop->Child("builder"))
            {
                const auto b = prop->ChildAs<std::string>("builder", "gpu");
                if (b == "gpu") builder_ = LMB200_BUILD_GPU_LBVH;
                else if (b == "host") builder_ = LMB200_BUILD_HOST_SAH;
                else if (b == "gpu_sah") builder_ = LMB200_BUILD_GPU_LBVH_SAH;
                else if (b == "ploc") builder_ = LMB200_BUILD_GPU_PLOC;
                else { LM_LOG_ERROR("renderer::lmb200pt: unknown builder '" + b + "' (gpu | host | gpu_sah | ploc)"); return false; }
            }
        }
        if (mode == "pt") mode_ = LMB200_MODE_PT;
        else if (mode == "ptdirect") mode_ = LMB200_MODE_PTDIRECT;
        else if (mode == "ptmis") mode_ = LMB200_MODE_PTMIS;
        else if (mode == "normal") mode_ = LMB200_MODE_NORMAL;
        else { LM_LOG_ERROR("renderer::lmb200pt: unknown mode '" + mode + "' (pt | ptdirect | ptmis | normal)"); return false; }
        if (numGpus_ < 1) numGpus_ = 1;
        if (lmb200_device_count() < device_ + numGpus_ && !std::getenv("LMB200_DUMP_SCENE"))
        {
            LM_LOG_ERROR("renderer::lmb200pt: needs " + std::to_string(numGpus_) + " CUDA device(s) starting at " + std::to_string(device_) +
                         ", found " + std::to_string(lmb200_device_count()) + " (there is no CPU fallback)");
            return false;
        }
        return true;
    };

    LM_IMPL_F(Render) = [this](const Scene* scene_, Random* initRng, const std::string& outputPath) -> void
    {
        const auto* scene = static_cast<const Scene3*>(scene_);
        const auto* sensorPrim = scene->GetSensor();
        auto* film = static_cast<const Sensor*>(sensorPrim->emitter)->GetFilm();   // renderer_pt.cpp:67

        // ---- flatten geometry ----
        std::vector<float> verts, normals, uvs;
        std::vector<uint32_t> primOfTri, faceOfTri;
        std::vector<lmb200_primitive> prims;
        const auto tFlatten0 = std::chrono::steady_clock::now();
        lmb200plugin::FlattenTriangles(scene, verts, &normals, primOfTri, faceOfTri, &prims, &uvs);
        const double flattenSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tFlatten0).count();
        curScene_ = scene;
        textures_.clear(); textureData_.clear(); textureIndex_.clear();
        bool anyNormals = false;
        for (const auto& p : prims) anyNormals = anyNormals || p.has_normals;

        // ---- materials and lights through the reference interfaces ----
        std::vector<lmb200_bsdf> bsdfs;
        std::vector<lmb200_light> lights;
        std::map<const BSDF*, int> bsdfIndex;
        std::map<const Light*, int> lightBound;
        const int np = scene->NumPrimitives();
        for (int i = 0; i < np; i++)
        {
            const auto* prim = scene->PrimitiveAt(i);
            int bi = -1;
            if (prim->bsdf)
            {
                auto it = bsdfIndex.find(prim->bsdf);
                if (it != bsdfIndex.end()) bi = it->second;
                else
                {
                    lmb200_bsdf b;
                    if (!ConvertBSDF(prim->bsdf, b)) return;
                    bi = (int)bsdfs.size();
                    bsdfs.push_back(b);
                    bsdfIndex[prim->bsdf] = bi;
                }
            }
            if (bi < 0)
            {
                lmb200_bsdf b; memset(&b, 0, sizeof(b)); b.type = LMB200_BSDF_NULL;
                bi = (int)bsdfs.size(); bsdfs.push_back(b);
            }
            prims[i].bsdf = bi;
            prims[i].light = -1;
            if (prim->light)
            {
                const std::string limpl = prim->light->implName;
                if (limpl == "Light_Point")
                {
                    // light_point.cpp:47-105: position and Le through the Emitter interface
                    SurfaceGeometry gp, gprev;
                    prim->light->SamplePositionGivenPreviousPosition(Vec2(), gprev, gp);
                    const auto Le = prim->light->EvaluateDirection(gp, SurfaceInteractionType::L, Vec3(), Vec3(0_f, 0_f, 1_f), TransportDirection::LE, false).ToRGB();
                    lmb200_light L; memset(&L, 0, sizeof(L));
                    L.Le[0] = Le.x; L.Le[1] = Le.y; L.Le[2] = Le.z; L.primitive = i; L.kind = LMB200_LIGHT_POINT;
                    L.position[0] = gp.p.x; L.position[1] = gp.p.y; L.position[2] = gp.p.z;
                    prims[i].light = (int)lights.size();
                    lights.push_back(L);
                    continue;
                }
                if (limpl == "Light_Directional" || limpl == "Light_EnvLight")
                {
                    // light_directional.cpp:147-162 / light_env.cpp:148-168: the sampled direction of a directional light is
                    // its (transformed, normalised) direction_; Le through EvaluateDirection (constant for light::env, whose
                    // envmap branch cannot load: light_env.cpp:107-113)
                    SurfaceGeometry gp;
                    Vec3 wo;
                    prim->light->SamplePositionAndDirection(Vec2(0.5_f, 0.5_f), Vec2(0.5_f, 0.5_f), gp, wo);
                    const auto Le = prim->light->EvaluateDirection(gp, SurfaceInteractionType::L, Vec3(), wo, TransportDirection::LE, false).ToRGB();
                    lmb200_light L; memset(&L, 0, sizeof(L));
                    L.Le[0] = Le.x; L.Le[1] = Le.y; L.Le[2] = Le.z; L.primitive = i;
                    if (limpl == "Light_Directional")
                    {
                        L.kind = LMB200_LIGHT_DIRECTIONAL;
                        L.direction[0] = wo.x; L.direction[1] = wo.y; L.direction[2] = wo.z;
                    }
                    else
                    {
                        L.kind = LMB200_LIGHT_ENV;
                        if (mode_ == LMB200_MODE_PT || mode_ == LMB200_MODE_PTMIS)
                        {
                            LM_LOG_ERROR("renderer::lmb200pt: light::env is supported in mode ptdirect only (renderer::pt / ptmis of the reference "
                                         "dereference a null primitive when a ray escapes to the env emitter shape)");
                            return;
                        }
                    }
                    prims[i].light = (int)lights.size();
                    lights.push_back(L);
                    continue;
                }
                if (limpl != "Light_Area")
                {
                    LM_LOG_ERROR("renderer::lmb200pt: unsupported light '" + limpl + "' (light::area | light::point | light::directional | light::env)");
                    return;
                }
                const auto Le = prim->light->Emittance().ToRGB();
                // A light asset is loaded once, with the first primitive that references it
                // (assets.cpp:50-122; light_area.cpp:50-55 binds mesh, transform and area
                // distribution at Load): primitives sharing the asset all sample that first mesh.
                auto bound = lightBound.find(prim->light);
                if (bound == lightBound.end()) bound = lightBound.emplace(prim->light, i).first;
                lmb200_light L; memset(&L, 0, sizeof(L));
                L.Le[0] = Le.x; L.Le[1] = Le.y; L.Le[2] = Le.z; L.primitive = bound->second; L.kind = LMB200_LIGHT_AREA;
                prims[i].light = (int)lights.size();
                lights.push_back(L);
            }
        }

        // ---- sensor::pinhole (sensor_pinhole.cpp:47-61) / sensor::thinlens (sensor_thinlens.cpp:44-68) ----
        const std::string simpl = sensorPrim->sensor->implName;
        if (simpl != "Sensor_Pinhole" && simpl != "Sensor_ThinLens")
        {
            LM_LOG_ERROR("renderer::lmb200pt: unsupported sensor '" + simpl + "' (sensor::pinhole | sensor::thinlens)");
            return;
        }
        lmb200_scene_desc d;
        memset(&d, 0, sizeof(d));
        {
            // Scene3::GetSphereBound (scene3.cpp:56-78), the virtual-disk geometry of directional / env lights
            const auto sb = scene->GetSphereBound();
            d.sphere_center[0] = sb.center.x; d.sphere_center[1] = sb.center.y; d.sphere_center[2] = sb.center.z;
            d.sphere_radius = sb.radius;
        }
        {
            const Vec3 pos(sensorPrim->transform * Vec4(0_f, 0_f, 0_f, 1_f));
            const Vec3 vx(sensorPrim->transform[0]), vy(sensorPrim->transform[1]), vz(sensorPrim->transform[2]);
            d.camera.position[0] = pos.x; d.camera.position[1] = pos.y; d.camera.position[2] = pos.z;
            d.camera.vx[0] = vx.x; d.camera.vx[1] = vx.y; d.camera.vx[2] = vx.z;
            d.camera.vy[0] = vy.x; d.camera.vy[1] = vy.y; d.camera.vy[2] = vy.z;
            d.camera.vz[0] = vz.x; d.camera.vz[1] = vz.y; d.camera.vz[2] = vz.z;
            // fov is not exposed by the interface: the YAML value if reachable, else recovered from
            // GetProjectionMatrix()[1][1] = 1/tan(fov/2) (sensor_pinhole.cpp:192-202)
            Float fovDeg = -1_f;
            if (const auto* ap = AssetParams(sensorPrim->sensor)) fovDeg = ap->ChildAs<Float>("fov", 45_f);
            if (fovDeg > 0_f) d.camera.fov = Math::Radians(fovDeg);
            else d.camera.fov = 2_f * std::atan(1_f / sensorPrim->sensor->GetProjectionMatrix(1_f, 2_f)[1][1]);
            d.camera.width = film->Width();
            d.camera.height = film->Height();
            d.camera.kind = LMB200_CAMERA_PINHOLE;
            if (simpl == "Sensor_ThinLens")
            {
                d.camera.kind = LMB200_CAMERA_THINLENS;
                if (const auto* ap = AssetParams(sensorPrim->sensor))
                {
                    d.camera.lens_radius = ap->ChildAs<Float>("lens_radius", 0.1_f);      // sensor_thinlens.cpp:64-65
                    d.camera.focal_distance = ap->ChildAs<Float>("focal_distance", 1_f);
                }
                else
                {
                    // not reachable through the YAML tree: recover both from one sample through the Sensor interface.
                    // Lens sample (1,.5) maps to lensUV = (radius, 0) (sampler.h:44-60), raster centre to rayDir = -vz,
                    // so p = pos + R vx and wo ~ -(F vz + R vx)  (sensor_thinlens.cpp:87-106)
                    SurfaceGeometry gp;
                    Vec3 wo;
                    sensorPrim->sensor->SamplePositionAndDirection(Vec2(0.5_f, 0.5_f), Vec2(1_f, 0.5_f), gp, wo);
                    const Float R = Math::Dot(gp.p - pos, vx) / Math::Dot(vx, vx);
                    d.camera.lens_radius = R;
                    d.camera.focal_distance = R * Math::Dot(wo, -vz) / Math::Dot(wo, -vx);
                    LM_LOG_WARN("renderer::lmb200pt: thin-lens parameters recovered through the Sensor interface (asset params not reachable)");
                }
            }
        }
        d.num_tris = verts.size() / 9;
        d.verts = verts.data();
        d.normals = (anyNormals && !normals.empty()) ? normals.data() : nullptr;      // FlattenTriangles leaves it empty when no mesh has normals
        d.tri_prim = primOfTri.data();
        d.num_prims = (uint32_t)prims.size();
        d.prims = prims.data();
        d.num_bsdfs = (uint32_t)bsdfs.size();
        d.bsdfs = bsdfs.data();
        d.num_lights = (uint32_t)lights.size();
        d.lights = lights.data();
        for (size_t t = 0; t < textures_.size(); t++) textures_[t].rgb = textureData_[t].data();
        d.num_textures = (uint32_t)textures_.size();
        d.textures = textures_.empty() ? nullptr : textures_.data();
        d.uvs = (textures_.empty() || uvs.empty()) ? nullptr : uvs.data();            // no texture coordinates anywhere: (0, 0), as the reference (intersectionutils.h:107-115)

        LM_LOG_INFO("renderer::lmb200pt: " + std::to_string(d.num_tris) + " triangles of " + std::to_string(d.num_prims) + " primitives flattened in " +
                    std::to_string(flattenSeconds) + " s, scene read in " +
                    std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - tFlatten0).count()) + " s");
        if (const char* dump = std::getenv("LMB200_DUMP_SCENE"))
        {
            // debugging aid: the flattened scene exactly as handed to lmb200_scene_create
            if (FILE* f = fopen(dump, "wb"))
            {
                const uint64_t hdr[5] = { d.num_tris, d.num_prims, d.num_bsdfs, d.num_lights, d.normals ? 1u : 0u };
                fwrite(hdr, sizeof(hdr), 1, f);
                fwrite(&d.camera, sizeof(d.camera), 1, f);
                fwrite(d.sphere_center, sizeof(float), 3, f);
                fwrite(&d.sphere_radius, sizeof(float), 1, f);
                fwrite(d.verts, sizeof(float), 9 * d.num_tris, f);
                fwrite(d.tri_prim, sizeof(uint32_t), d.num_tris, f);
                fwrite(d.prims, sizeof(lmb200_primitive), d.num_prims, f);
                fwrite(d.bsdfs, sizeof(lmb200_bsdf), d.num_bsdfs, f);
                fwrite(d.lights, sizeof(lmb200_light), d.num_lights, f);
                const uint64_t tail[2] = { d.num_textures, d.uvs ? 1u : 0u };
                fwrite(tail, sizeof(tail), 1, f);
                for (uint32_t t = 0; t < d.num_textures; t++)
                {
                    const int32_t wh[2] = { d.textures[t].width, d.textures[t].height };
                    fwrite(wh, sizeof(wh), 1, f);
                    fwrite(d.textures[t].rgb, sizeof(float), 3 * (size_t)wh[0] * wh[1], f);
                }
                if (d.uvs) fwrite(d.uvs, sizeof(float), 6 * d.num_tris, f);
                if (d.normals) fwrite(d.normals, sizeof(float), 9 * d.num_tris, f);
                fclose(f);
            }
        }

        // ---- render ----
        // If the YAML selected accel::lmb200 too, its device BVH (same triangle list, same order) is reused
        // on the GPU it lives on (if that is one of ours); the BVH is built ONCE (by accel::lmb200 or by scene 0) and
        // replicated device to device onto the other GPUs.
        lmb200_accel* sharedAccel = lmb200_registry_get(scene->GetAccel());
        const int sharedDev = sharedAccel ? lmb200_accel_device(sharedAccel) : -1;
        std::vector<lmb200_scene*> scenes(numGpus_, nullptr);
        std::vector<lmb200_accel*> replicas;
        auto cleanup = [&]() {
            for (auto* t : scenes) if (t) lmb200_scene_destroy(t);
            for (auto* r : replicas) lmb200_accel_destroy(r);
        };
        auto fail = [&]() {
            LM_LOG_ERROR(std::string("renderer::lmb200pt: ") + lmb200_last_error());
            cleanup();
        };
        lmb200_accel* source = nullptr;     // a built device BVH to replicate from
        if (sharedAccel && sharedDev >= device_ && sharedDev < device_ + numGpus_)
        {
            scenes[sharedDev - device_] = lmb200_scene_create_shared(&d, sharedAccel);
            if (!scenes[sharedDev - device_]) { fail(); return; }
            LM_LOG_INFO("renderer::lmb200pt: reusing the BVH of accel::lmb200 on device " + std::to_string(sharedDev));
            source = sharedAccel;
        }
        else if (sharedAccel) source = sharedAccel;      // lives on a GPU outside [device, device + num_gpus): replicate from it
        for (int g = 0; g < numGpus_; g++)
        {
            if (scenes[g]) continue;
            if (source)
            {
                lmb200_accel* r = lmb200_accel_replicate(source, device_ + g);
                if (!r) { fail(); return; }
                replicas.push_back(r);
                scenes[g] = lmb200_scene_create_shared(&d, r);
            }
            else
            {
                scenes[g] = lmb200_scene_create_ex(device_ + g, &d, builder_);
                if (scenes[g]) source = lmb200_scene_accel(scenes[g]);
            }
            if (!scenes[g]) { fail(); return; }
        }
        lmb200_render_params p;
        memset(&p, 0, sizeof(p));
        p.mode = mode_;
        p.num_samples = numSamples_;
        p.sample_begin = 0;
        p.sample_end = numSamples_;
        p.max_num_vertices = maxNumVertices_;
        p.min_num_vertices = minNumVertices_;
        // one seed from the host RNG replaces the per-thread seeds of scheduler.cpp:157-164
        p.seed = (uint64_t)initRng->NextUInt();
        p.pool_size = poolSize_;
        p.tile_partition = tilePartitioning_ ? 1 : 0;
        p.primary_tile = primaryTile_;
        const int W = film->Width(), H = film->Height();
        std::vector<float> rgba((size_t)W * H * 4, 0.f);
        lmb200_render_stats st;
        memset(&st, 0, sizeof(st));
        int rc;
        if (renderTime_ > 0 || progressImageInterval_ > 0)
        {
            // Scheduler_ semantics: passes of grain_size*1000 samples until render_time has elapsed,
            // "progress_%010d" images every progress_image_update_interval seconds (scheduler.cpp:108,221-255)
            ProgressCtx ctx{ film, W, H };
            rc = lmb200_render_timed(scenes.data(), numGpus_, &p, renderTime_, grainSize_ * 1000, progressImageInterval_,
                                     &Renderer_LMB200PT::OnProgress, &ctx, rgba.data(), &st);
        }
        else
        {
            rc = numGpus_ > 1 ? lmb200_render_multi(scenes.data(), numGpus_, &p, rgba.data(), &st)
                              : lmb200_render(scenes[0], &p, rgba.data(), &st);
        }
        cleanup();
        if (rc != LMB200_OK)
        {
            LM_LOG_ERROR(std::string("renderer::lmb200pt: ") + lmb200_last_error());
            return;
        }
        LM_LOG_INFO("renderer::lmb200pt: " + std::to_string(st.samples) + " samples, " + std::to_string(st.extend_rays) + " extend rays, " +
                    std::to_string(st.shadow_rays) + " shadow rays in " + std::to_string(st.seconds) + " s on " + std::to_string(numGpus_) + " GPU(s)");

        // ---- hand the image back through the Film interface (film.h) ----
        StoreFilm(film, W, H, rgba.data());
        {
            LM_LOG_INFO("Saving image");
            LM_LOG_INDENTER();
            film->Save(outputPath);
        }
    };

private:

    struct ProgressCtx { Film* film; int W, H; };

    static void StoreFilm(Film* film, int W, int H, const float* rgba)
    {
        film->Clear();
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
            {
                const float* c = &rgba[4 * ((size_t)y * W + x)];
                film->SetPixel(x, y, SPD::FromRGB(Vec3(c[0], c[1], c[2])));
            }
    }

    static int OnProgress(void* user, const float* rgba, int64_t samplesDone, int64_t tick)
    {
        auto* ctx = static_cast<ProgressCtx*>(user);
        StoreFilm(ctx->film, ctx->W, ctx->H, rgba);
        char name[64];
        snprintf(name, sizeof(name), "progress_%010lld", (long long)tick);   // scheduler.cpp:236-240
        LM_LOG_INFO("Saving progress: " + std::string(name) + " (" + std::to_string(samplesDone) + " samples)");
        ctx->film->Save(name);
        return 0;
    }

    // "lightmetrica/assets[/params]/<asset id>/params" of the loaded YAML, or nullptr
    auto AssetParams(const Asset* asset) const -> const PropertyNode*
    {
        if (!prop_ || !prop_->Tree() || !prop_->Tree()->Root()) return nullptr;
        const auto* lm = prop_->Tree()->Root()->Child("lightmetrica");
        if (!lm) return nullptr;
        const auto* assets = lm->Child("assets");
        if (!assets) return nullptr;
        if (assets->Child("type") && assets->Child("params")) assets = assets->Child("params");
        const auto* a = assets->Child(asset->ID());
        return a ? a->Child("params") : nullptr;
    }

    auto ConvertBSDF(const BSDF* bsdf, lmb200_bsdf& b) -> bool
    {
        memset(&b, 0, sizeof(b));
        const std::string impl = bsdf->implName;
        if (impl == "BSDF_Null") { b.type = LMB200_BSDF_NULL; return true; }
        if (impl == "BSDF_ReflectAll" || impl == "BSDF_RefractAll" || impl == "BSDF_Flesnel")
        {
            // delta BSDFs expose neither R nor the indices of refraction through the interface
            // (bsdf_reflectall.cpp:48-51, bsdf_refractall.cpp:46-52): read the YAML the asset was loaded from
            const auto* sp = AssetParams(bsdf);
            if (!sp)
            {
                LM_LOG_ERROR("renderer::lmb200pt: cannot reach the scene tree to read the parameters of '" + bsdf->ID() + "' (" + impl + ")");
                return false;
            }
            const Vec3 R = sp->ChildAs<Vec3>("R", Vec3());
            b.R[0] = R.x; b.R[1] = R.y; b.R[2] = R.z;
            b.eta1 = sp->ChildAs<Float>("eta1", 1_f);
            b.eta2 = sp->ChildAs<Float>("eta2", 2_f);
            b.type = impl == "BSDF_ReflectAll" ? LMB200_BSDF_REFLECT_ALL : impl == "BSDF_RefractAll" ? LMB200_BSDF_REFRACT_ALL : LMB200_BSDF_FLESNEL;
            return true;
        }
        if (impl != "BSDF_Diffuse" && impl != "BSDF_CookTorrance")
        {
            LM_LOG_ERROR("renderer::lmb200pt: unsupported BSDF '" + impl + "' (diffuse | cook_torrance | reflect_all | refract_all | flesnel | null)");
            return false;
        }
        const auto* ap = AssetParams(bsdf);
        if (ap && ap->Child("TexR"))
        {
            // bsdf_diffuse.cpp:48-53 / bsdf_cooktorrance.cpp:50-55: R comes from a texture asset. The Texture interface only
            // has Evaluate(uv) (texture.h:55), so the texture is baked at texel centres of a texture_resolution^2 grid and looked
            // up on the device the way texture::bitmap does (texture_bitmap.cpp:162-168): exact for bitmaps whose size
            // divides the resolution, an approximation along the edges of procedural textures otherwise.
            const auto id = ap->ChildAs<std::string>("TexR", "");
            auto* assets = const_cast<Assets*>(curScene_->GetAssets());
            const auto* tex = assets ? static_cast<const Texture*>(assets->AssetByIDAndType(id, "texture", nullptr)) : nullptr;
            if (!tex)
            {
                LM_LOG_ERROR("renderer::lmb200pt: cannot resolve texture '" + id + "' of BSDF '" + bsdf->ID() + "'");
                return false;
            }
            auto it = textureIndex_.find(tex);
            if (it == textureIndex_.end())
            {
                const int res = textureResolution_ > 0 ? textureResolution_ : 1024;
                std::vector<float> data((size_t)res * res * 3);
                for (int y = 0; y < res; y++)
                    for (int x = 0; x < res; x++)
                    {
                        const Vec3 c = tex->Evaluate(Vec2(((Float)x + 0.5_f) / (Float)res, ((Float)y + 0.5_f) / (Float)res));
                        float* o = &data[3 * ((size_t)res * y + x)];
                        o[0] = c.x; o[1] = c.y; o[2] = c.z;
                    }
                lmb200_texture t; t.width = res; t.height = res; t.rgb = nullptr;   // rgb is bound after all textures are baked
                textures_.push_back(t);
                textureData_.push_back(std::move(data));
                it = textureIndex_.emplace(tex, (int)textures_.size()).first;
            }
            b.texR = it->second;
        }
        else
        {
            const auto R = bsdf->Reflectance().ToRGB();
            b.R[0] = R.x; b.R[1] = R.y; b.R[2] = R.z;
        }
        if (impl == "BSDF_Diffuse") { b.type = LMB200_BSDF_DIFFUSE; return true; }
        b.type = LMB200_BSDF_COOKTORRANCE;
        b.roughness = bsdf->Glossiness();                           // bsdf_cooktorrance.cpp:157-162
        // eta/k are not exposed by the interface (bsdf_cooktorrance.cpp:297-301): YAML values or the defaults of :61-62
        const Vec3 etaDef(0.140000_f, 0.129000_f, 0.158500_f), kDef(4.586250_f, 3.348125_f, 2.329375_f);
        Vec3 eta = etaDef, k = kDef;
        if (ap)
        {
            // const defaults select ChildAs(name, const T& def) -> T, not the bool-returning overload
            eta = ap->ChildAs<Vec3>("eta", etaDef);
            k = ap->ChildAs<Vec3>("k", kDef);
        }
        else
        {
            LM_LOG_WARN("renderer::lmb200pt: scene tree not reachable, using default eta/k for cook_torrance '" + bsdf->ID() + "'");
        }
        b.eta[0] = eta.x; b.eta[1] = eta.y; b.eta[2] = eta.z;
        b.k[0] = k.x; b.k[1] = k.y; b.k[2] = k.z;
        return true;
    }

private:

    const PropertyNode* prop_ = nullptr;
    int mode_ = LMB200_MODE_PTDIRECT;
    long long numSamples_ = 10000000L;
    int maxNumVertices_ = -1;
    int minNumVertices_ = 0;
    int numGpus_ = 1;
    int device_ = 0;
    int poolSize_ = 0;
    int textureResolution_ = 1024;
    bool tilePartitioning_ = false;                       // num_gpus > 1: each GPU samples its own horizontal strip of the raster
    int primaryTile_ = 0;                                 // lmb200_render_params::primary_tile: 0 = automatic, < 0 = independent raster positions
    const Scene3* curScene_ = nullptr;                    // valid during Render
    std::vector<lmb200_texture> textures_;                 // baked TexR textures of the scene being rendered
    std::vector<std::vector<float>> textureData_;
    std::map<const Texture*, int> textureIndex_;           // texture -> index + 1
    int builder_ = LMB200_BUILD_DEFAULT;
    double renderTime_ = -1.0;
    double progressImageInterval_ = -1.0;
    long long grainSize_ = 10000;

};

LM_COMPONENT_REGISTER_IMPL(Renderer_LMB200PT, "renderer::lmb200pt");

LM_NAMESPACE_END
