// accel::lmb200 — Lightmetrica v2 plugin that drops in for the in-tree accels
// (/root/reference/src/liblightmetrica/accel/accel_qbvh.cpp et al.) through the reference's own
// component mechanism: built as plugin/accel_lmb200.so, registered at dlopen time with
// LM_COMPONENT_REGISTER_IMPL (component.h:660-667), selected from the scene YAML with
//     accel: {type: lmb200, params: {device: 0, builder: gpu}}
// All compute happens in liblmb200.so (CUDA, include/lmb200.h); this file only flattens the
// scene the way the reference accels do and fills the Intersection the way they do.
#include <lightmetrica/lightmetrica.h>
#include <lightmetrica/intersectionutils.h>
#include <vector>
#include <cstring>
#include "lmb200.h"
#include "flatten.h"

LM_NAMESPACE_BEGIN

class Accel_LMB200 final : public Accel3
{
public:

    LM_IMPL_CLASS(Accel_LMB200, Accel3);

public:

    ~Accel_LMB200()
    {
        // runs before dlclose (component.h:640-646): CUDA resources are released here
        lmb200_registry_put(static_cast<const Accel*>(this), nullptr);
        if (accel_) { lmb200_accel_destroy(accel_); accel_ = nullptr; }
    }

    // prop is the "params" child and may be nullptr (main.cpp:767; test_accel3.cpp:280)
    LM_IMPL_F(Initialize) = [this](const PropertyNode* prop) -> bool
    {
        device_ = (prop && prop->Child("device")) ? prop->ChildAs<int>("device", 0) : 0;
        // builder: gpu (Morton radix tree on the device, milliseconds, default) | host (binned SAH on the host, seconds) |
        //          gpu_sah (radix tree + SAH-optimal collapse) | ploc (clustering + SAH-optimal collapse)
        builder_ = LMB200_BUILD_DEFAULT;
        if (prop && prop->Child("builder"))
        {
            const auto b = prop->ChildAs<std::string>("builder", "gpu");
            if (b == "gpu") builder_ = LMB200_BUILD_GPU_LBVH;
            else if (b == "host") builder_ = LMB200_BUILD_HOST_SAH;
            else if (b == "gpu_sah") builder_ = LMB200_BUILD_GPU_LBVH_SAH;
            else if (b == "ploc") builder_ = LMB200_BUILD_GPU_PLOC;
            else { LM_LOG_ERROR("accel::lmb200: unknown builder '" + b + "' (gpu | host | gpu_sah | ploc)"); return false; }
        }
        if (accel_) { lmb200_accel_destroy(accel_); accel_ = nullptr; }
        accel_ = lmb200_accel_create(device_);
        if (!accel_)
        {
            LM_LOG_ERROR(std::string("accel::lmb200: ") + lmb200_last_error());
            return false;
        }
        return true;
    };

    LM_IMPL_F(Build) = [this](const Scene* scene_) -> bool
    {
        const auto* scene = static_cast<const Scene3*>(scene_);
        std::vector<float> verts;
        lmb200plugin::FlattenTriangles(scene, verts, nullptr, primOfTri_, faceOfTri_, nullptr);
        if (lmb200_accel_build_ex(accel_, verts.data(), verts.size() / 9, builder_) != LMB200_OK)
        {
            LM_LOG_ERROR(std::string("accel::lmb200: ") + lmb200_last_error());
            return false;
        }
        // publish the device BVH so that renderer::lmb200pt can reuse it instead of building its own
        lmb200_registry_put(static_cast<const Accel*>(this), accel_);
        lmb200_accel_stats st;
        if (lmb200_accel_get_stats(accel_, &st) == LMB200_OK)
        {
            LM_LOG_INFO("accel::lmb200: " + std::to_string(st.num_triangles) + " triangles, " + std::to_string(st.num_nodes) +
                        " wide nodes, build " + std::to_string(st.build_seconds) + " s, upload " + std::to_string(st.upload_seconds) + " s");
        }
        return true;
    };

    // Called concurrently from all render threads (scheduler.cpp:146-175): re-entrant, read-only.
    LM_IMPL_F(Intersect) = [this](const Scene* scene_, const Ray& ray, Intersection& isect, Float minT, Float maxT) -> bool
    {
        lmb200_ray r;
        r.ox = ray.o.x; r.oy = ray.o.y; r.oz = ray.o.z; r.tmin = minT;
        r.dx = ray.d.x; r.dy = ray.d.y; r.dz = ray.d.z; r.tmax = maxT;
        lmb200_hit h;
        if (lmb200_trace_closest_one(accel_, &r, &h) != LMB200_OK)
        {
            LM_LOG_ERROR(std::string("accel::lmb200: ") + lmb200_last_error());
            return false;
        }
        if (h.tri == LMB200_MISS) return false;
        const auto* scene = static_cast<const Scene3*>(scene_);
        // same epilogue as accel_qbvh.cpp:486-493
        isect = IntersectionUtils::CreateTriangleIntersection(
            scene->PrimitiveAt((int)primOfTri_[h.tri]),
            ray.o + ray.d * h.t,
            Vec2(h.u, h.v),
            (int)faceOfTri_[h.tri]);
        return true;
    };

private:

    int device_ = 0;
    int builder_ = LMB200_BUILD_DEFAULT;
    lmb200_accel* accel_ = nullptr;
    std::vector<uint32_t> primOfTri_;
    std::vector<uint32_t> faceOfTri_;

};

LM_COMPONENT_REGISTER_IMPL(Accel_LMB200, "accel::lmb200");

LM_NAMESPACE_END
