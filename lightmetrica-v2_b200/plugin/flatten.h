// Scene flattening shared by the two plugins: world-space triangles in the reference accels'
// order (primitive-major, face-minor; /root/reference/src/liblightmetrica/accel/accel_qbvh.cpp:161-194)
// computed with the reference's own math types so the vertices are bit-identical to what the
// in-tree accels feed TriAccelTriangle::Load.
#pragma once
#include <lightmetrica/lightmetrica.h>
#include <vector>
#include <cstdint>
#include "lmb200.h"

namespace lmb200plugin {

using namespace lightmetrica_v2;

// prims (optional): one lmb200_primitive per scene primitive with first_tri/num_tris/has_normals filled.
// uvs (optional): 6 floats per triangle, TriangleMesh::Texcoords of the three vertices (0 for meshes without).
inline void FlattenTriangles(const Scene3* scene, std::vector<float>& verts, std::vector<float>* normals,
                             std::vector<uint32_t>& primOfTri, std::vector<uint32_t>& faceOfTri,
                             std::vector<lmb200_primitive>* prims, std::vector<float>* uvs = nullptr)
{
    verts.clear(); primOfTri.clear(); faceOfTri.clear();
    if (normals) normals->clear();
    if (uvs) uvs->clear();
    const int np = scene->NumPrimitives();
    if (prims) prims->assign((size_t)np, lmb200_primitive{0, -1, 0u, 0u, 0});
    for (int i = 0; i < np; i++)
    {
        const auto* prim = scene->PrimitiveAt(i);
        const auto* mesh = prim->mesh;
        if (prims) (*prims)[i].first_tri = (uint32_t)primOfTri.size();
        if (!mesh) continue;
        const auto* ps = mesh->Positions();
        const auto* ns = mesh->Normals();
        const auto* faces = mesh->Faces();
        const auto* tc = uvs ? mesh->Texcoords() : nullptr;
        const int nf = mesh->NumFaces();
        for (int j = 0; j < nf; j++)
        {
            const unsigned int idx[3] = { faces[3 * j], faces[3 * j + 1], faces[3 * j + 2] };
            for (int k = 0; k < 3; k++)
            {
                const Vec3 p(prim->transform * Vec4(ps[3 * idx[k]], ps[3 * idx[k] + 1], ps[3 * idx[k] + 2], 1_f));
                verts.push_back(p.x); verts.push_back(p.y); verts.push_back(p.z);
                if (normals)
                {
                    // intersectionutils.h:88-90
                    Vec3 n;
                    if (ns) n = prim->normalTransform * Vec3(ns[3 * idx[k]], ns[3 * idx[k] + 1], ns[3 * idx[k] + 2]);
                    normals->push_back(n.x); normals->push_back(n.y); normals->push_back(n.z);
                }
                if (uvs)
                {
                    // intersectionutils.h:107-115
                    uvs->push_back(tc ? tc[2 * idx[k]] : 0.f);
                    uvs->push_back(tc ? tc[2 * idx[k] + 1] : 0.f);
                }
            }
            primOfTri.push_back((uint32_t)i);
            faceOfTri.push_back((uint32_t)j);
        }
        if (prims)
        {
            (*prims)[i].num_tris = (uint32_t)nf;
            (*prims)[i].has_normals = ns ? 1 : 0;
        }
    }
}

}  // namespace lmb200plugin
