#include "bounding_volume_hierarchy.h"
#include "rt_b200.h"
#include <stdexcept>
#include <string>

namespace {
void check(int rc, const char* what)
{
    if (rc != RT_OK)
        throw std::runtime_error(std::string(what) + ": " + rt_last_error());
}
}

BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene* pScene, int bvhMode, int device)
    : m_pScene(pScene)
{
    std::vector<rt_material> mats;
    std::vector<float> uv;                 // Vertex::texCoord of every corner
    std::vector<const Image*> images;      // distinct Material::kdTexture objects, by file
    std::vector<int> matTex;
    int meshIndex = 0;
    for (const Mesh& mesh : pScene->meshes) {
        for (const Triangle& tri : mesh.triangles) {
            const Vertex* v[3] = { &mesh.vertices[tri.x], &mesh.vertices[tri.y], &mesh.vertices[tri.z] };
            for (int k = 0; k < 3; k++) {
                m_pos.insert(m_pos.end(), { v[k]->p.x, v[k]->p.y, v[k]->p.z });
                m_nrm.insert(m_nrm.end(), { v[k]->n.x, v[k]->n.y, v[k]->n.z });
                uv.insert(uv.end(), { v[k]->texCoord.x, v[k]->texCoord.y });
            }
            m_meshId.push_back(meshIndex);
        }
        const Material& m = mesh.material;
        mats.push_back(rt_material { { m.kd.x, m.kd.y, m.kd.z }, { m.ks.x, m.ks.y, m.ks.z }, m.shininess, m.transparency });
        int tex = -1;
        if (m.kdTexture) {
            for (size_t k = 0; k < images.size() && tex < 0; k++)
                if (images[k]->path() == m.kdTexture->path())
                    tex = (int)k;
            if (tex < 0) {
                tex = (int)images.size();
                images.push_back(&*m.kdTexture);
            }
        }
        matTex.push_back(tex);
        meshIndex++;
    }
    check(rt_create(device, &m_ctx), "rt_create");
    check(rt_upload_scene(m_ctx, m_pos.data(), m_nrm.data(), m_meshId.data(), (int64_t)m_meshId.size(), mats.data(), (int)mats.size()), "rt_upload_scene");
    // textures travel with the geometry (they are loaded by loadMesh, src/mesh.cpp:138-145); whether they are used is a
    // per-frame knob (useTextures, render.h)
    if (!m_meshId.empty())
        check(rt_set_texcoords(m_ctx, uv.data()), "rt_set_texcoords");
    if (!images.empty()) {
        std::vector<rt_texture> tex;
        static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "texels must be packed float3");
        for (const Image* img : images)
            tex.push_back(rt_texture { img->width(), img->height(), &img->pixels()[0].x });
        check(rt_set_textures(m_ctx, tex.data(), (int)tex.size(), matTex.data(), (int)matTex.size()), "rt_set_textures");
    }
    std::vector<rt_sphere> sp;
    for (const Sphere& s : pScene->spheres) {
        const Material& m = s.material;
        sp.push_back(rt_sphere { { s.center.x, s.center.y, s.center.z }, s.radius,
            rt_material { { m.kd.x, m.kd.y, m.kd.z }, { m.ks.x, m.ks.y, m.ks.z }, m.shininess, m.transparency } });
    }
    check(rt_set_spheres(m_ctx, sp.data(), (int)sp.size()), "rt_set_spheres");
    check(rt_build_bvh(m_ctx, bvhMode), "rt_build_bvh");
    // depth of a balanced binary tree over the triangles, for callers that print it
    size_t n = m_meshId.size();
    while (n > 1) {
        n = (n + 1) / 2;
        m_numLevels++;
    }
    m_numLevels++;
}

BoundingVolumeHierarchy::~BoundingVolumeHierarchy()
{
    if (m_ctx)
        rt_destroy(m_ctx);
}

bool BoundingVolumeHierarchy::intersect(Ray& ray, HitInfo& hitInfo, bool useBVH) const
{
    const float r[6] = { ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z };
    int id = -1;
    float t = 0.0f;
    check(rt_intersect(m_ctx, r, 1, useBVH ? 1 : 0, &id, &t), "rt_intersect");
    if (id < 0 || !(t < ray.t))
        return false;
    ray.t = t;
    hitInfo.triangle_index = id;
    if (id >= (int)m_meshId.size()) { // sphere primitive (src/ray_tracing.cpp:199-204)
        const Sphere& sp = m_pScene->spheres.at(size_t(id) - m_meshId.size());
        hitInfo.is_triangle = false;
        hitInfo.hitPoint = ray.origin + t * ray.direction;
        hitInfo.normal = glm::normalize(hitInfo.hitPoint - sp.center);
        hitInfo.sphere_material = sp.material;
        return true;
    }
    hitInfo.is_triangle = true;
    hitInfo.material_index = m_meshId[id];
    hitInfo.hitPoint = ray.origin + ray.direction * t;
    // shading normal: barycentric blend of the corner normals, flipped to the geometric side
    // (src/ray_tracing.cpp:147-160, 276-308; barycentrics always evaluated, see DESIGN.md "defined barycentrics")
    const float* p = &m_pos[9 * size_t(id)];
    const float* nn = &m_nrm[9 * size_t(id)];
    const glm::vec3 v0(p[0], p[1], p[2]), v1(p[3], p[4], p[5]), v2(p[6], p[7], p[8]);
    const glm::vec3 faceN = glm::normalize(glm::cross(v0 - v2, v1 - v2));
    const glm::vec3 hp = hitInfo.hitPoint;
    const float total = glm::length(glm::cross(v1 - v0, v2 - v0));
    const float c0 = glm::length(glm::cross(v1 - hp, v2 - hp)) / total;
    const float c1 = glm::length(glm::cross(hp - v0, v2 - v0)) / total;
    const float c2 = glm::length(glm::cross(v1 - v0, hp - v0)) / total;
    glm::vec3 n = glm::vec3(nn[0], nn[1], nn[2]) * c0 + glm::vec3(nn[3], nn[4], nn[5]) * c1 + glm::vec3(nn[6], nn[7], nn[8]) * c2;
    if (glm::dot(n, faceN) < 0)
        n = -n;
    hitInfo.normal = n;
    return true;
}
