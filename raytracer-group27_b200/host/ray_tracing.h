// HitInfo — what a closest-hit query reports (reference: src/ray_tracing.h:6-37).  The device keeps
// (t, triangle id) per ray and recomputes the rest; this host struct is filled by
// BoundingVolumeHierarchy::intersect for single-ray queries.
#pragma once
#include "scene.h"

struct HitInfo {
    glm::vec3 normal;    // interpolated shading normal, flipped to the geometric side (src/ray_tracing.cpp:147-160)
    glm::vec3 hitPoint;  // origin + direction * t (src/ray_tracing.cpp:111)
    int material_index = -1; // index of the mesh owning the material
    Material sphere_material; // spheres carry their own material
    bool is_triangle = false; // false: a sphere was hit
    int triangle_index = -1;  // global triangle index in Scene order, or #triangles + sphere index (not in the reference)

    Material& getMaterial(Scene& scene)
    {
        if (is_triangle)
            return scene.meshes[material_index].material;
        return sphere_material;
    }
};
