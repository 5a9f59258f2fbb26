// HitInfo — what a closest-hit query reports (reference: src/ray_tracing.h:6-37).  The device keeps
// (t, triangle id) per ray and recomputes the rest; this host struct is filled by
// BoundingVolumeHierarchy::intersect for single-ray queries.
#pragma once
#include "scene.h"

struct HitInfo {
    glm::vec3 normal;    // interpolated shading normal, flipped to the geometric side (src/ray_tracing.cpp:147-160)
    glm::vec3 hitPoint;  // origin + direction * t (src/ray_tracing.cpp:111)
    int material_index = -1; // index of the mesh owning the material
    bool is_triangle = false;
    int triangle_index = -1; // global triangle index in Scene order (not in the reference; handy for parity)

    Material& getMaterial(Scene& scene) { return scene.meshes[material_index].material; }
};
