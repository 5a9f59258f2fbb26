// Ray — same fields as the reference's hot-path view of `struct Ray` (framework/include/ray.h:11-14).
// The ray-differential members of the reference (ray.h:19-28) only feed texture mip selection
// (src/ray_differentials.cpp), which is outside the rebuilt path, so they are not carried.
#pragma once
#include <glm/vec3.hpp>
#include <limits>

struct Ray {
    glm::vec3 origin { 0.0f }, direction { 0.0f, 0.0f, -1.0f };
    float t { std::numeric_limits<float>::max() }; // closest hit so far along normalize(direction); max = nothing hit
};
