// The reference's GL debug drawer (src/draw.h) has no meaning on a headless GPU box; the two hooks the render
// path calls are kept as no-ops so reference-style call sites still compile.
#pragma once
#include "scene.h"
enum class DrawMode { Filled, Wireframe };
inline bool enableDrawRay = false;
inline void drawRay(const Ray&, const glm::vec3& = glm::vec3(1.0f)) {}
inline void drawAABB(const AxisAlignedBox&, DrawMode = DrawMode::Filled, const glm::vec3& = glm::vec3(1.0f), float = 1.0f) {}
