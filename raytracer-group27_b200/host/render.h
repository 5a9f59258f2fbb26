// renderRayTracing — the render entry main.cpp calls (src/main.cpp:340-341), same parameter list.  The knobs
// the reference keeps as file-scope statics in main.cpp (src/main.cpp:58-60,123-127) are exported here under
// the same names so a main.cpp that includes this header can drop its own definitions.
#pragma once
#include "bounding_volume_hierarchy.h"
#include "scene.h"
#include "screen.h"
#include "trackball.h"

extern int max_reflection_level;   // default 5
extern int sphere_light_ray_count; // default 10
extern int plane_light_1D_ray_count; // default 3
extern int glossy_ray_count;       // default 1 here (mirror ray only; the reference's 10 draws from rand()); > 1 uses the defined stream of rt_b200.h
extern float refraction_factor;    // default 0.8
extern bool useBVH;                // reference default false: which of two objects at exactly the same t is reported (rt_b200.h)
// texture knobs (src/main.cpp:54-58)
extern TextureFiltering textureFiltering;  // default NearestNeighbor; the mip-mapped modes sample level of detail 0 (rt_b200.h)
extern OutOfBoundsRule outOfBoundsRuleX, outOfBoundsRuleY; // default Border
extern glm::vec3 textureBorderColor;       // default black
extern bool useTextures;                   // default false

struct RenderTimings {
    float gpu_ms = 0.0f;
    unsigned long long rays = 0; // primary + shadow queries + reflection/refraction rays
};
extern RenderTimings lastRenderTimings;

void renderRayTracing(Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen,
    bool textureDebugging = false, bool anti_aliasing = false, bool multipleRays = false, int sampleSize = 4);
