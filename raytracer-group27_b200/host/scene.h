// Scene data model of the drop-in host layer.  Type and member names are the reference's (src/scene.h:14-97), because
// main.cpp, the ImGui panels and the students' code address them by name; everything a frame needs from a Scene is read
// by renderRayTracing (render.cpp) and BoundingVolumeHierarchy (geometry, spheres, textures) and handed to the device.
#pragma once
#include "mesh.h"
#include "ray.h"
#include <filesystem>
#include <glm/vec3.hpp>
#include <vector>

// preset selector of loadScene; the order is the reference's (main.cpp indexes its combo box with it)
enum SceneType { SingleTriangle, Bookeshelf, Cube, CornellBox, CornellBoxSphericalLight, CornellBoxPlaneLight, Monkey, Teapot, Dragon, Spheres, ChessBoard,
    Custom, AndreasScene, CatalinScene, MikeScene, MikeScene2 };

// geometric helpers used by ray_tracing.h / draw.h call sites
struct Plane { float D = 0.0f; glm::vec3 normal { 0.0f, 1.0f, 0.0f }; };
struct AxisAlignedBox { glm::vec3 lower { 0.0f }, upper { 1.0f }; };

// primitive traced next to the triangles (rt_set_spheres); its material is re-read every frame like the lights
struct Sphere { glm::vec3 center { 0.0f }; float radius = 1.0f; Material material; };

// the four light kinds of shadow.cpp: hard shadows, ring-sampled disc, cone, n x n grid on a parallelogram
struct PointLight { glm::vec3 position, color; };
struct SphericalLight { glm::vec3 position; float radius; glm::vec3 color; };
struct SpotLight { glm::vec3 position, direction; float angle /* degrees */; glm::vec3 color; };
struct PlaneLight {
    glm::vec3 position, width, height, color;
    glm::vec3 center() const { return position + 0.5f * (width + height); }
};

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Sphere> spheres;
    std::vector<PointLight> pointLights;
    std::vector<SphericalLight> sphericalLight;
    std::vector<PlaneLight> planeLight;
    std::vector<SpotLight> spotLight;
};

// Presets with the reference's meshes and light placements (src/scene.cpp:4-150).
Scene loadScene(SceneType type, const std::filesystem::path& dataDir);
