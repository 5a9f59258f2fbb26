// Scene data model and presets — mirrors src/scene.h:14-97.
#pragma once
#include "mesh.h"
#include "ray.h"
#include <filesystem>
#include <glm/vec3.hpp>
#include <vector>

enum SceneType {
    SingleTriangle,
    Bookeshelf,
    Cube,
    CornellBox,
    CornellBoxSphericalLight,
    CornellBoxPlaneLight,
    Monkey,
    Teapot,
    Dragon,
    Spheres,
    ChessBoard,
    Custom,
    AndreasScene,
    CatalinScene,
    MikeScene,
    MikeScene2
};

struct Plane {
    float D = 0.0f;
    glm::vec3 normal { 0.0f, 1.0f, 0.0f };
};

struct AxisAlignedBox {
    glm::vec3 lower { 0.0f };
    glm::vec3 upper { 1.0f };
};

struct Sphere {
    glm::vec3 center { 0.0f };
    float radius = 1.0f;
    Material material;
};

struct PointLight {
    glm::vec3 position;
    glm::vec3 color;
};

struct SphericalLight {
    glm::vec3 position;
    float radius;
    glm::vec3 color;
};

struct SpotLight {
    glm::vec3 position;
    glm::vec3 direction;
    float angle;
    glm::vec3 color;
};

struct PlaneLight {
    glm::vec3 position;
    glm::vec3 width;
    glm::vec3 height;
    glm::vec3 color;
    glm::vec3 center() const { return position + 0.5f * (width + height); }
};

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Sphere> spheres; // traced (rt_set_spheres), re-read every frame like the lights

    std::vector<PointLight> pointLights;
    std::vector<SphericalLight> sphericalLight;
    std::vector<PlaneLight> planeLight;
    std::vector<SpotLight> spotLight;
};

// Presets with the reference's light placements (src/scene.cpp:4-150).
Scene loadScene(SceneType type, const std::filesystem::path& dataDir);
