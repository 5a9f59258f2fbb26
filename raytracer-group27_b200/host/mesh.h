// Vertex / Material / Mesh / loadMesh — the reference's mesh data model (src/mesh.h:14-46).
#pragma once
#include "image.h"
#include <filesystem>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <optional>
#include <vector>

struct Vertex {
    glm::vec3 p; // position
    glm::vec3 n; // normal
    glm::vec2 texCoord;
};

struct Material {
    glm::vec3 kd; // diffuse colour
    glm::vec3 ks { 0.0f };
    float shininess { 1.0f };
    float transparency { 1.0f };
    std::optional<Image> kdTexture;
};

using Triangle = glm::uvec3;

struct Mesh {
    std::vector<Vertex> vertices;    // one per face corner (nothing is welded, as with assimp without JoinIdenticalVertices)
    std::vector<Triangle> triangles; // indices into vertices
    Material material;
};

// Wavefront OBJ + MTL import.  Throws std::runtime_error when the file is missing or unparsable (the reference
// throws std::exception, src/mesh.cpp:60-73).  normalize = centre on the vertex mean and scale by the
// largest distance (src/mesh.cpp:164-188).
[[nodiscard]] std::vector<Mesh> loadMesh(const std::filesystem::path& file, bool normalize = false);
