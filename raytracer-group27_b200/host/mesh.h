// Vertex / Material / Mesh / loadMesh — the reference's mesh data model (src/mesh.h:14-46).
#pragma once
#include "image.h"
#include <filesystem>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <optional>
#include <vector>

// One corner of a face: position, normal, texture coordinate (aggregate order as in the reference: { p, n, texCoord }).
struct Vertex { glm::vec3 p, n; glm::vec2 texCoord; };

// Phong material of a mesh.  The defaults are what loadMesh leaves when the MTL does not say otherwise: no specular
// term, shininess 1, opaque (transparency 1 = the MTL's `d`), no texture.
struct Material {
    glm::vec3 kd, ks { 0.0f };
    float shininess { 1.0f }, transparency { 1.0f };
    std::optional<Image> kdTexture;
};

using Triangle = glm::uvec3; // three indices into Mesh::vertices

// One material group of the file.  Nothing is welded: every face corner is its own Vertex (as with assimp without
// JoinIdenticalVertices), so triangles[i] = (3i, 3i+1, 3i+2) after triangulation.
struct Mesh { std::vector<Vertex> vertices; std::vector<Triangle> triangles; Material material; };

// Wavefront OBJ + MTL import.  Throws std::runtime_error when the file is missing or unparsable (the reference
// throws std::exception, src/mesh.cpp:60-73).  normalize = centre on the vertex mean and scale by the
// largest distance (src/mesh.cpp:164-188).
[[nodiscard]] std::vector<Mesh> loadMesh(const std::filesystem::path& file, bool normalize = false);
