// loadMesh — Wavefront OBJ/MTL import for the drop-in Scene API.
//
// The reference imports through assimp 5.0.1 with aiProcess_GenNormals | aiProcess_Triangulate
// (src/mesh.cpp:66-67) and flattens the node tree with a LIFO stack (src/mesh.cpp:77-155).  assimp is a
// build-time download of the reference and is not available here, so the behaviour of its OBJ importer
// for the files in data/ is restated (recalled behaviour of 5.0.1, see SURVEY Appendix B.11 — this is the
// DEFINITION shared by the CPU oracle and the GPU path, not verified byte-fidelity to assimp):
//   * one vertex per face corner, nothing welded (no aiProcess_JoinIdenticalVertices);
//   * `o` and `g` start a new object (= node); `usemtl` starts a new mesh inside the current object when the
//     current mesh already has faces with another material; empty meshes are dropped;
//   * objects become children of the root node and the reference's stack pops them LAST-FIRST, so meshes
//     come out in reverse object order (meshes inside one object keep file order);
//   * quads are split (0,1,2),(0,2,3), fanning from the concave corner if there is one; larger polygons fan
//     from corner 0;
//   * a mesh without `vn` gets per-triangle face normals, cross(v1-v0, v2-v0) normalised, written to the
//     triangle's three corners in triangle order (later triangles overwrite shared quad corners);
//   * MTL: Kd -> kd, Ks -> ks, Ns -> shininess, d -> transparency, Tr -> 1-transparency; a material that
//     never sets them keeps kd 0.6, ks 0, shininess 0, transparency 1 (src/mesh.cpp:144-147 reads those keys).
// Number parsing uses strtof (correctly rounded).
#include "mesh.h"
#include <iostream>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>

namespace {

struct MtlEntry {
    glm::vec3 kd { 0.6f };
    glm::vec3 ks { 0.0f };
    float shininess = 0.0f;
    float opacity = 1.0f;
    std::string mapKd;
};

struct Corner {
    int v = -1, vt = -1, vn = -1;
};

struct RawMesh {
    std::string material; // empty: importer default material
    bool hasMaterial = false;
    std::vector<std::vector<Corner>> faces;
    bool hasNormals = false;
};

struct RawObject {
    std::vector<int> meshes;
};

std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos)
        return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

bool readFloats(const char* p, float* out, int n)
{
    for (int i = 0; i < n; i++) {
        char* end = nullptr;
        out[i] = std::strtof(p, &end);
        if (end == p)
            return false;
        p = end;
    }
    return true;
}

std::map<std::string, MtlEntry> parseMtl(const std::filesystem::path& file)
{
    std::map<std::string, MtlEntry> lib;
    std::ifstream in(file);
    if (!in)
        return lib; // a missing material library is not fatal for assimp either
    std::string line;
    MtlEntry* cur = nullptr;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#')
            continue;
        const size_t sp = line.find_first_of(" \t");
        const std::string key = line.substr(0, sp);
        const std::string rest = sp == std::string::npos ? "" : trim(line.substr(sp));
        if (key == "newmtl") {
            cur = &lib[rest];
            continue;
        }
        if (!cur)
            continue;
        float f[3];
        if (key == "Kd" && readFloats(rest.c_str(), f, 3))
            cur->kd = glm::vec3(f[0], f[1], f[2]);
        else if (key == "Ks" && readFloats(rest.c_str(), f, 3))
            cur->ks = glm::vec3(f[0], f[1], f[2]);
        else if (key == "Ns" && readFloats(rest.c_str(), f, 1))
            cur->shininess = f[0];
        else if (key == "d" && readFloats(rest.c_str(), f, 1))
            cur->opacity = f[0];
        else if (key == "Tr" && readFloats(rest.c_str(), f, 1))
            cur->opacity = 1.0f - f[0];
        else if (key == "map_Kd")
            cur->mapKd = rest;
    }
    return lib;
}

int resolveIndex(long idx, size_t count)
{
    if (idx > 0)
        return int(idx - 1);
    if (idx < 0)
        return int(long(count) + idx);
    return -1;
}

// Index of the corner to fan a quad from: the concave corner if any, else 0.
int quadStartCorner(const glm::vec3 q[4])
{
    for (int i = 0; i < 4; i++) {
        const glm::vec3 v = q[i];
        glm::vec3 left = glm::normalize(q[(i + 3) % 4] - v);
        glm::vec3 diag = glm::normalize(q[(i + 2) % 4] - v);
        glm::vec3 right = glm::normalize(q[(i + 1) % 4] - v);
        const float angle = std::acos(glm::dot(left, diag)) + std::acos(glm::dot(right, diag));
        if (angle > 3.14159265358979323846f)
            return i;
    }
    return 0;
}

void centerAndScaleToUnitMesh(std::vector<Mesh>& meshes) // src/mesh.cpp:164-188
{
    std::vector<glm::vec3> positions;
    for (const Mesh& mesh : meshes)
        for (const Vertex& v : mesh.vertices)
            positions.push_back(v.p);
    if (positions.empty())
        return;
    const glm::vec3 center = std::accumulate(positions.begin(), positions.end(), glm::vec3(0.0f)) / static_cast<float>(positions.size());
    float maxD = 0.0f;
    for (const glm::vec3& p : positions)
        maxD = std::max(glm::length(p - center), maxD);
    for (Mesh& mesh : meshes)
        for (Vertex& v : mesh.vertices)
            v.p = (v.p - center) / maxD;
}

} // namespace

std::vector<Mesh> loadMesh(const std::filesystem::path& file, bool normalize)
{
    std::ifstream in(file);
    if (!in)
        throw std::runtime_error("loadMesh: file " + file.string() + " does not exist");

    std::vector<glm::vec3> positions, normals;
    std::vector<glm::vec2> texCoords;
    std::map<std::string, MtlEntry> materials;
    std::vector<RawMesh> meshes;
    std::vector<RawObject> objects;
    int curObject = -1, curMesh = -1;
    std::string curMaterial; // last `usemtl` that named a known material
    bool haveMaterial = false;
    std::string activeGroup;

    auto createMesh = [&]() {
        meshes.emplace_back();
        curMesh = int(meshes.size()) - 1;
        if (curObject >= 0)
            objects[curObject].meshes.push_back(curMesh);
    };
    auto createObject = [&]() {
        objects.emplace_back();
        curObject = int(objects.size()) - 1;
        createMesh();
        if (haveMaterial) {
            meshes[curMesh].material = curMaterial;
            meshes[curMesh].hasMaterial = true;
        }
    };

    std::string line;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#')
            continue;
        const size_t sp = line.find_first_of(" \t");
        const std::string key = line.substr(0, sp);
        const std::string rest = sp == std::string::npos ? "" : trim(line.substr(sp));
        float f[3];
        if (key == "v") {
            if (!readFloats(rest.c_str(), f, 3))
                throw std::runtime_error("loadMesh: bad vertex line in " + file.string());
            positions.emplace_back(f[0], f[1], f[2]);
        } else if (key == "vn") {
            if (!readFloats(rest.c_str(), f, 3))
                throw std::runtime_error("loadMesh: bad normal line in " + file.string());
            normals.emplace_back(f[0], f[1], f[2]);
        } else if (key == "vt") {
            if (!readFloats(rest.c_str(), f, 2))
                throw std::runtime_error("loadMesh: bad texcoord line in " + file.string());
            texCoords.emplace_back(f[0], f[1]);
        } else if (key == "mtllib") {
            auto lib = parseMtl(file.parent_path() / rest);
            materials.insert(lib.begin(), lib.end());
        } else if (key == "o") {
            createObject();
        } else if (key == "g") {
            if (rest != activeGroup) {
                createObject();
                activeGroup = rest;
            }
        } else if (key == "usemtl") {
            if (haveMaterial && curMaterial == rest)
                continue; // same material as the active one: ignored
            if (materials.find(rest) == materials.end())
                continue; // unknown material: faces keep what they had
            const bool needNew = curMesh < 0
                || (meshes[curMesh].hasMaterial && meshes[curMesh].material != rest && !meshes[curMesh].faces.empty());
            curMaterial = rest;
            haveMaterial = true;
            if (needNew)
                createMesh();
            meshes[curMesh].material = rest;
            meshes[curMesh].hasMaterial = true;
        } else if (key == "f") {
            if (curObject < 0)
                createObject();
            if (curMesh < 0)
                createMesh();
            std::vector<Corner> face;
            std::istringstream ss(rest);
            std::string tok;
            while (ss >> tok) {
                Corner c;
                long idx[3] = { 0, 0, 0 };
                int field = 0;
                const char* p = tok.c_str();
                while (*p && field < 3) {
                    if (*p == '/') {
                        field++;
                        p++;
                        continue;
                    }
                    char* end = nullptr;
                    idx[field] = std::strtol(p, &end, 10);
                    if (end == p)
                        throw std::runtime_error("loadMesh: bad face line in " + file.string());
                    p = end;
                }
                c.v = resolveIndex(idx[0], positions.size());
                c.vt = resolveIndex(idx[1], texCoords.size());
                c.vn = resolveIndex(idx[2], normals.size());
                if (c.v < 0 || c.v >= int(positions.size()))
                    throw std::runtime_error("loadMesh: face index out of range in " + file.string());
                if (c.vn >= 0)
                    meshes[curMesh].hasNormals = true;
                face.push_back(c);
            }
            if (face.size() >= 3)
                meshes[curMesh].faces.push_back(std::move(face));
        }
        // `s` (smoothing groups) and everything else: ignored
    }

    std::vector<Mesh> out;
    // children of the root are popped from a stack: last object first (src/mesh.cpp:77-81,151-154)
    for (int o = int(objects.size()) - 1; o >= 0; o--) {
        for (int mi : objects[o].meshes) {
            const RawMesh& raw = meshes[mi];
            if (raw.faces.empty())
                continue;
            Mesh mesh;
            for (const auto& face : raw.faces) {
                const unsigned base = unsigned(mesh.vertices.size());
                for (const Corner& c : face) {
                    Vertex v;
                    v.p = positions[c.v];
                    v.n = (raw.hasNormals && c.vn >= 0 && c.vn < int(normals.size())) ? normals[c.vn] : glm::vec3(0.0f);
                    v.texCoord = (c.vt >= 0 && c.vt < int(texCoords.size())) ? texCoords[c.vt] : glm::vec2(0.0f);
                    mesh.vertices.push_back(v);
                }
                const unsigned n = unsigned(face.size());
                if (n == 3) {
                    mesh.triangles.emplace_back(base, base + 1, base + 2);
                } else if (n == 4) {
                    const glm::vec3 q[4] = { mesh.vertices[base].p, mesh.vertices[base + 1].p, mesh.vertices[base + 2].p, mesh.vertices[base + 3].p };
                    const unsigned s = unsigned(quadStartCorner(q));
                    const unsigned t0 = base + s, t1 = base + (s + 1) % 4, t2 = base + (s + 2) % 4, t3 = base + (s + 3) % 4;
                    mesh.triangles.emplace_back(t0, t1, t2);
                    mesh.triangles.emplace_back(t0, t2, t3);
                } else {
                    for (unsigned k = 1; k + 1 < n; k++)
                        mesh.triangles.emplace_back(base, base + k, base + k + 1);
                }
            }
            if (!raw.hasNormals) {
                for (const Triangle& t : mesh.triangles) {
                    const glm::vec3& a = mesh.vertices[t.x].p;
                    const glm::vec3& b = mesh.vertices[t.y].p;
                    const glm::vec3& c = mesh.vertices[t.z].p;
                    glm::vec3 n = glm::cross(b - a, c - a);
                    const float len = glm::length(n);
                    if (len > 0.0f)
                        n = n / len;
                    mesh.vertices[t.x].n = mesh.vertices[t.y].n = mesh.vertices[t.z].n = n;
                }
            }
            MtlEntry m; // importer default material
            if (raw.hasMaterial) {
                auto it = materials.find(raw.material);
                if (it != materials.end())
                    m = it->second;
            }
            mesh.material.kd = m.kd;
            mesh.material.ks = m.ks;
            mesh.material.shininess = m.shininess;
            mesh.material.transparency = m.opacity;
            if (!m.mapKd.empty()) {
                // The reference lets Image::Image's exception end the program (src/mesh.cpp:140-145).  Here a texture that
                // is missing or cannot be decoded is reported and the material keeps its kd: the reference snapshot itself
                // lacks some of the files its MTLs name, and the geometry of those scenes is still wanted.
                try {
                    mesh.material.kdTexture = Image(std::filesystem::absolute(file).parent_path() / m.mapKd);
                } catch (const ImageError& e) {
                    std::cerr << e.what() << std::endl;
                }
            }
            out.push_back(std::move(mesh));
        }
    }
    if (out.empty())
        throw std::runtime_error("loadMesh: no faces in " + file.string());
    if (normalize)
        centerAndScaleToUnitMesh(out);
    return out;
}
