// Image — placeholder for the reference's texture class (src/image.h:33-47).  Texture sampling is outside
// the rebuilt hot path (useTextures defaults to false, src/main.cpp:58); the type only records the path
// named by an MTL `map_Kd` so that Material keeps its reference shape (src/mesh.h:29).
#pragma once
#include <filesystem>

class Image {
public:
    explicit Image(const std::filesystem::path& filePath) : m_path(filePath) {}
    const std::filesystem::path& path() const { return m_path; }

private:
    std::filesystem::path m_path;
};
