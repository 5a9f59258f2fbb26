// Image — the reference's texture class (src/image.h:33-47) as far as the rebuilt path needs it: the texels, decoded
// like Image::Image does (8-bit RGB through the image reader, each byte / 255.0f, src/image.cpp:36-59), and the knobs
// main.cpp sets on a texture before sampling it (src/main.cpp:159-162).  Sampling itself — Image::getPixel with all five
// filters, the mip pyramid and the level of detail from the ray differentials — runs on the device (csrc/rt_kernels.cu:
// sample_texture, level_of_detail; the pyramid is built by rt_set_textures).  stb_image is not vendored by the reference; PNG files (every
// texture the reference ships) are decoded here with zlib.
#pragma once
#include <exception>
#include <filesystem>
#include <glm/vec3.hpp>
#include <string>
#include <vector>

enum class OutOfBoundsRule { Border, Clamp, Repeat };
enum class TextureFiltering { NearestNeighbor, Bilinear, MipMappingNearestLevelNearestNeighbor, MipMappingNearestLevelBilinear, Trilinear };

class ImageError : public std::exception { // the reference throws a bare std::exception after a message on stderr
public:
    explicit ImageError(std::string what) : m_what(std::move(what)) {}
    const char* what() const noexcept override { return m_what.c_str(); }

private:
    std::string m_what;
};

class Image {
public:
    // throws ImageError when the file is missing, cannot be decoded or has fewer than 3 colour channels (src/image.cpp:38-54)
    explicit Image(const std::filesystem::path& filePath);

    void setBorderColor(const glm::vec3 color) { _borderColor = color; }
    void setOutOfBoundsRuleX(const OutOfBoundsRule rule) { _outOfBoundRuleX = rule; }
    void setOutOfBoundsRuleY(const OutOfBoundsRule rule) { _outOfBoundRuleY = rule; }
    void setTextureFilteringMethod(const TextureFiltering method) { _filteringMethod = method; }
    // true if the texture is 2^N x 2^N (src/image.cpp:411-413)
    bool canUseMipmapping() const { return ((m_height & (m_height - 1)) == 0) && ((m_width & (m_width - 1)) == 0) && (m_width == m_height); }

    [[nodiscard]] int width() const { return m_width; }
    [[nodiscard]] int height() const { return m_height; }
    [[nodiscard]] const std::vector<glm::vec3>& pixels() const { return m_pixels; } // top row first
    [[nodiscard]] const std::filesystem::path& path() const { return m_path; }
    [[nodiscard]] glm::vec3 borderColor() const { return _borderColor; }
    [[nodiscard]] OutOfBoundsRule outOfBoundsRuleX() const { return _outOfBoundRuleX; }
    [[nodiscard]] OutOfBoundsRule outOfBoundsRuleY() const { return _outOfBoundRuleY; }
    [[nodiscard]] TextureFiltering textureFilteringMethod() const { return _filteringMethod; }

private:
    std::filesystem::path m_path;
    int m_width = 0, m_height = 0;
    std::vector<glm::vec3> m_pixels;
    OutOfBoundsRule _outOfBoundRuleX = OutOfBoundsRule::Border, _outOfBoundRuleY = OutOfBoundsRule::Border;
    glm::vec3 _borderColor = glm::vec3(0);
    TextureFiltering _filteringMethod = TextureFiltering::NearestNeighbor;
};
