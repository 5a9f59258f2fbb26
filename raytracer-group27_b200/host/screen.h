// Screen — the reference's HDR framebuffer (src/screen.h:34-81).  Pixels are float RGB, row 0 at the top
// (setPixel flips y, src/screen.cpp:32-38).  No GL texture is created; draw() is a no-op on a headless box.
// Bloom / tone-mapping / gamma (src/screen.cpp:56-395) are a "next" row: the setters are
// kept so reference call sites compile, and postprocessImage() does what the reference does with its defaults
// (nothing).
#pragma once
#include <filesystem>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <vector>

enum class FilteringOption { None, Bloom, BloomWithReinhardHdr, BloomWithExposureHdr, OnlyLight, OnlyLightWithKernel };
enum class Kernel { BoxKernel, GaussianKernel };

class Screen {
public:
    explicit Screen(const glm::ivec2& resolution);

    void clear(const glm::vec3& color);
    void setPixel(int x, int y, const glm::vec3& color);
    // 32-bit BMP, colours clamped to [0,1] and truncated to 8 bits (src/screen.cpp:40-53)
    void writeBitmapToFile(const std::filesystem::path& filePath);
    void draw() {}
    void postprocessImage() {}

    void setBloomFilterLive(bool) {}
    void setBloomFilter(FilteringOption) {}
    void setKernel(Kernel) {}
    void setKernelNumRepetitions(int) {}
    void setGammaValue(float) {}
    void enableGammaCorrection(float) {}
    void setSigma(float) {}
    void setExposure(float) {}
    void setFilterSize(int) {}

    [[nodiscard]] glm::ivec2 resolution() const { return m_resolution; }
    [[nodiscard]] std::vector<glm::vec3>& pixels() { return m_textureData; }
    [[nodiscard]] const std::vector<glm::vec3>& pixels() const { return m_textureData; }

private:
    glm::ivec2 m_resolution;
    std::vector<glm::vec3> m_textureData;
};
