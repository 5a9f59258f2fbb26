// Screen — the reference's HDR framebuffer (src/screen.h:34-81).  Pixels are float RGB, row 0 at the top
// (setPixel flips y, src/screen.cpp:32-38).  No GL texture is created; draw() is a no-op on a headless box.
// Bloom / tone-mapping / gamma (src/screen.cpp:56-69, 172-395) keep the reference's setters and defaults; the work runs
// on the GPU (rt_postprocess, csrc/rt_post.cu): inside the frame when renderRayTracing renders into this Screen, or on
// the Screen's own pixels when postprocessImage() / writeBitmapToFile() are called directly.
#pragma once
#include "rt_b200.h"
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
    // applies the bloom (when live) and the gamma curve to the pixels (src/screen.cpp:56-69)
    void postprocessImage();

    void setBloomFilterLive(bool bloomFilterLive) { m_post.bloom_live = bloomFilterLive ? 1 : 0; }
    void setBloomFilter(FilteringOption option) { m_post.filtering_option = static_cast<int>(option); }
    void setKernel(Kernel kernel) { m_post.kernel = static_cast<int>(kernel); }
    void setKernelNumRepetitions(int repetitions) { m_post.kernel_repetitions = repetitions < 1 ? 1 : repetitions; }
    void setGammaValue(float gamma) { m_post.gamma = gamma; }
    void enableGammaCorrection(float gammaCorrection) { m_post.gamma_correction = gammaCorrection != 0.0f ? 1 : 0; }
    void setSigma(float sigma) { m_post.sigma = sigma < 0.001f ? 0.001f : sigma; }
    void setExposure(float exposure) { m_post.exposure = exposure; }
    void setFilterSize(int filterSize) { m_post.filter_size = filterSize; }

    // the settings above in the C ABI's form (what renderRayTracing hands to rt_set_postprocess)
    [[nodiscard]] const rt_post_params& postSettings() const { return m_post; }

    [[nodiscard]] glm::ivec2 resolution() const { return m_resolution; }
    [[nodiscard]] std::vector<glm::vec3>& pixels() { return m_textureData; }
    [[nodiscard]] const std::vector<glm::vec3>& pixels() const { return m_textureData; }

    // The pixel storage is page-locked and mapped into the CUDA devices while the Screen lives (when there is a device), so that
    // renderRayTracing's kernels can store the frame into it themselves (rt_render, rt_b200.h); hence no copies of a Screen.
    ~Screen();
    Screen(const Screen&) = delete;
    Screen& operator=(const Screen&) = delete;

private:
    glm::ivec2 m_resolution;
    std::vector<glm::vec3> m_textureData;
    bool m_pageLocked = false;
    // defaults of src/screen.h:84-101: no bloom, box kernel applied once, size 5, sigma 2, exposure 0.5, gamma 2.2 off
    rt_post_params m_post { RT_FILTER_NONE, RT_KERNEL_BOX, 1, 5, 2.0f, 0.5f, 0, 2.2f, 0 };
};
