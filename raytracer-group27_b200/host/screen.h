// Screen — the reference's HDR framebuffer (src/screen.h:34-81).  Pixels are float RGB, row 0 at the top
// (setPixel flips y, src/screen.cpp:32-38).  No GL texture is created; draw() is a no-op on a headless box.
// Bloom / tone-mapping / gamma (src/screen.cpp:56-69, 172-395) keep the reference's setters and defaults; the work runs
// on the GPU (rt_postprocess, csrc/rt_post.cu): inside the frame when renderRayTracing renders into this Screen, or on
// the Screen's own pixels when postprocessImage() / writeBitmapToFile() are called directly.
#pragma once
#include "rt_b200.h"
#include <cstddef>
#include <cstdlib>
#include <filesystem>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>
#include <new>
#include <vector>

enum class FilteringOption { None, Bloom, BloomWithReinhardHdr, BloomWithExposureHdr, OnlyLight, OnlyLightWithKernel };
enum class Kernel { BoxKernel, GaussianKernel };

// Allocator of the Screen's pixels: page-locked, device-mapped memory from the CUDA allocator (rt_host_alloc) when there is a device,
// so that renderRayTracing's kernels can store the frame into it themselves (rt_render's fast path, rt_b200.h) at the link's full rate —
// ordinary pages registered afterwards (rt_host_register on a std::vector's storage, the first version) take the same path 4 % slower
// (tools/host_mem_probe.py: 2.66 against 2.56 ms for the 4K frame) — and plain memory where there is none (the frame then arrives through
// staged copies).  A 64-byte header in front of the pixels remembers which of the two it was.
template <typename T> struct ScreenAllocator {
    using value_type = T;
    ScreenAllocator() = default;
    template <typename U> ScreenAllocator(const ScreenAllocator<U>&) {}
    static constexpr std::size_t kHeader = 64;
    T* allocate(std::size_t n)
    {
        const std::size_t bytes = kHeader + n * sizeof(T);
        void* p = nullptr;
        const bool pinned = rt_host_alloc(bytes, &p) == RT_OK;
        if (!pinned && !(p = std::malloc(bytes)))
            throw std::bad_alloc();
        *static_cast<int*>(p) = pinned ? 1 : 0;
        return reinterpret_cast<T*>(static_cast<char*>(p) + kHeader);
    }
    void deallocate(T* q, std::size_t) noexcept
    {
        void* p = reinterpret_cast<char*>(q) - kHeader;
        if (*static_cast<int*>(p))
            rt_host_free(p);
        else
            std::free(p);
    }
    template <typename U> bool operator==(const ScreenAllocator<U>&) const { return true; }
    template <typename U> bool operator!=(const ScreenAllocator<U>&) const { return false; }
};

class Screen {
public:
    using Pixels = std::vector<glm::vec3, ScreenAllocator<glm::vec3>>;

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
    [[nodiscard]] Pixels& pixels() { return m_textureData; }
    [[nodiscard]] const Pixels& pixels() const { return m_textureData; }

    // (the pixels are page-locked memory, a scarce resource: no accidental copies of a Screen)
    Screen(const Screen&) = delete;
    Screen& operator=(const Screen&) = delete;

private:
    glm::ivec2 m_resolution;
    Pixels m_textureData; // src/screen.h:85, in ScreenAllocator's memory
    // defaults of src/screen.h:84-101: no bloom, box kernel applied once, size 5, sigma 2, exposure 0.5, gamma 2.2 off
    rt_post_params m_post { RT_FILTER_NONE, RT_KERNEL_BOX, 1, 5, 2.0f, 0.5f, 0, 2.2f, 0 };
};
