#include "screen.h"
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>

Screen::Screen(const glm::ivec2& resolution)
    : m_resolution(resolution), m_textureData(size_t(resolution.x) * size_t(resolution.y), glm::vec3(0.0f))
{
}

void Screen::clear(const glm::vec3& color) { std::fill(m_textureData.begin(), m_textureData.end(), color); }

void Screen::setPixel(int x, int y, const glm::vec3& color)
{
    // (0,0) is the bottom-left pixel for callers; storage starts at the top row (src/screen.cpp:32-38)
    m_textureData[size_t(m_resolution.y - 1 - y) * m_resolution.x + x] = color;
}

namespace {
// Screen has no link to a BoundingVolumeHierarchy, so its own post-processing calls share one lazily created context,
// on the CUDA device that is current for the calling thread at that moment (device 0 unless the caller chose another).
rt_ctx* postContext()
{
    static rt_ctx* ctx = nullptr;
    if (!ctx) {
        if (rt_create(rt_current_device(), &ctx) != RT_OK)
            throw std::runtime_error(std::string("Screen: ") + rt_last_error());
    }
    return ctx;
}
}

void Screen::postprocessImage()
{
    if (!((m_post.bloom_live && m_post.filtering_option != RT_FILTER_NONE) || m_post.gamma_correction) || m_textureData.empty())
        return;
    static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "Screen pixels must be packed float3");
    if (rt_postprocess(postContext(), &m_post, &m_textureData[0].x, m_resolution.x, m_resolution.y, 0, nullptr) != RT_OK)
        throw std::runtime_error(std::string("Screen::postprocessImage: ") + rt_last_error());
}

namespace {
void put32(std::ofstream& f, uint32_t v) { f.write(reinterpret_cast<const char*>(&v), 4); }
void put16(std::ofstream& f, uint16_t v) { f.write(reinterpret_cast<const char*>(&v), 2); }
}

// The reference applies the bloom to its pixels (whether or not it is live, src/screen.cpp:42), then hands RGBA8 rows (top
// row first) to stbi_write_bmp with comp=4 (src/screen.cpp:44-52).  stb is not vendored; this writes an equivalent
// uncompressed 32-bit BMP (BGRA, bottom-up rows).
void Screen::writeBitmapToFile(const std::filesystem::path& filePath)
{
    const int w = m_resolution.x, h = m_resolution.y;
    if (m_post.filtering_option != RT_FILTER_NONE && !m_textureData.empty()
        && rt_postprocess(postContext(), &m_post, &m_textureData[0].x, w, h, 1, nullptr) != RT_OK)
        throw std::runtime_error(std::string("Screen::writeBitmapToFile: ") + rt_last_error());
    std::ofstream f(filePath, std::ios::binary);
    if (!f)
        throw std::runtime_error("Screen::writeBitmapToFile: cannot open " + filePath.string());
    const uint32_t dataSize = uint32_t(w) * uint32_t(h) * 4u;
    f.put('B').put('M');
    put32(f, 14 + 40 + dataSize);
    put32(f, 0);
    put32(f, 14 + 40);
    put32(f, 40);
    put32(f, uint32_t(w));
    put32(f, uint32_t(h));
    put16(f, 1);
    put16(f, 32);
    put32(f, 0);
    put32(f, dataSize);
    put32(f, 2835);
    put32(f, 2835);
    put32(f, 0);
    put32(f, 0);
    std::vector<uint8_t> row(size_t(w) * 4);
    for (int y = h - 1; y >= 0; y--) {
        for (int x = 0; x < w; x++) {
            const glm::vec3 c = glm::clamp(m_textureData[size_t(y) * w + x], 0.0f, 1.0f);
            row[4 * x + 0] = uint8_t(c.z * 255.0f);
            row[4 * x + 1] = uint8_t(c.y * 255.0f);
            row[4 * x + 2] = uint8_t(c.x * 255.0f);
            row[4 * x + 3] = 255;
        }
        f.write(reinterpret_cast<const char*>(row.data()), std::streamsize(row.size()));
    }
}
