#include "screen.h"
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <stdexcept>

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
void put32(std::ofstream& f, uint32_t v) { f.write(reinterpret_cast<const char*>(&v), 4); }
void put16(std::ofstream& f, uint16_t v) { f.write(reinterpret_cast<const char*>(&v), 2); }
}

// The reference hands RGBA8 rows (top row first) to stbi_write_bmp with comp=4 (src/screen.cpp:44-52).  stb is
// not vendored; this writes an equivalent uncompressed 32-bit BMP (BGRA, bottom-up rows).
void Screen::writeBitmapToFile(const std::filesystem::path& filePath)
{
    const int w = m_resolution.x, h = m_resolution.y;
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
