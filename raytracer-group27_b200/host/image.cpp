// PNG reader behind Image::Image.  The reference calls stbi_load(path, &w, &h, &channels, STBI_rgb) (src/image.cpp:45):
// 8-bit RGB whatever the file holds (palette expanded, grey replicated, 16-bit samples reduced to their high byte, alpha
// dropped), `channels` = what the file has (palette 3, or 4 with a tRNS chunk).  The same conversions are done here.
#include "image.h"
#include <cstdint>
#include <cstring>
#include <fstream>
#include <zlib.h>

namespace {

uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

struct Png {
    int w = 0, h = 0, channels = 0; // channels as stb reports them
    std::vector<unsigned char> rgb;
};

Png decodePng(const std::filesystem::path& path)
{
    std::ifstream f(path, std::ios::binary);
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0)
        throw ImageError("Failed to read texture " + path.string() + ": not a PNG file (the only format this reader decodes)");
    int depth = 0, ctype = 0, interlace = 0;
    Png out;
    std::vector<unsigned char> palette, idat;
    bool has_trns = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + (size_t)len > file.size())
            throw ImageError("Failed to read texture " + path.string() + ": truncated chunk");
        const unsigned char* data = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            out.w = (int)be32(data);
            out.h = (int)be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            has_trns = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (out.w <= 0 || out.h <= 0 || interlace != 0)
        throw ImageError("Failed to read texture " + path.string() + ": empty or interlaced PNG");
    const int samples = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!samples || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) || (ctype == 3 && (depth == 16 || palette.empty())))
        throw ImageError("Failed to read texture " + path.string() + ": unsupported PNG colour type / bit depth");
    out.channels = ctype == 3 ? (has_trns ? 4 : 3) : (samples + ((ctype == 0 || ctype == 2) && has_trns ? 1 : 0));
    const size_t row_bytes = ((size_t)out.w * samples * depth + 7) / 8;
    const int bpp = std::max(1, samples * depth / 8);
    std::vector<unsigned char> raw((row_bytes + 1) * (size_t)out.h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        throw ImageError("Failed to read texture " + path.string() + ": corrupt image data");
    std::vector<unsigned char> prev(row_bytes, 0), cur(row_bytes);
    out.rgb.resize((size_t)out.w * out.h * 3);
    for (int y = 0; y < out.h; y++) {
        const unsigned char* line = &raw[(row_bytes + 1) * (size_t)y];
        const int filter = line[0];
        for (size_t i = 0; i < row_bytes; i++) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int v = line[1 + i];
            switch (filter) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) / 2; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            cur[i] = (unsigned char)v;
        }
        for (int x = 0; x < out.w; x++) {
            unsigned char s[4] = { 0, 0, 0, 0 };
            for (int k = 0; k < samples; k++) {
                const size_t bit = ((size_t)x * samples + k) * depth;
                if (depth == 16)
                    s[k] = cur[bit / 8]; // high byte
                else if (depth == 8)
                    s[k] = cur[bit / 8];
                else {
                    const int v = (cur[bit / 8] >> (8 - depth - (int)(bit % 8))) & ((1 << depth) - 1);
                    s[k] = ctype == 3 ? (unsigned char)v : (unsigned char)(v * 255 / ((1 << depth) - 1)); // grey samples scale to 0..255
                }
            }
            unsigned char* o = &out.rgb[3 * ((size_t)y * out.w + x)];
            if (ctype == 3) {
                const size_t e = 3 * (size_t)s[0];
                for (int k = 0; k < 3; k++)
                    o[k] = e + k < palette.size() ? palette[e + k] : 0;
            } else if (ctype == 0 || ctype == 4) {
                o[0] = o[1] = o[2] = s[0];
            } else {
                o[0] = s[0];
                o[1] = s[1];
                o[2] = s[2];
            }
        }
        prev.swap(cur);
    }
    return out;
}

} // namespace

Image::Image(const std::filesystem::path& filePath)
    : m_path(filePath)
{
    if (!std::filesystem::exists(filePath)) // src/image.cpp:38-41
        throw ImageError("Texture file " + filePath.string() + " does not exists!");
    Png png = decodePng(filePath);
    if (png.channels < 3) // src/image.cpp:47-50
        throw ImageError("Only textures with 3 or more color channels are supported. " + filePath.string() + " has " + std::to_string(png.channels) + " channels");
    m_width = png.w;
    m_height = png.h;
    // The reference walks the RGB bytes with a stride of `channels` (src/image.cpp:57-59), which for 4-channel files reads
    // shifted texels and runs off the buffer; here every pixel contributes its own R, G, B.
    m_pixels.reserve((size_t)m_width * m_height);
    for (size_t i = 0; i < (size_t)m_width * m_height; i++)
        m_pixels.emplace_back(png.rgb[3 * i] / 255.0f, png.rgb[3 * i + 1] / 255.0f, png.rgb[3 * i + 2] / 255.0f);
}
