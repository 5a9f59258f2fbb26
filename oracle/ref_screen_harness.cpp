// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Drives the reference's own Screen post-processing (src/screen.cpp, compiled verbatim into oracle/_ref) on a caller's
// image: postprocessImage() (src/screen.cpp:56-69, called at the end of renderRayTracing, src/main.cpp:397-398) or the
// bloom + 8-bit conversion of writeBitmapToFile() (src/screen.cpp:40-53).  Screen keeps its pixels private and offers
// no getter, so this file — and only this file — sees the class with `private` opened; the layout is unchanged.
#include <cmath>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <vector>
#include <glm/glm.hpp> // everything screen.h includes is seen before `private` is redefined
#define private public
#include "screen.h"
#undef private
#include "oracle_api.h"

thread_local std::vector<unsigned char> orc_bmp_rows; // filled by the stbi_write_bmp stand-in (ref_stubs.cpp)

extern "C" int oracle_postprocess(float* rgb, int w, int h, const orc_post* p, int via_write_bitmap, unsigned char* rgba8)
{
    if (!rgb || !p || w <= 0 || h <= 0)
        return 1;
    Screen screen(glm::ivec2(w, h));
    std::memcpy(screen.m_textureData.data(), rgb, sizeof(float) * 3 * (size_t)w * h);
    screen.setBloomFilter((FilteringOption)p->filtering_option);
    screen.setKernel((Kernel)p->kernel);
    screen.setKernelNumRepetitions(p->kernel_repetitions);
    screen.setFilterSize(p->filter_size);
    screen.setSigma(p->sigma);
    screen.setExposure(p->exposure);
    screen.setGammaValue(p->gamma);
    screen.enableGammaCorrection(p->gamma_correction ? 1.0f : 0.0f);
    screen.setBloomFilterLive(p->bloom_live != 0);
    if (via_write_bitmap) {
        screen.writeBitmapToFile("unused.bmp");
        if (rgba8 && orc_bmp_rows.size() == 4 * (size_t)w * h)
            std::memcpy(rgba8, orc_bmp_rows.data(), orc_bmp_rows.size());
    } else {
        screen.postprocessImage();
    }
    if (screen.m_textureData.size() != (size_t)w * h)
        return 2;
    std::memcpy(rgb, screen.m_textureData.data(), sizeof(float) * 3 * (size_t)w * h);
    return 0;
}
