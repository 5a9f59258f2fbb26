// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Link-time stand-ins for the three symbols of the reference's GL debug drawer (src/draw.h) that its
// ray-tracing translation units reference.  drawRay is a no-op in the reference whenever
// enableDrawRay == false (src/draw.cpp:185-187), which is the case for every rendered frame
// (src/main.cpp:736-738); here it additionally counts calls so the harness can report how many
// closest-hit queries the verbatim cansee loop issued (one drawRay per loop iteration).
#include "draw.h"

bool enableDrawRay = false;
thread_local unsigned long long orc_drawray_calls = 0;

void drawRay(const Ray&, const glm::vec3&) { orc_drawray_calls++; }
void drawAABB(const AxisAlignedBox&, DrawMode, const glm::vec3&, float) {}

// Screen::writeBitmapToFile's sink (see gl_standin/stb_image_write.h): keep the rows instead of writing a file.
#include <vector>
extern thread_local std::vector<unsigned char> orc_bmp_rows;
extern "C" int stbi_write_bmp(char const*, int w, int h, int comp, const void* data)
{
    const unsigned char* b = static_cast<const unsigned char*>(data);
    orc_bmp_rows.assign(b, b + (size_t)w * h * comp);
    return 1;
}

// Image::Image's source of texels (see gl_standin/stb_image.h).  The registry is filled by oracle_set_textures.
#include <cstdlib>
#include <cstring>
#include <string>
struct OrcRegisteredTexture {
    int w, h;
    std::vector<unsigned char> rgb;
};
std::vector<OrcRegisteredTexture> orc_registered_textures;
extern "C" unsigned char* stbi_load(char const* filename, int* x, int* y, int* channels_in_file, int)
{
    const std::string name(filename);
    const size_t us = name.find_last_of('_');
    const int k = us == std::string::npos ? -1 : std::atoi(name.c_str() + us + 1);
    if (k < 0 || k >= (int)orc_registered_textures.size())
        return nullptr;
    const OrcRegisteredTexture& t = orc_registered_textures[k];
    *x = t.w;
    *y = t.h;
    *channels_in_file = 3;
    unsigned char* p = static_cast<unsigned char*>(std::malloc(t.rgb.size()));
    std::memcpy(p, t.rgb.data(), t.rgb.size());
    return p;
}
extern "C" void stbi_image_free(void* p) { std::free(p); }
