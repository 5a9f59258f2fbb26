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
