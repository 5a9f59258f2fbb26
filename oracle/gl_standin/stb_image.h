// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Stand-in for stb_image.h (not vendored by the reference): Image::Image (src/image.cpp:36-73) asks stbi_load for 8-bit
// RGB rows.  The oracle's definition (ref_stubs.cpp) serves images the harness registered (oracle_set_textures) under
// placeholder file names ending in "_<index>", so the reference's own Image code runs on caller-supplied texels.
#pragma once
typedef unsigned char stbi_uc;
enum { STBI_default = 0, STBI_grey = 1, STBI_grey_alpha = 2, STBI_rgb = 3, STBI_rgb_alpha = 4 };
extern "C" stbi_uc* stbi_load(char const* filename, int* x, int* y, int* channels_in_file, int desired_channels);
extern "C" void stbi_image_free(void* retval_from_stbi_load);
