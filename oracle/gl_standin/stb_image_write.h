// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Stand-in for stb_image_write.h (not vendored by the reference): Screen::writeBitmapToFile (src/screen.cpp:40-53) hands
// its 8-bit RGBA rows to stbi_write_bmp; the oracle's definition (ref_stubs.cpp) keeps them for the harness to return.
#pragma once
extern "C" int stbi_write_bmp(char const* filename, int w, int h, int comp, const void* data);
