// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Stand-in for the OpenGL loader header the reference's screen.cpp includes (framework/include/opengl_includes.h:5).
// The oracle only runs Screen's CPU post-processing (src/screen.cpp:40-69, 226-395); the GL calls in its constructor
// and in drawImage() become no-ops so the translation unit compiles and links without a GL context.
#pragma once
typedef unsigned int GLenum, GLuint, GLbitfield;
typedef int GLint, GLsizei;
typedef float GLfloat;
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_NEAREST 0x2600
#define GL_ALL_ATTRIB_BITS 0xFFFFFFFF
#define GL_RGB32F 0x8815
#define GL_RGB 0x1907
#define GL_FLOAT 0x1406
#define GL_LIGHTING 0x0B50
#define GL_LIGHT0 0x4000
#define GL_COLOR_MATERIAL 0x0B57
#define GL_NORMALIZE 0x0BA1
#define GL_TEXTURE0 0x84C0
#define GL_MODELVIEW 0x1700
#define GL_PROJECTION 0x1701
#define GL_QUADS 0x0007
inline void glGenTextures(GLsizei, GLuint* t) { *t = 0; }
inline void glBindTexture(GLenum, GLuint) {}
inline void glTexParameteri(GLenum, GLenum, GLint) {}
inline void glTexImage2D(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void*) {}
inline void glPushAttrib(GLbitfield) {}
inline void glPopAttrib() {}
inline void glDisable(GLenum) {}
inline void glEnable(GLenum) {}
inline void glColor3f(GLfloat, GLfloat, GLfloat) {}
inline void glActiveTexture(GLenum) {}
inline void glMatrixMode(GLenum) {}
inline void glPushMatrix() {}
inline void glPopMatrix() {}
inline void glLoadIdentity() {}
inline void glBegin(GLenum) {}
inline void glEnd() {}
inline void glTexCoord2f(GLfloat, GLfloat) {}
inline void glVertex3f(GLfloat, GLfloat, GLfloat) {}
