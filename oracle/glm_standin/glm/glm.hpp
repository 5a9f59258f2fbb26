// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Minimal stand-in for the subset of glm 0.9.9.8 (pinned by the reference at
// framework/cmake/download_framework_packages.cmake:19-22) that the reference's ray-tracing
// translation units use.  glm itself is not vendored in /root/reference and there is no network,
// so the *scalar* formulas glm 0.9.9.8 evaluates (GLM_FORCE_INTRINSICS is not defined by the
// reference) are restated here, keeping glm's operation order because the float results feed
// bit-exact closest-hit comparisons.  Parity of this header against real glm is UNPINNED (no
// copy of glm is reachable from this sandbox); it is anchored on glm's published formulas:
//   dot(vec3)      = (a.x*b.x + a.y*b.y) + a.z*b.z           (detail::compute_dot<vec<3>>)
//   cross          = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   inversesqrt    = 1 / sqrt(x);  normalize(v) = v * inversesqrt(dot(v,v));  length = sqrt(dot)
//   reflect(I,N)   = I - N * dot(N,I) * 2
//   min(a,b)       = (b < a) ? b : a ;  max(a,b) = (a < b) ? b : a
//   quat(euler)    and quat * vec3 as in glm/detail/type_quat.inl
//   mat3 ops       as in glm/detail/type_mat3x3.inl (column major)
// Only oracle/ (the parity checker) includes this file.
#pragma once
#include <array> // real glm pulls this in transitively; src/ray_tracing.h relies on it
#include <cmath>
#include <cstddef>
#include <limits>

namespace glm {

using std::pow;   // glm's func_exponential.inl does `using std::pow;` for scalars
using std::sqrt;  // likewise `using std::sqrt;`
using std::sin;
using std::cos;
using std::log2;  // func_exponential.inl with GLM_HAS_CXX11_STL: `using std::log2;` (src/ray_differentials.cpp:138)
using std::exp;   // func_exponential.inl: `using std::exp;` for scalars (src/screen.cpp:325,363)

typedef int length_t;

template <typename T> struct tvec2 {
    union { T x; T r; T s; };
    union { T y; T g; T t; };
    constexpr tvec2() : x(0), y(0) {}
    constexpr explicit tvec2(T v) : x(v), y(v) {}
    constexpr tvec2(T a, T b) : x(a), y(b) {}
    template <typename A, typename B> constexpr tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
    template <typename U> constexpr tvec2(const tvec2<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    T& operator[](int i) { return i == 0 ? x : y; }
    const T& operator[](int i) const { return i == 0 ? x : y; }
};

template <typename T> struct tvec4;

template <typename T> struct tvec3 {
    union { T x; T r; T s; };
    union { T y; T g; T t; };
    union { T z; T b; T p; };
    constexpr tvec3() : x(0), y(0), z(0) {}
    constexpr explicit tvec3(T v) : x(v), y(v), z(v) {}
    constexpr tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    template <typename A, typename B, typename C>
    constexpr tvec3(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
    template <typename U> constexpr tvec3(const tvec3<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)) {}
    constexpr tvec3(const tvec4<T>& o);
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(const tvec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    template <typename U> tvec3& operator*=(U s_) { x *= static_cast<T>(s_); y *= static_cast<T>(s_); z *= static_cast<T>(s_); return *this; }
    template <typename U> tvec3& operator/=(U s_) { x /= static_cast<T>(s_); y /= static_cast<T>(s_); z /= static_cast<T>(s_); return *this; }
};

template <typename T> struct tvec4 {
    T x, y, z, w;
    constexpr tvec4() : x(0), y(0), z(0), w(0) {}
    constexpr explicit tvec4(T v) : x(v), y(v), z(v), w(v) {}
    constexpr tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    constexpr tvec4(const tvec3<T>& v, T d) : x(v.x), y(v.y), z(v.z), w(d) {}
    template <typename U> constexpr explicit tvec4(const tvec4<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)), z(static_cast<T>(o.z)), w(static_cast<T>(o.w)) {}
};
template <typename T> constexpr tvec3<T>::tvec3(const tvec4<T>& o) : x(o.x), y(o.y), z(o.z) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec4<unsigned char> u8vec4; // src/screen.cpp:44-49: glm::u8vec4(glm::vec4 * 255.0f) converts per component with static_cast
typedef tvec2<int> ivec2;
typedef tvec3<unsigned int> uvec3;
typedef tvec3<bool> bvec3;

// ---- vec2 arithmetic ----
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }

// ---- vec3 arithmetic (component-wise, no reciprocal tricks) ----
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3& a, const vec3& b) { return !(a == b); }

inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

// ---- common ----
inline float abs(float v) { return std::fabs(v); }
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline int min(int a, int b) { return (b < a) ? b : a; }
inline int max(int a, int b) { return (a < b) ? b : a; }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec3 clamp(const vec3& v, const vec3& lo, const vec3& hi) { return min(max(v, lo), hi); }
inline vec3 exp(const vec3& v) { return vec3(std::exp(v.x), std::exp(v.y), std::exp(v.z)); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline vec3 radians(const vec3& d) { return vec3(radians(d.x), radians(d.y), radians(d.z)); }
inline vec3 pow(const vec3& b, const vec3& e) { return vec3(std::pow(b.x, e.x), std::pow(b.y, e.y), std::pow(b.z, e.z)); }
inline vec3 cos(const vec3& v) { return vec3(std::cos(v.x), std::cos(v.y), std::cos(v.z)); }
inline vec3 sin(const vec3& v) { return vec3(std::sin(v.x), std::sin(v.y), std::sin(v.z)); }

// ---- geometric ----
inline float dot(const vec3& a, const vec3& b) {
    const vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
inline vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float dot(const vec2& a, const vec2& b) {
    const vec2 tmp(a.x * b.x, a.y * b.y);
    return tmp.x + tmp.y;
}
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }

// ---- vector relational ----
inline bvec3 equal(const vec3& a, const vec3& b) { return bvec3(a.x == b.x, a.y == b.y, a.z == b.z); }
inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bool all(const bvec3& v) { return v.x && v.y && v.z; }
inline bool any(const bvec3& v) { return v.x || v.y || v.z; }

// ---- mat3 (column major, glm/detail/type_mat3x3.inl) ----
struct mat3 {
    vec3 c[3];
    mat3() : c { vec3(1, 0, 0), vec3(0, 1, 0), vec3(0, 0, 1) } {}
    explicit mat3(float s) : c { vec3(s, 0, 0), vec3(0, s, 0), vec3(0, 0, s) } {}
    mat3(const vec3& c0, const vec3& c1, const vec3& c2) : c { c0, c1, c2 } {}
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 operator*(const mat3& m, float s) { return mat3(m[0] * s, m[1] * s, m[2] * s); }
inline mat3 operator+(const mat3& a, const mat3& b) { return mat3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(
        m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
        m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
        m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
inline mat3 operator*(const mat3& m1, const mat3& m2) {
    const float A00 = m1[0][0], A01 = m1[0][1], A02 = m1[0][2];
    const float A10 = m1[1][0], A11 = m1[1][1], A12 = m1[1][2];
    const float A20 = m1[2][0], A21 = m1[2][1], A22 = m1[2][2];
    const float B00 = m2[0][0], B01 = m2[0][1], B02 = m2[0][2];
    const float B10 = m2[1][0], B11 = m2[1][1], B12 = m2[1][2];
    const float B20 = m2[2][0], B21 = m2[2][1], B22 = m2[2][2];
    mat3 R;
    R[0][0] = A00 * B00 + A10 * B01 + A20 * B02;
    R[0][1] = A01 * B00 + A11 * B01 + A21 * B02;
    R[0][2] = A02 * B00 + A12 * B01 + A22 * B02;
    R[1][0] = A00 * B10 + A10 * B11 + A20 * B12;
    R[1][1] = A01 * B10 + A11 * B11 + A21 * B12;
    R[1][2] = A02 * B10 + A12 * B11 + A22 * B12;
    R[2][0] = A00 * B20 + A10 * B21 + A20 * B22;
    R[2][1] = A01 * B20 + A11 * B21 + A21 * B22;
    R[2][2] = A02 * B20 + A12 * B21 + A22 * B22;
    return R;
}

// ---- quaternion (glm/detail/type_quat.inl) ----
struct quat {
    float x, y, z, w;
    quat() : x(0), y(0), z(0), w(1) {}
    explicit quat(const vec3& eulerAngle) {
        const vec3 c = glm::cos(eulerAngle * 0.5f);
        const vec3 s = glm::sin(eulerAngle * 0.5f);
        w = c.x * c.y * c.z + s.x * s.y * s.z;
        x = s.x * c.y * c.z - c.x * s.y * s.z;
        y = c.x * s.y * c.z + s.x * c.y * s.z;
        z = c.x * c.y * s.z - s.x * s.y * c.z;
    }
};
inline vec3 operator*(const quat& q, const vec3& v) {
    const vec3 QuatVector(q.x, q.y, q.z);
    const vec3 uv(cross(QuatVector, v));
    const vec3 uuv(cross(QuatVector, uv));
    return v + ((uv * q.w) + uuv) * 2.0f;
}

} // namespace glm
