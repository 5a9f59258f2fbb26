// glm-compatible math subset for the host-side drop-in API.
//
// The reference's public types (Vertex, Material, Scene, Ray, Screen, Trackball ...) are spelled with
// glm::vec3 & co. (src/mesh.h:14-31, src/scene.h:36-94, framework/include/ray.h:11-29).  glm 0.9.9.8 is a
// build-time download of the reference (framework/cmake/download_framework_packages.cmake:19-22) and is not
// available offline, so the host layer ships this small source-compatible subset: just the types and
// functions the kept API surface, the OBJ loader and the camera set-up need.  All device-side arithmetic
// lives in csrc/ and does not use this header.  Scalar evaluation order follows glm's non-SIMD code path
// (dot = (x+y)+z, normalize = v * (1/sqrt(dot)), quat*vec3 via two cross products).
#pragma once
#include <cmath>
#include <cstdint>

namespace glm {

template <typename T> struct vec2_t {
    T x {}, y {};
    constexpr vec2_t() = default;
    constexpr explicit vec2_t(T s) : x(s), y(s) {}
    constexpr vec2_t(T x_, T y_) : x(x_), y(y_) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T> struct vec4_t;

template <typename T> struct vec3_t {
    T x {}, y {}, z {};
    constexpr vec3_t() = default;
    constexpr explicit vec3_t(T s) : x(s), y(s), z(s) {}
    constexpr vec3_t(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    template <typename A, typename B, typename C>
    constexpr vec3_t(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    constexpr vec3_t(const vec4_t<T>& v);
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec3_t& operator+=(const vec3_t& o) { x += o.x; y += o.y; z += o.z; return *this; }
    vec3_t& operator-=(const vec3_t& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    vec3_t& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
};

template <typename T> struct vec4_t {
    T x {}, y {}, z {}, w {};
    constexpr vec4_t() = default;
    constexpr explicit vec4_t(T s) : x(s), y(s), z(s), w(s) {}
    constexpr vec4_t(T x_, T y_, T z_, T w_) : x(x_), y(y_), z(z_), w(w_) {}
    constexpr vec4_t(const vec3_t<T>& v, T w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};
template <typename T> constexpr vec3_t<T>::vec3_t(const vec4_t<T>& v) : x(v.x), y(v.y), z(v.z) {}

using vec2 = vec2_t<float>;
using vec3 = vec3_t<float>;
using vec4 = vec4_t<float>;
using ivec2 = vec2_t<int>;
using uvec3 = vec3_t<unsigned int>;
using u8vec4 = vec4_t<std::uint8_t>;

inline vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator-(vec2 a, vec2 b) { return { a.x - b.x, a.y - b.y }; }
inline vec2 operator*(vec2 a, float s) { return { a.x * s, a.y * s }; }
inline vec2 operator*(float s, vec2 a) { return { s * a.x, s * a.y }; }

inline vec3 operator+(const vec3& a, const vec3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(const vec3& a, const vec3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator-(const vec3& a) { return { -a.x, -a.y, -a.z }; }
inline vec3 operator*(const vec3& a, const vec3& b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator*(const vec3& a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*(float s, const vec3& a) { return { s * a.x, s * a.y, s * a.z }; }
inline vec3 operator/(const vec3& a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

inline float radians(float degrees) { return degrees * 0.01745329251994329576923690768489f; }
inline vec3 radians(const vec3& d) { return { radians(d.x), radians(d.y), radians(d.z) }; }
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec3 min(const vec3& a, const vec3& b) { return { min(a.x, b.x), min(a.y, b.y), min(a.z, b.z) }; }
inline vec3 max(const vec3& a, const vec3& b) { return { max(a.x, b.x), max(a.y, b.y), max(a.z, b.z) }; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return { clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi) }; }

inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b)
{
    return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y };
}
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }

struct quat {
    float x = 0, y = 0, z = 0, w = 1;
    quat() = default;
    // rotation from XYZ euler angles (radians)
    explicit quat(const vec3& euler)
    {
        const float cx = std::cos(euler.x * 0.5f), cy = std::cos(euler.y * 0.5f), cz = std::cos(euler.z * 0.5f);
        const float sx = std::sin(euler.x * 0.5f), sy = std::sin(euler.y * 0.5f), sz = std::sin(euler.z * 0.5f);
        w = cx * cy * cz + sx * sy * sz;
        x = sx * cy * cz - cx * sy * sz;
        y = cx * sy * cz + sx * cy * sz;
        z = cx * cy * sz - sx * sy * cz;
    }
};
inline vec3 operator*(const quat& q, const vec3& v)
{
    const vec3 u(q.x, q.y, q.z);
    const vec3 uv = cross(u, v);
    const vec3 uuv = cross(u, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}

} // namespace glm
