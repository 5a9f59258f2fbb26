// Exact float3 arithmetic for the parity-critical geometry code.
//
// The reference evaluates everything through glm 0.9.9.8's scalar path on x86-64 without FMA (SURVEY §2 #17,
// Appendix A): every +,-,*,/ and sqrt is a single IEEE-754 binary32 operation in a fixed order.  To reproduce
// closest-hit ids and ray origins bit for bit the device code must not let the compiler contract a*b+c into an
// FMA, so the helpers below are spelled with the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn are
// never fused, and __fdiv_rn/__fsqrt_rn are IEEE regardless of -prec-div/-prec-sqrt/-use_fast_math).
// Conservative tests (BVH slabs) and colour arithmetic use ordinary operators and may fuse.
#pragma once
#include <cuda_runtime.h>

namespace rtb {

struct f3 {
    float x, y, z;
};

__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3 { x, y, z }; }
__device__ __forceinline__ f3 mk3(const float4& v) { return f3 { v.x, v.y, v.z }; }

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ f3 xadd(const f3& a, const f3& b) { return { xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z) }; }
__device__ __forceinline__ f3 xsub(const f3& a, const f3& b) { return { xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z) }; }
__device__ __forceinline__ f3 xmul(const f3& a, const f3& b) { return { xmul(a.x, b.x), xmul(a.y, b.y), xmul(a.z, b.z) }; }
__device__ __forceinline__ f3 xmul(const f3& a, float s) { return { xmul(a.x, s), xmul(a.y, s), xmul(a.z, s) }; }
__device__ __forceinline__ f3 xdiv(const f3& a, float s) { return { xdiv(a.x, s), xdiv(a.y, s), xdiv(a.z, s) }; }
__device__ __forceinline__ f3 xneg(const f3& a) { return { -a.x, -a.y, -a.z }; }

// glm::dot(vec3): products first, then (x + y) + z
__device__ __forceinline__ float xdot(const f3& a, const f3& b)
{
    return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z));
}
// glm::cross(x, y)
__device__ __forceinline__ f3 xcross(const f3& x, const f3& y)
{
    return { xsub(xmul(x.y, y.z), xmul(y.y, x.z)), xsub(xmul(x.z, y.x), xmul(y.z, x.x)), xsub(xmul(x.x, y.y), xmul(y.x, x.y)) };
}
// glm::length / glm::normalize (v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt)
__device__ __forceinline__ float xlength(const f3& v) { return xsqrt(xdot(v, v)); }
__device__ __forceinline__ f3 xnormalize(const f3& v) { return xmul(v, xdiv(1.0f, xsqrt(xdot(v, v)))); }
// glm::reflect(I, N) = I - N * dot(N, I) * 2
__device__ __forceinline__ f3 xreflect(const f3& I, const f3& N) { return xsub(I, xmul(xmul(N, xdot(N, I)), 2.0f)); }
// quat * vec3 (glm/detail/type_quat.inl): v + ((uv * w) + uuv) * 2
__device__ __forceinline__ f3 xquat_rotate(const f3& qv, float qw, const f3& v)
{
    const f3 uv = xcross(qv, v);
    const f3 uuv = xcross(qv, uv);
    return xadd(v, xmul(xadd(xmul(uv, qw), uuv), 2.0f));
}

} // namespace rtb
