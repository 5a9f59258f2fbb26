// Visiting order of the reference's own BVH, computed on the device.
//
// The reference breaks exact ties in t by visiting order: intersectRayWithPlane accepts only t < ray.t
// (src/ray_tracing.cpp:74), so of several objects at the same t the one tested first wins.  With useBVH = false the
// order is the mesh order (bounding_volume_hierarchy.cpp:51-71), i.e. the global object id; through the BVH it is the
// depth-first order of intersectBVH (bounding_volume_hierarchy.cpp:414-447: children and leaf objects in stored order,
// independent of the ray).  That order follows from constructBVH / sortObjects (bounding_volume_hierarchy.cpp:107-217,
// 299-320): objects = triangles in global order, then spheres; a node of level L < 4 with more than one object sorts
// its objects by std::pair(centroid[(L + 1) % 3], current position) and gives the first (n + 1) / 2 to its left child.
// Four rounds of a stable sort inside the segments of the previous round reproduce it exactly; the final position of
// an object is its visiting rank, which the traversal kernels use as the tie key (rt_trace.cuh).
#include "rt_kernels.h"
#include "rt_math.cuh"
#include <cub/cub.cuh>

namespace rtb {
namespace {

constexpr int kRefMaxLevel = 4; // bounding_volume_hierarchy.h:67

// Index of the level-`depth` node of the reference tree that holds position i of n objects (nodes numbered left to right).
__device__ __forceinline__ unsigned segment_of(long long i, long long n, int depth)
{
    long long lo = 0;
    unsigned seg = 0;
    for (int k = 0; k < depth; k++) {
        const long long half = (n + 1) / 2; // bounding_volume_hierarchy.cpp:176
        if (i - lo < half) {
            n = half;
            seg = seg * 2;
        } else {
            lo += half;
            n -= half;
            seg = seg * 2 + 1;
        }
    }
    return seg;
}

__global__ void k_order_init(int* perm, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        perm[i] = (int)i;
}

// getSortingAttributeTriangle / getSortingAttributeSphere (bounding_volume_hierarchy.cpp:322-366)
__global__ void k_order_keys(const int* __restrict__ perm, long long n, long long n_tris, const float* __restrict__ pos,
    const float4* __restrict__ spheres, int level, unsigned long long* keys)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int obj = perm[i];
    const int axis = level % 3;
    float a;
    if (obj < n_tris) {
        const float* p = pos + 9 * (size_t)obj + axis;
        a = xdiv(xadd(xadd(p[0], p[3]), p[6]), 3.0f);
    } else {
        const float4 c = spheres[3 * (obj - n_tris)];
        a = axis == 0 ? c.x : axis == 1 ? c.y : c.z;
    }
    a = xadd(a, 0.0f); // -0 and +0 compare equal in the reference's std::pair order
    const unsigned bits = (unsigned)__float_as_int(a);
    const unsigned ordered = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
    keys[i] = ((unsigned long long)segment_of(i, n, level - 1) << 32) | ordered;
}

__global__ void k_order_rank(const int* __restrict__ perm, long long n, int* rank)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        rank[perm[i]] = (int)i;
}

__global__ void k_apply_tie_keys(float4* v0, const float4* __restrict__ v2, const int* __restrict__ rank, long long n_slots, long long n_ranked)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots)
        return;
    const int g = __float_as_int(v2[kTriStride * i].w);
    v0[kTriStride * i].w = __int_as_float(g < n_ranked ? rank[g] : g);
}

} // namespace

int reference_visit_rank(cudaStream_t st, const float* d_pos, long long n_tris, const float4* d_spheres, int n_spheres, int* d_rank, const char** err)
{
    const long long n = n_tris + n_spheres;
    if (n <= 0)
        return 0;
    unsigned long long *keys = nullptr, *keys_out = nullptr;
    int *perm = nullptr, *perm_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    char* arena_to_free = nullptr;
    auto fail = [&](const char* what) {
        if (err)
            *err = what;
        cudaFree(arena_to_free);
        return 1;
    };
    if (n >= (1ll << 31))
        return fail("too many objects");
    // one allocation for all scratch (a cudaMalloc / cudaFree pair per array costs milliseconds)
    if (cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, perm, perm_out, (int)n, 0, 32 + kRefMaxLevel, st) != cudaSuccess)
        return fail("radix sort set-up failed (reference visiting order)");
    auto align = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t kb = align((size_t)n * sizeof(*keys)), pb = align((size_t)n * sizeof(int));
    char* arena = nullptr;
    if (cudaMalloc(&arena, 2 * kb + 2 * pb + align(tmp_bytes)) != cudaSuccess)
        return fail("out of device memory (reference visiting order)");
    keys = reinterpret_cast<unsigned long long*>(arena);
    keys_out = reinterpret_cast<unsigned long long*>(arena + kb);
    perm = reinterpret_cast<int*>(arena + 2 * kb);
    perm_out = reinterpret_cast<int*>(arena + 2 * kb + pb);
    tmp = arena + 2 * kb + 2 * pb;
    arena_to_free = arena;
    const unsigned grid = (unsigned)((n + 255) / 256);
    k_order_init<<<grid, 256, 0, st>>>(perm, n);
    for (int level = 1; level <= kRefMaxLevel; level++) {
        k_order_keys<<<grid, 256, 0, st>>>(perm, n, n_tris, d_pos, d_spheres, level, keys);
        // stable LSD radix sort: equal (segment, centroid) keep their current order, like the pair's second member
        if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, perm, perm_out, (int)n, 0, 32 + kRefMaxLevel, st) != cudaSuccess)
            return fail("radix sort failed (reference visiting order)");
        std::swap(perm, perm_out);
    }
    k_order_rank<<<grid, 256, 0, st>>>(perm, n, d_rank);
    if (cudaStreamSynchronize(st) != cudaSuccess)
        return fail("kernel failure (reference visiting order)");
    cudaFree(arena);
    return 0;
}

void launch_apply_tie_keys(cudaStream_t st, float4* v0, const float4* v2, const int* rank, long long n_slots, long long n_ranked)
{
    if (n_slots > 0)
        k_apply_tie_keys<<<(unsigned)((n_slots + 255) / 256), 256, 0, st>>>(v0, v2, rank, n_slots, n_ranked);
}

} // namespace rtb
