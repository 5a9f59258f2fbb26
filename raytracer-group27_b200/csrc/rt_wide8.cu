// An 8-wide tree walked by EIGHT LANES PER RAY, for small wavefronts (added in round 2; measurements in profiles/README.md).
//
// The per-level kernels of a small wavefront wait for their longest ray: a chain of ~200 dependent node-pair fetches at L2
// latency plus up to ~90 triangle tests one after the other (profiles/README.md, round 2).  Eight lanes per ray shorten that
// chain instead of hiding it: one step fetches and slab-tests the eight children of a node at once (2.5 times fewer steps than
// pairs), a leaf's triangles (up to eight) are tested side by side, and the four rays of a warp diverge only as groups.  The
// price is throughput — eight lanes do what one lane did — so this form is for the queues that cannot fill the GPU anyway.
//
// The 8-wide tree is collapsed, on the device (collapse_bvh_wide8_device), from the binary one the builders emit (boxes copied, never
// recomputed: the conservative padding carries over): a node takes the two children of a binary node and keeps replacing its largest inner child by that child's two
// until it has eight; binary subtrees of at most eight triangles become one leaf (their triangles are contiguous in leaf order).
// Node layout: child k at float4 [2k] = {lo.xyz, entry}, [2k + 1] = {hi.xyz, -}; 256 bytes per node; entry >= 0: node index,
// < 0: leaf ~((first << 3) | (count - 1)) like the binary tree's; a missing child is a box turned inside out.
#include "rt_kernels.h"
#include "rt_wide8.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <utility>

namespace rtb {

namespace {

// rt_intersect through the 8-wide tree: group g takes rays g, g + groups, ...
__global__ void __launch_bounds__(kWideBlock) k_intersect_wide(SceneDev s, const float4* __restrict__ wide, int root_entry, const float* __restrict__ rays,
    long long n, int* tri_id, float* t_out)
{
    __shared__ int s_stack[kWideBlock / kGroup][kWideStack];
    int* stack = s_stack[threadIdx.x / kGroup];
    const long long groups = (long long)gridDim.x * (kWideBlock / kGroup);
    for (long long i = (long long)blockIdx.x * (kWideBlock / kGroup) + threadIdx.x / kGroup; i < n; i += groups) {
        const f3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        const f3 d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        HitRec best = fresh_query();
        trace_wide<false>(s, wide, root_entry, o, d, best, stack);
        if ((threadIdx.x & (kGroup - 1)) == 0) {
            tri_id[i] = global_id(s, best);
            t_out[i] = best.t;
        }
    }
}

} // namespace

// ---- the same collapse on the device ----
namespace {

// Triangles below every binary node: a subtree's triangles are contiguous in leaf order, from the first triangle of its leftmost
// leaf to the last of its rightmost.
__global__ void k_wide8_spans(const float4* __restrict__ nodes2, int n_nodes2, int* count, int* first, int* n_big)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes2)
        return;
    int l = v, r = v, e;
    while ((e = __float_as_int(nodes2[2 * (size_t)l].w)) >= 0)
        l = e;
    const int lenc = ~e;
    while ((e = __float_as_int(nodes2[2 * (size_t)r].w)) >= 0)
        r = e + 1;
    const int renc = ~e;
    const int f = lenc >> 3, c = (renc >> 3) + (renc & 7) + 1 - f;
    first[v] = f;
    count[v] = c;
    if (c > 8 && v != 1) // (node 1 is a copy of the root)
        atomicAdd(n_big, 1);
}

struct Wide8Item {
    int bin, dst;
};

__device__ __forceinline__ float box_area(const float4* __restrict__ nodes2, int v)
{
    const float4 lo = nodes2[2 * (size_t)v], hi = nodes2[2 * (size_t)v + 1];
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// One thread fills one 8-wide node: the two children of its binary node, the largest subtree of more than 8 triangles opened again
// and again until there are eight children; subtrees of at most 8 triangles become leaves, the others nodes of the next level.
__global__ void k_wide8_level(const float4* __restrict__ nodes2, const int* __restrict__ count, const int* __restrict__ first, const Wide8Item* __restrict__ frontier,
    int n_items, float4* wide, int* n_wide, Wide8Item* next, int* n_next)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items)
        return;
    const Wide8Item it = frontier[i];
    int kids[8], n = 2;
    kids[0] = __float_as_int(nodes2[2 * (size_t)it.bin].w);
    kids[1] = kids[0] + 1;
    while (n < 8) {
        int best = -1;
        float ba = -1.0f;
        for (int k = 0; k < n; k++)
            if (count[kids[k]] > 8) {
                const float a = box_area(nodes2, kids[k]);
                if (a > ba) {
                    ba = a;
                    best = k;
                }
            }
        if (best < 0)
            break;
        const int e = __float_as_int(nodes2[2 * (size_t)kids[best]].w);
        kids[best] = e;
        kids[n++] = e + 1;
    }
    for (int k = 0; k < 8; k++) {
        float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.0f), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.0f);
        int entry = INT_MIN;
        if (k < n) {
            const int v = kids[k];
            lo = nodes2[2 * (size_t)v];
            hi = nodes2[2 * (size_t)v + 1];
            if (count[v] <= 8) {
                entry = ~((first[v] << 3) | (count[v] - 1));
            } else {
                entry = atomicAdd(n_wide, 1);
                next[atomicAdd(n_next, 1)] = Wide8Item { v, entry };
            }
        }
        lo.w = __int_as_float(entry);
        wide[16 * (size_t)it.dst + 2 * k] = lo;
        wide[16 * (size_t)it.dst + 2 * k + 1] = hi;
    }
}

} // namespace

// nodes: allocated by the callee (cudaMalloc); *n_nodes8 = 0 and a leaf root_entry8 when the whole scene is one leaf.
int collapse_bvh_wide8_device(cudaStream_t st, const float4* d_nodes2, int n_nodes2, float4** nodes8, int* n_nodes8, int* root_entry8, int* depth8, const char** err)
{
    *nodes8 = nullptr;
    *n_nodes8 = 0;
    *depth8 = 1;
    int *count = nullptr, *first = nullptr, *ctr = nullptr; // ctr: [0] subtrees of more than 8 triangles, [1] wide nodes, [2] next frontier
    Wide8Item *fa = nullptr, *fb = nullptr;
    float4* wide = nullptr;
    auto fail = [&](const char* what) {
        if (err)
            *err = what;
        cudaFree(count);
        cudaFree(first);
        cudaFree(ctr);
        cudaFree(fa);
        cudaFree(fb);
        cudaFree(wide);
        return 1;
    };
    if (cudaMalloc(&count, (size_t)n_nodes2 * sizeof(int)) != cudaSuccess || cudaMalloc(&first, (size_t)n_nodes2 * sizeof(int)) != cudaSuccess
        || cudaMalloc(&ctr, 3 * sizeof(int)) != cudaSuccess)
        return fail("out of device memory (8-wide BVH)");
    cudaMemsetAsync(ctr, 0, 3 * sizeof(int), st);
    k_wide8_spans<<<(n_nodes2 + 255) / 256, 256, 0, st>>>(d_nodes2, n_nodes2, count, first, ctr);
    int h[3] = { 0, 0, 0 }, root_span[2] = { 0, 0 };
    cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&root_span[0], count, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&root_span[1], first, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
        return fail("kernel failure (8-wide BVH, spans)");
    if (root_span[0] <= 8) { // the whole scene is one leaf
        *root_entry8 = ~((root_span[1] << 3) | (root_span[0] - 1));
        cudaFree(count);
        cudaFree(first);
        cudaFree(ctr);
        return 0;
    }
    const int n_big = h[0]; // every 8-wide node is rooted at a different one of these binary nodes
    if (cudaMalloc(&wide, (size_t)n_big * 16 * sizeof(float4)) != cudaSuccess || cudaMalloc(&fa, (size_t)n_big * sizeof(Wide8Item)) != cudaSuccess
        || cudaMalloc(&fb, (size_t)n_big * sizeof(Wide8Item)) != cudaSuccess)
        return fail("out of device memory (8-wide BVH)");
    const Wide8Item root { 0, 0 };
    const int init[3] = { n_big, 1, 0 };
    cudaMemcpyAsync(fa, &root, sizeof(root), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(ctr, init, sizeof(init), cudaMemcpyHostToDevice, st);
    int n_items = 1, depth = 1;
    while (n_items > 0) {
        depth++;
        k_wide8_level<<<(n_items + 127) / 128, 128, 0, st>>>(d_nodes2, count, first, fa, n_items, wide, ctr + 1, fb, ctr + 2);
        cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess)
            return fail("kernel failure (8-wide BVH, collapse)");
        n_items = h[2];
        if (h[1] > n_big || n_items > n_big)
            return fail("inconsistent binary BVH (8-wide BVH)");
        cudaMemsetAsync(ctr + 2, 0, sizeof(int), st);
        std::swap(fa, fb);
    }
    cudaFree(count);
    cudaFree(first);
    cudaFree(ctr);
    cudaFree(fa);
    cudaFree(fb);
    *nodes8 = wide;
    *n_nodes8 = h[1];
    *root_entry8 = 0;
    *depth8 = depth;
    return 0;
}

void launch_intersect_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int root_entry, const float* rays, long long n, int* tri_id,
    float* t_out)
{
    if (n <= 0)
        return;
    const long long want = (n + (kWideBlock / kGroup) - 1) / (kWideBlock / kGroup);
    const int grid = (int)std::min<long long>(want, (long long)sm_count * 16);
    k_intersect_wide<<<grid, kWideBlock, 0, st>>>(s, wide, root_entry, rays, n, tri_id, t_out);
}

#if RT_CHECKED
namespace {
__global__ void k_provoke_violation(int site) { (void)RT_GUARD(-1, 1, site); }
} // namespace
#endif

// One deliberate violation at site kChkWideStack from this translation unit (rt_violations_selftest): proves that what the kernels count is
// what add_violations_wide8 reads.
void provoke_violation_wide8(cudaStream_t st)
{
#if RT_CHECKED
    k_provoke_violation<<<1, 1, 0, st>>>(kChkWideStack);
#else
    (void)st;
#endif
}

void add_violations_wide8(unsigned int* out)
{
#if RT_CHECKED
    unsigned int h[kChkSites] = {};
    if (cudaMemcpyFromSymbol(h, g_rt_violations, sizeof(h)) == cudaSuccess)
        for (int k = 0; k < kChkSites; k++)
            out[k] += h[k];
#else
    (void)out;
#endif
}

} // namespace rtb
