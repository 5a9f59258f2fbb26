// Collapse of the binary BVH into a 4-wide one (device, both builders feed it).
//
// Input: the sibling-pair layout the builders emit (rt_types.h): node i = {lo.xyz, entry}{hi.xyz, count} at float4 2i, siblings
// adjacent; entry >= 0 is the index of the first of the two child nodes, < 0 a leaf.  Output: one 128-byte node per group of up to four boxes,
// laid out per axis so that one node = one cache line and the slab test of the four boxes reads whole float4s:
//   [0] lo.x[0..3]  [1] lo.y  [2] lo.z  [3] hi.x  [4] hi.y  [5] hi.z  [6] entry[0..3] (bits)  [7] unused
// A 4-wide node takes the two boxes of a pair and replaces each inner one by the two boxes of ITS pair (every other level
// of the binary tree disappears); missing children get an inverted box that no ray can hit.  The walk is breadth first,
// one kernel launch per level: a frontier of (binary pair, destination node) items, children allocated with one atomic.
// The boxes themselves are copied, never recomputed, so the conservative padding of the builders carries over.
#include "rt_kernels.h"
#include <cfloat>
#include <climits>
#include <utility>

namespace rtb {
namespace {

struct WideItem {
    int pair; // first node of the binary sibling pair to expand
    int dst;  // 4-wide node to fill
};

__global__ void k_collapse_level(const float4* __restrict__ nodes2, const WideItem* __restrict__ frontier, int n_items, float4* nodes4, int* n_nodes4,
    WideItem* next, int* n_next)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items)
        return;
    const WideItem it = frontier[i];
    float4 lo[4], hi[4];
    int n = 0;
    for (int side = 0; side < 2; side++) {
        const float4 l = nodes2[2 * ((size_t)it.pair + side)], h = nodes2[2 * ((size_t)it.pair + side) + 1];
        const int e = __float_as_int(l.w);
        if (e >= 0) { // inner: its two children move up
            for (int k = 0; k < 2; k++) {
                lo[n] = nodes2[2 * ((size_t)e + k)];
                hi[n] = nodes2[2 * ((size_t)e + k) + 1];
                n++;
            }
        } else {
            lo[n] = l;
            hi[n] = h;
            n++;
        }
    }
    int entry[4];
    for (int k = 0; k < 4; k++) {
        if (k >= n) { // no child: a box turned inside out
            lo[k] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.0f);
            hi[k] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.0f);
            entry[k] = INT_MIN;
            continue;
        }
        const int e = __float_as_int(lo[k].w);
        if (e >= 0) {
            const int idx = atomicAdd(n_nodes4, 1);
            next[atomicAdd(n_next, 1)] = WideItem { e, idx };
            entry[k] = idx;
        } else {
            entry[k] = e;
        }
    }
    float4* o = nodes4 + 8 * (size_t)it.dst;
    o[0] = make_float4(lo[0].x, lo[1].x, lo[2].x, lo[3].x);
    o[1] = make_float4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
    o[2] = make_float4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
    o[3] = make_float4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
    o[4] = make_float4(hi[0].y, hi[1].y, hi[2].y, hi[3].y);
    o[5] = make_float4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
    o[6] = make_float4(__int_as_float(entry[0]), __int_as_float(entry[1]), __int_as_float(entry[2]), __int_as_float(entry[3]));
    o[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

} // namespace

int collapse_bvh_wide_device(cudaStream_t st, const float4* d_nodes2, int n_nodes2, int root_entry2, WideBvh* out, const char** err)
{
    out->nodes = nullptr;
    out->n_nodes = 0;
    out->depth = 0;
    out->root_entry = root_entry2;
    if (root_entry2 < 0) // the whole scene is one leaf
        return 0;
    const int n_pairs = n_nodes2 / 2;
    float4* nodes4 = nullptr;
    WideItem *fa = nullptr, *fb = nullptr;
    int* counters = nullptr; // [0] nodes allocated, [1] next frontier size
    auto fail = [&](const char* what) {
        if (err)
            *err = what;
        cudaFree(nodes4);
        cudaFree(fa);
        cudaFree(fb);
        cudaFree(counters);
        return 1;
    };
    // every 4-wide node consumes at least one binary pair
    if (cudaMalloc(&nodes4, (size_t)n_pairs * 8 * sizeof(float4)) != cudaSuccess || cudaMalloc(&fa, (size_t)n_pairs * sizeof(WideItem)) != cudaSuccess
        || cudaMalloc(&fb, (size_t)n_pairs * sizeof(WideItem)) != cudaSuccess || cudaMalloc(&counters, 2 * sizeof(int)) != cudaSuccess)
        return fail("out of device memory (4-wide BVH)");
    const WideItem root { root_entry2, 0 };
    int h_counters[2] = { 1, 0 };
    if (cudaMemcpyAsync(fa, &root, sizeof(root), cudaMemcpyHostToDevice, st) != cudaSuccess
        || cudaMemcpyAsync(counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail("copy failed (4-wide BVH)");
    int n_items = 1, depth = 0;
    while (n_items > 0) {
        depth++;
        k_collapse_level<<<(n_items + 127) / 128, 128, 0, st>>>(d_nodes2, fa, n_items, nodes4, counters, fb, counters + 1);
        if (cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            return fail("kernel failure (4-wide BVH)");
        n_items = h_counters[1];
        if (n_items > n_pairs)
            return fail("inconsistent binary BVH (4-wide BVH)");
        const int zero = 0;
        if (cudaMemcpyAsync(counters + 1, &zero, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return fail("copy failed (4-wide BVH)");
        std::swap(fa, fb);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess)
        return fail("kernel failure (4-wide BVH)");
    cudaFree(fa);
    cudaFree(fb);
    cudaFree(counters);
    out->nodes = nodes4;
    out->n_nodes = h_counters[0];
    out->root_entry = 0;
    out->depth = depth;
    return 0;
}

} // namespace rtb
