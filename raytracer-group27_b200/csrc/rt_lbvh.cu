// K6: device LBVH build — Morton codes -> radix sort -> Karras hierarchy -> bottom-up refit -> collapse of small
// subtrees into leaves -> emission in the flattened sibling-pair layout of rt_types.h.
//
// Replaces BoundingVolumeHierarchy::constructBVH (src/bounding_volume_hierarchy.cpp:108-217: BFS median split with one
// std::sort per node, at most 5 levels).  The whole build runs on the GPU; the only library call is CUB's radix sort.
// T. Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", HPG 2012.
#include "rt_kernels.h"

#include <cfloat>
#include <cub/cub.cuh>

namespace rtb {

namespace {

constexpr int kLeafCollapse = 4; // Karras subtrees with at most this many triangles become one leaf

struct Bounds {
    int lo[3], hi[3]; // order-preserving int encoding of floats
};

__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_init_bounds(Bounds* b)
{
    for (int a = 0; a < 3; a++) {
        b->lo[a] = f2ord(FLT_MAX);
        b->hi[a] = f2ord(-FLT_MAX);
    }
}

// per-triangle boxes (stored for the refit) and the bounds of their centres
__global__ void k_tri_bounds(const float* __restrict__ pos, int n, float4* tlo, float4* thi, Bounds* b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = { 0, 0, 0 };
    const bool valid = i < n;
    if (valid) {
        const float* p = pos + 9 * (size_t)i;
        float lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
            hi[a] = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
            c[a] = 0.5f * (lo[a] + hi[a]);
        }
        tlo[i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        thi[i] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
    // warp-level min/max before the atomics
    for (int a = 0; a < 3; a++) {
        float mn = valid ? c[a] : FLT_MAX, mx = valid ? c[a] : -FLT_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&b->lo[a], f2ord(mn));
            atomicMax(&b->hi[a], f2ord(mx));
        }
    }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v)
{
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float4* __restrict__ tlo, const float4* __restrict__ thi, int n, const Bounds* b, unsigned long long* keys, int* vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float4 lo = tlo[i], hi = thi[i];
    const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
    unsigned long long q[3];
    for (int a = 0; a < 3; a++) {
        const float mn = ord2f(b->lo[a]), mx = ord2f(b->hi[a]);
        const float ext = mx - mn;
        float u = ext > 0.0f ? (c[a] - mn) / ext : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        q[a] = (unsigned long long)fminf(u * 2097152.0f, 2097151.0f);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// common-prefix length of sorted keys i and j, ties broken by the index (Karras sec. 4)
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n)
        return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b)
        return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras internal node i: range, split, children.  Children >= n-1 denote leaves (leaf k stored as n-1+k).
__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int* left, int* right, int* parent, int* first, int* last)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1)
        return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin)
        lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin)
            l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode)
            s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    first[i] = lo;
    last[i] = hi;
    if (i == 0)
        parent[0] = -1;
}

// Bottom-up refit: the second thread to reach a node merges its children's boxes and heights.
__global__ void k_refit(const int* __restrict__ vals, const float4* __restrict__ tlo, const float4* __restrict__ thi, int n, const int* __restrict__ left,
    const int* __restrict__ right, const int* __restrict__ parent, float4* blo, float4* bhi, int* height, unsigned* visit)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n)
        return;
    const int t = vals[k];
    const int leaf = n - 1 + k;
    blo[leaf] = tlo[t];
    bhi[leaf] = thi[t];
    height[leaf] = 0;
    __threadfence();
    int node = parent[leaf];
    while (node >= 0) {
        if (atomicAdd(&visit[node], 1u) == 0)
            return; // the sibling subtree is not finished yet
        __threadfence();
        const int l = left[node], r = right[node];
        const float4 al = blo[l], ah = bhi[l], cl = blo[r], ch = bhi[r];
        blo[node] = make_float4(fminf(al.x, cl.x), fminf(al.y, cl.y), fminf(al.z, cl.z), 0.0f);
        bhi[node] = make_float4(fmaxf(ah.x, ch.x), fmaxf(ah.y, ch.y), fmaxf(ah.z, ch.z), 0.0f);
        height[node] = 1 + max(height[l], height[r]);
        __threadfence();
        node = parent[node];
    }
}

__global__ void k_live(const int* __restrict__ first, const int* __restrict__ last, int n, int* live)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1)
        live[i] = (last[i] - first[i] + 1 > kLeafCollapse) ? 1 : 0;
}

__device__ __forceinline__ void emit_node(float4* nodes, int idx, float4 lo, float4 hi, float pad, int left_or_first, int count)
{
    const int entry = count ? ~((left_or_first << 3) | (count - 1)) : left_or_first; // see rt_types.h
    nodes[2 * (size_t)idx] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, __int_as_float(entry));
    nodes[2 * (size_t)idx + 1] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, __int_as_float(count));
}

// Every live internal node writes its two children into the pair slot (1 + rank among live nodes).
__global__ void k_emit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ first, const int* __restrict__ last,
    const int* __restrict__ live, const int* __restrict__ live_rank, const float4* __restrict__ blo, const float4* __restrict__ bhi, float pad, float4* nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !live[i])
        return;
    const int pair = 2 * (1 + live_rank[i]);
    const int ch[2] = { left[i], right[i] };
    for (int k = 0; k < 2; k++) {
        const int c = ch[k];
        if (c >= n - 1) // Karras leaf: one triangle
            emit_node(nodes, pair + k, blo[c], bhi[c], pad, c - (n - 1), 1);
        else if (live[c])
            emit_node(nodes, pair + k, blo[c], bhi[c], pad, 2 * (1 + live_rank[c]), 0);
        else // collapsed subtree: contiguous run of the sorted order
            emit_node(nodes, pair + k, blo[c], bhi[c], pad, first[c], last[c] - first[c] + 1);
    }
    if (i == 0) { // root and its twin
        emit_node(nodes, 0, blo[0], bhi[0], pad, pair, 0);
        emit_node(nodes, 1, blo[0], bhi[0], pad, pair, 0);
    }
}

// scenes with at most kLeafCollapse triangles: a single leaf
__global__ void k_emit_single(int n, const float4* __restrict__ tlo, const float4* __restrict__ thi, float pad, float4* nodes, int* perm)
{
    float4 lo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0), hi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0);
    for (int i = 0; i < n; i++) {
        lo = make_float4(fminf(lo.x, tlo[i].x), fminf(lo.y, tlo[i].y), fminf(lo.z, tlo[i].z), 0);
        hi = make_float4(fmaxf(hi.x, thi[i].x), fmaxf(hi.y, thi[i].y), fmaxf(hi.z, thi[i].z), 0);
        perm[i] = i;
    }
    emit_node(nodes, 0, lo, hi, pad, 0, n);
    emit_node(nodes, 1, lo, hi, pad, 0, n);
}

struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch()
    {
        for (void* p : ptrs)
            cudaFree(p);
    }
    template <typename T> T* alloc(size_t n, cudaError_t& e)
    {
        T* p = nullptr;
        if (e == cudaSuccess) {
            e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
            if (e == cudaSuccess)
                ptrs.push_back(p);
        }
        return p;
    }
};

} // namespace

int build_bvh_lbvh_device(cudaStream_t st, const float* d_pos, long long n_tris, float pad, DeviceBvh* out, const char** err)
{
    const int n = (int)n_tris;
    const int blk = 256, grid = (n + blk - 1) / blk;
    cudaError_t e = cudaSuccess;
    Scratch sc;
    float4* tlo = sc.alloc<float4>(n, e);
    float4* thi = sc.alloc<float4>(n, e);
    Bounds* bounds = sc.alloc<Bounds>(1, e);
    int* perm = nullptr;
    if (e == cudaSuccess)
        e = cudaMalloc(&perm, (size_t)n * sizeof(int));
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return 1;
    }
    k_init_bounds<<<1, 1, 0, st>>>(bounds);
    k_tri_bounds<<<grid, blk, 0, st>>>(d_pos, n, tlo, thi, bounds);

    if (n <= kLeafCollapse) {
        float4* nodes = nullptr;
        e = cudaMalloc(&nodes, 4 * sizeof(float4));
        if (e == cudaSuccess) {
            k_emit_single<<<1, 1, 0, st>>>(n, tlo, thi, pad, nodes, perm);
            e = cudaStreamSynchronize(st);
        }
        if (e != cudaSuccess) {
            cudaFree(perm);
            cudaFree(nodes);
            *err = cudaGetErrorString(e);
            return 1;
        }
        out->nodes = nodes;
        out->perm = perm;
        out->n_nodes = 2;
        out->root_entry = 0;
        out->depth = 1;
        return 0;
    }

    unsigned long long* keys_in = sc.alloc<unsigned long long>(n, e);
    unsigned long long* keys = sc.alloc<unsigned long long>(n, e);
    int* vals_in = sc.alloc<int>(n, e);
    int* left = sc.alloc<int>(n, e);
    int* right = sc.alloc<int>(n, e);
    int* parent = sc.alloc<int>(2 * (size_t)n, e);
    int* first = sc.alloc<int>(n, e);
    int* last = sc.alloc<int>(n, e);
    int* height = sc.alloc<int>(2 * (size_t)n, e);
    unsigned* visit = sc.alloc<unsigned>(n, e);
    float4* blo = sc.alloc<float4>(2 * (size_t)n, e);
    float4* bhi = sc.alloc<float4>(2 * (size_t)n, e);
    int* live = sc.alloc<int>(n, e);
    int* live_rank = sc.alloc<int>(n, e);
    size_t tmp_sort = 0, tmp_scan = 0;
    if (e == cudaSuccess)
        e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys_in, keys, vals_in, perm, n, 0, 63, st);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, live, live_rank, n - 1, st);
    void* tmp = sc.alloc<unsigned char>(std::max(tmp_sort, tmp_scan), e);
    if (e != cudaSuccess) {
        cudaFree(perm);
        *err = cudaGetErrorString(e);
        return 1;
    }
    k_morton<<<grid, blk, 0, st>>>(tlo, thi, n, bounds, keys_in, vals_in);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_sort, keys_in, keys, vals_in, perm, n, 0, 63, st);
    k_hierarchy<<<grid, blk, 0, st>>>(keys, n, left, right, parent, first, last);
    cudaMemsetAsync(visit, 0, (size_t)n * sizeof(unsigned), st);
    k_refit<<<grid, blk, 0, st>>>(perm, tlo, thi, n, left, right, parent, blo, bhi, height, visit);
    k_live<<<grid, blk, 0, st>>>(first, last, n, live);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_scan, live, live_rank, n - 1, st);
    int n_live_last[2] = { 0, 0 }, tree_height = 0;
    cudaMemcpyAsync(&n_live_last[0], live_rank + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&n_live_last[1], live + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&tree_height, height, sizeof(int), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(perm);
        *err = cudaGetErrorString(e);
        return 1;
    }
    const int n_live = n_live_last[0] + n_live_last[1];
    const int n_nodes = 2 * (1 + n_live);
    float4* nodes = nullptr;
    e = cudaMalloc(&nodes, 2 * (size_t)n_nodes * sizeof(float4));
    if (e == cudaSuccess) {
        k_emit<<<grid, blk, 0, st>>>(n, left, right, first, last, live, live_rank, blo, bhi, pad, nodes);
        e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(perm);
        cudaFree(nodes);
        *err = cudaGetErrorString(e);
        return 1;
    }
    out->nodes = nodes;
    out->perm = perm;
    out->n_nodes = n_nodes;
    out->root_entry = 2; // the root is live (n > kLeafCollapse): its children are the first emitted pair
    out->depth = tree_height + 1;
    return 0;
}

} // namespace rtb
