// K6b: device BVH build of SAH quality.  Bottom: parallel locally-ordered clustering — Morton codes -> radix sort -> repeated
// merging of mutual nearest neighbours (distance = surface area of the merged box, searched within +-kRadius positions of the
// Morton order) until at most kTopClusters clusters are left.  Top: a binned surface-area-heuristic build over those clusters,
// top-down, one thread block per node and one launch per level.  Then: leaf order by depth-first position -> collapse of small
// subtrees into leaves -> emission in the flattened sibling-pair layout of rt_types.h.
//
// Replaces BoundingVolumeHierarchy::constructBVH (src/bounding_volume_hierarchy.cpp:108-217).  Agglomeration builds very good
// small subtrees (every cluster joins the partner that grows its box least) but, like every bottom-up method, weaker top levels;
// the levels every ray walks through are therefore built the way the host SAH builder builds them, over a few thousand cluster
// boxes instead of all triangles.  The whole build runs on the GPU (CUB for the radix sort and the prefix sums).
// D. Meister, J. Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction", IEEE TVCG 24(3), 2018;
// I. Wald, "On fast Construction of SAH-based Bounding Volume Hierarchies", RT 2007 (binning).
#include "rt_kernels.h"

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <cub/cub.cuh>

namespace rtb {

namespace {

#ifndef RT_PLOC_LEAF
#define RT_PLOC_LEAF 2
#endif
constexpr int kLeafCollapse = RT_PLOC_LEAF; // subtrees with at most this many triangles become one leaf
#ifndef RT_PLOC_RADIUS
#define RT_PLOC_RADIUS 25 // the paper's best-quality setting; the search is 2 * radius box unions per cluster and iteration
#endif
constexpr int kRadius = RT_PLOC_RADIUS;
#ifndef RT_PLOC_TOP
#define RT_PLOC_TOP 32768 // clusters handed to the top-down SAH build
#endif
constexpr int kTopClusters = RT_PLOC_TOP;

struct Bounds {
    int lo[3], hi[3]; // order-preserving int encoding of floats
};

__device__ __forceinline__ int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
inline float ord2f_host(int i)
{
    const int v = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    std::memcpy(&f, &v, sizeof(f));
    return f;
}

__global__ void k_ploc_init_bounds(Bounds* b)
{
    for (int a = 0; a < 3; a++) {
        b->lo[a] = f2ord(FLT_MAX);
        b->hi[a] = f2ord(-FLT_MAX);
    }
}

__global__ void k_ploc_tri_bounds(const float* __restrict__ pos, int n, float4* tlo, float4* thi, Bounds* b)
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

__global__ void k_ploc_morton(const float4* __restrict__ tlo, const float4* __restrict__ thi, int n, const Bounds* b, unsigned long long* keys, int* vals)
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

// Node numbering: internal nodes 0 .. n-2 (0 = the root: ids are handed out downwards, the last merge gets 0), leaf k of the
// Morton order = n-1+k.  The first clusters are the leaves.
__global__ void k_ploc_leaves(const int* __restrict__ sorted, const float4* __restrict__ tlo, const float4* __restrict__ thi, int n, int* cluster, float4* clo,
    float4* chi, int* count, int* height)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n)
        return;
    const int t = sorted[k];
    cluster[k] = n - 1 + k;
    clo[k] = tlo[t];
    chi[k] = thi[t];
    count[n - 1 + k] = 1;
    height[n - 1 + k] = 0;
}

__global__ void k_box_union(const float4* __restrict__ lo, const float4* __restrict__ hi, int m, Bounds* b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m)
        return;
    const float4 l = lo[i], h = hi[i];
    atomicMin(&b->lo[0], f2ord(l.x));
    atomicMin(&b->lo[1], f2ord(l.y));
    atomicMin(&b->lo[2], f2ord(l.z));
    atomicMax(&b->hi[0], f2ord(h.x));
    atomicMax(&b->hi[1], f2ord(h.y));
    atomicMax(&b->hi[2], f2ord(h.z));
}

__device__ __forceinline__ float union_area(const float4& alo, const float4& ahi, const float4& blo, const float4& bhi)
{
    const float dx = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x), dy = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y), dz = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return dx * dy + dy * dz + dz * dx;
}

// Nearest neighbour of every cluster among the kRadius clusters before and after it: the one whose union with it has the
// smallest surface area (ties: the lower position).  The block stages its window of boxes in shared memory.
constexpr int kNnBlock = 256;
__global__ void __launch_bounds__(kNnBlock) k_ploc_nearest(const float4* __restrict__ clo, const float4* __restrict__ chi, int m, int* nn)
{
    __shared__ float4 slo[kNnBlock + 2 * kRadius], shi[kNnBlock + 2 * kRadius];
    const int base = blockIdx.x * kNnBlock - kRadius;
    for (int k = threadIdx.x; k < kNnBlock + 2 * kRadius; k += kNnBlock) {
        const int j = base + k;
        if (j >= 0 && j < m) {
            slo[k] = clo[j];
            shi[k] = chi[j];
        }
    }
    __syncthreads();
    const int i = blockIdx.x * kNnBlock + threadIdx.x;
    if (i >= m)
        return;
    const float4 lo = slo[threadIdx.x + kRadius], hi = shi[threadIdx.x + kRadius];
    float best = FLT_MAX;
    int best_j = -1;
    const int j0 = max(0, i - kRadius), j1 = min(m - 1, i + kRadius);
    for (int j = j0; j <= j1; j++) {
        if (j == i)
            continue;
        const float a = union_area(lo, hi, slo[j - base], shi[j - base]);
        if (a < best) {
            best = a;
            best_j = j;
        }
    }
    nn[i] = best_j;
}

// Mutual nearest neighbours merge: the lower position leads (it becomes the new node), the higher one disappears.
// flags[i] = (leads << 32) | stays, so that one exclusive sum ranks both; element m is a zero whose sum carries the totals.
__global__ void k_ploc_flags(const int* __restrict__ nn, int m, unsigned long long* flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m)
        return;
    if (i == m) {
        flags[m] = 0ull;
        return;
    }
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    const unsigned long long leads = mutual && i < j ? 1ull : 0ull, stays = mutual && i > j ? 0ull : 1ull;
    flags[i] = (leads << 32) | stays;
}

// next_id: id of the first node created by this iteration (ids go downwards from there).
__global__ void k_ploc_merge(const int* __restrict__ nn, const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ ranks, int m,
    int next_id, const int* __restrict__ cluster, const float4* __restrict__ clo, const float4* __restrict__ chi, int* cluster_out, float4* clo_out,
    float4* chi_out, int* left, int* right, int* parent, int* count, int* height, float4* blo, float4* bhi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !(flags[i] & 0xffffffffull))
        return;
    const int pos = (int)(ranks[i] & 0xffffffffull);
    if (!(flags[i] >> 32)) {
        cluster_out[pos] = cluster[i];
        clo_out[pos] = clo[i];
        chi_out[pos] = chi[i];
        return;
    }
    const int j = nn[i], id = next_id - (int)(ranks[i] >> 32);
    const int a = cluster[i], b = cluster[j];
    const float4 alo = clo[i], ahi = chi[i], clo_j = clo[j], chi_j = chi[j];
    const float4 lo = make_float4(fminf(alo.x, clo_j.x), fminf(alo.y, clo_j.y), fminf(alo.z, clo_j.z), 0.0f);
    const float4 hi = make_float4(fmaxf(ahi.x, chi_j.x), fmaxf(ahi.y, chi_j.y), fmaxf(ahi.z, chi_j.z), 0.0f);
    left[id] = a;
    right[id] = b;
    parent[a] = id;
    parent[b] = id;
    count[id] = count[a] + count[b];
    height[id] = 1 + max(height[a], height[b]);
    // boxes of the children, kept by node id for the emission
    blo[a] = alo;
    bhi[a] = ahi;
    blo[b] = clo_j;
    bhi[b] = chi_j;
    cluster_out[pos] = id;
    clo_out[pos] = lo;
    chi_out[pos] = hi;
}

// ---- top levels: binned SAH over the clusters the agglomeration left (Wald 2007), one block per node, one launch per level ----
constexpr int kBins = 16;
constexpr int kTopBlock = 128;
struct TopTask {
    int id, begin, end; // node to split and its range of the cluster index array
};
struct TopState {
    int next_id;     // next unused internal node id (ids grow downwards in the tree: a child's id exceeds its parent's)
    int n_tasks[2];  // task counts of the current / next level (ping-pong)
    int depth;       // deepest leaf of the finished tree seen so far (levels of the top part + height of the cluster below)
};

__device__ __forceinline__ float half_area(const int* lo, const int* hi)
{
    const float dx = ord2f(hi[0]) - ord2f(lo[0]), dy = ord2f(hi[1]) - ord2f(lo[1]), dz = ord2f(hi[2]) - ord2f(lo[2]);
    return dx * dy + dy * dz + dz * dx;
}

// One block splits one node: centroid bounds of its clusters -> kBins bins on each axis (box + triangle count per bin) -> the
// cheapest of the 3 * (kBins - 1) planes, cost = area(left) * triangles(left) + area(right) * triangles(right) -> partition of
// the index range -> children (a range of one cluster is that cluster's node; larger ranges become tasks of the next level).
__global__ void __launch_bounds__(kTopBlock) k_top_split(const TopTask* __restrict__ tasks, int level, int par, TopTask* next_tasks, TopState* state,
    const int* __restrict__ idx_in, int* idx_out, const int* __restrict__ cluster, const float4* __restrict__ clo, const float4* __restrict__ chi,
    int* left, int* right, int* parent, int* count, const int* __restrict__ height, float4* blo, float4* bhi)
{
    __shared__ int c_lo[3], c_hi[3];
    __shared__ int b_lo[3][kBins][3], b_hi[3][kBins][3], b_cnt[3][kBins], b_n[3][kBins];
    __shared__ float s_cost[3 * (kBins - 1)];
    __shared__ int s_best, s_nl, s_nr, s_fill[2];
    __shared__ int side_lo[2][3], side_hi[2][3], side_cnt[2];
    const TopTask t = tasks[blockIdx.x];
    const int n = t.end - t.begin;
    for (int k = threadIdx.x; k < 3; k += kTopBlock) {
        c_lo[k] = f2ord(FLT_MAX);
        c_hi[k] = f2ord(-FLT_MAX);
    }
    for (int k = threadIdx.x; k < 3 * kBins; k += kTopBlock) {
        const int a = k / kBins, b = k % kBins;
        for (int d = 0; d < 3; d++) {
            b_lo[a][b][d] = f2ord(FLT_MAX);
            b_hi[a][b][d] = f2ord(-FLT_MAX);
        }
        b_cnt[a][b] = 0;
        b_n[a][b] = 0;
    }
    if (threadIdx.x < 2) {
        s_fill[threadIdx.x] = 0;
        side_cnt[threadIdx.x] = 0;
        for (int d = 0; d < 3; d++) {
            side_lo[threadIdx.x][d] = f2ord(FLT_MAX);
            side_hi[threadIdx.x][d] = f2ord(-FLT_MAX);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kTopBlock) { // centroid bounds
        const int e = idx_in[t.begin + k];
        const float4 lo = clo[e], hi = chi[e];
        const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
        for (int a = 0; a < 3; a++) {
            atomicMin(&c_lo[a], f2ord(c[a]));
            atomicMax(&c_hi[a], f2ord(c[a]));
        }
    }
    __syncthreads();
    float cmin[3], scale[3];
    for (int a = 0; a < 3; a++) {
        cmin[a] = ord2f(c_lo[a]);
        const float ext = ord2f(c_hi[a]) - cmin[a];
        scale[a] = ext > 0.0f ? (float)kBins / ext : 0.0f;
    }
    for (int k = threadIdx.x; k < n; k += kTopBlock) { // binning
        const int e = idx_in[t.begin + k];
        const float4 lo = clo[e], hi = chi[e];
        const float c[3] = { 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z) };
        const int ilo[3] = { f2ord(lo.x), f2ord(lo.y), f2ord(lo.z) }, ihi[3] = { f2ord(hi.x), f2ord(hi.y), f2ord(hi.z) };
        const int tris = count[cluster[e]];
        for (int a = 0; a < 3; a++) {
            const int b = min(kBins - 1, max(0, (int)((c[a] - cmin[a]) * scale[a])));
            for (int d = 0; d < 3; d++) {
                atomicMin(&b_lo[a][b][d], ilo[d]);
                atomicMax(&b_hi[a][b][d], ihi[d]);
            }
            atomicAdd(&b_cnt[a][b], tris);
            atomicAdd(&b_n[a][b], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * (kBins - 1)) { // plane after bin `s` of axis `a`
        const int a = threadIdx.x / (kBins - 1), sp = threadIdx.x % (kBins - 1);
        int lo[3] = { f2ord(FLT_MAX), f2ord(FLT_MAX), f2ord(FLT_MAX) }, hi[3] = { f2ord(-FLT_MAX), f2ord(-FLT_MAX), f2ord(-FLT_MAX) };
        int cl = 0, nl = 0, cr = 0, nr = 0;
        for (int b = 0; b <= sp; b++) {
            for (int d = 0; d < 3; d++) {
                lo[d] = min(lo[d], b_lo[a][b][d]);
                hi[d] = max(hi[d], b_hi[a][b][d]);
            }
            cl += b_cnt[a][b];
            nl += b_n[a][b];
        }
        float cost = FLT_MAX;
        if (nl > 0 && nl < n) {
            const float al = half_area(lo, hi);
            for (int d = 0; d < 3; d++) {
                lo[d] = f2ord(FLT_MAX);
                hi[d] = f2ord(-FLT_MAX);
            }
            for (int b = sp + 1; b < kBins; b++) {
                for (int d = 0; d < 3; d++) {
                    lo[d] = min(lo[d], b_lo[a][b][d]);
                    hi[d] = max(hi[d], b_hi[a][b][d]);
                }
                cr += b_cnt[a][b];
                nr += b_n[a][b];
            }
            cost = al * (float)cl + half_area(lo, hi) * (float)cr;
        }
        s_cost[threadIdx.x] = cost;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = -1;
        float bc = FLT_MAX;
        for (int k = 0; k < 3 * (kBins - 1); k++)
            if (s_cost[k] < bc) {
                bc = s_cost[k];
                best = k;
            }
        s_best = best;
        int nl = 0;
        if (best >= 0) {
            const int a = best / (kBins - 1), sp = best % (kBins - 1);
            for (int b = 0; b <= sp; b++)
                nl += b_n[a][b];
        } else
            nl = n / 2; // all centroids coincide: halve the range as it stands
        s_nl = nl;
        s_nr = n - nl;
    }
    __syncthreads();
    const int best = s_best, nl = s_nl;
    const int axis = best >= 0 ? best / (kBins - 1) : 0, split = best >= 0 ? best % (kBins - 1) : 0;
    for (int k = threadIdx.x; k < n; k += kTopBlock) { // partition, and the boxes / triangle counts of the two sides
        const int e = idx_in[t.begin + k];
        const float4 lo = clo[e], hi = chi[e];
        int side;
        if (best >= 0) {
            const float c = axis == 0 ? 0.5f * (lo.x + hi.x) : (axis == 1 ? 0.5f * (lo.y + hi.y) : 0.5f * (lo.z + hi.z));
            side = min(kBins - 1, max(0, (int)((c - cmin[axis]) * scale[axis]))) <= split ? 0 : 1;
        } else
            side = k < nl ? 0 : 1;
        const int slot = atomicAdd(&s_fill[side], 1);
        idx_out[side == 0 ? t.begin + slot : t.begin + nl + slot] = e;
        const int ilo[3] = { f2ord(lo.x), f2ord(lo.y), f2ord(lo.z) }, ihi[3] = { f2ord(hi.x), f2ord(hi.y), f2ord(hi.z) };
        for (int d = 0; d < 3; d++) {
            atomicMin(&side_lo[side][d], ilo[d]);
            atomicMax(&side_hi[side][d], ihi[d]);
        }
        atomicAdd(&side_cnt[side], count[cluster[e]]);
    }
    __syncthreads();
    if (threadIdx.x < 2) { // the two children
        const int side = threadIdx.x, b = side == 0 ? t.begin : t.begin + nl, e = side == 0 ? t.begin + nl : t.end;
        int child;
        if (e - b == 1) {
            child = cluster[idx_out[b]];
            atomicMax(&state->depth, level + 1 + height[child]);
        } else {
            child = atomicAdd(&state->next_id, 1);
            count[child] = side_cnt[side];
            const int slot = atomicAdd(&state->n_tasks[par ^ 1], 1);
            next_tasks[slot] = TopTask { child, b, e };
        }
        (side == 0 ? left : right)[t.id] = child;
        parent[child] = t.id;
        blo[child] = make_float4(ord2f(side_lo[side][0]), ord2f(side_lo[side][1]), ord2f(side_lo[side][2]), 0.0f);
        bhi[child] = make_float4(ord2f(side_hi[side][0]), ord2f(side_hi[side][1]), ord2f(side_hi[side][2]), 0.0f);
    }
}

__global__ void k_top_init(int m, int total_tris, int* idx, int* count, TopTask* tasks, TopState* state)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m)
        idx[i] = i;
    if (i == 0) {
        tasks[0] = TopTask { 0, 0, m };
        count[0] = total_tris;
        state->next_id = 1;
        state->n_tasks[0] = 1;
        state->n_tasks[1] = 0;
        state->depth = 0;
    }
}

__global__ void k_top_next_level(TopState* state, int par)
{
    state->n_tasks[par] = 0; // the level just processed; its slot counts the level after next
}

// Depth-first position of the first leaf below every node: the sum, over the ancestors in whose right subtree the node lies,
// of the leaf count of the left subtree.  Leaves then know their slot in the final triangle order.
__global__ void k_ploc_first(int n_nodes_total, const int* __restrict__ left, const int* __restrict__ parent, const int* __restrict__ count, int* first)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes_total)
        return;
    int pos = 0;
    for (int c = v, p = parent[v]; p >= 0; c = p, p = parent[p])
        if (left[p] != c)
            pos += count[left[p]];
    first[v] = pos;
}

__global__ void k_ploc_perm(const int* __restrict__ sorted, const int* __restrict__ first, int n, int* perm)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n)
        perm[first[n - 1 + k]] = sorted[k];
}

__global__ void k_ploc_live(const int* __restrict__ count, int n, int* live)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1)
        live[i] = count[i] > kLeafCollapse ? 1 : 0;
}

__device__ __forceinline__ void emit_node(float4* nodes, int idx, float4 lo, float4 hi, float pad, int left_or_first, int count)
{
    const int entry = count ? ~((left_or_first << 3) | (count - 1)) : left_or_first; // see rt_types.h
    nodes[2 * (size_t)idx] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, __int_as_float(entry));
    nodes[2 * (size_t)idx + 1] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, __int_as_float(count));
}

// Every live internal node writes its two children into the pair slot (1 + rank among live nodes); node 0 is the root.
__global__ void k_ploc_emit(int n, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ first, const int* __restrict__ count,
    const int* __restrict__ live, const int* __restrict__ live_rank, const float4* __restrict__ blo, const float4* __restrict__ bhi, float4 root_lo, float4 root_hi,
    float pad, float4* nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || !live[i])
        return;
    const int pair = 2 * (1 + live_rank[i]);
    const int ch[2] = { left[i], right[i] };
    for (int k = 0; k < 2; k++) {
        const int c = ch[k];
        if (c < n - 1 && live[c])
            emit_node(nodes, pair + k, blo[c], bhi[c], pad, 2 * (1 + live_rank[c]), 0);
        else // a triangle, or a collapsed subtree: a contiguous run of the final order
            emit_node(nodes, pair + k, blo[c], bhi[c], pad, first[c], count[c]);
    }
    if (i == 0) { // root and its twin
        emit_node(nodes, 0, root_lo, root_hi, pad, pair, 0);
        emit_node(nodes, 1, root_lo, root_hi, pad, pair, 0);
    }
}

// All scratch of a build comes out of ONE allocation (a cudaMalloc per array costs more than the build: ~40 of them took 40-80 ms).
struct Scratch {
    char* base = nullptr;
    size_t capacity = 0, used = 0;
    cudaError_t reserve(size_t bytes, BuildScratch* keep) // the caller's arena, grown if it is too small
    {
        if (keep->capacity < bytes) {
            cudaFree(keep->base);
            keep->base = nullptr;
            keep->capacity = 0;
            const cudaError_t e = cudaMalloc(&keep->base, bytes);
            if (e != cudaSuccess)
                return e;
            keep->capacity = bytes;
        }
        base = keep->base;
        capacity = keep->capacity;
        return cudaSuccess;
    }
    template <typename T> T* alloc(size_t n, cudaError_t& e)
    {
        const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256;
        if (e != cudaSuccess)
            return nullptr;
        if (used + bytes > capacity) {
            e = cudaErrorMemoryAllocation;
            return nullptr;
        }
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
};

} // namespace

// Returns 0 on success, 1 on a CUDA error (*err set), 2 when the scene is too small for this builder (the caller uses the LBVH
// path, which handles single-leaf scenes).
int build_bvh_ploc_device(cudaStream_t st, const float* d_pos, long long n_tris, float pad, DeviceBvh* out, const char** err, BuildScratch* keep)
{
    const int n = (int)n_tris;
    if (n <= kLeafCollapse)
        return 2;
    const bool trace = std::getenv("RTB200_TRACE_BUILD") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    double t_alloc = 0, t_sorted = 0, t_clustered = 0, t_top = 0;
    const int blk = 256, grid = (n + blk - 1) / blk;
    cudaError_t e = cudaSuccess;
    size_t tmp_sort = 0, tmp_scan = 0, tmp_scan2 = 0;
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n, 0, 63, st);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (unsigned long long*)nullptr, (unsigned long long*)nullptr, n + 1, st);
    if (e == cudaSuccess)
        e = cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan2, (int*)nullptr, (int*)nullptr, n, st);
    tmp_scan = std::max(tmp_scan, tmp_scan2);
    Scratch sc;
    if (e == cudaSuccess) // 292 bytes of arrays per triangle (listed below) + the library's temporaries + alignment slack
        e = sc.reserve((size_t)n * 300 + std::max(tmp_sort, tmp_scan) + (64 << 10), keep);
    float4* tlo = sc.alloc<float4>(n, e);
    float4* thi = sc.alloc<float4>(n, e);
    Bounds* bounds = sc.alloc<Bounds>(1, e);
    unsigned long long* keys_in = sc.alloc<unsigned long long>(n, e);
    unsigned long long* keys = sc.alloc<unsigned long long>(n, e);
    int* vals_in = sc.alloc<int>(n, e);
    int* sorted = sc.alloc<int>(n, e);
    int* cluster[2] = { sc.alloc<int>(n, e), sc.alloc<int>(n, e) };
    float4* clo[2] = { sc.alloc<float4>(n, e), sc.alloc<float4>(n, e) };
    float4* chi[2] = { sc.alloc<float4>(n, e), sc.alloc<float4>(n, e) };
    int* nn = sc.alloc<int>(n, e);
    unsigned long long* flags = sc.alloc<unsigned long long>((size_t)n + 1, e);
    unsigned long long* ranks = sc.alloc<unsigned long long>((size_t)n + 1, e);
    const size_t total = 2 * (size_t)n - 1; // nodes: n-1 internal + n leaves
    int* left = sc.alloc<int>(n, e);
    int* right = sc.alloc<int>(n, e);
    int* parent = sc.alloc<int>(total, e);
    int* count = sc.alloc<int>(total, e);
    int* height = sc.alloc<int>(total, e);
    int* first = sc.alloc<int>(total, e);
    float4* blo = sc.alloc<float4>(total, e);
    float4* bhi = sc.alloc<float4>(total, e);
    int* live = sc.alloc<int>(n, e);
    int* live_rank = sc.alloc<int>(n, e);
    int* perm = nullptr;
    void* tmp = sc.alloc<unsigned char>(std::max(tmp_sort, tmp_scan), e);
    int* idx[2] = { sc.alloc<int>(n, e), sc.alloc<int>(n, e) };                     // top-level build: at most min(n, kTopClusters) entries are used
    TopTask* tasks[2] = { sc.alloc<TopTask>(n, e), sc.alloc<TopTask>(n, e) };
    TopState* state = sc.alloc<TopState>(1, e);
    Bounds* rb = sc.alloc<Bounds>(1, e);
    if (e == cudaSuccess)
        e = cudaMalloc(&perm, (size_t)n * sizeof(int));
    if (e != cudaSuccess) {
        *err = cudaGetErrorString(e);
        return 1;
    }
    auto fail = [&](cudaError_t ce) {
        cudaFree(perm);
        *err = cudaGetErrorString(ce);
        return 1;
    };
    t_alloc = since();
    k_ploc_init_bounds<<<1, 1, 0, st>>>(bounds);
    k_ploc_tri_bounds<<<grid, blk, 0, st>>>(d_pos, n, tlo, thi, bounds);
    k_ploc_morton<<<grid, blk, 0, st>>>(tlo, thi, n, bounds, keys_in, vals_in);
    cub::DeviceRadixSort::SortPairs(tmp, tmp_sort, keys_in, keys, vals_in, sorted, n, 0, 63, st);
    k_ploc_leaves<<<grid, blk, 0, st>>>(sorted, tlo, thi, n, cluster[0], clo[0], chi[0], count, height);
    cudaMemsetAsync(parent, 0xff, sizeof(int), st); // the root (node 0) has no parent

    if (trace) {
        cudaStreamSynchronize(st);
        t_sorted = since();
    }
    int m = n, next_id = n - 2, cur = 0, iterations = 0;
    unsigned long long* h_totals = keep->host_word; // pinned: one 8-byte read-back per iteration
    while (m > kTopClusters) {
        const int g = (m + 1 + blk - 1) / blk;
        k_ploc_nearest<<<(m + kNnBlock - 1) / kNnBlock, kNnBlock, 0, st>>>(clo[cur], chi[cur], m, nn);
        k_ploc_flags<<<g, blk, 0, st>>>(nn, m, flags);
        cub::DeviceScan::ExclusiveSum(tmp, tmp_scan, flags, ranks, m + 1, st);
        k_ploc_merge<<<g, blk, 0, st>>>(nn, flags, ranks, m, next_id, cluster[cur], clo[cur], chi[cur], cluster[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1], left,
            right, parent, count, height, blo, bhi);
        cudaMemcpyAsync(h_totals, ranks + m, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return fail(e);
        const int merged = (int)(*h_totals >> 32), stayed = (int)(*h_totals & 0xffffffffull);
        if (merged <= 0 || stayed != m - merged) { // the closest pair of the whole array is always mutual
            cudaFree(perm);
            *err = "PLOC iteration made no progress";
            return 1;
        }
        next_id -= merged;
        m = stayed;
        cur ^= 1;
        iterations++;
    }
    t_clustered = since();
    // ---- top levels over the m clusters that are left: internal nodes 0 .. m-2 remain to be made (next_id == m - 2) ----
    if (next_id != m - 2) {
        cudaFree(perm);
        *err = "PLOC node numbering is inconsistent";
        return 1;
    }
    int tree_height = 0, top_levels = 0;
    float4 root_lo, root_hi;
    {

        k_top_init<<<(m + blk - 1) / blk, blk, 0, st>>>(m, n, idx[0], count, tasks[0], state);
        TopState hs;
        int n_tasks = m > 1 ? 1 : 0;
        // every level reads the index array of the level before: ranges that are not split any more keep their (final) order
        // in whichever buffer they were last written to, which nothing reads afterwards (children of a range of one are nodes)
        for (int par = 0; n_tasks > 0; par ^= 1, top_levels++) {
            k_top_split<<<n_tasks, kTopBlock, 0, st>>>(tasks[par], top_levels, par, tasks[par ^ 1], state, idx[par], idx[par ^ 1], cluster[cur], clo[cur],
                chi[cur], left, right, parent, count, height, blo, bhi);
            k_top_next_level<<<1, 1, 0, st>>>(state, par);
            cudaMemcpyAsync(&hs, state, sizeof(TopState), cudaMemcpyDeviceToHost, st);
            e = cudaStreamSynchronize(st);
            if (e != cudaSuccess)
                return fail(e);
            n_tasks = hs.n_tasks[par ^ 1];
            if (top_levels > 4 * 64) { // cannot happen: every split leaves both sides non-empty
                cudaFree(perm);
                *err = "top-level SAH build does not terminate";
                return 1;
            }
        }
        if (m > 1 && hs.next_id != m - 1) {
            cudaFree(perm);
            *err = "top-level SAH build made the wrong number of nodes";
            return 1;
        }
        tree_height = m > 1 ? hs.depth : 0;
        // the root's box: union of the remaining clusters (a short reduction through the ordered-int atomics of k_ploc_tri_bounds' kind)
        k_ploc_init_bounds<<<1, 1, 0, st>>>(rb);
        k_box_union<<<(m + blk - 1) / blk, blk, 0, st>>>(clo[cur], chi[cur], m, rb);
        Bounds hb;
        cudaMemcpyAsync(&hb, rb, sizeof(Bounds), cudaMemcpyDeviceToHost, st);
        if (m == 1)
            cudaMemcpyAsync(&tree_height, height, sizeof(int), cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return fail(e);
        root_lo = make_float4(ord2f_host(hb.lo[0]), ord2f_host(hb.lo[1]), ord2f_host(hb.lo[2]), 0.0f);
        root_hi = make_float4(ord2f_host(hb.hi[0]), ord2f_host(hb.hi[1]), ord2f_host(hb.hi[2]), 0.0f);
    }
    t_top = since();
    if (trace)
        std::fprintf(stderr, "[ploc] %d triangles, radius %d: %d clustering iterations, then %d SAH levels over %d clusters | ms: alloc %.2f, sort %.2f, clustering %.2f, top %.2f\n",
            n, kRadius, iterations, top_levels, m, t_alloc, t_sorted - t_alloc, t_clustered - t_sorted, t_top - t_clustered);
    const int gt = (int)((total + blk - 1) / blk);
    k_ploc_first<<<gt, blk, 0, st>>>((int)total, left, parent, count, first);
    k_ploc_perm<<<grid, blk, 0, st>>>(sorted, first, n, perm);
    k_ploc_live<<<grid, blk, 0, st>>>(count, n, live);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_scan, live, live_rank, n - 1, st);
    int n_live_last[2] = { 0, 0 };
    cudaMemcpyAsync(&n_live_last[0], live_rank + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&n_live_last[1], live + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
        return fail(e);
    const int n_live = n_live_last[0] + n_live_last[1];
    const int n_nodes = 2 * (1 + n_live);
    float4* nodes = nullptr;
    e = cudaMalloc(&nodes, 2 * (size_t)n_nodes * sizeof(float4));
    if (e == cudaSuccess) {
        k_ploc_emit<<<grid, blk, 0, st>>>(n, left, right, first, count, live, live_rank, blo, bhi, root_lo, root_hi, pad, nodes);
        e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(nodes);
        return fail(e);
    }
    if (trace)
        std::fprintf(stderr, "[ploc] order + emit %.2f ms, total %.2f ms (the scratch arena is kept by the caller)\n", since() - t_top, since());
    out->nodes = nodes;
    out->perm = perm;
    out->n_nodes = n_nodes;
    out->root_entry = 2; // the root is live (n > kLeafCollapse): its children are the first emitted pair
    out->depth = tree_height + 1;
    return 0;
}

} // namespace rtb
