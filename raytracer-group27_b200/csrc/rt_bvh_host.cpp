// Host-side binned-SAH BVH builder (RT_BVH_SAH_HOST).
//
// Replaces BoundingVolumeHierarchy::constructBVH (src/bounding_volume_hierarchy.cpp:108-217), whose median-split
// tree is capped at 5 levels / 16 leaves (bounding_volume_hierarchy.h:67).  Any conservative hierarchy returns the
// same closest hits, so the tree shape is free: this builder minimises the surface-area
// heuristic with 16 bins per axis and emits the flattened sibling-pair layout described in rt_types.h.
#include "rt_kernels.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace rtb {

namespace {

struct Box {
    float lo[3], hi[3];
    void reset()
    {
        for (int a = 0; a < 3; a++) {
            lo[a] = FLT_MAX;
            hi[a] = -FLT_MAX;
        }
    }
    void grow(const Box& b)
    {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    void grow(const float* p)
    {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    float half_area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0)
            return 0.0f;
        return dx * dy + dy * dz + dz * dx;
    }
};

#ifndef RT_SAH_BINS
#define RT_SAH_BINS 16
#endif
#ifndef RT_SAH_TRI_COST
#define RT_SAH_TRI_COST 1.2f
#endif
constexpr int kBins = RT_SAH_BINS;
constexpr float kTraversalCost = 1.0f;
constexpr float kTriangleCost = RT_SAH_TRI_COST;

struct Work {
    int node;
    long long begin, end;
    int depth;
};

void write_node(std::vector<float4>& nodes, int idx, const Box& b, float pad, int left_or_first, int count)
{
    float4 a, c;
    a.x = b.lo[0] - pad; a.y = b.lo[1] - pad; a.z = b.lo[2] - pad;
    c.x = b.hi[0] + pad; c.y = b.hi[1] + pad; c.z = b.hi[2] + pad;
    // lo.w carries the node's traversal-stack entry ready-made: inner node -> index of its left child (the sibling
    // pair to fetch), leaf -> ~((first << 3) | (count - 1)); hi.w keeps the triangle count (0 = inner)
    const int entry = count ? ~((left_or_first << 3) | (count - 1)) : left_or_first;
    std::memcpy(&a.w, &entry, 4);
    std::memcpy(&c.w, &count, 4);
    nodes[2 * (size_t)idx] = a;
    nodes[2 * (size_t)idx + 1] = c;
}

} // namespace

HostBvh build_bvh_sah_host(const float* pos, long long n, float pad)
{
    HostBvh out;
    std::vector<Box> tb((size_t)n);
    std::vector<float> cen((size_t)n * 3);
    for (long long i = 0; i < n; i++) {
        Box b;
        b.reset();
        b.grow(pos + 9 * i);
        b.grow(pos + 9 * i + 3);
        b.grow(pos + 9 * i + 6);
        tb[(size_t)i] = b;
        for (int a = 0; a < 3; a++)
            cen[(size_t)i * 3 + a] = 0.5f * (b.lo[a] + b.hi[a]);
    }
    out.perm.resize((size_t)n);
    for (long long i = 0; i < n; i++)
        out.perm[(size_t)i] = (int)i;
    std::vector<int>& idx = out.perm;
    out.nodes.resize(4); // root + its twin slot
    std::vector<Work> stack;
    stack.push_back({ 0, 0, n, 1 });
    int max_depth = 1;
    bool root_is_leaf = false;

    while (!stack.empty()) {
        const Work w = stack.back();
        stack.pop_back();
        max_depth = std::max(max_depth, w.depth);
        const long long cnt = w.end - w.begin;
        Box nb, cb;
        nb.reset();
        cb.reset();
        for (long long i = w.begin; i < w.end; i++) {
            nb.grow(tb[(size_t)idx[(size_t)i]]);
            cb.grow(&cen[(size_t)idx[(size_t)i] * 3]);
        }
        bool make_leaf = cnt <= 1;
        long long mid = -1;
        if (!make_leaf) {
            // binned SAH over the three axes
            float best_cost = FLT_MAX;
            int best_axis = -1, best_bin = -1;
            if (w.depth < kStackDepth - 8) {
                for (int axis = 0; axis < 3; axis++) {
                    const float ext = cb.hi[axis] - cb.lo[axis];
                    if (!(ext > 0.0f))
                        continue;
                    Box bins[kBins];
                    long long bc[kBins];
                    for (int b = 0; b < kBins; b++) {
                        bins[b].reset();
                        bc[b] = 0;
                    }
                    const float scale = kBins / ext;
                    for (long long i = w.begin; i < w.end; i++) {
                        const int t = idx[(size_t)i];
                        int b = (int)((cen[(size_t)t * 3 + axis] - cb.lo[axis]) * scale);
                        b = std::min(std::max(b, 0), kBins - 1);
                        bins[b].grow(tb[(size_t)t]);
                        bc[b]++;
                    }
                    float right_area[kBins];
                    long long right_cnt[kBins];
                    Box acc;
                    acc.reset();
                    long long c = 0;
                    for (int b = kBins - 1; b > 0; b--) {
                        acc.grow(bins[b]);
                        c += bc[b];
                        right_area[b] = acc.half_area();
                        right_cnt[b] = c;
                    }
                    acc.reset();
                    c = 0;
                    for (int b = 0; b < kBins - 1; b++) {
                        acc.grow(bins[b]);
                        c += bc[b];
                        if (c == 0 || right_cnt[b + 1] == 0)
                            continue;
                        const float cost = acc.half_area() * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                        if (cost < best_cost) {
                            best_cost = cost;
                            best_axis = axis;
                            best_bin = b;
                        }
                    }
                }
            }
            const float parent_area = nb.half_area();
            const float split_cost = best_axis >= 0 && parent_area > 0.0f ? kTraversalCost + kTriangleCost * best_cost / parent_area : FLT_MAX;
            const float leaf_cost = kTriangleCost * (float)cnt;
            if (cnt <= kMaxLeafTris && !(split_cost < leaf_cost)) {
                make_leaf = true; // small enough and splitting does not pay
            } else if (best_axis >= 0) {
                const float ext = cb.hi[best_axis] - cb.lo[best_axis];
                const float scale = kBins / ext;
                const float lo = cb.lo[best_axis];
                auto it = std::partition(idx.begin() + w.begin, idx.begin() + w.end, [&](int t) {
                    int b = (int)((cen[(size_t)t * 3 + best_axis] - lo) * scale);
                    b = std::min(std::max(b, 0), kBins - 1);
                    return b <= best_bin;
                });
                mid = it - idx.begin();
            }
            if (!make_leaf && (mid <= w.begin || mid >= w.end)) {
                // degenerate (coincident centroids) or depth guard: balanced median split on the widest axis
                int axis = 0;
                for (int a = 1; a < 3; a++)
                    if (cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis])
                        axis = a;
                mid = w.begin + cnt / 2;
                std::nth_element(idx.begin() + w.begin, idx.begin() + mid, idx.begin() + w.end,
                    [&](int x, int y) { return cen[(size_t)x * 3 + axis] < cen[(size_t)y * 3 + axis]; });
            }
        }
        if (make_leaf) {
            write_node(out.nodes, w.node, nb, pad, (int)w.begin, (int)cnt);
            if (w.node == 0)
                root_is_leaf = true;
            continue;
        }
        const int left = (int)(out.nodes.size() / 2);
        out.nodes.resize(out.nodes.size() + 4);
        write_node(out.nodes, w.node, nb, pad, left, 0);
        stack.push_back({ left + 1, mid, w.end, w.depth + 1 });
        stack.push_back({ left, w.begin, mid, w.depth + 1 });
    }
    if (n == 0) {
        Box e;
        e.lo[0] = e.lo[1] = e.lo[2] = 1e30f;
        e.hi[0] = e.hi[1] = e.hi[2] = 1e30f;
        write_node(out.nodes, 0, e, 0.0f, 0, 0);
        root_is_leaf = true;
    }
    // node 1 mirrors node 0 so pair 0 can be fetched like any other pair (duplicates lose the (t, id) tie rule)
    out.nodes[2] = out.nodes[0];
    out.nodes[3] = out.nodes[1];
    if (root_is_leaf) {
        out.root_entry = 0;
    } else {
        int left;
        std::memcpy(&left, &out.nodes[0].w, 4);
        out.root_entry = left;
    }
    out.depth = max_depth;
    return out;
}

} // namespace rtb
