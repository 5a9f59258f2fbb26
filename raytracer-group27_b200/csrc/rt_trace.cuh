// Closest-hit / any-hit traversal of the flattened BVH and the exact triangle test.
//
// Replaces BoundingVolumeHierarchy::intersect / intersectBVH / intersectNode / intersectObject
// (src/bounding_volume_hierarchy.cpp:49-78, 395-448) and intersectRayWithTriangle / trianglePlane /
// intersectRayWithPlane / pointInTriangle (src/ray_tracing.cpp:42-128).
//
// Semantics kept from the reference: a triangle is hit when the plane parameter t (measured along
// normalize(direction), ray_tracing.cpp:65-71) satisfies 0 <= t < ray.t and the point origin + direction * t
// (UN-normalised direction, ray_tracing.cpp:111) passes the three edge-sign tests, all >= 0 or all < 0.  The
// reference keeps the first strictly smaller t, i.e. the winner is the lexicographic minimum of (t, visiting order):
// the global object id in its brute-force loop (useBVH = false), the depth-first rank of its own BVH otherwise
// (rt_reforder.cu); a traversal in any order reproduces either with the tie rule below (SceneDev::tie_by_id selects the key).  Box tests only have
// to be conservative (never cull a triangle the reference would accept): boxes are padded at build time and the
// slab comparison carries a rounding guard; the reference's own box test is not reproduced (its result never
// changes which triangle wins).  The BVH assumes |direction| = 1 up to rounding, which holds
// for every ray the renderer creates; for other rays only the exhaustive search reproduces the reference.
#pragma once
#include "rt_math.cuh"
#include "rt_types.h"
#include <cfloat>
#include <climits>

namespace rtb {

// Checked build (RT_CHECKED, rt_types.h): RT_GUARD(i, n, site) is i when 0 <= i < n; otherwise the violation is counted and the
// access goes to element 0.  One counter table per translation unit that includes this header (rt_kernels.cu, rt_wide8.cu), read
// back by add_violations_*() (rt_kernels.h).  Default build: the identity.
#if RT_CHECKED
static __device__ unsigned int g_rt_violations[kChkSites];
template <typename T> __device__ __forceinline__ T rt_guard(T i, long long n, int site)
{
    if ((long long)i < 0 || (long long)i >= n) {
        atomicAdd(&g_rt_violations[site], 1u);
        return (T)0;
    }
    return i;
}
#define RT_GUARD(i, n, site) rtb::rt_guard((i), (long long)(n), (site))
#else
#define RT_GUARD(i, n, site) (i)
#endif

struct TraceStats {
    unsigned int nodes = 0, tris = 0, tris_full = 0;
    unsigned int ray_start_nodes = 0, ray_start_tris = 0; // counters when the lane's current query began (instrumented builds)
    unsigned int max_ray_nodes = 0, max_ray_tris = 0;     // most boxes / triangles any one query of this lane touched
    __device__ __forceinline__ void query_begins() { ray_start_nodes = nodes; ray_start_tris = tris; }
    __device__ __forceinline__ void query_ends()
    {
        max_ray_nodes = max(max_ray_nodes, nodes - ray_start_nodes);
        max_ray_tris = max(max_ray_tris, tris - ray_start_tris);
    }
};

struct HitRec {
    float t;   // best t so far (search bound on entry)
    int key;   // tie key of the best hit: its visiting rank in the reference's BVH, or its global id (tie_by_id searches)
    int ti;    // BVH-order index of the best triangle hit; -1 = none; -2 - k = sphere k
};

// One ray against one triangle.  `o`/`d` as stored in the ray, `dn` = normalize(d); `pl` = the triangle's plane record.
#ifndef RT_TRI_EAGER_V1
#define RT_TRI_EAGER_V1 0 // 1: fetch v1 together with v0 / v2 instead of after the tie check (one round trip less per full test)
#endif
#ifndef RT_LEAF_PREFETCH
#define RT_LEAF_PREFETCH 0 // 1: the plane of the leaf's next triangle is requested before the current one is tested
#endif
template <bool COUNT>
__device__ __forceinline__ bool test_triangle(const SceneDev& s, int ti, const float4& pl, const f3& o, const f3& d, const f3& dn, HitRec& best, TraceStats& st)
{
    ti = RT_GUARD(ti, s.n_tris, kChkTri);
    const f3 n = mk3(pl);
    if (COUNT)
        st.tris++;
    const float nd = xdot(dn, n);
    if (nd == 0.0f)
        return false;
    const float t = xdiv(xsub(pl.w, xdot(o, n)), nd);
    if (!(t >= 0.0f) || !(t <= best.t))
        return false;
    const float4 a = __ldg(&s.tri_v0[kTriStride * ti]);
    const float4 c = __ldg(&s.tri_v2[kTriStride * ti]);
#if RT_TRI_EAGER_V1
    const float4 b = __ldg(&s.tri_v1[kTriStride * ti]);
#endif
    const int key = __float_as_int(s.tie_by_id ? c.w : a.w);
    if (t == best.t && key >= best.key)
        return false; // equal t: the object the reference visits first wins
#if !RT_TRI_EAGER_V1
    const float4 b = __ldg(&s.tri_v1[kTriStride * ti]);
#endif
    if (COUNT)
        st.tris_full++;
    const f3 v0 = mk3(a), v1 = mk3(b), v2 = mk3(c);
    const f3 p = xadd(o, xmul(d, t));
    const bool s0 = xdot(xcross(xsub(p, v0), xsub(v2, v0)), n) >= 0.0f;
    const bool s1 = xdot(xcross(xsub(p, v2), xsub(v1, v2)), n) >= 0.0f;
    const bool s2 = xdot(xcross(xsub(p, v1), xsub(v0, v1)), n) >= 0.0f;
    if ((s0 && s1 && s2) || (!s0 && !s1 && !s2)) {
        best.t = t;
        best.key = key;
        best.ti = ti;
        return true;
    }
    return false;
}
template <bool COUNT>
__device__ __forceinline__ bool test_triangle(const SceneDev& s, int ti, const f3& o, const f3& d, const f3& dn, HitRec& best, TraceStats& st)
{
    ti = RT_GUARD(ti, s.n_tris, kChkTri);
    return test_triangle<COUNT>(s, ti, __ldg(&s.tri_plane[kTriStride * ti]), o, d, dn, best, st);
}

// Sphere primitives: intersectRayWithShape(const Sphere&, Ray&, HitInfo&) (src/ray_tracing.cpp:182-209).  The reference's
// glm::pow(float, int) is std::pow and returns double, so the sums of squares and the discriminant are formed in double
// and rounded to float; x*x of a float is exact in double, so pow(x, 2) is a plain product.  The parameter is along the
// ray's own (un-normalised) direction.  A hit needs t < ray.t strictly; in the reference's brute-force order spheres come
// after all triangles (ids n_tris + k), in its BVH they are objects like the triangles (sphere_rank).
__device__ __forceinline__ bool test_spheres(const SceneDev& s, const f3& o, const f3& d, HitRec& best)
{
    bool any = false;
    for (int k = 0; k < s.n_spheres; k++) {
        const float4 cr = __ldg(&s.spheres[3 * k]);
        const f3 m = xsub(o, mk3(cr));
        const float A = (float)__dadd_rn(__dadd_rn(__dmul_rn(d.x, d.x), __dmul_rn(d.y, d.y)), __dmul_rn(d.z, d.z));
        const float B = xmul(2.0f, xadd(xadd(xmul(d.x, m.x), xmul(d.y, m.y)), xmul(d.z, m.z)));
        const float C = (float)__dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(m.x, m.x), __dmul_rn(m.y, m.y)), __dmul_rn(m.z, m.z)), __dmul_rn(cr.w, cr.w));
        const float disc = (float)__dsub_rn(__dmul_rn(B, B), (double)xmul(xmul(4.0f, A), C));
        if (!(disc >= 0.0f))
            continue;
        const float sq = xsqrt(disc), den = xmul(2.0f, A);
        float t0 = xdiv(xadd(-B, sq), den), t1 = xdiv(xsub(-B, sq), den);
        if (t0 < 0.0f)
            t0 = t1;
        if (t1 < 0.0f)
            t1 = t0;
        const float t = (t1 < t0) ? t1 : t0;
        if (!(t > 0.0f) || !(t <= best.t))
            continue;
        const int key = s.tie_by_id ? s.sphere_id_base + k : __ldg(&s.sphere_rank[k]);
        if (t == best.t && key >= best.key)
            continue;
        best.t = t;
        best.key = key;
        best.ti = -2 - k;
        any = true;
    }
    return any;
}

// Prune bound for box tests derived from the best t: t is measured along normalize(d) while the accepted point
// uses the un-normalised d (|d| = 1 +- a few ulp), so leave a relative and an absolute margin.
__device__ __forceinline__ float prune_limit(float best_t) { return best_t * 1.000004f + 1e-5f; }

// Initial HitRec for a fresh closest-hit query: nothing may be accepted at t == FLT_MAX (reference: t < ray.t).
__device__ __forceinline__ HitRec fresh_query() { return HitRec { FLT_MAX, INT_MIN, -1 }; }
// Query bounded by an inclusive limit (shadow rays: blockers have t <= distance - 0.001).
__device__ __forceinline__ HitRec bounded_query(float limit) { return HitRec { limit, INT_MAX, -1 }; }

// Global id of a finished query's winner: triangle id, n_tris + sphere index, or -1.
__device__ __forceinline__ int global_id(const SceneDev& s, const HitRec& best)
{
    if (best.ti >= 0)
        return __float_as_int(__ldg(&s.tri_v2[kTriStride * RT_GUARD(best.ti, s.n_tris, kChkTri)]).w);
    return best.ti == -1 ? -1 : s.sphere_id_base + (-2 - best.ti);
}

// Exhaustive search: every object (order irrelevant thanks to the tie rule).
template <int ANYHIT, bool COUNT>
__device__ __forceinline__ void trace_exhaustive(const SceneDev& s, const f3& o, const f3& d, HitRec& best, TraceStats& st)
{
    const f3 dn = xnormalize(d);
    if (test_spheres(s, o, d, best) && ANYHIT == 1)
        return;
    for (int ti = 0; ti < s.n_tris; ti++) {
        if (test_triangle<COUNT>(s, ti, o, d, dn, best, st) && ANYHIT == 1)
            return;
    }
}

// ---- warp-synchronous BVH traversal engine -----------------------------------------------------------------
// Stack entries (= the `entry` word of a node, rt_types.h): >= 0 -> node index of a sibling pair to fetch;
// < 0 -> ~((first << 3) | (count - 1)), a leaf.  The builders keep the tree depth below kStackDepth (rt_build_bvh
// fails otherwise), so the per-lane stack cannot overflow.
constexpr int kTravDone = INT_MIN;      // not a valid leaf encoding (triangle count is limited to 2^28 - 2)
// ANYHIT template modes of the engine: 0 = closest hit, 1 = any hit (the first blocker found ends the query), 2 = decided per
// query by what it asks for: a bounded_query() is an any-hit query, a fresh_query() a closest-hit one (k_paths mixes both).
constexpr int kAnyHitPerQuery = 2;
#ifndef RT_REFILL_THRESHOLD
#define RT_REFILL_THRESHOLD 0 // measured on B200: refilling only when the whole warp is done beats mid-traversal refills (profiles/README.md)
#endif
#ifndef RT_STACK_TMIN
#define RT_STACK_TMIN 0
#endif
constexpr int kRefillThreshold = RT_REFILL_THRESHOLD; // the traversal loop yields for a refill when fewer lanes than this are still busy

struct Trav {
    f3 o, d, dn;            // ray as stored, and normalize(d)
    // Slab planes are picked by the sign of the direction, not by min/max: with r = 1/dn (clamped, see trav_begin),
    // (nl, nh) = (r, 0) if r >= 0 else (0, r), and t_near = fma(lo, nl, fma(hi, nh, -o*r)), t_far = fma(lo, nh, fma(hi, nl, -o*r)).
    // One of the two products is an exact zero, so these are bit for bit min / max of fma(lo, r, -o*r) and fma(hi, r, -o*r),
    // but they run on the FP32 FMA pipe instead of the half-rate ALU pipe that FMNMX uses (profiles/README.md, r01d).
    float nlx, nly, nlz, nhx, nhy, nhz;
    float oix, oiy, oiz;    // o * r
    HitRec best;
    float tlimit;
    int cur;                // current entry, kTravDone when the traversal is complete
    int sp;
    bool anyhit;            // engines in kAnyHitPerQuery mode: this query ends with the first blocker
};

__device__ __forceinline__ void trav_begin(Trav& tv, const f3& o, const f3& d, const HitRec& query, int root_entry)
{
    tv.o = o;
    tv.d = d;
    tv.dn = xnormalize(d);
    // Reciprocal direction, clamped to +-1e30: with an infinite reciprocal (axis-parallel ray) the fused form
    // box * inf - o * inf is NaN and loses on which side of the origin the slab plane lies; with 1e30 the FMA is exact
    // before rounding, so the signs survive and the slab covers every finite t when the origin is inside it.
    const float rx = fabsf(tv.dn.x) > 1e-30f ? 1.0f / tv.dn.x : copysignf(1e30f, tv.dn.x);
    const float ry = fabsf(tv.dn.y) > 1e-30f ? 1.0f / tv.dn.y : copysignf(1e30f, tv.dn.y);
    const float rz = fabsf(tv.dn.z) > 1e-30f ? 1.0f / tv.dn.z : copysignf(1e30f, tv.dn.z);
    tv.nlx = rx >= 0.0f ? rx : 0.0f; tv.nhx = rx >= 0.0f ? 0.0f : rx;
    tv.nly = ry >= 0.0f ? ry : 0.0f; tv.nhy = ry >= 0.0f ? 0.0f : ry;
    tv.nlz = rz >= 0.0f ? rz : 0.0f; tv.nhz = rz >= 0.0f ? 0.0f : rz;
    tv.oix = o.x * rx;
    tv.oiy = o.y * ry;
    tv.oiz = o.z * rz;
    tv.best = query;
    tv.cur = root_entry;
    tv.sp = 0;
}

// Start a query: the (few) sphere primitives first — their hit bounds the BVH walk, or ends an any-hit query outright.
template <int ANYHIT>
__device__ __forceinline__ void trav_start(const SceneDev& s, Trav& tv, const f3& o, const f3& d, const HitRec& query, int root_entry)
{
    trav_begin(tv, o, d, query, root_entry);
    if (ANYHIT == kAnyHitPerQuery)
        tv.anyhit = query.key == INT_MAX; // bounded_query()
    if (s.n_spheres > 0 && test_spheres(s, o, d, tv.best) && (ANYHIT == 1 || (ANYHIT == kAnyHitPerQuery && tv.anyhit)))
        tv.cur = kTravDone;
    tv.tlimit = prune_limit(tv.best.t);
}

// Per-thread traversal stack.  RT_SMEM_STACK > 0: the first RT_SMEM_STACK entries live in shared memory (one column per thread,
// entry i of thread t at [i][t]: conflict-free), deeper ones spill to local memory; 0: all of it in local memory.
#ifndef RT_TRACE_BLOCK
#define RT_TRACE_BLOCK 128
#endif
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 0
#endif
constexpr int kStackWords = kStackDepth * (RT_STACK_TMIN ? 2 : 1);
struct TravStack {
#if RT_SMEM_STACK > 0
    int* sh; // this thread's column of the block's shared array
    int lo[kStackWords > RT_SMEM_STACK ? kStackWords - RT_SMEM_STACK : 1];
    __device__ __forceinline__ int operator[](int i) const { return i < RT_SMEM_STACK ? sh[i * RT_TRACE_BLOCK] : lo[i - RT_SMEM_STACK]; }
    __device__ __forceinline__ void put(int i, int v)
    {
        if (i < RT_SMEM_STACK)
            sh[i * RT_TRACE_BLOCK] = v;
        else
            lo[i - RT_SMEM_STACK] = v;
    }
#else
    int lo[kStackWords];
    __device__ __forceinline__ int operator[](int i) const { return lo[RT_GUARD(i, kStackWords, kChkStack)]; }
    __device__ __forceinline__ void put(int i, int v) { lo[RT_GUARD(i, kStackWords, kChkStack)] = v; }
#endif
};
// Binds the shared part of the stack; must be called by every thread of a block of at most RT_TRACE_BLOCK threads.
__device__ __forceinline__ void trav_stack_init(TravStack& st)
{
#if RT_SMEM_STACK > 0
    __shared__ int s_stack[RT_SMEM_STACK][RT_TRACE_BLOCK];
    st.sh = &s_stack[0][threadIdx.x];
#endif
}

// Fetch of one sibling pair = 64 aligned bytes.  RT_LDG256: two 256-bit loads (LDG.E.ENL2.256, sm_100) instead of four 128-bit ones.
#ifndef RT_LDG256
#define RT_LDG256 0
#endif
__device__ __forceinline__ void load_node_pair(const float4* np, float4& a0, float4& a1, float4& b0, float4& b1)
{
#if RT_LDG256
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a0.x), "=f"(a0.y), "=f"(a0.z), "=f"(a0.w), "=f"(a1.x), "=f"(a1.y), "=f"(a1.z), "=f"(a1.w)
                 : "l"(np));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w), "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w)
                 : "l"(np + 2));
#else
    a0 = __ldg(np);
    a1 = __ldg(np + 1);
    b0 = __ldg(np + 2);
    b1 = __ldg(np + 3);
#endif
}

// Pop the next entry; with RT_STACK_TMIN the entry distance saved at push time lets entries that have fallen behind the
// current best hit be skipped without fetching them.
__device__ __forceinline__ int trav_pop(Trav& tv, const TravStack& stack)
{
#if RT_STACK_TMIN
    while (tv.sp) {
        --tv.sp;
        if (__int_as_float(stack[2 * tv.sp + 1]) <= tv.tlimit)
            return stack[2 * tv.sp];
    }
    return kTravDone;
#else
    return tv.sp ? stack[--tv.sp] : kTravDone;
#endif
}

#ifndef RT_BVH_WIDE
#define RT_BVH_WIDE 0
#endif

#if RT_BVH_WIDE
// One node step of the 4-wide tree (rt_wide.cu): fetch the 128-byte node `tv.cur`, slab-test its four boxes, go on with the
// nearest hit child and push the others, farthest first; or pop.  An any-hit query does not care about the order.
template <int ANYHIT, bool COUNT>
__device__ __forceinline__ void trav_node_step(const SceneDev& s, Trav& tv, TravStack& stack, TraceStats& st)
{
    const float4* np = s.nodes + 8 * (size_t)tv.cur;
    const float4 lx = __ldg(np), ly = __ldg(np + 1), lz = __ldg(np + 2), hx = __ldg(np + 3), hy = __ldg(np + 4), hz = __ldg(np + 5);
    const float4 eb = __ldg(np + 6);
    if (COUNT)
        st.nodes += 4;
    float key[4];
    int e[4] = { __float_as_int(eb.x), __float_as_int(eb.y), __float_as_int(eb.z), __float_as_int(eb.w) };
    int n = 0;
#define RT_SLAB4(k, c)                                                                                                                  \
    {                                                                                                                                   \
        const float nx = fmaf(lx.c, tv.nlx, fmaf(hx.c, tv.nhx, -tv.oix)), fx = fmaf(lx.c, tv.nhx, fmaf(hx.c, tv.nlx, -tv.oix));           \
        const float ny = fmaf(ly.c, tv.nly, fmaf(hy.c, tv.nhy, -tv.oiy)), fy = fmaf(ly.c, tv.nhy, fmaf(hy.c, tv.nly, -tv.oiy));           \
        const float nz = fmaf(lz.c, tv.nlz, fmaf(hz.c, tv.nhz, -tv.oiz)), fz = fmaf(lz.c, tv.nhz, fmaf(hz.c, tv.nlz, -tv.oiz));           \
        const float tn = fmaxf(fmaxf(fmaxf(nx, ny), nz), 0.0f), tf = fminf(fminf(fminf(fx, fy), fz), tv.tlimit);                        \
        const bool hit = tn <= tf * 1.0000005f;                                                                                         \
        key[k] = hit ? tn : INFINITY;                                                                                                   \
        n += hit ? 1 : 0;                                                                                                               \
    }
    RT_SLAB4(0, x)
    RT_SLAB4(1, y)
    RT_SLAB4(2, z)
    RT_SLAB4(3, w)
#undef RT_SLAB4
    if (n == 0) {
        tv.cur = trav_pop(tv, stack);
        return;
    }
    // sort the four (key, entry) pairs by key; missed boxes carry +inf and end up behind the hits
#define RT_CSWAP(a, b)                          \
    {                                           \
        const bool sw = key[b] < key[a];        \
        const float ka = key[a], kb = key[b];   \
        const int ea = e[a], eb_ = e[b];        \
        key[a] = sw ? kb : ka;                  \
        key[b] = sw ? ka : kb;                  \
        e[a] = sw ? eb_ : ea;                   \
        e[b] = sw ? ea : eb_;                   \
    }
    RT_CSWAP(0, 1)
    RT_CSWAP(2, 3)
    RT_CSWAP(0, 2)
    RT_CSWAP(1, 3)
    RT_CSWAP(1, 2)
#undef RT_CSWAP
    if (n > 3)
        stack.put(tv.sp++, e[3]);
    if (n > 2)
        stack.put(tv.sp++, e[2]);
    if (n > 1)
        stack.put(tv.sp++, e[1]);
    tv.cur = e[0];
}
#else
// One node step: fetch the sibling pair `tv.cur` (one aligned 64-byte read), slab-test both boxes, descend into the
// nearer hit child and push the other one, or pop.
template <int ANYHIT, bool COUNT>
__device__ __forceinline__ void trav_node_step(const SceneDev& s, Trav& tv, TravStack& stack, TraceStats& st)
{
    const float4* np = s.nodes + 2 * (size_t)RT_GUARD(tv.cur, s.n_nodes - 1, kChkNode); // (a fetch reads nodes cur and cur + 1)
    float4 a0, a1, b0, b1;
    load_node_pair(np, a0, a1, b0, b1);
    if (COUNT)
        st.nodes += 2;
    // near / far slab distances of both boxes (all finite: boxes are finite, |r| <= 1e30)
    const float anx = fmaf(a0.x, tv.nlx, fmaf(a1.x, tv.nhx, -tv.oix)), afx = fmaf(a0.x, tv.nhx, fmaf(a1.x, tv.nlx, -tv.oix));
    const float any_ = fmaf(a0.y, tv.nly, fmaf(a1.y, tv.nhy, -tv.oiy)), afy = fmaf(a0.y, tv.nhy, fmaf(a1.y, tv.nly, -tv.oiy));
    const float anz = fmaf(a0.z, tv.nlz, fmaf(a1.z, tv.nhz, -tv.oiz)), afz = fmaf(a0.z, tv.nhz, fmaf(a1.z, tv.nlz, -tv.oiz));
    const float bnx = fmaf(b0.x, tv.nlx, fmaf(b1.x, tv.nhx, -tv.oix)), bfx = fmaf(b0.x, tv.nhx, fmaf(b1.x, tv.nlx, -tv.oix));
    const float bny = fmaf(b0.y, tv.nly, fmaf(b1.y, tv.nhy, -tv.oiy)), bfy = fmaf(b0.y, tv.nhy, fmaf(b1.y, tv.nly, -tv.oiy));
    const float bnz = fmaf(b0.z, tv.nlz, fmaf(b1.z, tv.nhz, -tv.oiz)), bfz = fmaf(b0.z, tv.nhz, fmaf(b1.z, tv.nlz, -tv.oiz));
    const float amin = fmaxf(fmaxf(fmaxf(anx, any_), anz), 0.0f), amax = fminf(fminf(fminf(afx, afy), afz), tv.tlimit);
    const float bmin = fmaxf(fmaxf(fmaxf(bnx, bny), bnz), 0.0f), bmax = fminf(fminf(fminf(bfx, bfy), bfz), tv.tlimit);
    const bool hitA = amin <= amax * 1.0000005f;
    const bool hitB = bmin <= bmax * 1.0000005f;
    const int ea = __float_as_int(a0.w), eb = __float_as_int(b0.w);
    if (hitA && hitB) {
        const bool aFirst = amin <= bmin;
#if RT_STACK_TMIN
        stack.put(2 * tv.sp, aFirst ? eb : ea);
        stack.put(2 * tv.sp + 1, __float_as_int(aFirst ? bmin : amin));
        tv.sp++;
#else
        stack.put(tv.sp++, aFirst ? eb : ea);
#endif
        tv.cur = aFirst ? ea : eb;
    } else if (hitA) {
        tv.cur = ea;
    } else if (hitB) {
        tv.cur = eb;
    } else {
        tv.cur = trav_pop(tv, stack);
    }
}
#endif

// Test the triangles of leaf entry `leaf`; the walk state (tv.cur, stack) is untouched unless an any-hit query is
// satisfied, which drops all remaining work.
template <int ANYHIT, bool COUNT>
__device__ __forceinline__ void trav_leaf_test(const SceneDev& s, Trav& tv, int leaf, TraceStats& st)
{
    const int enc = ~leaf;
    const int first = enc >> 3, count = (enc & 7) + 1;
    bool any = false;
#if RT_LEAF_PREFETCH
    float4 pl = __ldg(&s.tri_plane[kTriStride * first]);
    for (int i = 0; i < count; i++) {
        const float4 cur = pl;
        if (i + 1 < count)
            pl = __ldg(&s.tri_plane[kTriStride * (first + i + 1)]);
        any |= test_triangle<COUNT>(s, first + i, cur, tv.o, tv.d, tv.dn, tv.best, st);
    }
#else
    for (int i = 0; i < count; i++)
        any |= test_triangle<COUNT>(s, first + i, tv.o, tv.d, tv.dn, tv.best, st);
#endif
    if (any) {
        tv.tlimit = prune_limit(tv.best.t);
        if (ANYHIT == 1 || (ANYHIT == kAnyHitPerQuery && tv.anyhit)) { // the first blocker decides
            tv.cur = kTravDone;
            tv.sp = 0;
        }
    }
}

// Persistent-warp traversal of a work queue (speculative "while-while" with per-lane refill, after Aila & Laine,
// "Understanding the Efficiency of Ray Traversal on GPUs", HPG 2009).  Every lane walks inner nodes until it reaches a
// leaf, then tests the leaf's triangles, and repeats; when kRefillThreshold > 0 and fewer lanes than that are still inside
// the loop, the warp leaves it and the idle lanes pull new items (one atomicAdd per warp per refill).
// fetch(item, o, d, query) -> false if the item needs no ray;  finish(item, best, o, d, query) -> true to continue the
// same item with a new segment (shadow rays passing a transparent surface; the next query of a path, k_paths); o, d hold the
// finished query's ray on entry.
// STATIC: no work cursor — warp w takes items [(k * total_warps + w) * 32, + 32) for k = 0, 1, ...; only for queues whose items never
// continue (finish() returns false), so that a warp refills with all 32 lanes idle.  Level-0 extend uses it: a 4K frame would
// otherwise issue 260 K same-address atomics, which L2 serialises at about 1 ns each.
template <int ANYHIT, bool COUNT, bool STATIC = false, typename Fetch, typename Finish>
__device__ __forceinline__ void trace_queue(const SceneDev& s, int root_entry, bool exhaustive, unsigned* cursor, unsigned n_items,
    TraceStats& st, Fetch fetch, Finish finish, int max_quota = 32, int min_quota = 1)
{
    constexpr unsigned kFullMask = 0xffffffffu;
    TravStack stack;
    trav_stack_init(stack);
    Trav tv;
    tv.cur = kTravDone;
    tv.sp = 0;
    bool active = false, more = n_items > 0;
    unsigned my = 0;
    const int lane = threadIdx.x & 31;
    // Small queues (deep bounce levels, or one GPU's share of a frame split over many GPUs) cannot fill every lane of
    // every resident warp; spreading them thinly over ALL warps (fewer rays per warp, more warps busy) hides the
    // node-fetch latency far better than packing 32 rays into a few warps and leaving most SM warp slots empty.
    const unsigned total_warps = gridDim.x * (blockDim.x >> 5);
#ifndef RT_MAX_QUOTA
#define RT_MAX_QUOTA 32
#endif
    const int quota = STATIC ? 32 : (int)min((unsigned)max_quota, max((unsigned)min_quota, (n_items + total_warps - 1) / total_warps));
    // RT_LANE_DEPTH > 1: when the queue holds plenty of work a warp claims that many items per lane at once; a lane whose query
    // ends starts on its next item at once instead of idling until the slowest lane of the warp is done.
#ifndef RT_LANE_DEPTH
#define RT_LANE_DEPTH 1
#endif
    constexpr int kLaneDepth = STATIC ? 1 : RT_LANE_DEPTH;
    const bool deep_claims = kLaneDepth > 1 && !exhaustive && n_items >= (unsigned)kLaneDepth * 32u * total_warps;
    unsigned nxt = 0, nxt_end = 0, nxt_stride = 0; // the lane's next claimed item, end of the warp's claim, distance between a lane's items
    unsigned static_next = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
    if (STATIC)
        more = static_next < n_items;
    for (;;) {
        // ---- refill: idle lanes take the next unclaimed items (at most `quota` rays in flight per warp) ----
        const unsigned idle_all = __ballot_sync(kFullMask, !active);
        const int room = quota - (32 - __popc(idle_all));
        // the `room` lowest idle lanes refill
        const unsigned idle = __ballot_sync(kFullMask, !active && __popc(idle_all & ((1u << lane) - 1u)) < room);
        if (idle && more) {
            const int leader = __ffs(idle) - 1;
            const unsigned cnt = (unsigned)__popc(idle);
            unsigned base = 0;
            if (STATIC) {
                base = static_next;
                static_next += total_warps * 32u;
                more = static_next < n_items;
            } else {
                const unsigned claim = deep_claims && cnt == 32u ? cnt * (unsigned)kLaneDepth : cnt;
                if (lane == leader)
                    base = atomicAdd(cursor, claim);
                base = __shfl_sync(kFullMask, base, leader);
                more = base + claim < n_items;
                if (kLaneDepth > 1) {
                    nxt_stride = cnt;
                    nxt_end = min(base + claim, n_items);
                    nxt = base + cnt + (unsigned)__popc(idle & ((1u << lane) - 1u));
                }
            }
            if (!active && ((idle >> lane) & 1u)) {
                const unsigned item = base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (item < n_items) {
                    f3 o, d;
                    HitRec q;
                    if (fetch(item, o, d, q)) {
                        my = item;
                        active = true;
                        if (exhaustive) { // reference useBVH=false semantics (tests): loop over every primitive
                            bool again;
                            do {
                                trav_begin(tv, o, d, q, root_entry);
                                trace_exhaustive<ANYHIT, COUNT>(s, tv.o, tv.d, tv.best, st);
                                again = finish(my, tv.best, o, d, q);
                            } while (again);
                            active = false;
                            tv.cur = kTravDone;
                        } else {
                            if (COUNT)
                                st.query_begins();
                            trav_start<ANYHIT>(s, tv, o, d, q, root_entry);
                        }
                    }
                }
            }
        }
        if (idle_all == kFullMask && !__any_sync(kFullMask, active)) { // nothing in flight: stop, or fetch again
            if (!more)
                break;
            continue;
        }
        // ---- traverse ----
        // the lane's query is complete: hand it in; true if the lane goes on (its item continues with another segment, or the next
        // item of its claim starts)
        auto complete = [&]() -> bool {
            f3 o = tv.o, d = tv.d; // in: the ray of the finished query; out: the next segment when finish() asks for one
            HitRec q;
            if (COUNT)
                st.query_ends();
            if (finish(my, tv.best, o, d, q)) {
                if (COUNT)
                    st.query_begins();
                trav_start<ANYHIT>(s, tv, o, d, q, root_entry);
                return true;
            }
            if (kLaneDepth > 1) {
                while (nxt < nxt_end) {
                    const unsigned item = nxt;
                    nxt += nxt_stride;
                    if (fetch(item, o, d, q)) {
                        my = item;
                        if (COUNT)
                            st.query_begins();
                        trav_start<ANYHIT>(s, tv, o, d, q, root_entry);
                        return true;
                    }
                }
            }
            return false;
        };
        if (active) {
            const int min_active = more ? kRefillThreshold : 0;
            // Speculative while-while (Aila & Laine): a lane that reaches a leaf parks it and keeps walking nodes until
            // every lane of the warp that is still walking has found a leaf too; then all parked leaves are tested
            // together.  Parking only delays pruning; the (t, id) tie rule makes the result order independent.
            int parked = kTravDone;
            while (tv.cur != kTravDone || parked != kTravDone) {
                bool searching = parked == kTravDone;
                while (tv.cur >= 0) {
                    trav_node_step<ANYHIT, COUNT>(s, tv, stack, st);
                    if (tv.cur < 0 && tv.cur != kTravDone && parked == kTravDone) {
                        parked = tv.cur;
                        tv.cur = trav_pop(tv, stack);
                        searching = false;
                    }
                    if (!__any_sync(__activemask(), searching))
                        break;
                }
                if (tv.cur < 0 && tv.cur != kTravDone && parked == kTravDone) { // the entry itself was a leaf
                    parked = tv.cur;
                    tv.cur = trav_pop(tv, stack);
                }
                while (parked != kTravDone) {
                    trav_leaf_test<ANYHIT, COUNT>(s, tv, parked, st);
                    parked = kTravDone;
                    if (tv.cur < 0 && tv.cur != kTravDone) { // the next entry is a leaf as well: test it right away
                        parked = tv.cur;
                        tv.cur = trav_pop(tv, stack);
                    }
                }
                if (kLaneDepth > 1 && tv.cur == kTravDone && !complete()) {
                    active = false;
                    break;
                }
                if (min_active > 0 && __popc(__activemask()) < min_active)
                    break;
            }
            if (active && tv.cur == kTravDone && !complete())
                active = false;
        }
        __syncwarp();
    }
}

// One-shot traversal of a single ray per thread (rt_intersect); lanes are independent here.
template <int ANYHIT, bool COUNT>
__device__ __forceinline__ void trace_bvh(const SceneDev& s, int root_entry, const f3& o, const f3& d, HitRec& best, TraceStats& st)
{
    TravStack stack;
    trav_stack_init(stack);
    Trav tv;
    trav_start<ANYHIT>(s, tv, o, d, best, root_entry);
    while (tv.cur != kTravDone) {
        if (tv.cur >= 0) {
            trav_node_step<ANYHIT, COUNT>(s, tv, stack, st);
        } else {
            const int leaf = tv.cur;
            tv.cur = trav_pop(tv, stack);
            trav_leaf_test<ANYHIT, COUNT>(s, tv, leaf, st);
        }
    }
    best = tv.best;
}

} // namespace rtb
