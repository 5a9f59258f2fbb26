// Wavefront kernels of the per-pixel hot path (sm_100a): K0 tri_setup, K1 generate, K2 extend, K3 shade,
// K4 shadow (point / spherical light), K5 resolve.  See DESIGN.md for the data flow; every kernel cites the
// reference lines it replaces.
#include "rt_kernels.h"
#include "rt_trace.cuh"

namespace rtb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Warp-aggregated dynamic work fetch: one atomicAdd per warp hands out 32 consecutive items.
__device__ __forceinline__ unsigned warp_fetch(unsigned* cursor)
{
    unsigned base = 0;
    if (lane_id() == 0)
        base = atomicAdd(cursor, 32u);
    return __shfl_sync(kFull, base, 0);
}

// Warp-ballot compaction: lanes with `want` get consecutive slots of a queue; one atomicAdd per warp.
// Must be called by all 32 lanes.  Returns the slot or 0xffffffff (not wanted / queue full -> overflow flagged).
__device__ __forceinline__ unsigned warp_alloc(unsigned* counter, bool want, unsigned capacity, unsigned* overflow)
{
    const unsigned mask = __ballot_sync(kFull, want);
    if (mask == 0)
        return 0xffffffffu;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane_id() == leader)
        base = atomicAdd(counter, (unsigned)__popc(mask));
    base = __shfl_sync(kFull, base, leader);
    if (!want)
        return 0xffffffffu;
    const unsigned slot = base + __popc(mask & ((1u << lane_id()) - 1u));
    if (slot >= capacity) {
        atomicExch(overflow, 1u);
        return 0xffffffffu;
    }
    return slot;
}

__device__ __forceinline__ void warp_add_u64(unsigned long long* dst, unsigned v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFull, v, o);
    if (lane_id() == 0 && v)
        atomicAdd(dst, (unsigned long long)v);
}

// local padded pixel index -> pixel coordinates (tile interleaving across ranks, 8x4 warp blocks inside a tile)
__device__ __forceinline__ void local_to_pixel(const FrameParams& fp, unsigned lp, int& px, int& py)
{
    const unsigned j = lp / kTilePixels, k = lp % kTilePixels;
    const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world; // global tile id
    const unsigned tx = g % (unsigned)fp.tiles_x, ty = g / (unsigned)fp.tiles_x;
    const unsigned w = k >> 5, l = k & 31;
    px = (int)(tx * kTileW + (w & 3) * 8 + (l & 7));
    py = (int)(ty * kTileH + (w >> 2) * 4 + (l >> 3));
}

struct Shading { // what shade and the transparent-shadow branch need at a hit
    f3 p;      // hit point
    f3 N;      // interpolated normal, flipped to the geometric side (not normalised)
    int mesh;
};

// Hit point and shading normal: src/ray_tracing.cpp:111, 147-160 with barycentricCoordinates (276-308) evaluated
// unconditionally ("defined barycentrics": the reference leaves barCoords uninitialised when its epsilon tests
// fail).  `o`, `d` are the ray as traced, t its hit parameter.
__device__ __forceinline__ Shading shading_at(const SceneDev& s, int ti, const f3& o, const f3& d, float t)
{
    const float4 pl = __ldg(&s.tri_plane[ti]);
    const float4 a = __ldg(&s.tri_v0[ti]), b = __ldg(&s.tri_v1[ti]), c = __ldg(&s.tri_v2[ti]);
    const f3 v0 = mk3(a), v1 = mk3(b), v2 = mk3(c), fn = mk3(pl);
    Shading sh;
    sh.mesh = __float_as_int(b.w);
    sh.p = xadd(o, xmul(d, t));
    const float total = xlength(xcross(xsub(v1, v0), xsub(v2, v0)));
    const float c0 = xdiv(xlength(xcross(xsub(v1, sh.p), xsub(v2, sh.p))), total);
    const float c1 = xdiv(xlength(xcross(xsub(sh.p, v0), xsub(v2, v0))), total);
    const float c2 = xdiv(xlength(xcross(xsub(v1, v0), xsub(sh.p, v0))), total);
    const f3 n0 = mk3(__ldg(&s.tri_n0[ti])), n1 = mk3(__ldg(&s.tri_n1[ti])), n2 = mk3(__ldg(&s.tri_n2[ti]));
    f3 N = xadd(xadd(xmul(n0, c0), xmul(n1, c1)), xmul(n2, c2));
    if (xdot(N, fn) < 0.0f)
        N = xneg(N);
    sh.N = N;
    return sh;
}

// Schlick term of the reference, evaluated in double like `R0 + (1 - R0) * std::pow(1 - c, 5)` with float c, R0
// (src/main.cpp:279, src/shadow.cpp:59): pow(float,int) promotes to double.
__device__ __forceinline__ double schlick(float R0, float c)
{
    const double omc = (double)(1.0f - c);
    const double p5 = omc * omc * omc * omc * omc;
    return (double)R0 + (double)(1.0f - R0) * p5;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// K0 tri_setup: trianglePlane (src/ray_tracing.cpp:91-100) once per triangle instead of once per test, and the
// scatter of the triangle soup into BVH leaf order.  perm[i] = global id of the triangle stored at slot i.
__global__ void k_tri_setup(const float* __restrict__ pos, const float* __restrict__ nrm, const int* __restrict__ mesh_id,
    const int* __restrict__ perm, int n, float4* plane, float4* v0o, float4* v1o, float4* v2o, float4* n0o, float4* n1o, float4* n2o)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const int g = perm ? perm[i] : i;
    const float* p = pos + 9 * (size_t)g;
    const f3 v0 = mk3(p[0], p[1], p[2]), v1 = mk3(p[3], p[4], p[5]), v2 = mk3(p[6], p[7], p[8]);
    const f3 nn = xnormalize(xcross(xsub(v0, v2), xsub(v1, v2)));
    const float D = xdot(nn, v0);
    plane[i] = make_float4(nn.x, nn.y, nn.z, D);
    v0o[i] = make_float4(v0.x, v0.y, v0.z, __int_as_float(g));
    v1o[i] = make_float4(v1.x, v1.y, v1.z, __int_as_float(mesh_id ? mesh_id[g] : 0));
    v2o[i] = make_float4(v2.x, v2.y, v2.z, 0.0f);
    const float* q = nrm + 9 * (size_t)g;
    n0o[i] = make_float4(q[0], q[1], q[2], 0.0f);
    n1o[i] = make_float4(q[3], q[4], q[5], 0.0f);
    n2o[i] = make_float4(q[6], q[7], q[8], 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// Reset of the per-level cursors / queue fills (one thread).
__global__ void k_level_reset(Counters* c, int next_q, int clear_current)
{
    c->work[0] = c->work[1] = c->work[2] = c->work[3] = 0;
    c->n_shadow_pt = 0;
    c->n_shadow_sp = 0;
    c->n_rays[next_q] = 0;
    if (clear_current)
        c->n_rays[next_q ^ 1] = 0;
}

// ---------------------------------------------------------------------------------------------------------
// K1 generate: pixel -> NDC (src/main.cpp:350-353), sub-pixel sample positions (358-375 / 309-335, 377-385) and
// Trackball::generateRay (framework/src/trackball.cpp:87-98).  One thread per (pixel, sample) of the batch.
__global__ void __launch_bounds__(256) k_generate(FrameParams fp, BatchDev b, unsigned first_lp, unsigned n_lp, int qi)
{
    const unsigned n = n_lp * (unsigned)fp.spp;
    const unsigned n_round = (n + 31u) & ~31u;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += gridDim.x * blockDim.x) {
        bool valid = idx < n;
        unsigned lp = 0;
        int px = 0, py = 0, sidx = 0;
        if (valid) {
            lp = first_lp + idx / (unsigned)fp.spp;
            sidx = (int)(idx % (unsigned)fp.spp);
            local_to_pixel(fp, lp, px, py);
            valid = px < fp.W && py < fp.H;
        }
        f3 dir = mk3(0.f, 0.f, 0.f);
        if (valid) {
            float nx = xsub(xmul(xdiv((float)px, (float)fp.W), 2.0f), 1.0f);
            float ny = xsub(xmul(xdiv((float)py, (float)fp.H), 2.0f), 1.0f);
            if (fp.sample_mode == 1) {
                nx = (sidx & 1) ? xadd(nx, fp.aa_off_x) : xsub(nx, fp.aa_off_x);
                ny = (sidx & 2) ? xsub(ny, fp.aa_off_y) : xadd(ny, fp.aa_off_y);
            } else if (fp.sample_mode == 2) {
                const int k = (fp.ms_moves + 1) / 2; // odd steps 1,3,.. <= moves
                const int quad = sidx / (k * k), rem = sidx % (k * k);
                const int sx = 1 + 2 * (rem / k), sy = 1 + 2 * (rem % k);
                const float sgx = (quad & 1) ? 1.0f : -1.0f, sgy = (quad & 2) ? -1.0f : 1.0f;
                nx = xadd(nx, xmul(xmul(fp.ms_off_x, sgx), (float)sx));
                ny = xadd(ny, xmul(xmul(fp.ms_off_y, sgy), (float)sy));
            }
            const f3 cam = xnormalize(mk3(xmul(-nx, fp.halfW), xmul(ny, fp.halfH), 1.0f));
            dir = xquat_rotate(mk3(fp.qx, fp.qy, fp.qz), fp.qw, cam);
        }
        const unsigned slot = warp_alloc(&b.counters->n_rays[qi], valid, b.ray_capacity, &b.counters->overflow);
        if (slot != 0xffffffffu) {
            b.q[qi].o_pix[slot] = make_float4(fp.ox, fp.oy, fp.oz, __int_as_float((int)((lp << 1) | (sidx == 0 ? 1u : 0u))));
            b.q[qi].d[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
            b.q[qi].w[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        }
        warp_add_u64(&b.counters->primary_rays, valid ? 1u : 0u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2 extend: closest hit of every ray in queue qi (BoundingVolumeHierarchy::intersect,
// src/bounding_volume_hierarchy.cpp:49-78).  Persistent warps pull 32 rays at a time.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_extend(SceneDev s, int root_entry, FrameParams fp, BatchDev b, int qi, int level)
{
    const unsigned n = b.counters->n_rays[qi];
    TraceStats st;
    while (true) {
        const unsigned base = warp_fetch(&b.counters->work[0]);
        if (base >= n)
            break;
        const unsigned i = base + lane_id();
        if (i < n) {
            const float4 op = b.q[qi].o_pix[i];
            const float4 dd = b.q[qi].d[i];
            const f3 o = mk3(op), d = mk3(dd);
            HitRec best = fresh_query();
            bool ok = true;
            if (fp.exhaustive)
                trace_exhaustive<false, COUNT>(s, o, d, best, st);
            else
                ok = trace_bvh<false, COUNT>(s, root_entry, o, d, best, st);
            if (!ok)
                atomicExch(&b.counters->overflow, 2u);
            b.q[qi].hit[i] = make_int2(__float_as_int(best.t), best.ti);
            if (level == 0 && b.prim_id) {
                const int tag = __float_as_int(op.w);
                if (tag & 1) { // first sample of the pixel
                    b.prim_id[tag >> 1] = best.ti >= 0 ? best.id : -1;
                    b.prim_t[tag >> 1] = best.t;
                }
            }
        }
        __syncwarp();
    }
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
        warp_add_u64(&b.counters->ext_node_visits, st.nodes);
        warp_add_u64(&b.counters->ext_tri_tests, st.tris);
        warp_add_u64(&b.counters->ext_tri_tests_full, st.tris_full);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3 shade: getFinalColor without the recursion (src/main.cpp:129-190), light set-up of getPointLights /
// getSpherelights (src/shadow.cpp:106-131, 139-226: everything except the cansee calls), calcColor
// (src/main.cpp:112-121) folded into per-light coefficients, and the spawn of mirror (191-256 with
// glossy_ray_count == 1 => weight ks*ks) and dielectric (257-290) children with a throughput instead of recursion.
__global__ void __launch_bounds__(128) k_shade(SceneDev s, FrameParams fp, BatchDev b, int qi, int level)
{
    const unsigned n = b.counters->n_rays[qi];
    const int qo = qi ^ 1;
    while (true) {
        const unsigned base = warp_fetch(&b.counters->work[1]);
        if (base >= n)
            break;
        const unsigned i = base + lane_id();
        bool hit = false;
        int pix = 0;
        f3 w = mk3(0, 0, 0), refl = mk3(0, 0, 0), refr = mk3(0, 0, 0), dn = mk3(0, 0, 0), Nn = mk3(0, 0, 0);
        Shading sh;
        sh.p = mk3(0, 0, 0);
        sh.N = mk3(0, 0, 0);
        sh.mesh = 0;
        float4 m0 = make_float4(0, 0, 0, 0), m1 = make_float4(0, 0, 0, 1);
        if (i < n) {
            const int2 h = b.q[qi].hit[i];
            if (h.y >= 0) {
                hit = true;
                const float4 op = b.q[qi].o_pix[i];
                const float4 dd = b.q[qi].d[i];
                const float4 ww = b.q[qi].w[i];
                pix = __float_as_int(op.w) >> 1;
                w = mk3(ww);
                const f3 o = mk3(op), d = mk3(dd);
                sh = shading_at(s, h.y, o, d, __int_as_float(h.x));
                dn = xnormalize(d);
                Nn = xnormalize(sh.N);
                refl = xreflect(dn, Nn); // main.cpp:141
                m0 = __ldg(&s.mats[2 * sh.mesh]);
                m1 = __ldg(&s.mats[2 * sh.mesh + 1]);
            }
        }
        const f3 kd = mk3(m0), ks = mk3(m1);
        const float shininess = m0.w, transparency = m1.w;

        // direct light: one record per light; the shadow kernels add A * intensity + B when the light is visible
        const int n_lights = fp.n_point + fp.n_sphere;
        for (int li = 0; li < n_lights; li++) {
            const bool is_point = li < fp.n_point;
            const float4* L = is_point ? (s.point_lights + 2 * li) : (s.sphere_lights + 2 * (li - fp.n_point));
            f3 A = mk3(0, 0, 0), B = mk3(0, 0, 0);
            if (hit) {
                const f3 lp = mk3(__ldg(L)), lc = mk3(__ldg(L + 1));
                const f3 ldir = xnormalize(xsub(lp, sh.p));
                const float cosNL = fabsf(xdot(Nn, ldir));                         // shadow.cpp:125 / 218
                const float cosRL = fmaxf(0.0f, xdot(xnormalize(refl), ldir));      // shadow.cpp:126 / 219
                A = mk3(w.x * kd.x * lc.x * cosNL, w.y * kd.y * lc.y * cosNL, w.z * kd.z * lc.z * cosNL);
                if (shininess > 0.0f) {
                    const float sp = powf(cosRL, shininess);
                    B = mk3(w.x * lc.x * ks.x * sp, w.y * lc.y * ks.y * sp, w.z * lc.z * ks.z * sp);
                }
            }
            ShadowQueue& sq = is_point ? b.sq_point : b.sq_sphere;
            unsigned* cnt = is_point ? &b.counters->n_shadow_pt : &b.counters->n_shadow_sp;
            const unsigned cap = is_point ? b.shadow_pt_capacity : b.shadow_sp_capacity;
            const unsigned slot = warp_alloc(cnt, hit, cap, &b.counters->overflow);
            if (slot != 0xffffffffu) {
                sq.p_pix[slot] = make_float4(sh.p.x, sh.p.y, sh.p.z, __int_as_float(pix));
                sq.a_light[slot] = make_float4(A.x, A.y, A.z, __int_as_float(is_point ? li : li - fp.n_point));
                sq.b[slot] = make_float4(B.x, B.y, B.z, 0.0f);
            }
        }

        // children
        bool want0 = false, want1 = false;
        f3 w0 = mk3(0, 0, 0), w1 = mk3(0, 0, 0);
        if (hit && level < fp.max_level) { // main.cpp:187
            if (transparency == 1.0f) {
                if (ks.x > 0.0f || ks.y > 0.0f || ks.z > 0.0f) { // main.cpp:194
                    want0 = true;
                    w0 = mk3(w.x * ks.x * ks.x, w.y * ks.y * ks.y, w.z * ks.z * ks.z);
                }
            } else { // main.cpp:257-290
                const float r = fp.refraction;
                const float c = fabsf(xdot(dn, Nn));
                const float k = xmul(xmul(r, r), xsub(1.0f, xmul(c, c)));
                const float coef = xsub(xmul(r, c), xsqrt(xsub(1.0f, k)));
                refr = xnormalize(xadd(xmul(dn, r), xmul(Nn, coef)));
                const float R = (float)schlick(transparency, c);
                want0 = true;
                w0 = mk3(w.x * R, w.y * R, w.z * R);
                if (k <= 1.0f) {
                    want1 = true;
                    const float T = 1.0f - R;
                    w1 = mk3(w.x * T, w.y * T, w.z * T);
                }
            }
        }
        unsigned slot = warp_alloc(&b.counters->n_rays[qo], want0, b.ray_capacity, &b.counters->overflow);
        if (slot != 0xffffffffu) {
            const f3 o2 = xadd(sh.p, xmul(refl, 0.01f)); // main.cpp:199,286
            b.q[qo].o_pix[slot] = make_float4(o2.x, o2.y, o2.z, __int_as_float(pix << 1));
            b.q[qo].d[slot] = make_float4(refl.x, refl.y, refl.z, 0.0f);
            b.q[qo].w[slot] = make_float4(w0.x, w0.y, w0.z, 0.0f);
        }
        slot = warp_alloc(&b.counters->n_rays[qo], want1, b.ray_capacity, &b.counters->overflow);
        if (slot != 0xffffffffu) {
            const f3 o2 = xadd(sh.p, xmul(refr, 0.01f)); // main.cpp:288
            b.q[qo].o_pix[slot] = make_float4(o2.x, o2.y, o2.z, __int_as_float(pix << 1));
            b.q[qo].d[slot] = make_float4(refr.x, refr.y, refr.z, 0.0f);
            b.q[qo].w[slot] = make_float4(w1.x, w1.y, w1.z, 0.0f);
        }
        warp_add_u64(&b.counters->secondary_rays, (want0 ? 1u : 0u) + (want1 ? 1u : 0u));
    }
}

// ---------------------------------------------------------------------------------------------------------
// cansee (src/shadow.cpp:32-69): closest-hit loop from p1 towards p2 with transparent pass-through.  Returns
// visibility; `intensity` is attenuated by every transparent surface crossed (also when finally blocked, which
// getSpherelights relies on for its centre sample).  `queries` counts loop iterations.
template <bool COUNT>
__device__ __forceinline__ bool cansee(const SceneDev& s, int root_entry, const FrameParams& fp, const f3& p1, const f3& p2,
    float& intensity, unsigned& queries, TraceStats& st, bool& ok)
{
    f3 d = xsub(p2, p1);
    float distance = xlength(d);
    d = xnormalize(d);
    f3 o = xadd(p1, xmul(d, 0.0005f));
    while (distance > 0.0005f) {
        queries++;
        // blockers have t <= distance - 2*SHADOW_ERROR_OFFSET; anything farther means "visible" (shadow.cpp:44)
        HitRec best = bounded_query(xsub(distance, 0.001f));
        if (fp.exhaustive) {
            if (fp.any_transparent)
                trace_exhaustive<false, COUNT>(s, o, d, best, st);
            else
                trace_exhaustive<true, COUNT>(s, o, d, best, st);
        } else {
            // with opaque materials only, the first blocker found decides (any-hit); otherwise the closest one does
            if (fp.any_transparent)
                ok &= trace_bvh<false, COUNT>(s, root_entry, o, d, best, st);
            else
                ok &= trace_bvh<true, COUNT>(s, root_entry, o, d, best, st);
        }
        if (best.ti < 0)
            return true;
        if (!fp.any_transparent)
            return false;
        const Shading sh = shading_at(s, best.ti, o, d, best.t);
        const float R0 = __ldg(&s.mats[2 * sh.mesh + 1]).w;
        if (R0 == 1.0f)
            return false;
        distance = xsub(distance, best.t);                    // shadow.cpp:51
        o = xadd(sh.p, xmul(d, 0.0005f));                      // shadow.cpp:53
        const float c = fabsf(xdot(d, sh.N));                  // shadow.cpp:55 (normal not re-normalised)
        intensity = (float)((double)intensity * (1.0 - schlick(R0, c)));
    }
    return true;
}

// K4a shadow rays to point lights (getPointLights' cansee call, src/shadow.cpp:120).
template <bool COUNT>
__global__ void __launch_bounds__(128) k_shadow_point(SceneDev s, int root_entry, FrameParams fp, BatchDev b)
{
    const unsigned n = b.counters->n_shadow_pt;
    TraceStats st;
    unsigned queries = 0;
    bool ok = true;
    while (true) {
        const unsigned base = warp_fetch(&b.counters->work[2]);
        if (base >= n)
            break;
        const unsigned i = base + lane_id();
        if (i < n) {
            const float4 pp = b.sq_point.p_pix[i];
            const float4 al = b.sq_point.a_light[i];
            const f3 lp = mk3(__ldg(&s.point_lights[2 * __float_as_int(al.w)]));
            float intensity = 1.0f;
            if (cansee<COUNT>(s, root_entry, fp, mk3(pp), lp, intensity, queries, st, ok)) {
                const float4 bb = b.sq_point.b[i];
                float* acc = reinterpret_cast<float*>(&b.accum[__float_as_int(pp.w)]);
                atomicAdd(acc + 0, al.x * intensity + bb.x);
                atomicAdd(acc + 1, al.y * intensity + bb.y);
                atomicAdd(acc + 2, al.z * intensity + bb.z);
            }
        }
        __syncwarp();
    }
    if (!ok)
        atomicExch(&b.counters->overflow, 2u);
    warp_add_u64(&b.counters->shadow_queries, queries);
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
    }
}

// K4b spherical lights (getSpherelights, src/shadow.cpp:139-226): sl_group lanes share one (hit, light) record and
// split its 1 + m*n samples; sample positions follow the reference's sequential `perp = rotate * perp`.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_shadow_sphere(SceneDev s, int root_entry, FrameParams fp, BatchDev b)
{
    const unsigned n = b.counters->n_shadow_sp;
    const int G = fp.sl_group;             // lanes per record
    const unsigned per_warp = 32u / (unsigned)G; // records per warp per fetch
    TraceStats st;
    unsigned queries = 0;
    bool ok = true;
    while (true) {
        unsigned base = 0;
        if (lane_id() == 0)
            base = atomicAdd(&b.counters->work[3], per_warp);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n)
            break;
        const unsigned i = base + (unsigned)lane_id() / (unsigned)G;
        const int sub = lane_id() % G;
        float sum = 0.0f;
        int hits = 0;
        float4 pp = make_float4(0, 0, 0, 0), al = make_float4(0, 0, 0, 0);
        if (i < n) {
            pp = b.sq_sphere.p_pix[i];
            al = b.sq_sphere.a_light[i];
            const float4 L = __ldg(&s.sphere_lights[2 * __float_as_int(al.w)]);
            const f3 p = mk3(pp), lpos = mk3(L);
            const float radius = L.w;
            f3 d = xnormalize(xsub(lpos, p)); // shadow.cpp:153-155
            f3 notd = d;                       // shadow.cpp:158-166
            if (d.x != 0.0f) {
                notd.y = -d.x;
                notd.x = d.y;
            } else {
                notd.y = -d.z;
                notd.z = d.y;
            }
            const f3 perp0 = xmul(xnormalize(xcross(d, notd)), radius); // shadow.cpp:169
            // rotate = I + C*sin + (C*C)*(1-cos), C columns {0,dz,-dy},{-dz,0,dx},{dy,-dx,0} (shadow.cpp:134-137)
            const float C[3][3] = { { 0.0f, d.z, -d.y }, { -d.z, 0.0f, d.x }, { d.y, -d.x, 0.0f } }; // C[col][row]
            float R[3][3];
#pragma unroll
            for (int col = 0; col < 3; col++)
#pragma unroll
                for (int row = 0; row < 3; row++) {
                    const float cc = xadd(xadd(xmul(C[0][row], C[col][0]), xmul(C[1][row], C[col][1])), xmul(C[2][row], C[col][2]));
                    const float ident = (col == row) ? 1.0f : 0.0f;
                    R[col][row] = xadd(xadd(ident, xmul(C[col][row], fp.sl_sin)), xmul(cc, fp.sl_omc));
                }
            for (int k = sub; k < fp.sl_rc; k += G) {
                f3 target;
                if (k == 0) {
                    target = lpos; // centre sample (shadow.cpp:148)
                } else {
                    const int spoke = (k - 1) / fp.sl_m, ring = (k - 1) % fp.sl_m;
                    f3 perp = perp0;
                    for (int r = 0; r < spoke; r++) // perp = rotate * perp (shadow.cpp:207)
                        perp = mk3(xadd(xadd(xmul(R[0][0], perp.x), xmul(R[1][0], perp.y)), xmul(R[2][0], perp.z)),
                            xadd(xadd(xmul(R[0][1], perp.x), xmul(R[1][1], perp.y)), xmul(R[2][1], perp.z)),
                            xadd(xadd(xmul(R[0][2], perp.x), xmul(R[1][2], perp.y)), xmul(R[2][2], perp.z)));
                    const float frac = xdiv((float)(fp.sl_m - ring), (float)fp.sl_m); // (m-j)/(float)m
                    target = xadd(lpos, xmul(perp, frac));
                }
                float intensity = 1.0f;
                const bool vis = cansee<COUNT>(s, root_entry, fp, p, target, intensity, queries, st, ok);
                if (k == 0) {
                    sum += intensity; // intensitySum starts at 1 and carries the centre ray's attenuation (shadow.cpp:145-150)
                    hits += vis ? 1 : 0;
                } else if (vis) {
                    sum += intensity;
                    hits++;
                }
            }
        }
        // segmented reduction over the G lanes of a record
        for (int o = G >> 1; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(kFull, sum, o);
            hits += __shfl_xor_sync(kFull, hits, o);
        }
        if (i < n && sub == 0 && hits > 0) { // shadow.cpp:212-221
            const float intensity = sum / (float)fp.sl_rc;
            const float4 bb = b.sq_sphere.b[i];
            float* acc = reinterpret_cast<float*>(&b.accum[__float_as_int(pp.w)]);
            atomicAdd(acc + 0, al.x * intensity + bb.x);
            atomicAdd(acc + 1, al.y * intensity + bb.y);
            atomicAdd(acc + 2, al.z * intensity + bb.z);
        }
        __syncwarp();
    }
    if (!ok)
        atomicExch(&b.counters->overflow, 2u);
    warp_add_u64(&b.counters->shadow_queries, queries);
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5 resolve: sample average (src/main.cpp:374,384) and Screen::setPixel's y flip (src/screen.cpp:32-38).  One
// warp writes one 32-pixel tile row = 512 contiguous bytes of float4; `out` may be a peer-mapped framebuffer of
// another GPU (the gather of finished tiles fused into this store).
__global__ void __launch_bounds__(256) k_resolve(FrameParams fp, const float4* __restrict__ accum, const int* __restrict__ prim_id,
    const float* __restrict__ prim_t, float4* out, int* out_id, float* out_t)
{
    const unsigned n = (unsigned)fp.n_local_tiles * kTilePixels;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const unsigned j = idx / kTilePixels, k = idx % kTilePixels;
        const unsigned x = k % kTileW, y = k / kTileW; // row-major inside the tile for coalesced stores
        const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world;
        const int px = (int)((g % (unsigned)fp.tiles_x) * kTileW + x), py = (int)((g / (unsigned)fp.tiles_x) * kTileH + y);
        if (px >= fp.W || py >= fp.H)
            continue;
        const unsigned lp = j * kTilePixels + (((y >> 2) * 4 + (x >> 3)) << 5) + ((y & 3) << 3) + (x & 7);
        const float4 a = accum[lp];
        const size_t o = (size_t)(fp.H - 1 - py) * fp.W + px;
        out[o] = make_float4(a.x * fp.sample_scale, a.y * fp.sample_scale, a.z * fp.sample_scale, 1.0f);
        if (out_id) {
            out_id[o] = prim_id[lp];
            out_t[o] = prim_t[lp];
        }
    }
}

// float4 framebuffer -> packed float3 (the host Screen's std::vector<glm::vec3>), written as float4 words.
__global__ void __launch_bounds__(256) k_pack_rgb(const float4* __restrict__ in, float* __restrict__ out, size_t n_pixels)
{
    const size_t n_words = (n_pixels * 3 + 3) / 4; // float4 words of the packed image
    for (size_t wi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; wi < n_words; wi += (size_t)gridDim.x * blockDim.x) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const size_t f = wi * 4 + c;
            const size_t p = f / 3;
            const int ch = (int)(f % 3);
            float val = 0.0f;
            if (p < n_pixels) {
                const float4 px = in[p];
                val = ch == 0 ? px.x : (ch == 1 ? px.y : px.z);
            }
            v[c] = val;
        }
        if (wi * 4 + 3 < n_pixels * 3)
            reinterpret_cast<float4*>(out)[wi] = make_float4(v[0], v[1], v[2], v[3]);
        else
            for (int c = 0; c < 4; c++)
                if (wi * 4 + c < n_pixels * 3)
                    out[wi * 4 + c] = v[c];
    }
}

// Closest hit for caller-supplied rays (rt_intersect).
__global__ void __launch_bounds__(128) k_intersect(SceneDev s, int root_entry, const float* __restrict__ rays, long long n, int use_bvh,
    int* tri_id, float* t_out, unsigned* overflow)
{
    TraceStats st;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const f3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        const f3 d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        HitRec best = fresh_query();
        if (use_bvh) {
            if (!trace_bvh<false, false>(s, root_entry, o, d, best, st))
                atomicExch(overflow, 2u);
        } else {
            trace_exhaustive<false, false>(s, o, d, best, st);
        }
        tri_id[i] = best.ti >= 0 ? best.id : -1;
        t_out[i] = best.t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// launch wrappers
static int grid_for(long long n, int block, int max_blocks)
{
    long long g = (n + block - 1) / block;
    if (g < 1)
        g = 1;
    if (g > max_blocks)
        g = max_blocks;
    return (int)g;
}

void launch_tri_setup(cudaStream_t st, const float* pos, const float* nrm, const int* mesh_id, const int* perm, int n,
    float4* plane, float4* v0, float4* v1, float4* v2, float4* n0, float4* n1, float4* n2)
{
    if (n <= 0)
        return;
    k_tri_setup<<<(n + 255) / 256, 256, 0, st>>>(pos, nrm, mesh_id, perm, n, plane, v0, v1, v2, n0, n1, n2);
}

void launch_level_reset(cudaStream_t st, Counters* c, int next_q, int clear_current)
{
    k_level_reset<<<1, 1, 0, st>>>(c, next_q, clear_current);
}

void launch_generate(cudaStream_t st, int sm_count, const FrameParams& fp, const BatchDev& b, unsigned first_lp, unsigned n_lp, int qi)
{
    const long long n = (long long)n_lp * fp.spp;
    k_generate<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(fp, b, first_lp, n_lp, qi);
}

void launch_extend(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int qi,
    int level, bool count)
{
    const int grid = sm_count * 8;
    if (count)
        k_extend<true><<<grid, 128, 0, st>>>(s, root_entry, fp, b, qi, level);
    else
        k_extend<false><<<grid, 128, 0, st>>>(s, root_entry, fp, b, qi, level);
}

void launch_shade(cudaStream_t st, int sm_count, const SceneDev& s, const FrameParams& fp, const BatchDev& b, int qi, int level)
{
    k_shade<<<sm_count * 8, 128, 0, st>>>(s, fp, b, qi, level);
}

void launch_shadow_point(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count)
{
    const int grid = sm_count * 8;
    if (count)
        k_shadow_point<true><<<grid, 128, 0, st>>>(s, root_entry, fp, b);
    else
        k_shadow_point<false><<<grid, 128, 0, st>>>(s, root_entry, fp, b);
}

void launch_shadow_sphere(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count)
{
    const int grid = sm_count * 8;
    if (count)
        k_shadow_sphere<true><<<grid, 128, 0, st>>>(s, root_entry, fp, b);
    else
        k_shadow_sphere<false><<<grid, 128, 0, st>>>(s, root_entry, fp, b);
}

void launch_resolve(cudaStream_t st, int sm_count, const FrameParams& fp, const float4* accum, const int* prim_id, const float* prim_t,
    float4* out, int* out_id, float* out_t)
{
    const long long n = (long long)fp.n_local_tiles * kTilePixels;
    k_resolve<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(fp, accum, prim_id, prim_t, out, out_id, out_t);
}

void launch_pack_rgb(cudaStream_t st, int sm_count, const float4* in, float* out, size_t n_pixels)
{
    const long long n = (long long)((n_pixels * 3 + 3) / 4);
    k_pack_rgb<<<grid_for(n, 256, sm_count * 8), 256, 0, st>>>(in, out, n_pixels);
}

void launch_intersect(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const float* rays, long long n, int use_bvh,
    int* tri_id, float* t_out, unsigned* overflow)
{
    if (n <= 0)
        return;
    k_intersect<<<grid_for(n, 128, sm_count * 16), 128, 0, st>>>(s, root_entry, rays, n, use_bvh, tri_id, t_out, overflow);
}

} // namespace rtb
