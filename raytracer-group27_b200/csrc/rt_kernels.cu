// Wavefront kernels of the per-pixel hot path (sm_100a): K0 tri_setup, K1 generate (fused into the level-0 extend),
// K2 extend, K3 shade, K4 shadow (point / spherical light), K5 resolve.  See DESIGN.md for the data flow; every
// kernel cites the reference lines it replaces.
#include "rt_kernels.h"
#include "rt_trace.cuh"
#include "rt_wide8.cuh"
#include <algorithm>
#include <cstdlib>

namespace rtb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kShadeBlock = 256;
#ifndef RT_STATIC_L0
#define RT_STATIC_L0 0
#endif
#ifndef RT_TRACE_GRID_MULT
#define RT_TRACE_GRID_MULT 8
#endif
#ifndef RT_TRACE_MIN_BLOCKS
#define RT_TRACE_MIN_BLOCKS 8 // 64 registers per thread: the whole persistent grid (8 blocks per SM) is resident in one wave
#endif

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void warp_add_u64(unsigned long long* dst, unsigned v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFull, v, o);
    if (lane_id() == 0 && v)
        atomicAdd(dst, (unsigned long long)v);
}

// Block-aggregated queue allocation for NQ queues at once: every thread says whether it wants a slot in each queue;
// one atomicAdd per queue per block.  Must be called by all threads of the block.  slot[q] = 0xffffffff when nothing
// was requested or the queue is full (overflow flagged).
template <int NQ>
__device__ __forceinline__ void block_alloc(unsigned* const (&counter)[NQ], const bool (&want)[NQ], const unsigned (&capacity)[NQ],
    unsigned* overflow, unsigned (&slot)[NQ], unsigned (*smem)[kShadeBlock / 32 + 1])
{
    const int warp = threadIdx.x >> 5, lane = lane_id(), nw = blockDim.x >> 5;
    unsigned rank[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const unsigned m = __ballot_sync(kFull, want[q]);
        rank[q] = (unsigned)__popc(m & ((1u << lane) - 1u));
        if (lane == 0)
            smem[q][warp] = (unsigned)__popc(m);
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        const int q = threadIdx.x;
        unsigned total = 0;
        for (int w = 0; w < nw; w++)
            total += smem[q][w];
        unsigned base = total ? atomicAdd(counter[q], total) : 0u;
        for (int w = 0; w < nw; w++) {
            const unsigned c = smem[q][w];
            smem[q][w] = base;
            base += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        unsigned sl = 0xffffffffu;
        if (want[q]) {
            sl = smem[q][warp] + rank[q];
            if (sl >= capacity[q]) {
                atomicExch(overflow, 1u);
                sl = 0xffffffffu;
            }
        }
        slot[q] = sl;
    }
    __syncthreads();
}

// The batch's overflow flag, one answer per warp (it may rise while a kernel of the side stream is starting up).
__device__ __forceinline__ bool batch_overflowed(const BatchDev& b) { return __shfl_sync(kFull, b.counters->overflow, 0) != 0u; }

// local padded pixel index -> pixel coordinates (tile interleaving across ranks, 8x4 warp blocks inside a tile)
__device__ __forceinline__ void local_to_pixel(const FrameParams& fp, unsigned lp, int& px, int& py)
{
    const unsigned j = lp / kTilePixels, k = lp % kTilePixels;
    const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world; // global tile id
    unsigned tx, ty;
    tile_xy(g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, tx, ty);
    const unsigned w = k >> 5, l = k & 31;
    px = (int)(tx * kTileW + (w & 3) * 8 + (l & 7));
    py = (int)(ty * kTileH + (w >> 2) * 4 + (l >> 3));
}

// false: the camera rays of this pixel cannot meet the scene (or the pixel is padding outside the image)
__device__ __forceinline__ bool pixel_sees_scene(const FrameParams& fp, int px, int py)
{
    return px >= fp.vis_x0 && px < fp.vis_x1 && py >= fp.vis_y0 && py < fp.vis_y1;
}

// K1 generate: pixel -> NDC (src/main.cpp:350-353), sub-pixel sample positions (358-375 / 309-335, 377-385) and
// Trackball::generateRay (framework/src/trackball.cpp:87-98) for primary ray `idx` of the batch starting at local
// pixel first_lp.  Returns false for rays of padding pixels outside the image and of pixels that cannot see the scene.  tag = (local pixel << 1) | first-sample.
__device__ __forceinline__ bool generate_ray(const FrameParams& fp, unsigned first_lp, unsigned idx, f3& o, f3& dir, int& tag)
{
    const unsigned lp = first_lp + idx / (unsigned)fp.spp;
    const int sidx = (int)(idx % (unsigned)fp.spp);
    int px, py;
    local_to_pixel(fp, lp, px, py);
    if (!pixel_sees_scene(fp, px, py))
        return false;
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
    o = mk3(fp.ox, fp.oy, fp.oz);
    tag = (int)((lp << 1) | (sidx == 0 ? 1u : 0u));
    return true;
}

struct Shading { // what shade and the transparent-shadow branch need at a hit
    f3 p;      // hit point
    f3 N;      // triangle: interpolated normal, flipped to the geometric side (not normalised); sphere: unit normal
    float4 m0, m1; // material {kd, shininess}{ks, transparency} (HitInfo::getMaterial, src/ray_tracing.h:21-27)
    float c0, c1, c2; // barycentric weights of a triangle hit (interpolateProperty, src/ray_tracing.cpp:313-316)
    int mesh, gid;    // mesh (material) index and global triangle id of a triangle hit; -1 for spheres
};

// Hit point and shading normal: src/ray_tracing.cpp:111, 147-160 with barycentricCoordinates (276-308) evaluated
// unconditionally ("defined barycentrics": the reference leaves barCoords uninitialised when its epsilon tests
// fail).  `o`, `d` are the ray as traced, t its hit parameter.
__device__ __forceinline__ Shading shading_at(const SceneDev& s, int ti, const f3& o, const f3& d, float t)
{
    Shading sh;
    sh.p = xadd(o, xmul(d, t));
    if (ti <= -2) { // sphere primitive (src/ray_tracing.cpp:199-204)
        const int k = RT_GUARD(-2 - ti, s.n_spheres, kChkTable);
        sh.N = xnormalize(xsub(sh.p, mk3(__ldg(&s.spheres[3 * k]))));
        sh.m0 = __ldg(&s.spheres[3 * k + 1]);
        sh.m1 = __ldg(&s.spheres[3 * k + 2]);
        sh.c0 = sh.c1 = sh.c2 = 0.0f;
        sh.mesh = sh.gid = -1;
        return sh;
    }
    ti = RT_GUARD(ti, s.n_tris, kChkTri);
    const float4 pl = __ldg(&s.tri_plane[kTriStride * ti]);
    const float4 a = __ldg(&s.tri_v0[kTriStride * ti]), b = __ldg(&s.tri_v1[kTriStride * ti]), c = __ldg(&s.tri_v2[kTriStride * ti]);
    const f3 v0 = mk3(a), v1 = mk3(b), v2 = mk3(c), fn = mk3(pl);
    const int mesh = RT_GUARD(__float_as_int(b.w), s.n_mats, kChkTable);
    sh.m0 = __ldg(&s.mats[2 * mesh]);
    sh.m1 = __ldg(&s.mats[2 * mesh + 1]);
    const float total = xlength(xcross(xsub(v1, v0), xsub(v2, v0)));
    const float c0 = xdiv(xlength(xcross(xsub(v1, sh.p), xsub(v2, sh.p))), total);
    const float c1 = xdiv(xlength(xcross(xsub(sh.p, v0), xsub(v2, v0))), total);
    const float c2 = xdiv(xlength(xcross(xsub(v1, v0), xsub(sh.p, v0))), total);
    const f3 n0 = mk3(__ldg(&s.tri_n0[kTriStride * ti])), n1 = mk3(__ldg(&s.tri_n1[kTriStride * ti])), n2 = mk3(__ldg(&s.tri_n2[kTriStride * ti]));
    f3 N = xadd(xadd(xmul(n0, c0), xmul(n1, c1)), xmul(n2, c2));
    if (xdot(N, fn) < 0.0f)
        N = xneg(N);
    sh.N = N;
    sh.c0 = c0;
    sh.c1 = c1;
    sh.c2 = c2;
    sh.mesh = mesh;
    sh.gid = __float_as_int(c.w);
    return sh;
}

// Image::getPixel for the NearestNeighbor and Bilinear filters (src/image.cpp:75-108) with the out-of-bounds rules
// (110-198), toImageCoordinates (118-131: rows start at the top), nearestNeighbor (217-246), bilinearInterpolation (249-268)
// and linearInterpolation (366-375), operation for operation.
__device__ __forceinline__ bool tex_out_of_bounds(float c) { return c < 0.0f || c > 1.0f; }

__device__ __forceinline__ float tex_rule(float c, int rule)
{
    if (rule == 1)
        return c > 1.0f ? 1.0f : (c < 0.0f ? 0.0f : c);
    if (rule == 2)
        return tex_out_of_bounds(c) ? xsub(c, floorf(c)) : c;
    return c;
}

__device__ __forceinline__ f3 tex_lerp(float low, float high, const f3& cl, const f3& ch, float p)
{
    if ((double)fabsf(xsub(high, low)) < 1e-6)
        return cl;
    const float c = xdiv(xsub(p, low), xsub(high, low));
    return xadd(xmul(cl, xsub(1.0f, c)), xmul(ch, c));
}

__device__ __forceinline__ f3 tex_nearest(const float4* px, unsigned w, unsigned h, float ix, float iy) // nearestNeighbor, image.cpp:201-228
{
    unsigned x = (unsigned)roundf(ix), y = (unsigned)roundf(iy);
    if (x >= w)
        x = w - 1u;
    if (y >= h)
        y = h - 1u;
    return mk3(__ldg(&px[RT_GUARD((size_t)y * w + x, (size_t)w * h, kChkTexel)]));
}

__device__ __forceinline__ f3 tex_bilinear(const float4* px, unsigned w, unsigned h, float ix, float iy) // bilinearInterpolation, image.cpp:231-251
{
    const float xl = floorf(ix), xh = ceilf(ix), yl = floorf(iy), yh = ceilf(iy);
    const size_t n = (size_t)w * h; // (bound of the checked build)
    (void)n;
    const f3 ll = mk3(__ldg(&px[RT_GUARD((size_t)(unsigned)yl * w + (unsigned)xl, n, kChkTexel)])), lr = mk3(__ldg(&px[RT_GUARD((size_t)(unsigned)yl * w + (unsigned)xh, n, kChkTexel)]));
    const f3 hl = mk3(__ldg(&px[RT_GUARD((size_t)(unsigned)yh * w + (unsigned)xl, n, kChkTexel)])), hr = mk3(__ldg(&px[RT_GUARD((size_t)(unsigned)yh * w + (unsigned)xh, n, kChkTexel)]));
    const f3 low = tex_lerp(xl, xh, ll, lr, ix), high = tex_lerp(xl, xh, hl, hr, ix);
    return tex_lerp(yl, yh, low, high, iy);
}

// First texel of pyramid level `level` of a w x w texture whose levels lie back to back: sum of (w >> i)^2 for i < level.
__device__ __forceinline__ unsigned mip_offset(unsigned w, unsigned level) { return (4u * (w * w - (w >> level) * (w >> level))) / 3u; }

// Image::getPixel (src/image.cpp:75-112) with the out-of-bounds rules (110-198) and all five filters; lod: computeLevelOfDetails.
__device__ __forceinline__ f3 sample_texture(const SceneDev& s, const FrameParams& fp, int tex, float u, float v, float lod)
{
    const f3 border = mk3(fp.tex_border_r, fp.tex_border_g, fp.tex_border_b);
    if (fp.tex_oob_x == 0 && tex_out_of_bounds(u))
        return border;
    if (fp.tex_oob_y == 0 && tex_out_of_bounds(v))
        return border;
    u = tex_rule(u, fp.tex_oob_x);
    v = tex_rule(v, fp.tex_oob_y);
    const int4 tb = __ldg(&s.tex_table[tex]);
    const unsigned w = (unsigned)tb.y, h = (unsigned)tb.z;
    const float4* px = s.tex_texels + tb.x;
    if (fp.tex_filter < 2) { // level 0 of the image itself; toImageCoordinates (118-131): rows start at the top
        const float ix = xmul(u, (float)(w - 1u)), iy = xmul(xsub(1.0f, v), (float)(h - 1u));
        return fp.tex_filter == 0 ? tex_nearest(px, w, h, ix, iy) : tex_bilinear(px, w, h, ix, iy);
    }
    // The mip-mapped filters: nearestLevelMipmapping (255-277), nearestLevelBilinear (280-302), trilinearInterpolation (305-363) with
    // getBestLevelMipmap (506-541).  tb.w = number of pyramid levels; 0: the texture is not a square power of two and has none
    // (canUseMipmapping, 400-402), the filters then answer white, Trilinear black.
    const int n = tb.w;
    if (n == 0)
        return fp.tex_filter == 4 ? mk3(0.0f, 0.0f, 0.0f) : mk3(1.0f, 1.0f, 1.0f);
    if (fp.tex_filter != 4) {
        unsigned level;
        if (xsub(lod, floorf(lod)) < xsub(ceilf(lod), lod))
            level = (unsigned)(int)fmaxf(0.0f, floorf(lod));
        else
            level = (unsigned)(int)fminf((float)n - 1.0f, ceilf(lod));
        if (level >= (unsigned)n) // a level of detail beyond the pyramid rounded DOWN: getWidthHeightForLevel fails and the filters answer white (image.cpp:268-271, 293-296)
            return mk3(1.0f, 1.0f, 1.0f);
        const unsigned wl = w >> level;
        const float ix = xmul(u, (float)(wl - 1u)), iy = xmul(xsub(1.0f, v), (float)(wl - 1u));
        const float4* pl = px + mip_offset(w, level);
        return fp.tex_filter == 2 ? tex_nearest(pl, wl, wl, ix, iy) : tex_bilinear(pl, wl, wl, ix, iy);
    }
    const unsigned high = (unsigned)(int)fminf((float)n - 1.0f, ceilf(lod)), low = (unsigned)(int)fmaxf(0.0f, floorf(lod));
    if (low >= (unsigned)n || high >= (unsigned)n) // getWidthHeightForLevel fails
        return mk3(0.0f, 0.0f, 0.0f);
    const unsigned wlo = w >> low, whi = w >> high;
    const f3 cl = tex_bilinear(px + mip_offset(w, low), wlo, wlo, xmul(u, (float)(wlo - 1u)), xmul(xsub(1.0f, v), (float)(wlo - 1u)));
    const f3 ch = tex_bilinear(px + mip_offset(w, high), whi, whi, xmul(u, (float)(whi - 1u)), xmul(xsub(1.0f, v), (float)(whi - 1u)));
    return tex_lerp((float)low, (float)high, cl, ch, lod);
}

// computeDerivativeOfBarycentricCoordinate, src/ray_differentials.cpp:38-48
__device__ __forceinline__ float bary_derivative(const f3& a, const f3& b, const f3& p, const f3& pd, float area)
{
    const f3 term1 = xadd(xcross(pd, xsub(p, b)), xcross(xsub(p, a), pd));
    const f3 term2 = xcross(xsub(a, p), xsub(b, p));
    const float nominator = xadd(xdot(term1, term2), xdot(term2, term1));
    const float denominator = xmul(xmul(2.0f, area), xsqrt(xdot(term2, term2)));
    return xdiv(nominator, denominator);
}

// Level of detail of a textured triangle hit: the ray differentials of framework/include/ray.h:19-28 (their initial state DEFINED
// as in oracle/ref_harness.cpp: right = (1,0,0), up = (0,-1,0), evaluated on (0,0,-1) for camera rays — Trackball::generateRay fills
// a default-constructed Ray — and on the ray's own direction for reflection / refraction / glossy rays, which are fresh objects and
// inherit nothing), transfer_ray_differentials (src/ray_differentials.cpp:5-16) and computeLevelOfDetails (121-139) with
// computeTexturePartialDerivativeInInterpolatedTrianglePoint (73-88).  d, t: the ray as traced and its hit parameter.
// std::pow / std::log2 of the reference are evaluated in double and rounded (glibc's float versions can differ from that in the last place).
__device__ __forceinline__ float level_of_detail(const SceneDev& s, const Shading& sh, int ti, const f3& d, float t, bool camera_ray)
{
    const f3 right = mk3(1.0f, 0.0f, 0.0f), up = mk3(0.0f, -1.0f, 0.0f);
    const f3 dir0 = camera_ray ? mk3(0.0f, 0.0f, -1.0f) : d;
    const float dd = xdot(dir0, dir0);
    const float p15 = (float)pow((double)dd, 1.5);
    const f3 dDx = xdiv(xsub(xmul(right, dd), xmul(dir0, xdot(dir0, right))), p15);
    const f3 dDy = xdiv(xsub(xmul(up, dd), xmul(dir0, xdot(dir0, up))), p15);
    const f3 N = xnormalize(sh.N), D = xnormalize(d);
    const f3 zero = mk3(0.0f, 0.0f, 0.0f);
    const f3 ax = xadd(zero, xmul(dDx, t)), ay = xadd(zero, xmul(dDy, t)); // dP + t * dD with dP = 0
    const float dn = xdot(D, N);
    const float dt_dx = xdiv(-xdot(ax, N), dn), dt_dy = xdiv(-xdot(ay, N), dn);
    const f3 dPx = xadd(ax, xmul(D, dt_dx)), dPy = xadd(ay, xmul(D, dt_dy));
    const f3 v0 = mk3(__ldg(&s.tri_v0[kTriStride * ti])), v1 = mk3(__ldg(&s.tri_v1[kTriStride * ti])), v2 = mk3(__ldg(&s.tri_v2[kTriStride * ti]));
    float2 t0 = make_float2(0.0f, 0.0f), t1 = t0, t2 = t0;
    if (s.tri_uv) {
        t0 = __ldg(&s.tri_uv[3 * (size_t)sh.gid]);
        t1 = __ldg(&s.tri_uv[3 * (size_t)sh.gid + 1]);
        t2 = __ldg(&s.tri_uv[3 * (size_t)sh.gid + 2]);
    }
    const float area = xlength(xcross(xsub(v2, v0), xsub(v1, v0)));
    float len[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const f3 pd = k == 0 ? dPx : dPy;
        const float a = bary_derivative(v2, v1, sh.p, pd, area), b = bary_derivative(v0, v2, sh.p, pd, area), c = bary_derivative(v1, v0, sh.p, pd, area);
        const float tx = xmul(1.0f, xadd(xadd(xmul(a, t0.x), xmul(b, t1.x)), xmul(c, t2.x)));
        const float ty = xmul(1.0f, xadd(xadd(xmul(a, t0.y), xmul(b, t1.y)), xmul(c, t2.y)));
        len[k] = xsqrt(xadd(xmul(tx, tx), xmul(ty, ty)));
    }
    const float m = (len[0] < len[1]) ? len[1] : len[0]; // glm::max(a, b) = (a < b) ? b : a
    const float l = (float)log2((double)m);
    return (0.0f < l) ? l : 0.0f;
}

// kd of a hit: the material's, or its texture at the interpolated texture coordinate (getFinalColor, src/main.cpp:155-171)
__device__ __forceinline__ f3 diffuse_colour(const SceneDev& s, const FrameParams& fp, const Shading& sh, int ti, const f3& d, float t, bool camera_ray,
    bool force = false)
{
    if ((fp.tex_on || force) && sh.mesh >= 0) {
        const int tex = __ldg(&s.mat_tex[sh.mesh]);
        if (tex >= 0) {
            float2 t0 = make_float2(0.0f, 0.0f), t1 = t0, t2 = t0;
            if (s.tri_uv) {
                t0 = __ldg(&s.tri_uv[3 * (size_t)sh.gid]);
                t1 = __ldg(&s.tri_uv[3 * (size_t)sh.gid + 1]);
                t2 = __ldg(&s.tri_uv[3 * (size_t)sh.gid + 2]);
            }
            const float u = xadd(xadd(xmul(t0.x, sh.c0), xmul(t1.x, sh.c1)), xmul(t2.x, sh.c2));
            const float v = xadd(xadd(xmul(t0.y, sh.c0), xmul(t1.y, sh.c1)), xmul(t2.y, sh.c2));
            const float lod = fp.tex_filter >= 2 ? level_of_detail(s, sh, ti, d, t, camera_ray) : 0.0f;
            return sample_texture(s, fp, tex, u, v, lod);
        }
    }
    return mk3(sh.m0);
}

// Schlick term of the reference, evaluated in double like `R0 + (1 - R0) * std::pow(1 - c, 5)` with float c, R0
// (src/main.cpp:279, src/shadow.cpp:59): pow(float,int) promotes to double.
__device__ __forceinline__ double schlick(float R0, float c)
{
    const double omc = (double)(1.0f - c);
    const double p5 = omc * omc * omc * omc * omc;
    return (double)R0 + (double)(1.0f - R0) * p5;
}

__device__ __forceinline__ void accumulate(const BatchDev& b, int pix, float r, float g, float bl)
{
#if RT_CHECKED
    pix = RT_GUARD(pix, b.accum_pixels, kChkAccum);
#endif
    float* acc = reinterpret_cast<float*>(&b.accum[pix]);
    atomicAdd(acc + 0, r);
    atomicAdd(acc + 1, g);
    atomicAdd(acc + 2, bl);
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
    const size_t o = (size_t)kTriStride * i;
    plane[o] = make_float4(nn.x, nn.y, nn.z, D);
    v0o[o] = make_float4(v0.x, v0.y, v0.z, __int_as_float(g));
    v1o[o] = make_float4(v1.x, v1.y, v1.z, __int_as_float(mesh_id ? mesh_id[g] : 0));
    v2o[o] = make_float4(v2.x, v2.y, v2.z, __int_as_float(g));
    const float* q = nrm + 9 * (size_t)g;
    n0o[o] = make_float4(q[0], q[1], q[2], 0.0f);
    n1o[o] = make_float4(q[3], q[4], q[5], 0.0f);
    n2o[o] = make_float4(q[6], q[7], q[8], 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// Reset of the per-level cursors / queue fills (one thread).  n_current >= 0 sets the fill of the current queue
// (level 0: the primary rays are generated on the fly, the queue only provides the hit slots).
__global__ void k_level_reset(Counters* c, int next_q, long long n_current, unsigned long long add_primary, int par)
{
    c->work[0] = c->work[1] = 0;
    c->sh[par].n_pt = c->sh[par].n_sp = c->sh[par].n_pl = c->sh[par].work_pt = c->sh[par].work_sp = c->sh[par].work_pl = 0;
    c->n_rays[next_q] = 0;
    if (n_current >= 0)
        c->n_rays[next_q ^ 1] = (unsigned)n_current;
    c->primary_rays += add_primary;
}

// ---------------------------------------------------------------------------------------------------------
// K2 extend: closest hit of every ray of the level (BoundingVolumeHierarchy::intersect,
// src/bounding_volume_hierarchy.cpp:49-78) through the warp-synchronous engine of rt_trace.cuh.  At level 0 the ray is
// generated from the pixel index when a lane picks the item up (K1 fused).  Result: hit[i] = {bits(t), BVH-order triangle}.
template <bool LEVEL0, bool COUNT>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, RT_TRACE_MIN_BLOCKS) k_extend(SceneDev s, int root_entry, FrameParams fp, BatchDev b, int qi, unsigned first_lp, int level)
{
    // a queue that overflowed holds `capacity` items (block_alloc rejects the rest but still counts them); once the flag is up
    // the rest of the batch is skipped, rt_sync reports RT_ERR_OVERFLOW and the caller renders again with more head-room
    if (batch_overflowed(b))
        return;
    const unsigned n = LEVEL0 ? b.counters->n_rays[qi] : min(b.counters->n_rays[qi], b.ray_capacity); // (level 0: the batch's primary rays, set by the host)
    if (blockIdx.x == 0 && threadIdx.x == 0 && level < kLevelHistory)
        b.counters->level_ext[level] = n;
    TraceStats st;
    int tag = 0;
    trace_queue<false, COUNT, LEVEL0 && RT_STATIC_L0>(
        s, root_entry, fp.exhaustive != 0, &b.counters->work[0], n, st,
        [&](unsigned item, f3& o, f3& d, HitRec& q) {
            q = fresh_query();
            if (LEVEL0)
                return generate_ray(fp, first_lp, item, o, d, tag);
            item = RT_GUARD(item, b.ray_capacity, kChkQueue);
            const float4 op = b.q[qi].o_pix[item];
            o = mk3(op);
            d = mk3(b.q[qi].d[item]);
            tag = __float_as_int(op.w);
            return true;
        },
        [&](unsigned item, const HitRec& best, f3&, f3&, HitRec&) {
            b.q[qi].hit[RT_GUARD(item, b.hit_capacity, kChkQueue)] = make_int2(__float_as_int(best.t), best.ti);
            if (LEVEL0 && b.prim_id && (tag & 1)) { // first sample of the pixel
                b.prim_id[RT_GUARD(tag >> 1, b.accum_pixels, kChkAccum)] = global_id(s, best);
                b.prim_t[RT_GUARD(tag >> 1, b.accum_pixels, kChkAccum)] = best.t;
            }
            return false;
        },
        LEVEL0 ? 32 : RT_MAX_QUOTA, max(fp.min_quota, 1));
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
        warp_add_u64(&b.counters->ext_node_visits, st.nodes);
        warp_add_u64(&b.counters->ext_tri_tests, st.tris);
        warp_add_u64(&b.counters->ext_tri_tests_full, st.tris_full);
        atomicMax(&b.counters->max_ray_nodes, st.max_ray_nodes);
        atomicMax(&b.counters->max_ray_tris, st.max_ray_tris);
    }
}

// Glossy rays (src/main.cpp:204-250).  The reference draws their directions from rand(), one stream shared by all of its
// OpenMP threads, so it has no reproducible answer for glossy_ray_count > 1; the rebuilt path defines the stream: every ray
// carries a 32-bit path id (primary ray: mix(pixel index y * W + x, sample index); child: mix(parent id, child index) with
// mirror / reflection child 0, glossy child i = i, refraction child 0x4000), and uniform j of attempt a for glossy child i
// at a hit is the top 24 bits of mix(mix(mix(path, i), a), j) over 2^24.  The CPU port (oracle/oracle_port.cpp) restates
// the same integer arithmetic; parity for this feature is device == port, there is no reference run to pin it on.
__device__ __forceinline__ unsigned path_mix(unsigned a, unsigned b)
{
    unsigned h = (a * 0x9E3779B1u) ^ (b + 0x7F4A7C15u + (a << 6) + (a >> 2));
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ float path_uniform(unsigned path, unsigned child, unsigned attempt, unsigned j)
{
    return (float)(path_mix(path_mix(path_mix(path, child), attempt), j) >> 8) * (1.0f / 16777216.0f);
}

// A point-like light (point or spot, getPointLights / getSpotLichts, src/shadow.cpp:106-131, 229-252) seen from a hit.
// light_reaches: false for a spot light whose cone does not contain the hit.  light_terms: the coefficients of calcColor
// (src/main.cpp:112-121) folded with the path throughput w: the light adds A * intensity + B when it is visible.
__device__ __forceinline__ bool light_reaches(const SceneDev& s, int li, const f3& p)
{
    li = RT_GUARD(li, s.n_point_like, kChkTable);
    const float4 L0 = __ldg(s.point_lights + 3 * li);
    if (L0.w == 0.0f)
        return true;
    const float4 L1 = __ldg(s.point_lights + 3 * li + 1); // spot: dot(normalize(direction), normalize(p - position)) > cos(radians(angle))
    const f3 sd = mk3(__ldg(s.point_lights + 3 * li + 2));
    return xdot(xnormalize(sd), xnormalize(xsub(p, mk3(L0)))) > L1.w;
}

__device__ __forceinline__ void light_terms(const SceneDev& s, int li, const f3& p, const f3& Nn, const f3& reflN, const f3& w, const f3& kd, const f3& ks,
    float shininess, f3& A, f3& B)
{
    li = RT_GUARD(li, s.n_point_like, kChkTable);
    const f3 lp = mk3(__ldg(s.point_lights + 3 * li)), lc = mk3(__ldg(s.point_lights + 3 * li + 1));
    const f3 ldir = xnormalize(xsub(lp, p));
    const float cosNL = fabsf(xdot(Nn, ldir));                 // shadow.cpp:125 / 245
    const float cosRL = fmaxf(0.0f, xdot(reflN, ldir));         // shadow.cpp:126 / 246
    A = mk3(w.x * kd.x * lc.x * cosNL, w.y * kd.y * lc.y * cosNL, w.z * kd.z * lc.z * cosNL);
    B = mk3(0, 0, 0);
    if (shininess > 0.0f) {
        const float sp = powf(cosRL, shininess);
        B = mk3(w.x * lc.x * ks.x * sp, w.y * lc.y * ks.y * sp, w.z * lc.z * ks.z * sp);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3 shade: getFinalColor without the recursion (src/main.cpp:129-190), light set-up of getPointLights /
// getSpherelights (src/shadow.cpp:106-131, 139-226: everything except the cansee calls), calcColor
// (src/main.cpp:112-121) folded into per-light coefficients, and the spawn of mirror (191-256 with
// glossy_ray_count == 1 => weight ks*ks) and dielectric (257-290) children with a throughput instead of recursion.
// One thread per ray slot; output queues are filled through block-aggregated ballot compaction.
// EXTRAS: the texture branch and the glossy rays are compiled in (frames that use neither run the leaner kernels)
template <bool LEVEL0, bool EXTRAS>
__global__ void __launch_bounds__(kShadeBlock) k_shade(SceneDev s, FrameParams fp, BatchDev b, int qi, int level, unsigned first_lp)
{
    __shared__ unsigned smem[2][kShadeBlock / 32 + 1];
    __shared__ unsigned s_overflow;
    if (threadIdx.x == 0)
        s_overflow = b.counters->overflow; // read once per block: the flag may rise while the kernel runs, the barriers below need one answer
    __syncthreads();
    if (s_overflow)
        return;
    const unsigned n = LEVEL0 ? b.counters->n_rays[qi] : min(b.counters->n_rays[qi], b.ray_capacity);
    const int qo = qi ^ 1;
    const unsigned n_round = (n + kShadeBlock - 1) / kShadeBlock * kShadeBlock;
    unsigned n_secondary = 0;
    for (unsigned i = blockIdx.x * kShadeBlock + threadIdx.x; i < n_round; i += gridDim.x * kShadeBlock) {
        bool hit = false;
        int pix = 0;
        unsigned path = 0; // path id of the ray (glossy random stream; 0 and unused when glossy_ray_count is 1)
        f3 w = mk3(0, 0, 0), refl = mk3(0, 0, 0), refr = mk3(0, 0, 0), dn = mk3(0, 0, 0), Nn = mk3(0, 0, 0);
        f3 d_ray = mk3(0, 0, 0); // the ray as traced (EXTRAS: the mip-mapped texture filters take their level of detail from it)
        Shading sh;
        sh.p = mk3(0, 0, 0);
        sh.N = mk3(0, 0, 0);
        sh.m0 = make_float4(0, 0, 0, 0);
        sh.m1 = make_float4(0, 0, 0, 1);
        int2 h = make_int2(0, -1);
        if (i < n) {
            bool valid = true;
            if (LEVEL0) { // padding pixels of ragged tiles have no hit record
                int px, py;
                local_to_pixel(fp, first_lp + i / (unsigned)fp.spp, px, py);
                valid = pixel_sees_scene(fp, px, py);
            }
            if (valid)
                h = b.q[qi].hit[RT_GUARD(i, b.hit_capacity, kChkQueue)];
        }
        // most slots of a level-0 frame are misses: a block without any hit has nothing to allocate
        if (!__syncthreads_or(h.y != -1))
            continue;
        if (i < n) {
            f3 o, d;
            int tag = 0;
            if (h.y != -1) {
                hit = true;
                if (LEVEL0) {
                    generate_ray(fp, first_lp, i, o, d, tag); // K1 again: the primary ray is a function of the index
                    w = mk3(1.0f, 1.0f, 1.0f);
                    if (EXTRAS && fp.glossy > 1) {
                        int px, py;
                        local_to_pixel(fp, first_lp + i / (unsigned)fp.spp, px, py);
                        path = path_mix((unsigned)(py * fp.W + px), i % (unsigned)fp.spp);
                    }
                } else {
                    (void)RT_GUARD(i, b.ray_capacity, kChkQueue);
                    const float4 op = b.q[qi].o_pix[i];
                    const float4 dq = b.q[qi].d[i];
                    o = mk3(op);
                    d = mk3(dq);
                    w = mk3(b.q[qi].w[i]);
                    tag = __float_as_int(op.w);
                    path = (unsigned)__float_as_int(dq.w);
                }
                pix = tag >> 1;
                sh = shading_at(s, h.y, o, d, __int_as_float(h.x));
                if (EXTRAS)
                    d_ray = d;
                dn = xnormalize(d);
                Nn = xnormalize(sh.N);
                refl = xreflect(dn, Nn); // main.cpp:141
            }
        }
        if (EXTRAS && fp.tex_debug) { // getFinalColorNoRayTracingJustTextures (main.cpp:75-106): the texel, or white without a texture
            if (hit) {
                f3 c = mk3(1.0f, 1.0f, 1.0f);
                if (fp.tex_available && sh.mesh >= 0 && __ldg(&s.mat_tex[sh.mesh]) >= 0)
                    c = diffuse_colour(s, fp, sh, h.y, d_ray, __int_as_float(h.x), LEVEL0, true);
                accumulate(b, pix, c.x, c.y, c.z);
            }
            continue;
        }
        const f3 kd = (EXTRAS && hit) ? diffuse_colour(s, fp, sh, h.y, d_ray, __int_as_float(h.x), LEVEL0) : mk3(sh.m0), ks = mk3(sh.m1);
        const float shininess = sh.m0.w, transparency = sh.m1.w;

        // children
        bool want0 = false, want1 = false, glossy_here = false;
        f3 w0 = mk3(0, 0, 0), w1 = mk3(0, 0, 0);
        if (hit && level < fp.max_level) { // main.cpp:187
            if (transparency == 1.0f) {
                if (ks.x > 0.0f || ks.y > 0.0f || ks.z > 0.0f) { // main.cpp:194
                    want0 = true;
                    w0 = mk3(w.x * ks.x * ks.x, w.y * ks.y * ks.y, w.z * ks.z * ks.z);
                    if (EXTRAS && shininess != 0.0f && fp.glossy > 1) { // color += ks * reflectColor / glossy_ray_count, main.cpp:250
                        const float g = (float)fp.glossy;
                        w0 = mk3(w0.x / g, w0.y / g, w0.z / g);
                        glossy_here = true;
                    }
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
        n_secondary += (want0 ? 1u : 0u) + (want1 ? 1u : 0u);

        // one allocation round for: first child ray, spherical-light records (n_sphere per hit, contiguous); a second
        // round for the refraction child (dielectric hits only)
        unsigned* const counters[2] = { &b.counters->n_rays[qo], &b.counters->sh[b.par].n_sp };
        const bool want[2] = { want0, hit && fp.n_sphere > 0 };
        const unsigned cap[2] = { b.ray_capacity, b.shadow_sp_capacity / (unsigned)max(fp.n_sphere, 1) };
        unsigned slot[2];
        block_alloc<2>(counters, want, cap, &b.counters->overflow, slot, smem);
        unsigned slot1 = 0xffffffffu;
        if (fp.any_transparent) {
            unsigned* const c1[1] = { &b.counters->n_rays[qo] };
            const bool w1v[1] = { want1 };
            const unsigned cap1[1] = { b.ray_capacity };
            unsigned s1[1];
            block_alloc<1>(c1, w1v, cap1, &b.counters->overflow, s1, smem);
            slot1 = s1[0];
        }
        if (slot[0] != 0xffffffffu) {
            (void)RT_GUARD(slot[0], b.ray_capacity, kChkQueue);
            const f3 o2 = xadd(sh.p, xmul(refl, 0.01f)); // main.cpp:199,286
            b.q[qo].o_pix[slot[0]] = make_float4(o2.x, o2.y, o2.z, __int_as_float(pix << 1));
            b.q[qo].d[slot[0]] = make_float4(refl.x, refl.y, refl.z, __int_as_float((int)path_mix(path, 0u)));
            b.q[qo].w[slot[0]] = make_float4(w0.x, w0.y, w0.z, 0.0f);
        }
        if (slot1 != 0xffffffffu) {
            (void)RT_GUARD(slot1, b.ray_capacity, kChkQueue);
            const f3 o2 = xadd(sh.p, xmul(refr, 0.01f)); // main.cpp:288
            b.q[qo].o_pix[slot1] = make_float4(o2.x, o2.y, o2.z, __int_as_float(pix << 1));
            b.q[qo].d[slot1] = make_float4(refr.x, refr.y, refr.z, __int_as_float((int)path_mix(path, 0x4000u)));
            b.q[qo].w[slot1] = make_float4(w1.x, w1.y, w1.z, 0.0f);
        }
        // Glossy rays (main.cpp:207-249), glossy_ray_count - 1 candidates around the mirror direction with the defined
        // random stream; one block-wide allocation round per candidate (all threads take part, few have one).
        if (EXTRAS && fp.glossy > 1) {
            f3 pr1 = mk3(0, 0, 0), pr2 = mk3(0, 0, 0);
            float dev = 0.0f;
            if (glossy_here) {
                f3 notr = refl; // a vector that is not in line with reflect, main.cpp:211-219
                if (refl.x != 0.0f) {
                    notr.y = -refl.x;
                    notr.x = refl.y;
                } else {
                    notr.y = -refl.z;
                    notr.z = refl.y;
                }
                pr1 = xcross(refl, notr);
                pr2 = xcross(refl, pr1);
                dev = sh.mesh >= 0 ? __ldg(&s.mat_glossy_d[sh.mesh]) : __ldg(&s.sphere_glossy_d[-2 - h.y]); // main.cpp:224, from the host's libm
            }
            for (int gi = 1; gi < fp.glossy; gi++) {
                f3 shine = mk3(0, 0, 0);
                bool cast = false;
                if (glossy_here) {
                    int loopcount = 0;
                    unsigned attempt = 0;
                    do {
                        float a, bb;
                        do {
                            a = path_uniform(path, (unsigned)gi, attempt, 0u);
                            bb = path_uniform(path, (unsigned)gi, attempt, 1u);
                            attempt++;
                        } while (a == 0.0f && bb == 0.0f);
                        a = xmul(xsub(xmul(2.0f, a), 1.0f), dev);
                        bb = xmul(xsub(xmul(2.0f, bb), 1.0f), dev);
                        shine = xnormalize(xadd(xadd(refl, xmul(pr1, a)), xmul(pr2, bb)));
                        loopcount++;
                    } while (xdot(shine, sh.N) <= 0.0f && loopcount < fp.glossy / 4);
                    cast = xdot(shine, sh.N) > 0.0f;
                }
                unsigned* const cg[1] = { &b.counters->n_rays[qo] };
                const bool wg[1] = { cast };
                const unsigned capg[1] = { b.ray_capacity };
                unsigned sg[1];
                block_alloc<1>(cg, wg, capg, &b.counters->overflow, sg, smem);
                if (cast)
                    n_secondary++;
                if (sg[0] == 0xffffffffu)
                    continue;
                (void)RT_GUARD(sg[0], b.ray_capacity, kChkQueue);
                const float wgt = fmaxf(powf(xdot(refl, shine), shininess), 0.0f) / (float)fp.glossy; // main.cpp:245,250
                const f3 o2 = xadd(sh.p, xmul(shine, 0.01f));
                b.q[qo].o_pix[sg[0]] = make_float4(o2.x, o2.y, o2.z, __int_as_float(pix << 1));
                b.q[qo].d[sg[0]] = make_float4(shine.x, shine.y, shine.z, __int_as_float((int)path_mix(path, (unsigned)gi)));
                b.q[qo].w[sg[0]] = make_float4(w.x * ks.x * wgt, w.y * ks.y * wgt, w.z * ks.z * wgt, 0.0f);
            }
        }
        const f3 reflN = xnormalize(refl);
        // Direct light.  Point and spot lights (getPointLights / getSpotLichts, shadow.cpp:106-131, 229-252): one shadow
        // record per light, for a spot light only if the hit lies inside its cone; the shadow kernel adds
        // A * intensity + B when the light is visible (calcColor, main.cpp:112-121).
        for (int li = 0; li < fp.n_point; li++) {
            const bool lit = hit && light_reaches(s, li, sh.p);
            unsigned* const cl[1] = { &b.counters->sh[b.par].n_pt };
            const bool wl[1] = { lit };
            const unsigned capl[1] = { b.shadow_pt_capacity };
            unsigned sl1[1];
            block_alloc<1>(cl, wl, capl, &b.counters->overflow, sl1, smem);
            if (sl1[0] == 0xffffffffu)
                continue;
            (void)RT_GUARD(sl1[0], b.shadow_pt_capacity, kChkQueue);
            f3 A, B;
            light_terms(s, li, sh.p, Nn, reflN, w, kd, ks, shininess, A, B);
            b.sq_point.p_pix[sl1[0]] = make_float4(sh.p.x, sh.p.y, sh.p.z, __int_as_float(pix));
            b.sq_point.a_light[sl1[0]] = make_float4(A.x, A.y, A.z, __int_as_float(li));
            b.sq_point.b[sl1[0]] = make_float4(B.x, B.y, B.z, 0.0f);
        }
        // Spherical lights (getSpherelights, shadow.cpp:139-226): one record per light, sampled by k_shadow_sphere.
        if (hit && slot[1] != 0xffffffffu) {
            for (int lj = 0; lj < fp.n_sphere; lj++) {
                const float4* L = s.sphere_lights + 2 * lj;
                const f3 lp = mk3(__ldg(L)), lc = mk3(__ldg(L + 1));
                const f3 ldir = xnormalize(xsub(lp, sh.p));
                const float cosNL = fabsf(xdot(Nn, ldir));             // shadow.cpp:218
                const float cosRL = fmaxf(0.0f, xdot(reflN, ldir));     // shadow.cpp:219
                const f3 A = mk3(w.x * kd.x * lc.x * cosNL, w.y * kd.y * lc.y * cosNL, w.z * kd.z * lc.z * cosNL);
                f3 B = mk3(0, 0, 0);
                if (shininess > 0.0f) {
                    const float sp = powf(cosRL, shininess);
                    B = mk3(w.x * lc.x * ks.x * sp, w.y * lc.y * ks.y * sp, w.z * lc.z * ks.z * sp);
                }
                const unsigned sl = RT_GUARD(slot[1] * (unsigned)fp.n_sphere + (unsigned)lj, b.shadow_sp_capacity, kChkQueue);
                b.sq_sphere.p_pix[sl] = make_float4(sh.p.x, sh.p.y, sh.p.z, __int_as_float(pix));
                b.sq_sphere.a_light[sl] = make_float4(A.x, A.y, A.z, __int_as_float(lj));
                b.sq_sphere.b[sl] = make_float4(B.x, B.y, B.z, 0.0f);
                b.sphere_acc[sl] = make_float2(0.0f, 0.0f);
            }
        }
        // Plane lights (getPlaneLights, shadow.cpp:255-321): a record for every hit in front of the light; its pl_rc^2
        // samples are traced by k_shadow_plane, k_plane_finalize turns the sums into a Lighting and calcColor.
        for (int lj = 0; lj < fp.n_plane; lj++) {
            const f3 ppos = mk3(__ldg(s.plane_lights + 4 * lj)), pw = mk3(__ldg(s.plane_lights + 4 * lj + 1));
            const f3 ph = mk3(__ldg(s.plane_lights + 4 * lj + 2)), pc = mk3(__ldg(s.plane_lights + 4 * lj + 3));
            bool front = false;
            if (hit) {
                const f3 nrm = xnormalize(xcross(pw, ph));
                const f3 centre = xadd(ppos, xmul(xadd(pw, ph), 0.5f)); // position + 0.5f * (width + height)
                front = xdot(xnormalize(xsub(sh.p, centre)), nrm) > 0.0f;
            }
            unsigned* const cl[1] = { &b.counters->sh[b.par].n_pl };
            const bool wl[1] = { front };
            const unsigned capl[1] = { b.plane_capacity };
            unsigned sl1[1];
            block_alloc<1>(cl, wl, capl, &b.counters->overflow, sl1, smem);
            if (sl1[0] == 0xffffffffu)
                continue;
            (void)RT_GUARD(sl1[0], b.plane_capacity, kChkQueue);
            b.sq_plane.p_pix[sl1[0]] = make_float4(sh.p.x, sh.p.y, sh.p.z, __int_as_float(pix));
            b.sq_plane.a_light[sl1[0]] = make_float4(w.x * kd.x * pc.x, w.y * kd.y * pc.y, w.z * kd.z * pc.z, __int_as_float(lj));
            b.sq_plane.b_shin[sl1[0]] = make_float4(w.x * pc.x * ks.x, w.y * pc.y * ks.y, w.z * pc.z * ks.z, shininess);
            b.sq_plane.refl[sl1[0]] = make_float4(reflN.x, reflN.y, reflN.z, 0.0f);
            b.sq_plane.acc[sl1[0]] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    warp_add_u64(&b.counters->secondary_rays, n_secondary);
}

// ---------------------------------------------------------------------------------------------------------
// cansee (src/shadow.cpp:32-69) as a per-lane state machine: closest-hit loop from p1 towards p2 with transparent
// pass-through.  `intensity` is attenuated by every transparent surface crossed (also when finally blocked, which
// getSpherelights relies on for its centre sample).
struct CanSee {
    f3 o, d;           // current segment origin, unit direction
    float distance;
    float intensity;
};

__device__ __forceinline__ bool cansee_begin(CanSee& cs, const f3& p1, const f3& p2)
{
    f3 d = xsub(p2, p1);
    cs.distance = xlength(d);
    cs.d = xnormalize(d);
    cs.o = xadd(p1, xmul(cs.d, 0.0005f));
    cs.intensity = 1.0f;
    return cs.distance > 0.0005f; // false: the loop body never runs, the light is visible (shadow.cpp:41,67)
}

// Search bound of one loop iteration: blockers have t <= distance - 2*SHADOW_ERROR_OFFSET; a closest hit farther
// than that means "visible" (shadow.cpp:44), so nothing beyond it needs to be found.
__device__ __forceinline__ HitRec cansee_query(const CanSee& cs) { return bounded_query(xsub(cs.distance, 0.001f)); }

// After a traversal finished with `best`: 0 = visible, 1 = blocked, 2 = passed a transparent surface, go on.
__device__ __forceinline__ int cansee_step(const SceneDev& s, const FrameParams& fp, CanSee& cs, const HitRec& best)
{
    if (best.ti == -1)
        return 0;
    if (!fp.any_transparent)
        return 1;
    const Shading sh = shading_at(s, best.ti, cs.o, cs.d, best.t);
    const float R0 = sh.m1.w;
    if (R0 == 1.0f)
        return 1;
    cs.distance = xsub(cs.distance, best.t);               // shadow.cpp:51
    cs.o = xadd(sh.p, xmul(cs.d, 0.0005f));                // shadow.cpp:53
    const float c = fabsf(xdot(cs.d, sh.N));                // shadow.cpp:55 (normal not re-normalised)
    cs.intensity = (float)((double)cs.intensity * (1.0 - schlick(R0, c)));
    return cs.distance > 0.0005f ? 2 : 0;
}

// Shared body of the two shadow kernels.  Work item j < n_items; item_begin(j, p1, p2) gives the segment to test,
// item_end(j, visible, intensity) consumes the result.  Each cansee loop iteration is one traversal of the engine; a
// transparent blocker makes the lane continue the same item from behind the surface.
template <bool ANYHIT, bool COUNT, typename Begin, typename End>
__device__ __forceinline__ void shadow_loop(const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, unsigned* cursor,
    unsigned n_items, Begin item_begin, End item_end)
{
    TraceStats st;
    CanSee cs;
    unsigned queries = 0;
    trace_queue<ANYHIT, COUNT>(
        s, root_entry, fp.exhaustive != 0, cursor, n_items, st,
        [&](unsigned item, f3& o, f3& d, HitRec& q) {
            f3 p1, p2;
            item_begin(item, p1, p2);
            if (!cansee_begin(cs, p1, p2)) {
                item_end(item, true, 1.0f);
                return false;
            }
            queries++;
            o = cs.o;
            d = cs.d;
            q = cansee_query(cs);
            return true;
        },
        [&](unsigned item, const HitRec& best, f3& o, f3& d, HitRec& q) {
            if (ANYHIT) { // every material is opaque (the launcher's condition for ANYHIT): the first blocker decides, no segment state to keep
                item_end(item, best.ti == -1, 1.0f);
                return false;
            }
            const int r = cansee_step(s, fp, cs, best);
            if (r == 2) {
                queries++;
                o = cs.o;
                d = cs.d;
                q = cansee_query(cs);
                return true;
            }
            item_end(item, r == 0, cs.intensity);
            return false;
        },
        RT_MAX_QUOTA, max(fp.min_quota, 1));
    warp_add_u64(&b.counters->shadow_queries, queries);
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
        atomicMax(&b.counters->max_ray_nodes, st.max_ray_nodes);
        atomicMax(&b.counters->max_ray_tris, st.max_ray_tris);
    }
}

// K4a shadow rays to point lights (getPointLights' cansee call, src/shadow.cpp:120).
// ANYHIT: every material is opaque, so the first blocker found decides; otherwise the closest hit does.
template <bool ANYHIT, bool COUNT>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, RT_TRACE_MIN_BLOCKS) k_shadow_point(SceneDev s, int root_entry, FrameParams fp, BatchDev b, int level)
{
    if (batch_overflowed(b))
        return;
    const unsigned n = min(b.counters->sh[b.par].n_pt, b.shadow_pt_capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0 && level < kLevelHistory)
        b.counters->level_sh[level] = n;
    shadow_loop<ANYHIT, COUNT>(
        s, root_entry, fp, b, &b.counters->sh[b.par].work_pt, n,
        [&](unsigned i, f3& p1, f3& p2) {
            i = RT_GUARD(i, b.shadow_pt_capacity, kChkQueue);
            const float4 pp = b.sq_point.p_pix[i];
            const float4 al = b.sq_point.a_light[i];
            p1 = mk3(pp);
            p2 = mk3(__ldg(&s.point_lights[3 * RT_GUARD(__float_as_int(al.w), s.n_point_like, kChkTable)]));
        },
        [&](unsigned i, bool visible, float intensity) {
            if (!visible)
                return;
            const float4 pp = b.sq_point.p_pix[i];
            const float4 al = b.sq_point.a_light[i];
            const float4 bb = b.sq_point.b[i];
            accumulate(b, __float_as_int(pp.w), al.x * intensity + bb.x, al.y * intensity + bb.y, al.z * intensity + bb.z);
        });
}

// ---------------------------------------------------------------------------------------------------------
// K2w / K4w: extend and opaque point-light shadow queries of a SMALL queue through the 8-wide tree with eight lanes per ray
// (rt_wide8.cuh): same inputs, same outputs, bit for bit the same hits — the triangle test and the tie rule are the binary
// engine's — at 2.3 times shorter a chain for the longest ray (tools/wide8_probe.py: 1 K / 16 K / 64 K / 256 K incoherent rays:
// 48 / 81 / 145 / 444 us against 111 / 139 / 161 / 257 us), which is what a small queue's kernel waits for.  Group g of the
// grid takes items g, g + groups, ...: a small queue needs no work cursor.  enqueue_frame picks these kernels for the levels whose
// queues held at most kWideMaxRays items in the previous frame.
__global__ void __launch_bounds__(kWideBlock, 8) k_extend_wide(SceneDev s, const float4* __restrict__ wide, int wide_root, BatchDev b, int qi, int level)
{
    __shared__ int s_stack[kWideBlock / kGroup][kWideStack];
    if (batch_overflowed(b))
        return;
    const unsigned n = min(b.counters->n_rays[qi], b.ray_capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0 && level < kLevelHistory)
        b.counters->level_ext[level] = n;
    const unsigned groups = gridDim.x * (kWideBlock / kGroup);
    for (unsigned item = blockIdx.x * (kWideBlock / kGroup) + threadIdx.x / kGroup; item < n; item += groups) {
        (void)RT_GUARD(item, b.ray_capacity, kChkQueue);
        const f3 o = mk3(b.q[qi].o_pix[item]), d = mk3(b.q[qi].d[item]);
        HitRec best = fresh_query();
        trace_wide<false>(s, wide, wide_root, o, d, best, s_stack[threadIdx.x / kGroup]);
        if ((threadIdx.x & (kGroup - 1)) == 0)
            b.q[qi].hit[item] = make_int2(__float_as_int(best.t), best.ti);
    }
}

__global__ void __launch_bounds__(kWideBlock, 8) k_shadow_point_wide(SceneDev s, const float4* __restrict__ wide, int wide_root, BatchDev b, int level)
{
    __shared__ int s_stack[kWideBlock / kGroup][kWideStack];
    if (batch_overflowed(b))
        return;
    const unsigned n = min(b.counters->sh[b.par].n_pt, b.shadow_pt_capacity);
    if (blockIdx.x == 0 && threadIdx.x == 0 && level < kLevelHistory)
        b.counters->level_sh[level] = n;
    const bool leader = (threadIdx.x & (kGroup - 1)) == 0;
    const unsigned groups = gridDim.x * (kWideBlock / kGroup);
    unsigned queries = 0;
    for (unsigned item = blockIdx.x * (kWideBlock / kGroup) + threadIdx.x / kGroup; item < n; item += groups) {
        (void)RT_GUARD(item, b.shadow_pt_capacity, kChkQueue);
        const float4 pp = b.sq_point.p_pix[item], al = b.sq_point.a_light[item];
        CanSee cs;
        bool visible = true;
        if (cansee_begin(cs, mk3(pp), mk3(__ldg(&s.point_lights[3 * RT_GUARD(__float_as_int(al.w), s.n_point_like, kChkTable)])))) {
            queries += leader ? 1u : 0u;
            HitRec best = cansee_query(cs);
            trace_wide<true>(s, wide, wide_root, cs.o, cs.d, best, s_stack[threadIdx.x / kGroup]);
            visible = best.ti == -1;
        }
        if (leader && visible) {
            const float4 bb = b.sq_point.b[item];
            accumulate(b, __float_as_int(pp.w), al.x * 1.0f + bb.x, al.y * 1.0f + bb.y, al.z * 1.0f + bb.z);
        }
    }
    warp_add_u64(&b.counters->shadow_queries, queries);
}

// ---------------------------------------------------------------------------------------------------------
// K2-4 fused, for the bounce levels of SMALL wavefronts: one lane follows one path from its level-`first_level` ray to the
// end — closest hit (BoundingVolumeHierarchy::intersect), the hit's shading terms (getFinalColor, main.cpp:129-190), one
// any-hit shadow query per point-like light (getPointLights / getSpotLichts with cansee, shadow.cpp:32-69, 106-131, 229-252),
// then the mirror ray (main.cpp:191-203) — instead of one kernel per level and ray kind.  A level's kernel cannot end before its
// longest ray has (a chain of ~200 dependent node fetches at L2 latency: 50-100 us however few rays there are), and the
// per-level form pays that once per level and ray kind: 8 times for depth 3.  Here the longest PATH counts once.  For
// wavefronts that fill the machine the per-level kernels are faster (coherent batches, all lanes in the same phase), so
// enqueue_frame picks this kernel for small batches only (one GPU's share of a frame split over several).  Conditions (host):
// every material opaque, glossy_ray_count 1, no spherical / plane lights, no textures — a path then never splits and a shadow
// query never continues.  Queries of both kinds share the engine (kAnyHitPerQuery); the path state lives in shared memory.
struct PathState {
    float p[3], Nn[3], reflN[3], refl[3], w[3], kd[3], ks[3];
    float shininess;
    int pix, level, li;   // li: light of the shadow query in flight; -1: the closest-hit query is
};

constexpr int kPathBlocksPerSm = 4; // 128 registers: no spills; this kernel runs where latency, not occupancy, sets the time
template <bool COUNT>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, kPathBlocksPerSm) k_paths(SceneDev s, int root_entry, FrameParams fp, BatchDev b, int qi, int first_level)
{
    __shared__ PathState ps_all[RT_TRACE_BLOCK];
    if (batch_overflowed(b))
        return;
    PathState& ps = ps_all[threadIdx.x];
    const unsigned n = min(b.counters->n_rays[qi], b.ray_capacity);
    TraceStats st;
    unsigned queries = 0, n_secondary = 0;
    auto ld3 = [](const float* a) { return mk3(a[0], a[1], a[2]); };
    auto st3 = [](float* a, const f3& v) { a[0] = v.x; a[1] = v.y; a[2] = v.z; };
    // After the closest hit has been shaded, or a shadow query answered: the next shadow query of the hit, else the mirror ray, else the end.
    auto next_query = [&](f3& o, f3& d, HitRec& q) -> bool {
        const f3 p = ld3(ps.p);
        for (int li = ps.li + 1; li < fp.n_point; li++) {
            if (!light_reaches(s, li, p))
                continue;
            CanSee cs;
            if (!cansee_begin(cs, p, mk3(__ldg(&s.point_lights[3 * li])))) { // the light sits on the hit: visible without a query (shadow.cpp:41,67)
                f3 A, B;
                light_terms(s, li, p, ld3(ps.Nn), ld3(ps.reflN), ld3(ps.w), ld3(ps.kd), ld3(ps.ks), ps.shininess, A, B);
                accumulate(b, ps.pix, A.x + B.x, A.y + B.y, A.z + B.z);
                continue;
            }
            queries++;
            ps.li = li;
            o = cs.o;
            d = cs.d;
            q = cansee_query(cs);
            return true;
        }
        const f3 ks = ld3(ps.ks);
        if (ps.level < fp.max_level && (ks.x > 0.0f || ks.y > 0.0f || ks.z > 0.0f)) { // main.cpp:187, 194: the mirror ray, weight ks * ks
            const f3 w = ld3(ps.w), refl = ld3(ps.refl);
            st3(ps.w, mk3(w.x * ks.x * ks.x, w.y * ks.y * ks.y, w.z * ks.z * ks.z));
            ps.level++;
            ps.li = -1;
            n_secondary++;
            o = xadd(p, xmul(refl, 0.01f)); // main.cpp:199
            d = refl;
            q = fresh_query();
            return true;
        }
        return false;
    };
    trace_queue<kAnyHitPerQuery, COUNT>(
        s, root_entry, false, &b.counters->work[0], n, st,
        [&](unsigned item, f3& o, f3& d, HitRec& q) {
            item = RT_GUARD(item, b.ray_capacity, kChkQueue);
            const float4 op = b.q[qi].o_pix[item];
            o = mk3(op);
            d = mk3(b.q[qi].d[item]);
            st3(ps.w, mk3(b.q[qi].w[item]));
            ps.pix = __float_as_int(op.w) >> 1;
            ps.level = first_level;
            ps.li = -1;
            q = fresh_query();
            return true;
        },
        [&](unsigned, const HitRec& best, f3& o, f3& d, HitRec& q) {
            if (ps.li >= 0) { // a shadow query came back
                if (best.ti == -1) {
                    f3 A, B;
                    light_terms(s, ps.li, ld3(ps.p), ld3(ps.Nn), ld3(ps.reflN), ld3(ps.w), ld3(ps.kd), ld3(ps.ks), ps.shininess, A, B);
                    accumulate(b, ps.pix, A.x * 1.0f + B.x, A.y * 1.0f + B.y, A.z * 1.0f + B.z);
                }
                return next_query(o, d, q);
            }
            if (best.ti == -1) // the ray left the scene: black (main.cpp:295-301)
                return false;
            const Shading sh = shading_at(s, best.ti, o, d, best.t);
            const f3 dn = xnormalize(d), Nn = xnormalize(sh.N), refl = xreflect(dn, Nn); // main.cpp:141
            st3(ps.p, sh.p);
            st3(ps.Nn, Nn);
            st3(ps.refl, refl);
            st3(ps.reflN, xnormalize(refl));
            st3(ps.kd, mk3(sh.m0));
            st3(ps.ks, mk3(sh.m1));
            ps.shininess = sh.m0.w;
            return next_query(o, d, q);
        },
        RT_MAX_QUOTA, max(fp.min_quota, 1));
    warp_add_u64(&b.counters->shadow_queries, queries);
    warp_add_u64(&b.counters->secondary_rays, n_secondary);
    if (COUNT) {
        warp_add_u64(&b.counters->node_visits, st.nodes);
        warp_add_u64(&b.counters->tri_tests, st.tris);
        warp_add_u64(&b.counters->tri_tests_full, st.tris_full);
        atomicMax(&b.counters->max_ray_nodes, st.max_ray_nodes);
        atomicMax(&b.counters->max_ray_tris, st.max_ray_tris);
    }
}

// K4b spherical lights (getSpherelights, src/shadow.cpp:139-226).  Work item = (record, sample): sample 0 is the
// light centre, the others lie on rings of the disc facing the hit point; their positions follow the reference's
// sequential `perp = rotate * perp`.  Per-record sums go to sphere_acc = {sum of intensities, visible count}.
template <bool ANYHIT, bool COUNT>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, RT_TRACE_MIN_BLOCKS) k_shadow_sphere(SceneDev s, int root_entry, FrameParams fp, BatchDev b)
{
    if (batch_overflowed(b))
        return;
    const unsigned rc = (unsigned)fp.sl_rc;
    const unsigned n = min(b.counters->sh[b.par].n_sp, b.shadow_sp_capacity / (unsigned)max(fp.n_sphere, 1)) * (unsigned)fp.n_sphere * rc;
    shadow_loop<ANYHIT, COUNT>(
        s, root_entry, fp, b, &b.counters->sh[b.par].work_sp, n,
        [&](unsigned j, f3& p1, f3& p2) {
            const unsigned rec = RT_GUARD(j / rc, b.shadow_sp_capacity, kChkQueue);
            const int k = (int)(j % rc);
            const float4 pp = b.sq_sphere.p_pix[rec];
            const float4 al = b.sq_sphere.a_light[rec];
            const float4 L = __ldg(&s.sphere_lights[2 * __float_as_int(al.w)]);
            const f3 p = mk3(pp), lpos = mk3(L);
            p1 = p;
            p2 = lpos; // centre sample (shadow.cpp:148)
            if (k == 0)
                return;
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
            f3 perp = xmul(xnormalize(xcross(d, notd)), radius); // shadow.cpp:169
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
            const int spoke = (k - 1) / fp.sl_m, ring = (k - 1) % fp.sl_m;
            for (int r = 0; r < spoke; r++) // perp = rotate * perp (shadow.cpp:207)
                perp = mk3(xadd(xadd(xmul(R[0][0], perp.x), xmul(R[1][0], perp.y)), xmul(R[2][0], perp.z)),
                    xadd(xadd(xmul(R[0][1], perp.x), xmul(R[1][1], perp.y)), xmul(R[2][1], perp.z)),
                    xadd(xadd(xmul(R[0][2], perp.x), xmul(R[1][2], perp.y)), xmul(R[2][2], perp.z)));
            const float frac = xdiv((float)(fp.sl_m - ring), (float)fp.sl_m); // (m-j)/(float)m
            p2 = xadd(lpos, xmul(perp, frac));
        },
        [&](unsigned j, bool visible, float intensity) {
            const unsigned rec = j / rc;
            const bool centre = (j % rc) == 0;
            // intensitySum starts at 1 and carries the centre ray's attenuation even if that ray ends up blocked
            // (shadow.cpp:145-150); ring samples count only when visible (shadow.cpp:201-204)
            float* acc = reinterpret_cast<float*>(&b.sphere_acc[rec]);
            if (visible || centre)
                atomicAdd(acc, intensity);
            if (visible)
                atomicAdd(acc + 1, 1.0f);
        });
}

// K4c: per spherical-light record, Lighting::intensity = intensitySum / rayCount if any sample saw the light
// (shadow.cpp:212-221), then calcColor's A * intensity + B.
__global__ void __launch_bounds__(256) k_sphere_finalize(FrameParams fp, BatchDev b)
{
    if (batch_overflowed(b))
        return;
    const unsigned n = min(b.counters->sh[b.par].n_sp, b.shadow_sp_capacity / (unsigned)max(fp.n_sphere, 1)) * (unsigned)fp.n_sphere;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 acc = b.sphere_acc[i];
        if (acc.y > 0.0f) {
            const float intensity = acc.x / (float)fp.sl_rc;
            const float4 pp = b.sq_sphere.p_pix[i];
            const float4 al = b.sq_sphere.a_light[i];
            const float4 bb = b.sq_sphere.b[i];
            accumulate(b, __float_as_int(pp.w), al.x * intensity + bb.x, al.y * intensity + bb.y, al.z * intensity + bb.z);
        }
    }
}

// K4d plane (area) lights (getPlaneLights, src/shadow.cpp:255-321).  Work item = (record, sample): sample (i, j) sits at
// position + i*dy + j*dx, reached by the reference's sequential `py += dy`, `px += dx`.  Per record the kernel sums the
// intensities of the visible samples, the terms max(dot(normalize(p - px), normal), 0) / length(p - px), their number, and
// the largest cosine between the reflection direction and a visible sample.
template <bool ANYHIT, bool COUNT>
__global__ void __launch_bounds__(RT_TRACE_BLOCK, RT_TRACE_MIN_BLOCKS) k_shadow_plane(SceneDev s, int root_entry, FrameParams fp, BatchDev b)
{
    if (batch_overflowed(b))
        return;
    const unsigned rc = (unsigned)fp.pl_rc, per = rc * rc;
    const unsigned n = min(b.counters->sh[b.par].n_pl, b.plane_capacity) * per;
    f3 sample = mk3(0, 0, 0); // position of the sample the lane is tracing
    shadow_loop<ANYHIT, COUNT>(
        s, root_entry, fp, b, &b.counters->sh[b.par].work_pl, n,
        [&](unsigned j, f3& p1, f3& p2) {
            const unsigned rec = RT_GUARD(j / per, b.plane_capacity, kChkQueue), k = j % per;
            const int lj = __float_as_int(b.sq_plane.a_light[rec].w);
            const f3 ppos = mk3(__ldg(s.plane_lights + 4 * lj)), pw = mk3(__ldg(s.plane_lights + 4 * lj + 1)), ph = mk3(__ldg(s.plane_lights + 4 * lj + 2));
            const float step = xdiv(1.0f, (float)(fp.pl_rc - 1)); // 1.0f / (rayCount1D - 1)
            const f3 dx = xmul(pw, step), dy = xmul(ph, step);
            f3 px = ppos;
            for (unsigned a = 0; a < k / rc; a++)
                px = xadd(px, dy);
            for (unsigned a = 0; a < k % rc; a++)
                px = xadd(px, dx);
            sample = px;
            p1 = mk3(b.sq_plane.p_pix[rec]);
            p2 = px;
        },
        [&](unsigned j, bool visible, float intensity) {
            if (!visible)
                return;
            const unsigned rec = j / per;
            const int lj = __float_as_int(b.sq_plane.a_light[rec].w);
            const f3 nrm = xnormalize(xcross(mk3(__ldg(s.plane_lights + 4 * lj + 1)), mk3(__ldg(s.plane_lights + 4 * lj + 2))));
            const f3 p = mk3(b.sq_plane.p_pix[rec]), rn = mk3(b.sq_plane.refl[rec]);
            const f3 back = xsub(p, sample);
            const float term = xdiv(fmaxf(xdot(xnormalize(back), nrm), 0.0f), xlength(back));
            const float cosr = xdot(rn, xnormalize(xsub(sample, p)));
            float* acc = reinterpret_cast<float*>(&b.sq_plane.acc[rec]);
            atomicAdd(acc, intensity);
            atomicAdd(acc + 1, term);
            atomicAdd(acc + 2, 1.0f);
            atomicMax(reinterpret_cast<int*>(acc + 3), __float_as_int(fmaxf(cosr, 0.0f))); // maxCosAngle starts at 0
        });
}

// K4e: Lighting of a plane light (shadow.cpp:308-318) and calcColor (main.cpp:112-121) with cosLightSurfaceAngle = 1.
__global__ void __launch_bounds__(256) k_plane_finalize(FrameParams fp, BatchDev b)
{
    if (batch_overflowed(b))
        return;
    const unsigned n = min(b.counters->sh[b.par].n_pl, b.plane_capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 acc = b.sq_plane.acc[i];
        if (acc.y > 0.0f) {
            const float intensity = (acc.x / acc.z) * acc.y / (float)(fp.pl_rc * fp.pl_rc);
            const float4 pp = b.sq_plane.p_pix[i];
            const float4 al = b.sq_plane.a_light[i];
            const float4 bs = b.sq_plane.b_shin[i];
            const float sp = bs.w > 0.0f ? powf(acc.w, bs.w) : 0.0f;
            accumulate(b, __float_as_int(pp.w), al.x * intensity + bs.x * sp, al.y * intensity + bs.y * sp, al.z * intensity + bs.z * sp);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K5 resolve: sample average (src/main.cpp:374,384) and Screen::setPixel's y flip (src/screen.cpp:32-38).  One
// warp writes one 32-pixel tile row = 512 contiguous bytes of float4; `out` may be a peer-mapped framebuffer of
// another GPU (the gather of finished tiles fused into this store).
// row_flags (nullable, k_row_flags): sparse gather — tile rows none of whose camera rays hit anything are not stored; the gather
// root has pre-filled them with the background colour (rt_capi.cu, gather frames).
__global__ void __launch_bounds__(256) k_resolve(FrameParams fp, unsigned first_lp, unsigned n_lp, const float4* __restrict__ accum,
    const int* __restrict__ prim_id, const float* __restrict__ prim_t, float4* out, int* out_id, float* out_t, const unsigned char* __restrict__ row_flags)
{
    for (unsigned idx = first_lp + blockIdx.x * blockDim.x + threadIdx.x; idx < first_lp + n_lp; idx += gridDim.x * blockDim.x) {
        const unsigned j = idx / kTilePixels, k = idx % kTilePixels;
        const unsigned x = k % kTileW, y = k / kTileW; // row-major inside the tile for coalesced stores
        if (row_flags && !row_flags[(size_t)j * kTileH + y])
            continue;
        const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world;
        unsigned tx, ty;
        tile_xy(g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, tx, ty);
        const int px = (int)(tx * kTileW + x), py = (int)(ty * kTileH + y);
        if (px >= fp.W || py >= fp.H)
            continue;
        const unsigned lp = j * kTilePixels + (((y >> 2) * 4 + (x >> 3)) << 5) + ((y & 3) << 3) + (x & 7);
        const float4 a = accum[lp];
        const size_t o = RT_GUARD((size_t)(fp.H - 1 - py) * fp.W + px, (size_t)fp.W * fp.H, kChkPixel);
        out[o] = make_float4(a.x * fp.sample_scale, a.y * fp.sample_scale, a.z * fp.sample_scale, 1.0f);
        if (out_id) {
            const bool traced = pixel_sees_scene(fp, px, py);
            out_id[o] = traced ? prim_id[lp] : -1;
            out_t[o] = traced ? prim_t[lp] : FLT_MAX;
        }
    }
}

// float4 framebuffer -> packed float3 (the host Screen's std::vector<glm::vec3>) for pixels [p0, p1).  One thread packs
// a group of 4 pixels = 48 bytes = 3 aligned float4 stores; groups cut by the range ends fall back to scalar stores.
__global__ void __launch_bounds__(256) k_pack_rgb(const float4* __restrict__ in, float* __restrict__ out, size_t p0, size_t p1)
{
    const size_t g0 = p0 / 4, g1 = (p1 + 3) / 4;
    for (size_t g = g0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < g1; g += (size_t)gridDim.x * blockDim.x) {
        const size_t p = 4 * g;
        if (p >= p0 && p + 4 <= p1) {
            const float4 a = in[p], b = in[p + 1], c = in[p + 2], d = in[p + 3];
            float4* o = reinterpret_cast<float4*>(out + 3 * p);
            o[0] = make_float4(a.x, a.y, a.z, b.x);
            o[1] = make_float4(b.y, b.z, c.x, c.y);
            o[2] = make_float4(c.z, d.x, d.y, d.z);
        } else {
            for (size_t q = (p > p0 ? p : p0); q < p + 4 && q < p1; q++) {
                const float4 a = in[q];
                out[3 * q] = a.x;
                out[3 * q + 1] = a.y;
                out[3 * q + 2] = a.z;
            }
        }
    }
}

// The pixels of the tiles this rank owns, float4 framebuffer -> packed float3 in `out` (the whole image's buffer, Screen
// layout).  `out` may be page-locked host memory mapped into the device (rt_render_shard): the stores then cross PCIe
// directly, every rank over its own link, so they have to be whole lines: one warp takes one tile row (32 pixels), stages
// its 96 floats in shared memory and 24 lanes write them as float4, 384 contiguous bytes per store instruction (the first
// version, three 16-byte stores 48 bytes apart per thread, reached 8.6 GB/s over PCIe).  Rows cut by the image's right edge
// or not 16-byte aligned fall back to scalar stores.
__global__ void __launch_bounds__(256) k_pack_rgb_tiles(FrameParams fp, unsigned tile0, unsigned n_tiles, const unsigned char* __restrict__ flags,
    const float4* __restrict__ in, float* __restrict__ out)
{
    __shared__ __align__(16) float stage[8][3 * kTileW];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t n_rows = (size_t)n_tiles * kTileH;
    for (size_t r = (size_t)blockIdx.x * 8 + wib; r < n_rows; r += (size_t)gridDim.x * 8) {
        const unsigned j = tile0 + (unsigned)(r / kTileH), y = (unsigned)(r % kTileH);
        if (flags && !flags[(size_t)j * kTileH + y]) // rt_render, rows of background: already on the host (k_host_background)
            continue;
        const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world;
        unsigned tx, ty;
        tile_xy(g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, tx, ty);
        const int px0 = (int)(tx * kTileW), py = (int)(ty * kTileH + y);
        if (py >= fp.H)
            continue;
        const size_t p0 = RT_GUARD((size_t)(fp.H - 1 - py) * fp.W + px0, (size_t)fp.W * fp.H, kChkPixel);
        const bool inside = px0 + lane < fp.W;
        const float4 a = inside ? in[p0 + lane] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (px0 + (int)kTileW <= fp.W && ((3 * p0) & 3) == 0) {
            stage[wib][3 * lane] = a.x;
            stage[wib][3 * lane + 1] = a.y;
            stage[wib][3 * lane + 2] = a.z;
            __syncwarp();
            if (lane < 3 * (int)kTileW / 4)
                reinterpret_cast<float4*>(out + 3 * p0)[lane] = reinterpret_cast<const float4*>(stage[wib])[lane];
            __syncwarp();
        } else if (inside) {
            out[3 * (p0 + lane)] = a.x;
            out[3 * (p0 + lane) + 1] = a.y;
            out[3 * (p0 + lane) + 2] = a.z;
        }
    }
}

// rt_render into a device-mapped host image: which 32-pixel tile rows of the batch contain a pixel one of whose camera rays
// hit something.  The others are background — black, and final as soon as level 0 has been traced.
// One warp per tile row; flags[tile * kTileH + row] = 1 if any of its pixels was hit.
__global__ void __launch_bounds__(256) k_row_flags(FrameParams fp, unsigned first_lp, unsigned n_lp, const int2* __restrict__ hit, unsigned char* __restrict__ flags,
    unsigned* n_flagged)
{
    __shared__ unsigned s_flagged;
    if (threadIdx.x == 0)
        s_flagged = 0;
    __syncthreads();
    unsigned mine = 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned tile0 = first_lp / kTilePixels;
    const size_t n_rows = (size_t)(n_lp / kTilePixels) * kTileH;
    for (size_t r = (size_t)blockIdx.x * 8 + wib; r < n_rows; r += (size_t)gridDim.x * 8) {
        const unsigned jl = (unsigned)(r / kTileH), y = (unsigned)(r % kTileH);
        const unsigned g = (unsigned)fp.rank + (tile0 + jl) * (unsigned)fp.world;
        unsigned tx, ty;
        tile_xy(g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, tx, ty);
        const int px = (int)(tx * kTileW) + lane, py = (int)(ty * kTileH + y);
        const unsigned k = ((((y >> 2) * 4 + ((unsigned)lane >> 3)) << 5) + ((y & 3) << 3) + ((unsigned)lane & 7)); // slot of pixel (lane, y) in its tile
        bool was_hit = false;
        if (pixel_sees_scene(fp, px, py)) {
            const int2* h = hit + ((size_t)jl * kTilePixels + k) * (size_t)fp.spp; // the samples of a pixel are neighbours in the queue
            for (int sidx = 0; sidx < fp.spp; sidx++)
                was_hit |= h[sidx].y != -1;
        }
        const bool any = __any_sync(0xffffffffu, was_hit);
        if (lane == 0) {
            flags[(size_t)(tile0 + jl) * kTileH + y] = any ? 1 : 0;
            mine += any ? 1u : 0u;
        }
    }
    if (lane == 0 && mine)
        atomicAdd(&s_flagged, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_flagged && n_flagged)
        atomicAdd(n_flagged, s_flagged);
}

// The background rows of the batch (k_row_flags) go to the host image at once, as zeros, while the rest of the frame is
// still being traced: 384 contiguous bytes per warp store over PCIe.  The kernel is PACED: left alone, its warps queue stores
// as fast as the LSU takes them, the path to system memory backs up into L2 and the traversal kernels running beside it take
// 70 % longer.  Every warp therefore sends its k-th row no earlier than k * period_ns after it started (%globaltimer), with the
// period chosen so that all warps together stay just below what PCIe carries away: nothing queues, the warps sleep in between.
// flags == nullptr: the rows that lie outside the scene's projection altogether (known before anything is traced); otherwise the
// other rows whose flag says that no camera ray hit anything.
__global__ void __launch_bounds__(256) k_host_background(FrameParams fp, unsigned tile0, unsigned n_tiles, const unsigned char* __restrict__ flags, float* __restrict__ out,
    unsigned period_ns)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t n_rows = (size_t)n_tiles * kTileH;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned sent = 0;
    for (size_t r = (size_t)blockIdx.x * 8 + wib; r < n_rows; r += (size_t)gridDim.x * 8) {
        const unsigned j = tile0 + (unsigned)(r / kTileH), y = (unsigned)(r % kTileH);
        if (flags && flags[(size_t)j * kTileH + y])
            continue;
        const unsigned g = (unsigned)fp.rank + j * (unsigned)fp.world;
        unsigned tx, ty;
        tile_xy(g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, tx, ty);
        const int px0 = (int)(tx * kTileW), py = (int)(ty * kTileH + y);
        if (py >= fp.H)
            continue;
        const bool outside = py < fp.vis_y0 || py >= fp.vis_y1 || px0 + (int)kTileW <= fp.vis_x0 || px0 >= fp.vis_x1;
        if (outside != (flags == nullptr))
            continue;
        for (;;) { // wait for this warp's slot
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            const long long wait = (long long)(t0 + (unsigned long long)sent * period_ns) - (long long)now;
            if (wait <= 0)
                break;
            __nanosleep((unsigned)min(wait, 20000ll));
        }
        sent++;
        const size_t p0 = RT_GUARD((size_t)(fp.H - 1 - py) * fp.W + px0, (size_t)fp.W * fp.H, kChkPixel);
        if (px0 + (int)kTileW <= fp.W && ((3 * p0) & 3) == 0) {
            if (lane < 3 * (int)kTileW / 4)
                reinterpret_cast<float4*>(out + 3 * p0)[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else if (px0 + lane < fp.W) {
            out[3 * (p0 + lane)] = 0.0f;
            out[3 * (p0 + lane) + 1] = 0.0f;
            out[3 * (p0 + lane) + 2] = 0.0f;
        }
    }
}

// FP32 issue-rate microbenchmark (rt_measure_fp32_peak): the geometry code of this path is un-fused FMUL / FADD (rt_math.cuh), which
// cannot reach the FFMA peak; this kernel measures what the chip issues for exactly that instruction mix — 8 independent chains per
// thread, alternating __fmul_rn / __fadd_rn, nothing else in the loop.  16 * iters FP32 instructions per thread.
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f, x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = __fmul_rn(x0, a); x1 = __fmul_rn(x1, a); x2 = __fmul_rn(x2, a); x3 = __fmul_rn(x3, a);
        x4 = __fmul_rn(x4, a); x5 = __fmul_rn(x5, a); x6 = __fmul_rn(x6, a); x7 = __fmul_rn(x7, a);
        x0 = __fadd_rn(x0, b); x1 = __fadd_rn(x1, b); x2 = __fadd_rn(x2, b); x3 = __fadd_rn(x3, b);
        x4 = __fadd_rn(x4, b); x5 = __fadd_rn(x5, b); x6 = __fadd_rn(x6, b); x7 = __fadd_rn(x7, b);
    }
    const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 12345.678f) // never true for the arguments used; keeps the chains alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

double launch_fp32_peak(cudaStream_t st, int sm_count, float* scratch, int iters)
{
    const int grid = sm_count * 8;
    k_fp32_peak<<<grid, 256, 0, st>>>(scratch, iters, 0.999f, 0.001f);
    return (double)grid * 256.0 * 16.0 * iters; // FP32 instructions (per lane) of the launch
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
            trace_bvh<false, false>(s, root_entry, o, d, best, st);
        } else {
            trace_exhaustive<false, false>(s, o, d, best, st);
        }
        tri_id[i] = global_id(s, best);
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

// Blocks per SM of the persistent traversal kernels: 8 fill the SM (64 registers x 128 threads); frames whose bands are
// rendered by several lanes at once ask for fewer, so that kernels of different lanes are resident side by side instead of
// queueing behind each other (rt_capi.cu, enqueue_frame).  RTB200_GRID_MULT overrides both for experiments.
static int trace_grid_mult(const FrameParams& fp)
{
    static const int forced = [] {
        const char* e = std::getenv("RTB200_GRID_MULT");
        const int v = e ? std::atoi(e) : 0;
        return v >= 1 && v <= 16 ? v : 0;
    }();
    return forced ? forced : (fp.trace_grid_mult > 0 ? fp.trace_grid_mult : RT_TRACE_GRID_MULT);
}

void launch_level_reset(cudaStream_t st, Counters* c, int next_q, long long n_current, unsigned long long add_primary, int par)
{
    k_level_reset<<<1, 1, 0, st>>>(c, next_q, n_current, add_primary, par);
}

void launch_extend(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int qi,
    int level, unsigned first_lp, bool count)
{
    const int grid = sm_count * trace_grid_mult(fp);
    if (level == 0) {
        if (count)
            k_extend<true, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_lp, level);
        else
            k_extend<true, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_lp, level);
    } else {
        if (count)
            k_extend<false, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_lp, level);
        else
            k_extend<false, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_lp, level);
    }
}

void launch_shade(cudaStream_t st, int sm_count, const SceneDev& s, const FrameParams& fp, const BatchDev& b, int qi, int level, unsigned first_lp)
{
    // the texture branch (diffuse_colour) and the glossy rays are compiled out of the kernels used by frames without them
    const bool extras = fp.tex_on || fp.glossy > 1 || fp.tex_debug;
    if (level == 0) {
        if (extras)
            k_shade<true, true><<<sm_count * 8, kShadeBlock, 0, st>>>(s, fp, b, qi, level, first_lp);
        else
            k_shade<true, false><<<sm_count * 8, kShadeBlock, 0, st>>>(s, fp, b, qi, level, first_lp);
    } else {
        if (extras)
            k_shade<false, true><<<sm_count * 4, kShadeBlock, 0, st>>>(s, fp, b, qi, level, first_lp);
        else
            k_shade<false, false><<<sm_count * 4, kShadeBlock, 0, st>>>(s, fp, b, qi, level, first_lp);
    }
}

void launch_shadow_point(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int level, bool count)
{
    const int grid = sm_count * trace_grid_mult(fp);
    const bool anyhit = !fp.any_transparent;
    if (anyhit && count)
        k_shadow_point<true, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, level);
    else if (anyhit)
        k_shadow_point<true, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, level);
    else if (count)
        k_shadow_point<false, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, level);
    else
        k_shadow_point<false, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, level);
}

void launch_extend_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int wide_root, const BatchDev& b, int qi, int level)
{
    k_extend_wide<<<sm_count * 8, kWideBlock, 0, st>>>(s, wide, wide_root, b, qi, level);
}

void launch_shadow_point_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int wide_root, const BatchDev& b, int level)
{
    k_shadow_point_wide<<<sm_count * 8, kWideBlock, 0, st>>>(s, wide, wide_root, b, level);
}

void launch_paths(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int qi, int first_level, bool count)
{
    const int grid = sm_count * kPathBlocksPerSm;
    if (count)
        k_paths<true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_level);
    else
        k_paths<false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b, qi, first_level);
}

void launch_shadow_sphere(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count)
{
    const int grid = sm_count * trace_grid_mult(fp);
    const bool anyhit = !fp.any_transparent;
    if (anyhit && count)
        k_shadow_sphere<true, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else if (anyhit)
        k_shadow_sphere<true, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else if (count)
        k_shadow_sphere<false, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else
        k_shadow_sphere<false, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    k_sphere_finalize<<<sm_count * 4, 256, 0, st>>>(fp, b);
}

void launch_shadow_plane(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count)
{
    const int grid = sm_count * trace_grid_mult(fp);
    const bool anyhit = !fp.any_transparent;
    if (anyhit && count)
        k_shadow_plane<true, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else if (anyhit)
        k_shadow_plane<true, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else if (count)
        k_shadow_plane<false, true><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    else
        k_shadow_plane<false, false><<<grid, RT_TRACE_BLOCK, 0, st>>>(s, root_entry, fp, b);
    k_plane_finalize<<<sm_count * 4, 256, 0, st>>>(fp, b);
}

void launch_resolve(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned first_lp, unsigned n_lp, const float4* accum,
    const int* prim_id, const float* prim_t, float4* out, int* out_id, float* out_t, const unsigned char* row_flags)
{
    if (n_lp == 0)
        return;
    k_resolve<<<grid_for(n_lp, 256, sm_count * 8), 256, 0, st>>>(fp, first_lp, n_lp, accum, prim_id, prim_t, out, out_id, out_t, row_flags);
}

// Background fill of a gather target: every pixel (0, 0, 0, 1), what k_resolve stores for a pixel without a hit.
__global__ void __launch_bounds__(256) k_fill_background(float4* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
}

void launch_fill_background(cudaStream_t st, int sm_count, float4* out, size_t n)
{
    if (n)
        k_fill_background<<<grid_for((long long)n, 256, sm_count * 4), 256, 0, st>>>(out, n);
}

void launch_pack_rgb(cudaStream_t st, int sm_count, const float4* in, float* out, size_t p0, size_t p1)
{
    if (p1 <= p0)
        return;
    k_pack_rgb<<<grid_for((long long)((p1 - p0 + 3) / 4 + 1), 256, sm_count * 8), 256, 0, st>>>(in, out, p0, p1);
}

void launch_pack_rgb_tiles(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned tile0, unsigned n_tiles, const unsigned char* flags,
    const float4* in, float* out)
{
    if (n_tiles == 0)
        return;
    k_pack_rgb_tiles<<<grid_for((long long)n_tiles * kTileH * 32, 256, sm_count * 8), 256, 0, st>>>(fp, tile0, n_tiles, flags, in, out);
}

void launch_row_flags(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned first_lp, unsigned n_lp, const int2* hit, unsigned char* flags,
    unsigned* n_flagged)
{
    if (n_lp < (unsigned)kTilePixels)
        return;
    k_row_flags<<<grid_for((long long)(n_lp / kTilePixels) * kTileH * 32, 256, sm_count * 8), 256, 0, st>>>(fp, first_lp, n_lp, hit, flags, n_flagged);
}

// One block on every other SM; the pacing period follows from the number of warps and the rate to hold (`gbs`: a little under what
// the link carries, measured by the caller — with every rank of a sharded job storing at once when gbs_is_shared_rate, else alone, and then capped by this
// rank's share of what the host is assumed to absorb; RTB200_BG_GBS / RTB200_BG_BLOCKS override for experiments).
void launch_host_background(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned tile0, unsigned n_tiles, const unsigned char* flags, float* out,
    double gbs, bool gbs_is_shared_rate)
{
    if (n_tiles == 0)
        return;
    static const int blocks_env = [] {
        const char* e = std::getenv("RTB200_BG_BLOCKS");
        return e ? std::atoi(e) : 0;
    }();
    static const double gbs_env = [] {
        const char* e = std::getenv("RTB200_BG_GBS");
        return e ? std::atof(e) : 0.0;
    }();
    // All ranks of a sharded frame store into the same host memory at once, and the host takes in only so much: 8 ranks at
    // 46 GB/s each back up again (C3 end to end 1.75 ms), at 12 GB/s each they do not (1.21 ms).  RTB200_HOST_GBS: what the host
    // is assumed to absorb in total.
    static const double host_gbs = [] {
        const char* e = std::getenv("RTB200_HOST_GBS");
        const double v = e ? std::atof(e) : 0.0;
        return v > 0.0 ? v : 100.0;
    }();
    if (!gbs_is_shared_rate) // no measurement with all ranks storing at once (rt_set_host_store_rate): assume
        gbs = std::min(gbs, host_gbs / std::max(1, fp.world));
    if (gbs_env > 0.0)
        gbs = gbs_env;
    const int grid = grid_for((long long)n_tiles * kTileH * 32, 256, blocks_env > 0 ? blocks_env : std::max(1, sm_count / 2));
    const double period = (double)grid * 8.0 * (3.0 * kTileW * 4.0) / std::max(gbs, 1.0); // ns between two rows of one warp
    k_host_background<<<grid, 256, 0, st>>>(fp, tile0, n_tiles, flags, out, (unsigned)std::max(1.0, period));
}

void launch_intersect(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const float* rays, long long n, int use_bvh,
    int* tri_id, float* t_out, unsigned* overflow)
{
    if (n <= 0)
        return;
    k_intersect<<<grid_for(n, 128, sm_count * 16), 128, 0, st>>>(s, root_entry, rays, n, use_bvh, tri_id, t_out, overflow);
}

#if RT_CHECKED
namespace {
__global__ void k_provoke_violation(int site) { (void)RT_GUARD(-1, 1, site); }
} // namespace
#endif

// One deliberate violation at site kChkTable from this translation unit (rt_violations_selftest): proves that what the kernels count is
// what add_violations_kernels reads.
void provoke_violation_kernels(cudaStream_t st)
{
#if RT_CHECKED
    k_provoke_violation<<<1, 1, 0, st>>>(kChkTable);
#else
    (void)st;
#endif
}

void add_violations_kernels(unsigned int* out)
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
