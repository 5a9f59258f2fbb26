// Shared host/device plain-data types of the B200 ray-tracing path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rtb {

// Image tiling: the unit of multi-GPU interleaving is a 32x16-pixel tile (512 pixels = 16 warps); inside a tile
// pixels are ordered so that one warp owns an 8x4 block (coherent primary rays).
constexpr int kTileW = 32;
constexpr int kTileH = 16;
constexpr int kTilePixels = kTileW * kTileH;

// Triangle records.  RT_TRI_AOS = 1: one 64-byte record {plane, v0, v1, v2} per triangle (and {n0, n1, n2, -} for the corner
// normals), so that everything a triangle test reads sits in one cache line; 0: four separate float4 arrays.  The pointers
// of SceneDev address element 0 of each member either way; element ti of a member is at [kTriStride * ti].
#ifndef RT_TRI_AOS
#define RT_TRI_AOS 1
#endif
constexpr int kTriStride = RT_TRI_AOS ? 4 : 1;

constexpr int kMaxLeafTris = 8;   // leaf size must fit the 3-bit count of a packed stack entry
constexpr int kStackDepth = 64;   // per-thread traversal stack (ints); builders guarantee depth < kStackDepth
constexpr int kLevelHistory = 8;  // bounce levels whose queue fills are remembered from frame to frame (Counters::level_ext / level_sh)
constexpr int kMaxLanes = 4;       // batches of one frame in flight at once (rt_set_pipeline)
constexpr int kMaxPointLights = 16;
constexpr int kMaxSphereLights = 8;
constexpr int kMaxSpheres = 64;

// RT_CHECKED = 1 (make EXTRA=-DRT_CHECKED=1 OUT=librtb200_checked.so BUILD=build_checked): the "checked build".  Every index the
// kernels form from data — BVH node and triangle indices, traversal stack slots (per-lane and per-group), accumulator and framebuffer
// pixels, texel addresses, queue slots — is tested against its bound first; a violation is counted per site (rt_violations) and the access goes to
// element 0 instead.  The pool these kernels are developed on has no compute-sanitizer; the GPU test-suite run against this build
// (RTB200_LIB) is its stand-in.  In the default build RT_GUARD is the identity and the SASS is unchanged.
#ifndef RT_CHECKED
#define RT_CHECKED 0
#endif
enum CheckSite { kChkNode = 0, kChkTri, kChkStack, kChkWideNode, kChkWideStack, kChkAccum, kChkPixel, kChkTexel, kChkTable, kChkQueue, kChkSites };

// Tile g of the image (0 <= g < tiles_x * tiles_y; rank r of a sharded job owns the tiles with g % world == r) lies in tile row
// g / tiles_x, and inside that row the tiles are rotated by 3 columns per row.  Plain row-major numbering would hand every rank
// whole tile COLUMNS whenever tiles_x is a multiple of the world size (a 3840-wide frame has 120 tile columns: with 8 ranks, rank r
// would own the columns r, r + 8, ...), and columns that cross the object more often make their rank the slowest (measured on C3, 8
// ranks: 0.70-0.79 ms on seven ranks, 0.96 ms on the eighth).  With the rotation a rank's tiles form a diagonal lattice.
#if defined(__CUDACC__)
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif
RT_HD inline void tile_xy(unsigned g, unsigned tiles_x, unsigned rot, unsigned& tx, unsigned& ty)
{
    ty = g / tiles_x;
    tx = (g % tiles_x + rot * ty) % tiles_x;
}

// Flattened BVH in HBM: 32-byte nodes {lo.xyz, entry}{hi.xyz, count}, read as 2 x float4.
// Sibling nodes are adjacent (2k, 2k+1) so visiting an inner node is one aligned 64-byte fetch of both children.
// count == 0: inner node, entry = index of the left child (even, >= 0); count > 0: leaf over triangles
// [first, first + count) of the BVH-ordered triangle arrays, entry = ~((first << 3) | (count - 1)) (< 0) — i.e.
// `entry` is exactly what the traversal pushes on its stack.  Node 0 is the root, node 1 a copy of it, so that a
// root that is itself a leaf can be fetched like any other pair.
struct SceneDev {
    const float4* nodes;
    const float4* tri_plane; // {n.xyz, D}: trianglePlane() precomputed by K0 with the reference's op order
    const float4* tri_v0;    // {v0.xyz, bits(tie key: visiting rank in the reference's BVH, rt_reforder.cu)}
    const float4* tri_v1;    // {v1.xyz, bits(mesh id)}
    const float4* tri_v2;    // {v2.xyz, bits(global triangle id)}
    const float4* tri_n0;    // corner normals, read only by shade / transparent shadow hits
    const float4* tri_n1;
    const float4* tri_n2;
    const float4* mats;          // 2 x float4 per mesh: {kd.xyz, shininess}{ks.xyz, transparency}
    const float4* point_lights;  // point and spot lights, 3 x float4: {pos, kind 0|1}{color, cos(cut-off)}{spot direction, 0}
    const float4* plane_lights;  // 4 x float4: {position}{width}{height}{color}
    const float4* sphere_lights; // 2 x float4: {pos, radius}{color, 0}
    const float4* spheres;       // sphere primitives, 3 x float4: {centre, radius}{kd, shininess}{ks, transparency}
    // diffuse textures (src/image.cpp): texels of all textures back to back, top row first; table {first texel, width, height, 1 if it has a mip pyramid}
    const float4* tex_texels;
    const int4* tex_table;
    const int* mat_tex;          // texture of material (mesh) m, or -1
    const float2* tri_uv;        // Vertex::texCoord of the three corners of global triangle g at [3g .. 3g+2]; null = all zero
    // glossy rays: the cone half-width d of main.cpp:224, evaluated on the host per material / per sphere primitive
    const float* mat_glossy_d;
    const float* sphere_glossy_d;
    const int* sphere_rank;      // visiting rank of sphere k in the reference's BVH (tie key, rt_reforder.cu)
    int tie_by_id;               // tie key of this launch: 0 = visiting rank in the reference's BVH (every BVH search of the reference, i.e. all
                                 // shadow queries and, with useBVH, the other rays), 1 = global id (its useBVH = false loop)
    int n_spheres;
    int sphere_id_base;          // global id of sphere 0 = number of triangles the caller uploaded
    int n_tris;
    int n_nodes;
#if RT_CHECKED
    int n_wide_nodes;            // nodes of the 8-wide tree (rt_wide8.cu)
    int n_mats, n_point_like;    // entries of the material and point-like light tables
#endif
};

struct FrameParams {
    int W, H;
    int sample_mode; // 0 single, 1 four-tap AA, 2 multipleRays
    int sample_size;
    int spp;         // rays per pixel
    float sample_scale; // what the summed colour is multiplied with (1, 0.25, 1/sample_size)
    float aa_off_x, aa_off_y;   // main.cpp:360-361
    float ms_off_x, ms_off_y;   // main.cpp:311-312 (evaluated in double on the host, as the reference does)
    int ms_moves;               // main.cpp:324
    // camera (quaternion, origin and half extents are computed on the host with libm, like the reference)
    float qx, qy, qz, qw;
    float ox, oy, oz;
    float halfW, halfH;
    // sharding
    int tiles_x, tiles_y;
    int rank, world;
    int n_local_tiles;
    // shading
    int max_level;
    float refraction;
    int n_point, n_sphere;      // n_point counts point AND spot lights (one shadow segment per light at most)
    int n_plane;                // plane (area) lights
    int pl_rc;                  // plane_light_1D_ray_count (src/main.cpp:125): pl_rc x pl_rc samples per plane light
    int any_transparent;
    int exhaustive;
    // diffuse textures: useTextures and the knobs of src/main.cpp:54-58
    int tex_on, tex_filter, tex_oob_x, tex_oob_y;
    int tex_debug;      // texture-debug view (rt_params.texture_debug)
    int tex_available;  // the context holds textures (rt_set_textures)
    float tex_border_r, tex_border_g, tex_border_b;
    int glossy;      // glossy_ray_count (src/main.cpp:126): 1 = mirror ray only
    int tie_by_id;   // camera / reflection rays: equal t go to the lower global id (useBVH = false) instead of the BVH visiting rank
    // spherical-light ring sampling (shadow.cpp:190-196), host-computed and shared with the oracle
    int sl_m, sl_n, sl_rc;
    float sl_sin, sl_omc;
    int sl_group; // lanes cooperating on one (hit, light) pair: smallest power of two >= sl_rc, at most 32
    // Pixels [vis_x0, vis_x1) x [vis_y0, vis_y1) are the only ones whose camera rays can meet the scene (projection of its
    // bounding box, rt_capi.cu); the other primary rays are misses without being traced.  Whole image when unknown.
    int vis_x0, vis_x1, vis_y0, vis_y1;
    int tile_rot;        // columns by which the tiles of a tile row are rotated per row (tile_xy)
    int min_quota;       // fewest rays a warp of the traversal kernels takes per refill (small queues: fewer, fuller warps)
    int trace_grid_mult; // host side only: blocks per SM of the traversal kernels of this frame (0: the default, rt_kernels.cu)
};

// Device-resident counters of one wavefront batch.
struct Counters {
    unsigned int n_rays[2];      // ray queue fill (ping-pong)
    unsigned int work[2];        // dynamic work-fetch cursors: extend, (spare)
    unsigned int overflow;       // set when a queue would exceed its capacity
    unsigned int flagged_rows;   // tile rows with a hit counted by k_row_flags (frames that store rows selectively)
    // Shadow work is double-buffered by bounce-level parity so that the shadow kernels of level L (side stream) can run
    // concurrently with extend / shade of level L+1 (main stream).
    struct Shadow {
        unsigned int n_pt;       // point / spot light shadow records (one per hit and light inside whose cone the hit lies)
        unsigned int n_pl;       // plane-light records (one per hit in front of the light)
        unsigned int work_pl;
        unsigned int n_sp;       // hits with spherical-light records (n_sphere consecutive records per hit)
        unsigned int work_pt;    // work-fetch cursors of the two shadow kernels
        unsigned int work_sp;
    } sh[2];
    unsigned long long primary_rays;
    unsigned long long shadow_queries;
    unsigned long long secondary_rays;
    unsigned long long node_visits;
    unsigned long long tri_tests;
    unsigned long long tri_tests_full;
    unsigned long long ext_node_visits; // the same three, counted by the extend kernel alone
    unsigned long long ext_tri_tests;
    unsigned long long ext_tri_tests_full;
    unsigned int max_ray_nodes, max_ray_tris; // instrumented builds: most boxes / triangles one query touched
    // Queue fills of the batch the lane rendered last, by bounce level (levels past the table are not recorded): extend rays and
    // point-like shadow records.  The host reads them with the counters and uses them as the expected sizes of the NEXT frame's
    // queues, to pick the traversal form (one lane per ray / eight lanes per ray, rt_wide8.cuh) of every level.
    unsigned int level_ext[kLevelHistory], level_sh[kLevelHistory];
};

struct RayQueue {
    float4* o_pix; // {origin.xyz, bits(local pixel index)}
    float4* d;     // {direction.xyz, 0}
    float4* w;     // {throughput.rgb, 0}
    int2* hit;     // {bits(t), BVH-order triangle index or -1}
};

struct ShadowQueue {
    float4* p_pix;   // {hit point, bits(local pixel index)}
    float4* a_light; // {A.rgb (scaled by the shadow intensity), bits(light index)}
    float4* b;       // {B.rgb (added when the light is visible), 0}
};

struct PlaneQueue { // one record per (hit, plane light); the samples of a record share it
    float4* p_pix;   // {hit point, bits(local pixel index)}
    float4* a_light; // {throughput * kd * light colour, bits(light index)}
    float4* b_shin;  // {throughput * light colour * ks, shininess}
    float4* refl;    // {normalize(reflect direction), 0}
    float4* acc;     // {sum of intensities, sum of cos/length terms, visible samples, bits(max cos to the reflection)}
};

struct BatchDev {
    RayQueue q[2];
    ShadowQueue sq_point, sq_sphere;
    PlaneQueue sq_plane;
    unsigned int plane_capacity;
    unsigned int ray_capacity;
    unsigned int shadow_pt_capacity;
    unsigned int shadow_sp_capacity;
    Counters* counters;
    int par;          // bounce-level parity selecting counters->sh[par]; the shadow queue pointers above are set to match
    float2* sphere_acc; // per spherical-light record: {sum of sample intensities, number of visible samples}
    float4* accum;    // per local padded pixel, summed radiance
    int* prim_id;     // nullable: closest-hit global triangle id of the first primary ray of each local pixel
    float* prim_t;    // nullable
#if RT_CHECKED
    unsigned int accum_pixels; // entries of accum / prim_id / prim_t
    unsigned int hit_capacity; // entries of q[].hit (level 0 keeps a slot per primary ray whatever ray_capacity is)
#endif
};

} // namespace rtb
