/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Common C interface of the two CPU checkers for the per-pixel ray-tracing hot path:
 *   - oracle/_ref/libref_oracle.so : the reference's own translation units (ray_tracing.cpp,
 *     bounding_volume_hierarchy.cpp, shadow.cpp) compiled verbatim from /root/reference, plus a
 *     harness restating the pieces that live next to GL/ImGui code (src/main.cpp:112-121,129-301,
 *     309-335,340-400; framework/src/trackball.cpp:65-68,87-98).           kind = "reference"
 *   - oracle/liboracle_port.so     : a plain C++ restatement of the whole path.   kind = "port"
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load these libraries.  The product (raytracer-group27_b200/) never does.
 */
#ifndef ORACLE_API_H
#define ORACLE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float kd[3];
    float ks[3];
    float shininess;
    float transparency;
} orc_material; /* src/mesh.h:21-31 without the texture */

typedef struct {
    float look_at[3];
    float euler[3]; /* radians, Trackball::m_rotationEulerAngles */
    float dist;
    float fovy; /* radians */
} orc_camera; /* framework/include/trackball.h:44-52 */

typedef struct {
    int width, height;          /* windowResolution (src/main.cpp:33), parameterised            */
    int max_reflection_level;   /* src/main.cpp:123                                             */
    int sphere_light_ray_count; /* src/main.cpp:124                                             */
    int glossy_ray_count;       /* src/main.cpp:126.  1: mirror ray only.  > 1 (port only; the reference kind
                                   refuses): glossy rays with the DEFINED random stream of oracle_port.cpp —
                                   the reference draws from rand() shared by its threads and has no
                                   reproducible answer there                                                  */
    float refraction_factor;    /* src/main.cpp:127                                             */
    int use_bvh;                /* global useBVH (src/main.cpp:60) for primary/secondary rays   */
    int sample_mode;            /* 0: 1 ray/pixel, 1: anti_aliasing 4-tap, 2: multipleRays      */
    int sample_size;            /* 4 / 16 / 64 for sample_mode 2                                */
    int defined_bary;           /* 1: define the uninitialised-barycentric case (port only)     */
    int x0, y0, x_step, y_step; /* render only pixels x0+i*x_step, y0+j*y_step (timing subsets) */
    int num_threads;            /* OpenMP threads, <=0: all                                     */
    int shadow_exhaustive;      /* port only. 0: every BVH search runs as in the reference (cansee always uses the
                                   BVH, shadow.cpp:42; the other rays when use_bvh).  1: those searches test every
                                   object, in the depth-first order intersectBVH would reach them (so equal t
                                   resolve to the same object, bvh.cpp:414-447), without the box tests.  The
                                   reference's AABB slab test (ray_tracing.cpp:213-264) occasionally culls a box
                                   whose triangle the triangle test would accept (flat boxes, rounding); the
                                   cull-free answer is what the triangle arithmetic and the visiting order alone
                                   define and what a conservative BVH (the GPU path) returns.  tri_id then also
                                   follows the visiting order on equal t (lowest id otherwise).                  */
    int texture_debug;          /* renderRayTracing's textureDebugging argument (src/main.cpp:75-106, 355-356): every pixel shows
                                   the texture colour of what its corner ray hits (white where the material has no texture,
                                   black on a miss), no lighting; textures as set by oracle_set_textures, whatever use_textures */
} orc_params;

typedef struct {
    uint64_t primary_rays;
    uint64_t shadow_queries;   /* one per iteration of the cansee loop (src/shadow.cpp:41-66)  */
    uint64_t secondary_rays;   /* reflection + refraction rays                                 */
    double seconds;            /* wall time of the pixel loop                                  */
    int threads;
} orc_stats;

/* Geometry is a triangle soup: pos / nrm hold 9 floats per triangle (3 corners x xyz) in global
 * triangle order; mesh_id[i] is the index of the owning mesh (non-decreasing); mats[mesh_id].
 * point_lights: 6 floats each (position, colour); sphere_lights: 7 floats (position, radius, colour).
 * Outputs may be NULL.  rgb uses the Screen layout (src/screen.cpp:32-38): row (H-1-y), column x.
 * tri_id / t_hit use the same layout and describe the FIRST primary ray of the pixel (the pixel-corner ray, or the
 * first sub-pixel sample when sample_mode != 0); tri_id = -1 and t = FLT_MAX on a miss. */
int oracle_render(const float* pos, const float* nrm, const int* mesh_id, int n_tris,
    const orc_material* mats, int n_mats,
    const float* point_lights, int n_point, const float* sphere_lights, int n_sphere,
    const orc_camera* cam, const orc_params* prm,
    float* rgb, int* tri_id, float* t_hit, orc_stats* stats);

/* Closest hit for caller-supplied rays (6 floats each: origin, direction), exhaustive (use_bvh=0,
 * reference brute-force order) or through the reference-style BVH (use_bvh=1); port only: use_bvh=2 = every
 * object in the BVH's visiting order, no box tests. */
int oracle_closest_hit(const float* pos, const float* nrm, const int* mesh_id, int n_tris,
    const float* rays, int n_rays, int use_bvh, int* tri_id, float* t_hit);

/* Sphere primitives (src/scene.h:48-53, intersectRayWithShape(Sphere) src/ray_tracing.cpp:182-209) used by the following
 * oracle_render / oracle_closest_hit calls of the same thread group: 12 floats each = centre (3), radius, kd (3), ks (3),
 * shininess, transparency.  n = 0 clears them.  A sphere that wins the closest-hit query is reported with
 * tri_id = n_tris + sphere index. */
void oracle_set_spheres(const float* spheres, int n);

/* Spot and plane (area) lights (src/scene.h:68-86, getSpotLichts / getPlaneLights src/shadow.cpp:229-321) for the following
 * oracle_render calls.  spot: 10 floats each = position (3), direction (3), angle in degrees, colour (3).  plane: 12 floats
 * each = position (3), width (3), height (3), colour (3).  plane_ray_count_1d = plane_light_1D_ray_count (src/main.cpp:125). */
void oracle_set_extra_lights(const float* spot, int n_spot, const float* plane, int n_plane, int plane_ray_count_1d);

/* Diffuse textures (src/image.cpp; getFinalColor's texture branch, src/main.cpp:155-171; texture coordinates interpolated in
 * intersectRayWithTriangleWithInterpolation, src/ray_tracing.cpp:166-169) for the following oracle_render calls.
 * tri_uv: 6 floats per triangle (u, v of its three corners) in global triangle order.  textures: 8-bit RGB rows, top row first
 * (what stbi_load hands to Image::Image, which divides by 255).  mesh_tex[m]: texture of mesh m or -1.
 * filtering: TextureFiltering (src/image.h:24-31) 0 NearestNeighbor, 1 Bilinear, 2 MipMappingNearestLevelNearestNeighbor,
 * 3 MipMappingNearestLevelBilinear, 4 Trilinear.  The mip-mapped modes take their level from ray differentials that the reference
 * initialises from not-yet-constructed members (framework/include/ray.h:19-28), i.e. from whatever the stack held; both checkers
 * give those members the values of their declarations, right = (1,0,0) and up = (0,-1,0) (ref_harness.cpp: defineDifferentials;
 * oracle_port.cpp: initialDifferentials), and run the reference's src/ray_differentials.cpp from there.  oob_x / oob_y: OutOfBoundsRule (src/image.h:18-22) 0 Border, 1 Clamp, 2 Repeat.  use_textures is main.cpp's useTextures (the
 * materials keep their textures either way, the texture-debug view reads them regardless); n_textures = 0 removes them. */
typedef struct {
    int width, height;
    const unsigned char* rgb;
} orc_texture;
void oracle_set_textures(const float* tri_uv, int n_tris, const orc_texture* textures, int n_textures, const int* mesh_tex, int n_meshes,
    int use_textures, int filtering, int oob_x, int oob_y, const float* border_rgb);

/* Screen post-processing, the step renderRayTracing ends with (src/main.cpp:397-398; src/screen.cpp:56-69, 226-395). */
typedef struct {
    int filtering_option;   /* FilteringOption (src/screen.h:17-26): 0 None, 1 Bloom, 2 BloomWithReinhardHdr, 3 BloomWithExposureHdr,
                               4 OnlyLight, 5 OnlyLightWithKernel                                                                 */
    int kernel;             /* Kernel (src/screen.h:28-31): 0 box, 1 Gaussian                                                    */
    int kernel_repetitions; /* setKernelNumRepetitions (max(1, n))                                                               */
    int filter_size;        /* setFilterSize: taps run over [-size, size]^2                                                      */
    float sigma;            /* setSigma (max(0.001, s))                                                                          */
    float exposure;         /* setExposure                                                                                       */
    int gamma_correction;   /* enableGammaCorrection                                                                             */
    float gamma;            /* setGammaValue                                                                                     */
    int bloom_live;         /* setBloomFilterLive: postprocessImage applies the bloom only when set                              */
} orc_post;

/* rgb: W*H*3 floats in the Screen layout, processed in place.  via_write_bitmap = 0: Screen::postprocessImage();
 * 1: what Screen::writeBitmapToFile does before writing (bloom whatever bloom_live says, no gamma) and, if rgba8 != NULL,
 * the W*H*4 bytes it hands to the BMP encoder (clamp to [0,1], * 255, truncate; alpha 255). */
int oracle_postprocess(float* rgb, int w, int h, const orc_post* p, int via_write_bitmap, unsigned char* rgba8);

const char* oracle_kind(void); /* "reference" or "port" */

#ifdef __cplusplus
}
#endif
#endif
