/* rt_b200.h — C ABI of the B200-native per-pixel ray-tracing hot path.
 *
 * This is the drop-in boundary for catalinlup/RayTracer-Group27's render path.  The reference has no
 * FFI of its own: its seam is the free function renderRayTracing(Scene&, const Trackball&, const
 * BoundingVolumeHierarchy&, Screen&, ...) (src/main.cpp:340-341) plus the classes it touches.  The C++
 * host layer in raytracer-group27_b200/host/ keeps those class names and forwards to the entry points
 * below; everything device-side (sm_100a kernels, device memory, streams) stays behind this header.
 * Plain pointers and sizes only, no C++ or torch types.  Every function returns RT_OK (0) or an error
 * code; rt_last_error() returns a thread-local message for the last failure.  There is no CPU fallback:
 * without a usable CUDA device rt_create fails.
 */
#ifndef RT_B200_H
#define RT_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_ERR_INVALID 1 /* bad argument / call order                     */
#define RT_ERR_CUDA 2    /* CUDA runtime error (message in rt_last_error) */
#define RT_ERR_OVERFLOW 3 /* a ray queue overflowed its capacity          */
#define RT_ERR_IO 4      /* file not found / unparsable                   */

typedef struct rt_ctx rt_ctx; /* one per GPU (one process per GPU); owns all device buffers */

/* Material of a mesh — replaces `struct Material` (src/mesh.h:21-31) minus the texture. */
typedef struct {
    float kd[3];
    float ks[3];
    float shininess;
    float transparency; /* 1 = opaque; != 1 takes the dielectric branch (src/main.cpp:257-290) */
} rt_material;

typedef struct { float position[3]; float color[3]; } rt_point_light;                /* src/scene.h:55-59 */
typedef struct { float position[3]; float radius; float color[3]; } rt_sphere_light; /* src/scene.h:61-66 */
typedef struct { float position[3]; float direction[3]; float angle /* degrees */; float color[3]; } rt_spot_light; /* src/scene.h:68-73 */
typedef struct { float position[3]; float width[3]; float height[3]; float color[3]; } rt_plane_light;               /* src/scene.h:75-83 */

/* Sphere primitive — replaces `struct Sphere` (src/scene.h:48-53) with its own material; intersected with the
 * reference's quadratic (src/ray_tracing.cpp:182-209).  A sphere hit is reported as triangle id n_tris + sphere index. */
typedef struct { float center[3]; float radius; rt_material material; } rt_sphere;

/* Camera state — replaces the Trackball members read by position()/generateRay()
 * (framework/include/trackball.h:44-52, framework/src/trackball.cpp:65-68,87-98).  The aspect ratio is
 * width/height of rt_params, as Window::aspectRatio() (framework/src/window.cpp:334-337). */
typedef struct {
    float look_at[3];
    float euler[3]; /* radians */
    float dist;
    float fovy;     /* radians */
} rt_camera;

/* Per-frame knobs — replace the file-scope globals renderRayTracing reads (src/main.cpp:58-60,123-127)
 * and its arguments (src/main.cpp:340-341). */
typedef struct {
    int width, height;          /* windowResolution, src/main.cpp:33 (parameterised)                    */
    int max_reflection_level;   /* src/main.cpp:123                                                     */
    int sphere_light_ray_count; /* src/main.cpp:124                                                     */
    int glossy_ray_count;       /* src/main.cpp:126, 1..40.  1: mirror ray only — the setting every parity bar is
                                   defined on.  > 1: glossy rays (main.cpp:204-250) whose directions come from a
                                   DEFINED counter-based stream (csrc/rt_kernels.cu, path_uniform): the reference
                                   draws them from rand(), shared by its threads, and has no reproducible answer */
    float refraction_factor;    /* src/main.cpp:127                                                     */
    int sample_mode;            /* 0: one ray per pixel; 1: anti_aliasing (4 taps); 2: multipleRays     */
    int sample_size;            /* 4 / 16 / 64 when sample_mode == 2                                    */
    int use_bvh;                /* the reference's useBVH (src/main.cpp:60).  It never changes WHAT is hit, only which of
                                   two objects at exactly the same t a camera / reflection ray reports: 1 = the one
                                   its BVH visits first (bounding_volume_hierarchy.cpp:414-447), 0 = the one earlier
                                   in mesh order (ibid. 51-72).  Shadow queries always use the BVH order
                                   (shadow.cpp:42).  The device searches through its own BVH either way.      */
    int exhaustive;             /* debugging aid: 1 = test every object for every ray instead of walking the device
                                   BVH (same result, O(n) per ray)                                           */
    int plane_light_ray_count_1d; /* src/main.cpp:125 (default 3): n x n samples per plane light; values < 2 mean 3 */
    int texture_debug;          /* renderRayTracing's textureDebugging argument (src/main.cpp:75-106, 355-356): every pixel
                                   shows the texture colour of what its corner ray hits — white where the material has no
                                   texture, black on a miss — without lighting, sampling modes or bounces; it reads the
                                   textures and the knobs of rt_set_texturing even when texturing is off for rendering   */
} rt_params;

typedef struct {
    uint64_t primary_rays;
    uint64_t shadow_queries;  /* one per cansee loop iteration (src/shadow.cpp:41-66) */
    uint64_t secondary_rays;  /* reflection + refraction rays                        */
    uint64_t node_visits;     /* 32-byte BVH nodes fetched (0 unless rt_set_counters is on) */
    uint64_t tri_tests;       /* triangles fetched                                          */
    uint64_t tri_tests_full;  /* tests that also ran the three edge functions               */
    float gpu_ms;             /* device time of the frame (CUDA events on the context's stream) */
    int kernel_launches;      /* kernels launched for the frame                                 */
    int batches;              /* wavefront batches the frame was split into                     */
    uint64_t extend_node_visits;    /* the three traversal counters restricted to the extend kernel (primary + */
    uint64_t extend_tri_tests;      /* reflection / refraction rays); the shadow kernels account for the rest  */
    uint64_t extend_tri_tests_full;
    uint64_t traced_primary_rays;   /* primary rays that walked the BVH: those of pixels inside the projection of the scene's bounding
                                       box; the others are counted in primary_rays (the reference casts them) but answered as misses */
    uint64_t gather_bytes;          /* bytes this rank stored into a framebuffer that is not its own (rt_render_device with a
                                       caller-supplied / peer-mapped d_rgba): float4 pixels of its tiles, or of their rows with hits only */
} rt_stats;

#define RT_BVH_LBVH_DEVICE 0      /* Morton codes -> radix sort -> Karras hierarchy -> refit, all on the GPU */
#define RT_BVH_SAH_HOST 1         /* binned SAH built by the host and uploaded                               */
#define RT_BVH_AUTO 2             /* PLOC_DEVICE                                                              */
#define RT_BVH_PLOC_DEVICE 3      /* Morton order -> parallel locally-ordered clustering (nearest neighbours by merged
                                     surface area) -> leaf collapse, all on the GPU: SAH quality at device speed      */

/* ---- lifetime ---- */
int rt_create(int device, rt_ctx** out);
int rt_destroy(rt_ctx* ctx);
/* Launch all work of this context on `cuda_stream` (a cudaStream_t, e.g. torch's current stream). */
int rt_set_stream(rt_ctx* ctx, void* cuda_stream);

/* ---- scene (replaces BoundingVolumeHierarchy::BoundingVolumeHierarchy(Scene*), which copies every
 *      triangle out of the scene: bounding_volume_hierarchy.cpp:5-9,80-99) ----
 * pos / nrm: 9 floats per triangle (3 corners x xyz) in the reference's global triangle order (meshes in
 * Scene::meshes order, triangles in Mesh::triangles order); mesh_id[i] indexes `mats`. */
int rt_upload_scene(rt_ctx* ctx, const float* pos, const float* nrm, const int* mesh_id, int64_t n_tris,
    const rt_material* mats, int n_mats);
int rt_build_bvh(rt_ctx* ctx, int mode);
/* Lights and materials are read live from the Scene every frame in the reference (src/shadow.cpp:111,141;
 * src/ray_tracing.h:23-27): re-upload them before a render when they changed. */
int rt_set_materials(rt_ctx* ctx, const rt_material* mats, int n_mats);
int rt_set_lights(rt_ctx* ctx, const rt_point_light* point, int n_point, const rt_sphere_light* sphere, int n_sphere);

/* Diffuse textures: Material::kdTexture (src/mesh.h:29), Vertex::texCoord (src/mesh.h:18) and the texture branch of
 * getFinalColor (src/main.cpp:155-171) with Image::getPixel (src/image.cpp:75-108). */
#define RT_TEX_NEAREST 0   /* TextureFiltering (src/image.h:24-31): NearestNeighbor                                   */
#define RT_TEX_BILINEAR 1  /* Bilinear                                                                                 */
#define RT_TEX_MIP_NEAREST 2    /* MipMappingNearestLevelNearestNeighbor */
#define RT_TEX_MIP_BILINEAR 3   /* MipMappingNearestLevelBilinear        */
#define RT_TEX_TRILINEAR 4      /* Trilinear                             */
/* The three mip-mapped filters sample the pyramid Image::initMipmap builds for square power-of-two textures (src/image.cpp:377-430;
 * built by rt_set_textures; textures without one answer white, Trilinear black) at the level of detail computeLevelOfDetails
 * derives from the ray differentials at the hit (src/ray_differentials.cpp:5-16, 38-88, 121-139; main.cpp:137, 168).  One thing is
 * DEFINED: Ray's default member initialisers of dD_dx / dD_dy read the members `right` and `up` before these are constructed
 * (framework/include/ray.h:19-28) — undefined behaviour in the reference; here they have the values their declarations give them,
 * right = (1,0,0), up = (0,-1,0).  Everything else follows the reference to the letter, including that camera rays evaluate those
 * defaults on the default direction (0,0,-1) (Trackball::generateRay fills a default-constructed Ray) and that reflection /
 * refraction rays are fresh Ray objects that inherit no differentials. */
#define RT_OOB_BORDER 0    /* OutOfBoundsRule, src/image.h:18-22 */
#define RT_OOB_CLAMP 1
#define RT_OOB_REPEAT 2
typedef struct {
    int width, height;
    const float* rgb;      /* width*height*3 floats, top row first: Image::m_pixels, i.e. byte / 255.0f (src/image.cpp:57-59) */
} rt_texture;
typedef struct {
    int filtering;                          /* textureFiltering, src/main.cpp:54 */
    int out_of_bounds_x, out_of_bounds_y;   /* outOfBoundsRuleX / Y, src/main.cpp:55-56 */
    float border_color[3];                  /* textureBorderColor, src/main.cpp:57 */
} rt_texture_params;
/* Texture coordinates of the uploaded triangles: 6 floats per triangle (u, v of its three corners), global order.
 * NULL: all zero (what loadMesh stores for meshes without them, src/mesh.cpp:113-117).  Call after rt_upload_scene. */
int rt_set_texcoords(rt_ctx* ctx, const float* uv);
/* The scene's textures and, per material (mesh), which one it uses (-1: none).  n_textures = 0 removes them. */
int rt_set_textures(rt_ctx* ctx, const rt_texture* textures, int n_textures, const int* material_texture, int n_materials);
/* useTextures (src/main.cpp:58) and its knobs for the following frames; NULL = off and the knobs back at the reference's
 * defaults (nearest neighbour, border, black).  A texture-debug frame (rt_params.texture_debug) samples with these knobs
 * whether useTextures is on or not, as getFinalColorNoRayTracingJustTextures does (src/main.cpp:75-106). */
int rt_set_texturing(rt_ctx* ctx, const rt_texture_params* params);

/* Screen post-processing: the step renderRayTracing ends with (screen.postprocessImage(), src/main.cpp:397-398), and the
 * bloom + 8-bit conversion of Screen::writeBitmapToFile (src/screen.cpp:40-53).  Fields = Screen's private settings
 * (src/screen.h:84-101) as its setters leave them (src/screen.cpp:172-223). */
#define RT_FILTER_NONE 0              /* FilteringOption, src/screen.h:17-26 */
#define RT_FILTER_BLOOM 1
#define RT_FILTER_BLOOM_REINHARD 2
#define RT_FILTER_BLOOM_EXPOSURE 3
#define RT_FILTER_ONLY_LIGHT 4
#define RT_FILTER_ONLY_LIGHT_KERNEL 5
#define RT_KERNEL_BOX 0               /* Kernel, src/screen.h:28-31 */
#define RT_KERNEL_GAUSSIAN 1
typedef struct {
    int filtering_option;   /* RT_FILTER_* (default NONE)                                                     */
    int kernel;             /* RT_KERNEL_* (default BOX)                                                      */
    int kernel_repetitions; /* setKernelNumRepetitions: values < 1 mean 1                                     */
    int filter_size;        /* setFilterSize (default 5): taps run over [-size, size]^2; at most 64           */
    float sigma;            /* setSigma (default 2): values < 0.001 mean 0.001                                */
    float exposure;         /* setExposure (default 0.5)                                                      */
    int gamma_correction;   /* enableGammaCorrection (default off)                                            */
    float gamma;            /* setGammaValue (default 2.2)                                                    */
    int bloom_live;         /* setBloomFilterLive (default off): postprocessImage blooms only when set        */
} rt_post_params;

/* Post-processing applied by every following rt_render / rt_render_device at the end of the frame, on the device,
 * before the rows travel to the host (what renderRayTracing does with the Screen it renders into).  NULL switches it
 * off.  Frames sharded over several GPUs (rt_set_shard) are not post-processed: the bloom needs the gathered image, so
 * rank 0 calls rt_postprocess_device on its framebuffer after the ranks have met. */
int rt_set_postprocess(rt_ctx* ctx, const rt_post_params* post);
/* Screen::postprocessImage (via_write_bitmap = 0) or writeBitmapToFile's bloom + conversion (via_write_bitmap = 1) on
 * an image in host memory: rgb = W*H*3 floats in the Screen layout, processed in place; rgba8 (may be NULL) receives
 * the W*H*4 bytes the reference hands to its BMP encoder. */
int rt_postprocess(rt_ctx* ctx, const rt_post_params* post, float* rgb, int width, int height, int via_write_bitmap, unsigned char* rgba8);
/* The same on a device-resident float4 image (NULL = the context's own framebuffer), asynchronous on the context's
 * stream; bloom_live and gamma_correction are honoured as in postprocessImage. */
int rt_postprocess_device(rt_ctx* ctx, const rt_post_params* post, void* d_rgba, int width, int height);

/* Scene::spotLight / Scene::planeLight (src/scene.h:93-94; getSpotLichts / getPlaneLights, src/shadow.cpp:229-321). */
int rt_set_spot_lights(rt_ctx* ctx, const rt_spot_light* spot, int n_spot);
int rt_set_plane_lights(rt_ctx* ctx, const rt_plane_light* plane, int n_plane);
/* Scene::spheres (src/scene.h:88): read live every frame like the lights; at most 64.  Spheres are tested against every
 * ray before the triangle BVH is walked (the reference puts them into its BVH leaves, bounding_volume_hierarchy.cpp:283-292). */
int rt_set_spheres(rt_ctx* ctx, const rt_sphere* spheres, int n_spheres);
/* Number of 32-byte nodes and depth of the BVH built last. */
int rt_bvh_info(rt_ctx* ctx, int* n_nodes, int* depth);
/* Instrumented kernels: fill rt_stats.node_visits / tri_tests / tri_tests_full (slower; off by default). */
int rt_set_counters(rt_ctx* ctx, int enable);
/* Per-stage device timing of the next frames: CUDA events around every kernel launch (adds a little overhead, so
 * leave it off for frames whose total time is being measured).  rt_stage_times returns, for the frame completed by
 * the last rt_sync / rt_render, the summed milliseconds and launch counts of each stage. */
#define RT_STAGE_GENERATE 0
#define RT_STAGE_EXTEND 1
#define RT_STAGE_SHADE 2
#define RT_STAGE_SHADOW_POINT 3
#define RT_STAGE_SHADOW_SPHERE 4
#define RT_STAGE_RESOLVE 5
#define RT_STAGE_SHADOW_PLANE 6
#define RT_STAGE_POST 7
#define RT_STAGE_COUNT 8
int rt_set_stage_timing(rt_ctx* ctx, int enable);
int rt_stage_times(rt_ctx* ctx, float* ms /* [RT_STAGE_COUNT] */, int* launches /* [RT_STAGE_COUNT] */);
/* How the bounce levels >= 1 are traced.  Per level (the default for wavefronts that fill the GPU): one extend, one shade and one
 * shadow kernel per level, coherent batches.  As whole paths (k_paths): one launch follows every level-1 ray to the end of its
 * path — legal when no material is transparent, glossy_ray_count is 1, all lights are point or spot lights and no texture is
 * sampled; faster for very small wavefronts (frames of a few hundred pixels a side), whose per-level kernels each wait for their
 * longest ray.  mode: -1 automatic (paths for batches of up to 128 K primary rays), 0 never, 1 whenever legal.  Same image either
 * way (per-pixel sums in a different order). */
int rt_set_paths(rt_ctx* ctx, int mode);
/* Traversal form of the extend and (opaque scenes) point-light shadow kernels of the levels >= 1.  One lane per ray through the
 * binary tree: best throughput, what a full queue wants.  Eight lanes per ray through an 8-wide tree (collapsed from the binary one,
 * on the device, when a frame first needs it, for scenes of up to 2^22 triangles): 2.3 times shorter a chain for the longest ray, which is what the kernel of a
 * SMALL queue waits for (deep bounce levels, one GPU's share of a sharded frame), at half the throughput.  Same hits bit for bit.
 * mode: -1 automatic (per level: eight lanes when that level's queue held at most 80 000 shadow records / 130 000 rays — 60 000
 * beside a one-lane-per-ray shadow kernel — in the previous frame of the same shape), 0 never, 1 for every level >= 1. */
int rt_set_wide(rt_ctx* ctx, int mode);
/* Measurement aid (bench.py's roofline): the rate at which this device issues un-fused FP32 multiplies and adds — the instruction
 * mix of the path's parity-critical geometry code — in 1e9 lane-instructions per second, from a microbenchmark kernel (best of 3). */
int rt_measure_fp32_peak(rt_ctx* ctx, double* ginst_per_s);
/* Shadow kernels of bounce level L run on a side stream concurrently with extend / shade of level L+1 (default on). */
int rt_set_overlap(rt_ctx* ctx, int enable);
/* Batch pipelining: a frame is cut into about `batches_per_frame` batches (none smaller than min_batch_pixels) that are
 * processed by `lanes` (1..4) independent sets of queues and streams, `lanes` batches at a time.  With rt_render, each
 * finished batch (a band of complete image rows) is packed and copied to the host while later batches still render. */
int rt_set_pipeline(rt_ctx* ctx, int lanes /* 0 = choose automatically (default) */, int batches_per_frame, unsigned int min_batch_pixels);
/* Upper bound on primary rays per wavefront batch (default 2^24): ray-state memory is O(batch), not O(W*H*spp). */
int rt_set_batch_rays(rt_ctx* ctx, unsigned int max_primary_rays_per_batch);

/* ---- image sharding across GPUs (one context per rank; scene replicated) ----
 * The image is cut into 32x16-pixel tiles; this context renders tiles with tile_id % world == rank. */
int rt_set_shard(rt_ctx* ctx, int rank, int world);

/* ---- render (replaces renderRayTracing, src/main.cpp:340-400, and Screen::setPixel, src/screen.cpp:32-38) ----
 * rt_render: host buffers in, host buffers out, synchronous.  rgb_out: width*height*3 floats in the Screen
 * layout (row H-1-y, column x).  Any host memory will do; page-locked memory (cudaHostAlloc / cudaHostRegister, which
 * host/screen.cpp does for the Screen's pixels) is the fast path: the kernels then store the frame into it themselves, rows of
 * background while the frame is still being traced, instead of staging bands through the copy engine.  tri_id_out / t_out (nullable): closest-hit triangle id (global index, -1 =
 * miss) and t of the first primary ray of each pixel, same layout. */
int rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, float* rgb_out, int* tri_id_out,
    float* t_out, rt_stats* stats);
/* rt_render_device: same frame, result left in device memory, asynchronous on the context's stream.
 * d_rgba: device pointer to width*height float4 (Screen layout); it may be a peer-mapped pointer to another
 * GPU's framebuffer (tiles of this rank are stored straight into it — the gather fused into the resolve
 * kernel).  NULL selects the context's own framebuffer.  stats (nullable) is filled by rt_sync. */
int rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, void* d_rgba);
int rt_sync(rt_ctx* ctx, rt_stats* stats);
/* rt_render_shard: rt_render for a sharded context (rt_set_shard), synchronous.  rgb_host_mapped is the WHOLE image's
 * packed float3 buffer (Screen layout, width*height*3 floats) in page-locked host memory that is mapped into this
 * context's device (cudaHostRegister with cudaHostRegisterMapped | cudaHostRegisterPortable, or cudaHostAlloc): typically
 * a shared-memory segment that every rank's process registers.  The frame's last kernel stores the pixels of the tiles
 * this rank owns straight into it — no staging copy, every GPU over its own PCIe link — and touches nothing else, so after
 * all ranks have returned the buffer holds the frame renderRayTracing would have left in Screen::m_textureData. */
int rt_render_shard(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, float* rgb_host_mapped, rt_stats* stats);
/* Host memory of the caller as rt_render's fast path wants it: page-locked and mapped into every device (what host/screen.cpp does
 * for the Screen's pixels, so that nothing above this header calls the CUDA runtime).  rt_host_register fails without a device or
 * without the permission to lock pages; the caller then simply keeps pageable memory.  rt_current_device: the calling thread's
 * current CUDA device (0 when there is none). */
int rt_host_register(void* p, size_t bytes);
/* The same kind of memory from the CUDA allocator itself (cudaHostAlloc, portable + mapped), for callers that can choose where their
 * image lives; release with rt_host_free. */
int rt_host_alloc(size_t bytes, void** p);
int rt_host_free(void* p);
int rt_host_unregister(void* p);
int rt_current_device(void);
/* The rows of background leave for the host while the frame is still traced, paced to just under what the link carries (stores that
 * back up into L2 slow the traversal kernels down).  One context measures its own link; what a link carries while ALL ranks of a
 * job store into the same host memory depends on the box (shared PCIe switches, the host's ingest rate) and only the job can measure
 * it: every rank copies device-to-host at the same time and hands its rate in GB/s to its context.  0 = back to the built-in
 * assumption min(own link, 100 GB/s / world). */
int rt_set_host_store_rate(rt_ctx* ctx, double gbs);
/* Device framebuffer of this context (width*height float4 of the last rt_render_device with d_rgba=NULL). */
int rt_framebuffer(rt_ctx* ctx, void** d_rgba, int* width, int* height);
/* The gather — the one collective of a sharded frame: finished tiles land in the framebuffer of the ROOT rank, stored there by the
 * other ranks' resolve kernels over NVLink.
 *   root:  rt_framebuffer_ipc_handle exports its framebuffer for width x height (CUDA IPC handle, 64 bytes) and makes the context the
 *          gather root: its rt_render_device(..., NULL) calls with world > 1 are GATHER FRAMES.
 *   peers: rt_open_peer_framebuffer maps the handle (another process), or rt_set_gather_target takes the root's rt_framebuffer
 *          pointer as it stood right after the export (same process, peer access enabled by the caller); their
 *          rt_render_device(..., that pointer) calls with world > 1 are gather frames too.
 * Gather frames store only what has to travel.  The exported allocation holds two images used by alternate frames; while frame k
 * is rendered into one, the root fills the other with the background colour, so a peer leaves out every 32-pixel tile row none of
 * whose camera rays hit anything (C3: 16 % of the rows are stored).  What this asks of the callers, and what every per-frame
 * job loop does anyway: all ranks render the same sequence of gather frames, and a barrier (after rt_sync on every rank)
 * separates consecutive frames.  rt_framebuffer on the root returns the image of the last frame.  rt_stats.gather_bytes
 * reports what a peer stored. */
int rt_framebuffer_ipc_handle(rt_ctx* ctx, int width, int height, void* handle64);
int rt_open_peer_framebuffer(rt_ctx* ctx, const void* handle64, void** d_rgba);
int rt_set_gather_target(rt_ctx* ctx, void* d_rgba);
int rt_close_peer_framebuffer(rt_ctx* ctx, void* d_rgba);
/* Copy a device float4 framebuffer to a host float3 image (Screen::m_textureData layout). */
int rt_download_rgb(rt_ctx* ctx, const void* d_rgba, int width, int height, float* rgb_out);

/* Host-only (no GPU needed): the pixel rectangle [rect[0], rect[1]) x [rect[2], rect[3]) (x, then y counted from the bottom as in
 * src/main.cpp:350-353) outside of which no camera ray of a width x height frame can meet the box [lo, hi].  rt_render uses it
 * with the scene's bounds to answer those primary rays as misses without tracing them; exported so that its conservativeness can
 * be tested where there is no device. */
int rt_visible_rect(const rt_camera* cam, int width, int height, const float lo[3], const float hi[3], int rect[4]);

/* ---- closest hit for caller-supplied rays (replaces BoundingVolumeHierarchy::intersect(Ray&, HitInfo&,
 *      bool useBVH), bounding_volume_hierarchy.cpp:49-78) ----
 * rays: 6 floats each (origin, direction).  tri_id: global triangle index or -1; t: ray.t or FLT_MAX. */
int rt_intersect(rt_ctx* ctx, const float* rays, int64_t n_rays, int use_bvh, int* tri_id, float* t);
/* Device time of the traversal kernel of the last rt_intersect (developer aid). */
float rt_last_intersect_ms(rt_ctx* ctx);

/* ---- the checked build (the GPU counterpart of the reference's sanitizer options, framework/cmake/Sanitizers.cmake:7-37) ----
 * `make EXTRA=-DRT_CHECKED=1 OUT=librtb200_checked.so BUILD=build_checked` compiles the same kernels with every data-derived index
 * (BVH node, triangle, traversal-stack slot, 8-wide node and group-stack slot, accumulator / id pixel, framebuffer pixel, texel,
 * material / light / sphere table entry, ray / shadow / hit queue slot) tested against its bound before the access; a violation is counted and the access goes to
 * element 0.  rt_checked_build: 1 for such a library, 0 for the default one (RT_GUARD is then the identity, same SASS as without it).
 * rt_violations: violations since the library was loaded on the context's device, per site in the order above plus queue slots (10 sites; further
 * entries are zero); all zero from the default build.  Synchronises the device.
 * rt_violations_selftest: one deliberate violation from each translation unit that counts (site 8, table entry, and site 4, group-stack
 * slot), so that a test can tell a working counter from a dead one; does nothing in the default build. */
int rt_checked_build(void);
int rt_violations(rt_ctx* ctx, unsigned int* counts, int n_counts);
int rt_violations_selftest(rt_ctx* ctx);

/* ---- OBJ/MTL loading (replaces loadMesh, src/mesh.cpp:58-188; host-only, no GPU needed) ---- */
typedef struct rt_mesh_soup rt_mesh_soup;
int rt_load_obj(const char* path, int center_and_normalize, rt_mesh_soup** out);
int64_t rt_soup_num_triangles(const rt_mesh_soup* s);
int rt_soup_num_meshes(const rt_mesh_soup* s);
const float* rt_soup_positions(const rt_mesh_soup* s); /* 9 floats per triangle */
const float* rt_soup_normals(const rt_mesh_soup* s);   /* 9 floats per triangle */
const int* rt_soup_mesh_ids(const rt_mesh_soup* s);
const rt_material* rt_soup_materials(const rt_mesh_soup* s);
const float* rt_soup_texcoords(const rt_mesh_soup* s);          /* 6 floats per triangle: Vertex::texCoord of its corners        */
int rt_soup_num_textures(const rt_mesh_soup* s);                /* distinct map_Kd files that could be decoded (PNG)             */
const rt_texture* rt_soup_textures(const rt_mesh_soup* s);
const int* rt_soup_material_textures(const rt_mesh_soup* s);    /* per mesh: index into rt_soup_textures or -1 (Material::kdTexture) */
void rt_soup_free(rt_mesh_soup* s);

const char* rt_last_error(void);
const char* rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
