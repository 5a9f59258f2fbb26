// extern "C" layer of the B200 ray-tracing path (include/rt_b200.h): context, device memory, frame orchestration.
#include "rt_b200.h"
#include "rt_kernels.h"
#ifndef RT_BVH_WIDE
#define RT_BVH_WIDE 0
#endif

#include "../host/mesh.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <string>
#include <cstdlib>
#include <vector>

using namespace rtb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                    \
    } while (0)

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count)
    {
        if (count <= n)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess)
            n = count;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

} // namespace

struct rt_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // scene
    long long n_tris = 0;
    int n_mats = 0;
    std::vector<float> h_pos;
    float coord_max = 1.0f;
    float tri_lo[3] = { 0, 0, 0 }, tri_hi[3] = { 0, 0, 0 }; // bounds of the uploaded triangle corners
    std::vector<float> h_sphere_radii;
    bool cull_primary = true;           // RTB200_CULL_PRIMARY=0 (developer): trace every primary ray
    DevBuf<float> d_pos, d_nrm;
    DevBuf<int> d_mesh, d_perm, d_rank;    // d_rank: visiting rank of every object in the reference's BVH (rt_reforder.cu)
    std::vector<float> h_sphere_centres;   // what d_rank was computed for
    DevBuf<float4> d_plane, d_v0, d_v1, d_v2, d_n0, d_n1, d_n2, d_nodes, d_mats, d_point, d_sphere;
    float4* lbvh_nodes = nullptr; // owned when the device builder allocated them
    int* lbvh_perm = nullptr;
    float4* wide_nodes = nullptr; // 4-wide tree collapsed from the builder's binary one (RT_BVH_WIDE builds)
    float4* wide8_nodes = nullptr; // 8-wide tree for the eight-lanes-per-ray kernels of small queues (rt_wide8.cu), collapsed from the binary one at first need
    int wide8_root = 0, wide8_depth = 0, wide8_count = 0;
    bool wide8_built = false, wide8_tried = false;
    BuildScratch build_scratch;   // arena of the device builder, kept between builds (up to 1 GB)
    int wide_mode = -1;           // rt_set_wide: -1 automatic (levels whose queues were small in the previous frame), 0 never, 1 every level >= 1
    // queue fills per lane and bounce level of the previous frame (Counters::level_ext / level_sh) and what that frame looked like
    unsigned hist_ext[kMaxLanes][kLevelHistory] = {}, hist_sh[kMaxLanes][kLevelHistory] = {};
    unsigned long long hist_signature = 0, frame_signature = 0;
    bool hist_valid = false;
    float last_intersect_ms = 0.0f;
    const float4* nodes = nullptr;
    int n_nodes = 0, root_entry = 0, bvh_depth = 0;
    bool bvh_built = false;
    int n_point = 0, n_sphere = 0;
    bool any_transparent = false;          // meshes or spheres
    bool mats_transparent = false, spheres_transparent = false;
    DevBuf<float4> d_spheres, d_plane_lights;
    DevBuf<float> d_mat_glossy, d_sphere_glossy; // glossy cone half-width per material / sphere primitive (glossy_cone)
    // diffuse textures (rt_set_texcoords / rt_set_textures / rt_set_texturing)
    DevBuf<float4> d_tex_texels;
    DevBuf<int4> d_tex_table;
    DevBuf<int> d_mat_tex;
    DevBuf<float2> d_uv;
    bool have_uv = false;
    int n_textures = 0;
    bool band_order_outer_first = true; // host-path frames: outer bands first (RTB200_BAND_ORDER=0: top to bottom)
    bool trace_bands = false;           // RTB200_TRACE_BANDS=1 (developer): print when every band was rendered / had left
    std::vector<cudaEvent_t> trace_ev;
    std::vector<std::string> trace_what;
    bool tex_on = false;
    rt_texture_params tex_params {};
    int n_spheres = 0;
    std::vector<float4> h_point, h_spot; // point-like light table = point lights followed by spot lights
    int n_point_user = 0, n_spot = 0, n_plane = 0;
    long long user_tris = 0;               // triangles the caller uploaded (0 allowed: a never-hit dummy is traced instead)

    // sharding
    int rank = 0, world = 1;

    // frame state
    bool counters_enabled = false;
    unsigned batch_rays = 1u << 24;
    // Pipelined batches: a frame is cut into batches that run concurrently on up to kMaxLanes "lanes".  Each lane owns
    // its ray / shadow queues, counters, a main stream (extend, shade, resolve, per-batch download) and a side stream
    // (shadow kernels of level L overlap extend / shade of level L+1).  Tails of one batch's latency-bound deep levels
    // are filled by other batches' work, and a finished batch's rows travel to the host while the others still render.
    struct Lane {
        cudaStream_t main = nullptr, side = nullptr;
        cudaEvent_t ev_shade[2] = { nullptr, nullptr }, ev_shadow[2] = { nullptr, nullptr }, ev_done = nullptr, ev_packed = nullptr;
        DevBuf<float4> q_o[2], q_d[2], q_w[2], sp_p[2], sp_a[2], sp_b[2], ss_p[2], ss_a[2], ss_b[2];
        DevBuf<float4> pl_p[2], pl_a[2], pl_b[2], pl_r[2], pl_acc[2];
        DevBuf<int2> q_hit[2];
        DevBuf<float2> sphere_acc[2];
        DevBuf<Counters> counters;
        bool used = false;
    } lanes[kMaxLanes];
    int n_lanes = 0;            // 0 = automatic
    int batches_per_frame = 1;
    unsigned min_batch_pixels = 1u << 18;
    cudaEvent_t ev_start = nullptr, ev_copied = nullptr;
    cudaStream_t copy = nullptr; // band downloads: must not hold up the next batch on the lane that produced the band
    bool overlap = true;
    DevBuf<float4> accum, fb;
    DevBuf<float4> post_img, post_a, post_b; // post-processing: host image staging, light image ping-pong
    DevBuf<float> post_weights;
    int post_weights_f = -1;
    float post_weights_sigma = 0.0f;
    DevBuf<unsigned char> post_rgba8;
    bool post_on = false;
    rt_post_params post {};
    DevBuf<int> prim_id, out_id;
    DevBuf<float> prim_t, out_t, rgb;
    DevBuf<unsigned char> row_flags;    // per 32-pixel tile row: some camera ray hit (k_row_flags)
    double store_gbs = 0.0;             // rate the paced background stores hold: 85 % of the measured device-to-host copy rate
    int paths_mode = -1;                // rt_set_paths: -1 automatic, 0 never, 1 whenever legal
    double queue_scale = 1.0;           // RTB200_QUEUE_SCALE (tests): shrinks the queues' head-room to provoke the overflow path
    unsigned headroom_shift = 0;        // doublings of the head-room asked for by the overflow retry of rt_render / rt_render_shard
    double shared_store_gbs = 0.0;      // rt_set_host_store_rate: what this rank's link carries while every rank of the job stores into host memory
    bool zero_copy_host = true;         // RTB200_ZERO_COPY=0 (developer): rt_render always stages bands through the copy engine
    DevBuf<float> rays_in;
    DevBuf<unsigned> flag;
    Counters* h_counters = nullptr; // pinned, one per lane
    int fb_w = 0, fb_h = 0;
    // Gather frames of a sharded job (rt_framebuffer_ipc_handle / rt_open_peer_framebuffer): the root's framebuffer has two halves
    // used by alternate frames; while frame k is rendered into half k & 1 the root fills the other half with the background colour,
    // and the barrier that ends frame k orders that fill before the peers' stores of frame k + 1 — which can therefore leave out
    // every tile row without a hit.
    bool gather_root = false;        // this context's framebuffer was exported
    void* peer_base = nullptr;       // the root's framebuffer as mapped into this process
    unsigned long long gather_frames = 0; // gather frames rendered so far (every rank counts the same sequence)
    float4* fb_last = nullptr;       // where the last frame rendered into the own framebuffer lives (rt_framebuffer)
    int fb_last_w = 0, fb_last_h = 0;
    DevBuf<float4> fb_plain;         // gather root: framebuffer of its frames that are not gather frames
    int fb_plain_w = 0, fb_plain_h = 0;
    int last_launches = 0, last_batches = 0;
    unsigned long long last_traced_primary = 0;
    int last_gather_mode = 0;        // 0: frame stayed in this context's memory; 1: every pixel of its tiles stored into a foreign buffer; 2: rows with hits only
    size_t last_local_pixels = 0;
    unsigned last_overflow = 0;
    bool frame_pending = false;

    // optional per-stage device timing (events around every launch; off for timed frames)
    bool stage_timing = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_stage; // stage of the launch between ev_pool[2k] and ev_pool[2k+1]
    size_t ev_used = 0;
    float stage_ms[RT_STAGE_COUNT] = { 0 };
    int stage_launches[RT_STAGE_COUNT] = { 0 };

    SceneDev scene_dev() const
    {
        SceneDev s;
        s.nodes = nodes;
        s.tri_plane = d_plane.p;
        s.tri_n0 = d_n0.p;
        if (RT_TRI_AOS) { // one 64-byte record per triangle in d_plane / d_n0
            s.tri_v0 = d_plane.p + 1;
            s.tri_v1 = d_plane.p + 2;
            s.tri_v2 = d_plane.p + 3;
            s.tri_n1 = d_n0.p + 1;
            s.tri_n2 = d_n0.p + 2;
        } else {
            s.tri_v0 = d_v0.p;
            s.tri_v1 = d_v1.p;
            s.tri_v2 = d_v2.p;
            s.tri_n1 = d_n1.p;
            s.tri_n2 = d_n2.p;
        }
        s.mats = d_mats.p;
        s.point_lights = d_point.p;
        s.sphere_lights = d_sphere.p;
        s.plane_lights = d_plane_lights.p;
        s.spheres = d_spheres.p;
        s.mat_glossy_d = d_mat_glossy.p;
        s.sphere_glossy_d = d_sphere_glossy.p;
        s.tex_texels = d_tex_texels.p;
        s.tex_table = d_tex_table.p;
        s.mat_tex = d_mat_tex.p;
        s.tri_uv = have_uv ? d_uv.p : nullptr;
        s.sphere_rank = d_rank.p ? d_rank.p + user_tris : nullptr;
        s.tie_by_id = 0;
        s.n_spheres = n_spheres;
        s.sphere_id_base = (int)user_tris;
        s.n_tris = (int)n_tris;
        s.n_nodes = n_nodes;
#if RT_CHECKED
        s.n_wide_nodes = wide8_count;
        s.n_mats = n_mats;
        s.n_point_like = n_point;
#endif
        return s;
    }
};

namespace {

int use_device(rt_ctx* ctx)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    CK(cudaSetDevice(ctx->device));
    return RT_OK;
}

// Half-width of the glossy cone, `d` of src/main.cpp:224: std::pow(0.5f, -1 / shininess) * std::sqrt(1 - std::pow(0.5, 2 / shininess)),
// float pow, double pow and sqrt, product rounded to float — evaluated here with the host's libm, as the CPU reference would.
float glossy_cone(float shininess)
{
    if (shininess == 0.0f)
        return 0.0f; // never read: the reference skips the glossy branch for shininess 0 (main.cpp:204)
    return (float)(std::pow(0.5f, -1 / (float)shininess) * std::sqrt(1 - std::pow(0.5, 2 / (float)shininess)));
}

int upload_materials(rt_ctx* ctx, const rt_material* mats, int n_mats)
{
    if (!mats || n_mats <= 0)
        return fail(RT_ERR_INVALID, "materials: need at least one");
    std::vector<float4> h(2 * (size_t)n_mats);
    std::vector<float> gd((size_t)n_mats);
    bool any_t = false;
    for (int i = 0; i < n_mats; i++) {
        h[2 * i] = make_float4(mats[i].kd[0], mats[i].kd[1], mats[i].kd[2], mats[i].shininess);
        h[2 * i + 1] = make_float4(mats[i].ks[0], mats[i].ks[1], mats[i].ks[2], mats[i].transparency);
        gd[i] = glossy_cone(mats[i].shininess);
        any_t |= mats[i].transparency != 1.0f;
    }
    CK(ctx->d_mats.ensure(h.size()));
    CK(ctx->d_mat_glossy.ensure(gd.size()));
    CK(cudaMemcpyAsync(ctx->d_mat_glossy.p, gd.data(), gd.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_mats.p, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // h is a stack-lifetime staging vector
    ctx->n_mats = n_mats;
    ctx->mats_transparent = any_t;
    ctx->any_transparent = ctx->mats_transparent || ctx->spheres_transparent;
    return RT_OK;
}

// Host-side pieces of the camera and sampling set-up; they use libm exactly where the reference does.
// Camera of the frame (fp.W, fp.H set): glm::quat(eulerAngles) and Trackball::position() (trackball.cpp:65-68), generateRay's half
// extents (89-90), with the host's libm exactly as the reference evaluates them.
void camera_to_frame(const rt_camera* cam, FrameParams& fp)
{
    const float ex = cam->euler[0] * 0.5f, ey = cam->euler[1] * 0.5f, ez = cam->euler[2] * 0.5f;
    const float cx = std::cos(ex), cy = std::cos(ey), cz = std::cos(ez);
    const float sx = std::sin(ex), sy = std::sin(ey), sz = std::sin(ez);
    fp.qw = cx * cy * cz + sx * sy * sz;
    fp.qx = sx * cy * cz - cx * sy * sz;
    fp.qy = cx * sy * cz + sx * cy * sz;
    fp.qz = cx * cy * sz - sx * sy * cz;
    {
        // q * (0, 0, -dist): uv = cross(qv, v), uuv = cross(qv, uv), v + ((uv*w) + uuv) * 2
        const float vx = 0.0f, vy = 0.0f, vz = -cam->dist;
        const float uvx = fp.qy * vz - vy * fp.qz, uvy = fp.qz * vx - vz * fp.qx, uvz = fp.qx * vy - vx * fp.qy;
        const float uux = fp.qy * uvz - uvy * fp.qz, uuy = fp.qz * uvx - uvz * fp.qx, uuz = fp.qx * uvy - uvx * fp.qy;
        fp.ox = cam->look_at[0] + (vx + ((uvx * fp.qw) + uux) * 2.0f);
        fp.oy = cam->look_at[1] + (vy + ((uvy * fp.qw) + uuy) * 2.0f);
        fp.oz = cam->look_at[2] + (vz + ((uvz * fp.qw) + uuz) * 2.0f);
    }
    fp.halfH = std::tan(cam->fovy / 2.0f);
    fp.halfW = (float(fp.W) / float(fp.H)) * fp.halfH;
}

// Pixel rectangle outside of which no camera ray can meet the box [lo, hi]: its eight corners taken through the inverse of
// generate_ray's mapping (rt_kernels.cu: pixel -> NDC -> normalize(-nx * halfW, ny * halfH, 1) rotated by q), widened by three
// pixels for the sub-pixel sample offsets (< 1 pixel) and the float rounding of the device's ray set-up.  A corner beside or
// behind the eye, or non-finite input, leaves the whole image.  (The projection of a convex body in front of the eye is the
// hull of its projected corners, so their bounding rectangle contains it.)
void box_pixel_rect(const double lo[3], const double hi[3], FrameParams& fp)
{
    fp.vis_x0 = fp.vis_y0 = 0;
    fp.vis_x1 = fp.W;
    fp.vis_y1 = fp.H;
    double nx_min = 1e300, nx_max = -1e300, ny_min = 1e300, ny_max = -1e300;
    const double qx = -fp.qx, qy = -fp.qy, qz = -fp.qz, qw = fp.qw; // inverse rotation: world -> camera
    for (int c = 0; c < 8; c++) {
        const double vx = ((c & 1) ? hi[0] : lo[0]) - fp.ox, vy = ((c & 2) ? hi[1] : lo[1]) - fp.oy, vz = ((c & 4) ? hi[2] : lo[2]) - fp.oz;
        if (!std::isfinite(vx) || !std::isfinite(vy) || !std::isfinite(vz))
            return;
        const double uvx = qy * vz - vy * qz, uvy = qz * vx - vz * qx, uvz = qx * vy - vx * qy;
        const double uux = qy * uvz - uvy * qz, uuy = qz * uvx - uvz * qx, uuz = qx * uvy - uvx * qy;
        const double cx = vx + (uvx * qw + uux) * 2.0, cy = vy + (uvy * qw + uuy) * 2.0, cz = vz + (uvz * qw + uuz) * 2.0;
        if (!(cz > 1e-3 * std::max(1.0, std::fabs(vx) + std::fabs(vy) + std::fabs(vz)))) // beside or behind the eye: the projection says nothing
            return;
        const double nx = -cx / cz / (double)fp.halfW, ny = cy / cz / (double)fp.halfH;
        nx_min = std::min(nx_min, nx);
        nx_max = std::max(nx_max, nx);
        ny_min = std::min(ny_min, ny);
        ny_max = std::max(ny_max, ny);
    }
    const double x0 = std::floor((nx_min + 1.0) * 0.5 * fp.W) - 3.0, x1 = std::ceil((nx_max + 1.0) * 0.5 * fp.W) + 4.0;
    const double y0 = std::floor((ny_min + 1.0) * 0.5 * fp.H) - 3.0, y1 = std::ceil((ny_max + 1.0) * 0.5 * fp.H) + 4.0;
    if (!std::isfinite(x0) || !std::isfinite(x1) || !std::isfinite(y0) || !std::isfinite(y1))
        return;
    fp.vis_x0 = (int)std::min((double)fp.W, std::max(0.0, x0));
    fp.vis_x1 = (int)std::min((double)fp.W, std::max((double)fp.vis_x0, x1));
    fp.vis_y0 = (int)std::min((double)fp.H, std::max(0.0, y0));
    fp.vis_y1 = (int)std::min((double)fp.H, std::max((double)fp.vis_y0, y1));
}

// The rectangle for this context's scene: bounds of the triangle corners and the sphere primitives (the exact geometry, which the
// padded BVH boxes contain).  Whole image for the exhaustive search, which the full-size tests compare against.
void visible_pixel_rect(const rt_ctx* ctx, FrameParams& fp)
{
    fp.vis_x0 = fp.vis_y0 = 0;
    fp.vis_x1 = fp.W;
    fp.vis_y1 = fp.H;
    if (!ctx->cull_primary || fp.exhaustive)
        return;
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        lo[a] = ctx->tri_lo[a];
        hi[a] = ctx->tri_hi[a];
    }
    for (size_t k = 0; k < ctx->h_sphere_radii.size() && k < (size_t)ctx->n_spheres; k++)
        for (int a = 0; a < 3; a++) {
            const double r = std::fabs((double)ctx->h_sphere_radii[k]);
            lo[a] = std::min(lo[a], (double)ctx->h_sphere_centres[3 * k + a] - r);
            hi[a] = std::max(hi[a], (double)ctx->h_sphere_centres[3 * k + a] + r);
        }
    box_pixel_rect(lo, hi, fp);
}

constexpr int kTileRotation = 3; // see tile_xy (rt_types.h)

int make_frame_params(const rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, FrameParams& fp)
{
    if (!cam || !prm)
        return fail(RT_ERR_INVALID, "null camera / params");
    if (prm->width <= 0 || prm->height <= 0)
        return fail(RT_ERR_INVALID, "resolution must be positive");
    if (prm->glossy_ray_count < 1 || prm->glossy_ray_count > 40) // the reference's slider range (main.cpp:530)
        return fail(RT_ERR_INVALID, "glossy_ray_count must be between 1 and 40");
    if (prm->max_reflection_level < 0 || prm->max_reflection_level > 64)
        return fail(RT_ERR_INVALID, "max_reflection_level out of range");
    std::memset(&fp, 0, sizeof(fp));
    fp.W = prm->width;
    fp.H = prm->height;
    fp.sample_mode = prm->sample_mode;
    fp.sample_size = prm->sample_size;
    fp.spp = 1;
    fp.sample_scale = 1.0f;
    if (prm->sample_mode == 1) { // main.cpp:358-375
        fp.spp = 4;
        fp.aa_off_x = 1.0f / fp.W * 0.25f;
        fp.aa_off_y = 1.0f / fp.H * 0.25f;
        fp.sample_scale = 0.25f;
    } else if (prm->sample_mode == 2) { // main.cpp:309-335, 377-385
        if (prm->sample_size < 4)
            return fail(RT_ERR_INVALID, "sample_size must be >= 4 for multipleRays");
        const double root = std::sqrt((double)prm->sample_size);
        fp.ms_off_x = (float)((1.0f / fp.W) * (1.0f / (root * 2)));
        fp.ms_off_y = (float)((1.0f / fp.H) * (1.0f / (root * 2)));
        fp.ms_moves = (int)(root - 1);
        const int k = (fp.ms_moves + 1) / 2;
        fp.spp = 4 * k * k;
        fp.sample_scale = (float)(1.0f / prm->sample_size);
    } else if (prm->sample_mode != 0) {
        return fail(RT_ERR_INVALID, "sample_mode must be 0, 1 or 2");
    }
    camera_to_frame(cam, fp);
    fp.tiles_x = (fp.W + kTileW - 1) / kTileW;
    fp.tiles_y = (fp.H + kTileH - 1) / kTileH;
    fp.rank = ctx->rank;
    fp.world = ctx->world;
    const long long total_tiles = (long long)fp.tiles_x * fp.tiles_y;
    fp.n_local_tiles = total_tiles > ctx->rank ? (int)((total_tiles - ctx->rank + ctx->world - 1) / ctx->world) : 0;
    fp.max_level = prm->max_reflection_level;
    fp.refraction = prm->refraction_factor;
    fp.n_point = ctx->n_point;
    fp.n_sphere = ctx->n_sphere;
    fp.n_plane = ctx->n_plane;
    fp.pl_rc = prm->plane_light_ray_count_1d >= 2 ? prm->plane_light_ray_count_1d : 3;
    if (fp.pl_rc > 64)
        return fail(RT_ERR_INVALID, "plane_light_ray_count_1d must be <= 64");
    fp.any_transparent = ctx->any_transparent ? 1 : 0;
    fp.exhaustive = prm->exhaustive ? 1 : 0;
    fp.tie_by_id = prm->use_bvh ? 0 : 1;
    fp.glossy = prm->glossy_ray_count;
    fp.tex_available = ctx->n_textures > 0 && ctx->d_mat_tex.p ? 1 : 0;
    fp.tex_on = ctx->tex_on && fp.tex_available ? 1 : 0;
    fp.tex_debug = prm->texture_debug ? 1 : 0;
    if (fp.tex_debug) { // one corner ray per pixel, no lights, no bounces (main.cpp:355-356)
        fp.sample_mode = 0;
        fp.spp = 1;
        fp.sample_scale = 1.0f;
        fp.max_level = 0;
        fp.n_point = fp.n_sphere = fp.n_plane = 0;
        fp.glossy = 1;
        fp.tex_on = 0;
    }
    fp.tex_filter = ctx->tex_params.filtering;
    fp.tex_oob_x = ctx->tex_params.out_of_bounds_x;
    fp.tex_oob_y = ctx->tex_params.out_of_bounds_y;
    fp.tex_border_r = ctx->tex_params.border_color[0];
    fp.tex_border_g = ctx->tex_params.border_color[1];
    fp.tex_border_b = ctx->tex_params.border_color[2];
    // getSpherelights ring layout (shadow.cpp:190-196)
    int rc = prm->sphere_light_ray_count;
    if (ctx->n_sphere > 0 && rc < 1)
        return fail(RT_ERR_INVALID, "sphere_light_ray_count must be >= 1");
    if (rc < 1)
        rc = 1;
    const int m = std::max(1, (int)(rc / std::round(std::sqrt(2 * 3.14159365358979f * rc))));
    const int n = (rc - 1) / m;
    fp.sl_m = m;
    fp.sl_n = n;
    fp.sl_rc = m * n + 1;
    if (n > 0) {
        const float angle = 2 * 3.14159365358979f / n;
        fp.sl_sin = std::sin(angle);
        fp.sl_omc = 1 - std::cos(angle);
    }
    int g = 1;
    while (g < fp.sl_rc && g < 32)
        g <<= 1;
    fp.sl_group = g;
    visible_pixel_rect(ctx, fp);
    static const int min_quota_env = [] { // developer knob for A/B timing
        const char* e = std::getenv("RTB200_MIN_QUOTA");
        return e ? std::max(1, std::min(32, std::atoi(e))) : 0;
    }();
    fp.min_quota = min_quota_env > 0 ? min_quota_env : 1;
    static const int tile_rot_env = [] { // developer knob for A/B timing
        const char* e = std::getenv("RTB200_TILE_ROT");
        return e ? std::max(0, std::atoi(e)) : -1;
    }();
    fp.tile_rot = tile_rot_env >= 0 ? tile_rot_env : kTileRotation;
    return RT_OK;
}

// Point the batch at the shadow queues / counters of one bounce-level parity.
void set_parity(rt_ctx::Lane& ln, BatchDev& b, int par)
{
    b.par = par;
    b.sphere_acc = ln.sphere_acc[par].p;
    b.sq_point = ShadowQueue { ln.sp_p[par].p, ln.sp_a[par].p, ln.sp_b[par].p };
    b.sq_sphere = ShadowQueue { ln.ss_p[par].p, ln.ss_a[par].p, ln.ss_b[par].p };
    b.sq_plane = PlaneQueue { ln.pl_p[par].p, ln.pl_a[par].p, ln.pl_b[par].p, ln.pl_r[par].p, ln.pl_acc[par].p };
}

// Rays a primary ray can turn into at the next level: two at a dielectric hit, glossy_ray_count at a glossy one (main.cpp:204-290).
size_t queue_multiplier(const rt_ctx* ctx, const FrameParams& fp) { return (size_t)(ctx->any_transparent ? 2 : 1) * (size_t)std::max(fp.glossy, 1); }

// Size one lane's queues for batches of `batch_pixels` pixels and describe them in `b`.
int ensure_lane(rt_ctx* ctx, rt_ctx::Lane& ln, const FrameParams& fp, unsigned batch_pixels, bool want_ids, BatchDev& b)
{
    const size_t prim = (size_t)batch_pixels * fp.spp;
    // head-room of the ray queues: a dielectric hit spawns two rays, a glossy one up to glossy_ray_count (queue_multiplier); deeper
    // levels of dielectric-heavy views can still outgrow it: the frame then ends with RT_ERR_OVERFLOW, and rt_render / rt_render_shard
    // render it again with twice the head-room per attempt (ctx->headroom_shift) on half the batch
    const size_t cap = std::max<size_t>((size_t)kTilePixels,
        (size_t)((double)(prim * queue_multiplier(ctx, fp)) * ctx->queue_scale * (double)(1u << ctx->headroom_shift)));
    for (int k = 0; k < 2; k++) {
        CK(ln.q_hit[k].ensure(std::max(cap, prim))); // level 0 keeps a hit slot per primary ray whatever the head-room of the later levels
        b.q[k].hit = ln.q_hit[k].p;
        if (fp.max_level > 0) { // level 0 never stores rays (K1 is fused)
            CK(ln.q_o[k].ensure(cap));
            CK(ln.q_d[k].ensure(cap));
            CK(ln.q_w[k].ensure(cap));
        }
        b.q[k].o_pix = ln.q_o[k].p;
        b.q[k].d = ln.q_d[k].p;
        b.q[k].w = ln.q_w[k].p;
    }
    const size_t cap_pt = cap * (size_t)std::max(1, fp.n_point), cap_sp = cap * (size_t)std::max(1, fp.n_sphere);
    for (int p = 0; p < 2; p++) {
        if (fp.n_point > 0) {
            CK(ln.sp_p[p].ensure(cap_pt));
            CK(ln.sp_a[p].ensure(cap_pt));
            CK(ln.sp_b[p].ensure(cap_pt));
        }
        if (fp.n_sphere > 0) {
            CK(ln.ss_p[p].ensure(cap_sp));
            CK(ln.ss_a[p].ensure(cap_sp));
            CK(ln.ss_b[p].ensure(cap_sp));
            CK(ln.sphere_acc[p].ensure(cap_sp));
        }
        if (fp.n_plane > 0) {
            const size_t cap_pl = cap * (size_t)fp.n_plane;
            CK(ln.pl_p[p].ensure(cap_pl));
            CK(ln.pl_a[p].ensure(cap_pl));
            CK(ln.pl_b[p].ensure(cap_pl));
            CK(ln.pl_r[p].ensure(cap_pl));
            CK(ln.pl_acc[p].ensure(cap_pl));
        }
    }
    b.plane_capacity = (unsigned)std::min<size_t>(cap * (size_t)std::max(1, fp.n_plane), 0xfffffff0u);
    set_parity(ln, b, 0);
    b.ray_capacity = (unsigned)std::min<size_t>(cap, 0xfffffff0u);
    b.shadow_pt_capacity = (unsigned)std::min<size_t>(cap_pt, 0xfffffff0u);
    b.shadow_sp_capacity = (unsigned)std::min<size_t>(cap_sp, 0xfffffff0u);
    CK(ln.counters.ensure(1));
    b.counters = ln.counters.p;
    b.accum = ctx->accum.p;
    b.prim_id = want_ids ? ctx->prim_id.p : nullptr;
    b.prim_t = want_ids ? ctx->prim_t.p : nullptr;
#if RT_CHECKED
    b.accum_pixels = (unsigned)std::max<size_t>((size_t)fp.n_local_tiles * kTilePixels, 1); // what the frame uses of accum / prim_id / prim_t
    b.hit_capacity = (unsigned)std::min<size_t>(std::max(cap, prim), 0xfffffff0u);
#endif
    return RT_OK;
}

struct StageScope { // brackets one launch with events when stage timing is on
    rt_ctx* ctx;
    bool on;
    size_t first = 0;
    cudaStream_t stream;
    StageScope(rt_ctx* c, int stage, cudaStream_t st) : ctx(c), on(c->stage_timing), stream(st)
    {
        if (!on)
            return;
        while (ctx->ev_pool.size() < ctx->ev_used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ctx->ev_pool.push_back(e);
        }
        ctx->ev_stage.push_back(stage);
        first = ctx->ev_used;
        ctx->ev_used += 2;
        cudaEventRecord(ctx->ev_pool[first], stream);
    }
    ~StageScope()
    {
        if (!on)
            return;
        cudaEventRecord(ctx->ev_pool[first + 1], stream);
    }
};


// Screen's setters clamp two of the settings (src/screen.cpp:191-194, 213-216); setFilterSize does not (219-223).
rt_post_params normalised_post(const rt_post_params& in)
{
    rt_post_params p = in;
    p.kernel_repetitions = std::max(1, p.kernel_repetitions);
    p.sigma = std::max(0.001f, p.sigma);
    return p;
}

bool post_has_effect(const rt_post_params& p) { return (p.bloom_live && p.filtering_option != RT_FILTER_NONE) || p.gamma_correction; }

int check_post(const rt_post_params* p)
{
    if (!p)
        return fail(RT_ERR_INVALID, "post-processing: null settings");
    if (p->filtering_option < RT_FILTER_NONE || p->filtering_option > RT_FILTER_ONLY_LIGHT_KERNEL || (p->kernel != RT_KERNEL_BOX && p->kernel != RT_KERNEL_GAUSSIAN))
        return fail(RT_ERR_INVALID, "post-processing: unknown filtering option or kernel");
    if (p->filter_size > 64)
        return fail(RT_ERR_INVALID, "post-processing: filter_size above 64");
    return RT_OK;
}

// Weight table of the Gaussian kernel for the settings in use; uploaded outside frames (blocks on the context's stream).
// gaussianFunction (src/screen.cpp:324-326): (1 / (sigma * sigma * 2 * M_PI)) * glm::exp(-(x*x + y*y) / (2 * sigma * sigma)) with the
// reference's own `#define M_PI 3.1415926535893238` (src/screen.cpp:13): double prefactor, float exp (this host's libm, the
// one the CPU reference would call), rounded to float on return.
int ensure_post_weights(rt_ctx* ctx, const rt_post_params& p)
{
    const int f = p.filter_size;
    if (p.kernel != RT_KERNEL_GAUSSIAN || f < 0 || p.filtering_option == RT_FILTER_NONE || p.filtering_option == RT_FILTER_ONLY_LIGHT)
        return RT_OK;
    if (ctx->post_weights_f == f && ctx->post_weights_sigma == p.sigma && ctx->post_weights.p)
        return RT_OK;
    const int side = 2 * f + 1;
    std::vector<float> wts((size_t)side * side);
    const float sigma = p.sigma;
    for (int i = -f; i <= f; i++)
        for (int j = -f; j <= f; j++) {
            const float x = (float)i, y = (float)j;
            wts[(size_t)(i + f) * side + (j + f)] = (float)((1 / (sigma * sigma * 2 * 3.1415926535893238)) * std::exp(-(x * x + y * y) / (2 * sigma * sigma)));
        }
    CK(cudaStreamSynchronize(ctx->stream)); // an earlier frame may still read the old table
    CK(ctx->post_weights.ensure(wts.size()));
    CK(cudaMemcpyAsync(ctx->post_weights.p, wts.data(), wts.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->post_weights_f = f;
    ctx->post_weights_sigma = p.sigma;
    return RT_OK;
}

// Screen::applyBloomEffect (src/screen.cpp:226-268) on a device image, in place.
int enqueue_bloom(rt_ctx* ctx, cudaStream_t st, float4* img, int w, int h, const rt_post_params& p, int& launches)
{
    if (p.filtering_option == RT_FILTER_NONE)
        return RT_OK;
    const size_t n = (size_t)w * h;
    CK(ctx->post_a.ensure(n));
    CK(ctx->post_b.ensure(n));
    const int f = p.filter_size;
    const bool gauss = p.kernel == RT_KERNEL_GAUSSIAN;
    float4 *light = ctx->post_a.p, *other = ctx->post_b.p;
    launch_post_light(st, ctx->sm_count, img, light, n);
    launches++;
    if (p.filtering_option == RT_FILTER_ONLY_LIGHT) {
        CK(cudaMemcpyAsync(img, light, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        return RT_OK;
    }
    const int reps = p.filtering_option == RT_FILTER_ONLY_LIGHT_KERNEL ? 1 : p.kernel_repetitions;
    for (int r = 0; r < reps; r++) {
        launch_post_blur(st, light, other, w, h, f, gauss, ctx->post_weights.p);
        launches++;
        std::swap(light, other);
    }
    if (p.filtering_option == RT_FILTER_ONLY_LIGHT_KERNEL) {
        CK(cudaMemcpyAsync(img, light, n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        return RT_OK;
    }
    launch_post_combine(st, ctx->sm_count, img, light, n, p.filtering_option, p.exposure);
    launches++;
    return RT_OK;
}

// Screen::postprocessImage (src/screen.cpp:56-69) on a device image, in place.
int enqueue_postprocess(rt_ctx* ctx, cudaStream_t st, float4* img, int w, int h, const rt_post_params& p, int& launches)
{
    StageScope sc(ctx, RT_STAGE_POST, st);
    if (p.bloom_live) {
        int rc = enqueue_bloom(ctx, st, img, w, h, p, launches);
        if (rc)
            return rc;
    }
    if (p.gamma_correction) {
        launch_post_gamma(st, ctx->sm_count, img, (size_t)w * h, 1.0f / p.gamma);
        launches++;
    }
    CK(cudaGetLastError());
    return RT_OK;
}

// Where rt_render wants the finished rows: packed float3 image on the host (pinned for full PCIe speed).
struct HostTarget {
    // gather frames: `sparse` — store only tile rows with hits (peer ranks); `prefill` — background fill of that buffer while the frame renders (root)
    bool sparse = false;
    bool foreign = false; // the framebuffer is not this context's (caller-supplied device pointer, maybe on a peer GPU)
    float4* prefill = nullptr;
    size_t prefill_n = 0;
    float* rgb = nullptr;
    // rt_render_shard: the whole image's packed float3 buffer in page-locked host memory as THIS device sees it; the
    // frame's last kernel stores the pixels of this rank's tiles into it (no staging copy)
    float* mapped_rgb = nullptr;
    // rt_render into such a buffer: tile rows whose camera rays all missed are final (black) after
    // level 0 and are stored while the other levels are traced; the rows with hits follow when their batch is resolved
    bool early_background = false;
};

// The 8-wide tree beside the binary one (rt_wide8.cu), collapsed from it on the device (two span walks per binary node, then one launch
// per level of the wide tree).  Built when a frame first asks for the eight-lanes-per-ray kernels, for scenes of up to 2^22 triangles
// (larger scenes do not have small queues at the frame sizes they are rendered at).
int ensure_wide8(rt_ctx* ctx)
{
    if (ctx->wide8_tried)
        return RT_OK;
    ctx->wide8_tried = true;
    static const bool wide8_off = [] {
        const char* e = std::getenv("RTB200_WIDE8");
        return e && e[0] == '0';
    }();
    if (wide8_off || ctx->n_tris > (1ll << 22) || RT_BVH_WIDE || !ctx->bvh_built)
        return RT_OK;
    const auto t0 = std::chrono::steady_clock::now();
    if (ctx->wide8_nodes)
        cudaFree(ctx->wide8_nodes);
    ctx->wide8_nodes = nullptr;
    int n8 = 0;
    const char* err = nullptr;
    if (collapse_bvh_wide8_device(ctx->stream, ctx->nodes, ctx->n_nodes, &ctx->wide8_nodes, &n8, &ctx->wide8_root, &ctx->wide8_depth, &err) != 0)
        return fail(RT_ERR_CUDA, std::string("8-wide tree: ") + (err ? err : "failed"));
    ctx->wide8_count = n8;
    ctx->wide8_built = n8 > 0 && 7 * ctx->wide8_depth + 1 < 64; // (a step pushes up to 7 entries on the 64-entry group stack)
    if (std::getenv("RTB200_TRACE_BUILD"))
        std::fprintf(stderr, "[build] 8-wide tree: %d nodes, depth %d, %.2f ms (collapsed on the device)\n", n8, ctx->wide8_depth,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return RT_OK;
}

// Enqueue one frame.  `out`: float4 framebuffer (Screen layout), maybe on a peer GPU.  The frame starts and ends on the
// context's stream; in between, its batches run on the lanes' own streams.
constexpr int kBandGridMult = 4;
constexpr long long kWideMaxRays = 130000;      // extend queues of up to this many rays (in the previous frame) are traced with eight lanes per ray ...
constexpr long long kWideMaxRaysAlone = 60000;  // ... and only up to this many when the shadow kernel running beside them is a one-lane-per-ray one
constexpr long long kWideMaxShadowRays = 80000; // the same for the point-light shadow queues
constexpr long long kPathsMaxPrimaryRays = 1 << 17; // batches of up to 128 K primary rays trace their bounce levels as whole paths

int enqueue_frame(rt_ctx* ctx, const FrameParams& fp_in, float4* out, bool want_ids, unsigned batch_rays, const HostTarget* host)
{
    FrameParams fp = fp_in;
    const auto host_t0 = std::chrono::steady_clock::now();
    const size_t n_local = (size_t)fp.n_local_tiles * kTilePixels;
    // batch size: the frame split evenly over the lanes, in whole tile rows when this context owns the whole image (so
    // that a finished batch is a band of complete image rows), bounded below (tiny batches are all latency) and above
    // (ray-state memory is O(batch))
    const size_t unit = (fp.world == 1 ? (size_t)fp.tiles_x : 1) * kTilePixels;
    // Automatic pipeline shape (rt_set_pipeline(0, ...)), from measurements on C3 (profiles/README.md): every extra
    // sequential batch costs about 0.4 ms of latency-bound deep levels, so a device-resident frame is cut in two
    // concurrent halves at most; when rows have to travel to the host, six bands on three lanes let the first bands'
    // download overlap the later bands' rendering (C3 through rt_render: 1x1 4.10, 2x3 3.55, 3x6 3.48, 4x8 3.72 ms).
    int lanes_wanted = ctx->n_lanes, batches_auto = ctx->batches_per_frame;
    if (lanes_wanted <= 0) {
        const bool big = n_local >= ((size_t)1 << 22);
        if (host && host->rgb && fp.world == 1 && !(ctx->post_on && post_has_effect(ctx->post))) {
            lanes_wanted = big ? 3 : 1;
            batches_auto = big ? 6 : 1;
        } else {
            lanes_wanted = big ? 2 : 1;
            batches_auto = big ? 2 : 1;
        }
    }
    lanes_wanted = std::max(1, std::min(lanes_wanted, kMaxLanes));
    const size_t batches_wanted = (size_t)std::max(1, batches_auto);
    size_t batch_pixels = (n_local + batches_wanted - 1) / batches_wanted;
    batch_pixels = std::max<size_t>(batch_pixels, ctx->min_batch_pixels);
    batch_pixels = (batch_pixels + unit - 1) / unit * unit; // whole tile rows, rounded up: at most batches_wanted batches
    // rt_set_batch_rays bounds the rays of one batch IN FLIGHT: primary rays times what each can turn into at the next level
    const size_t cap_pixels = std::max<size_t>(batch_rays / ((size_t)fp.spp * queue_multiplier(ctx, fp)) * (ctx->any_transparent ? 2 : 1), kTilePixels);
    if (batch_pixels > cap_pixels)
        batch_pixels = std::max<size_t>(unit, cap_pixels / unit * unit);
    if (batch_pixels > n_local)
        batch_pixels = std::max<size_t>(n_local, kTilePixels);
    // post-processing (bloom) needs the whole image: it runs after the last batch, so rows cannot leave band by band
    const bool post = ctx->post_on && fp.world == 1 && post_has_effect(ctx->post) && n_local;
    const bool band_download = host && host->rgb && fp.world == 1 && !post && batch_pixels % ((size_t)fp.tiles_x * kTilePixels) == 0;
    auto trace_at = [&](cudaStream_t on, const std::string& what) {
        if (!ctx->trace_bands)
            return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, on);
        ctx->trace_ev.push_back(e);
        ctx->trace_what.push_back(what);
    };
    const bool early_bg = host && host->mapped_rgb && host->early_background && !post && n_local;
    const bool sparse = host && host->sparse && !early_bg && n_local;
    if (early_bg || sparse)
        CK(ctx->row_flags.ensure(n_local / kTilePixels * kTileH));
    // several lanes render bands side by side: traversal grids of 4 blocks per SM leave room for another lane's kernel
    // (C3 through rt_render, 3 lanes x 6 bands: 3.40 ms with 8 blocks per SM, 3.17 with 4, 3.16 with 3)
    fp.trace_grid_mult = band_download && lanes_wanted > 1 ? kBandGridMult : 0;
    // The batches of the frame: {first local pixel, pixels}.  Bands that travel to the host go from the outside in (top,
    // bottom, second from the top, ...): the camera looks at the scene, so the outer bands are mostly background, finish
    // early and keep the copy engine busy while the expensive middle of the image renders (profiles/README.md, host path).
    std::vector<std::pair<size_t, size_t>> plan;
    for (size_t first = 0; first < n_local; first += batch_pixels)
        plan.push_back({ first, std::min(batch_pixels, n_local - first) });
    if (band_download && ctx->band_order_outer_first) {
        std::vector<std::pair<size_t, size_t>> outer;
        for (size_t lo = 0, hi = plan.size(); lo < hi;) {
            outer.push_back(plan[--hi]);
            if (lo < hi)
                outer.push_back(plan[lo++]);
        }
        plan.swap(outer);
    }
    // Bounce levels as whole paths (k_paths) instead of per-level kernels: legal when a path never splits and a shadow query never
    // continues (every material opaque, glossy_ray_count 1), all lights are point-like and no texture is sampled.  The per-level
    // kernels of a small wavefront each wait for their longest ray (rt_kernels.cu), the path kernel waits once for the longest PATH —
    // which, measured, is made of the long rays of every level.  One GPU, depth 3, per level -> paths (tools/small_frame_probe.py):
    // 256x256 0.499 -> 0.476 (dragon stand-in), 0.256 -> 0.240 (monkey), 0.294 -> 0.224 ms (teapot); 512x512 0.717 -> 0.779, 0.300 -> 0.320,
    // 0.339 -> 0.278 ms; 1024x1024 and up: slower or equal.  A rank's share of the 4K frame split 8 ways (tools/rank_probe.py): 0.70-0.79 ms on
    // seven ranks instead of 0.75-0.81, but 0.95 instead of 0.83 ms on the slowest, which is the one that counts.  Hence: automatic only
    // for batches of up to 128 K primary rays.  rt_set_paths / RTB200_PATHS=0 / 1: never / whenever legal.
    static const int paths_env_default = [] {
        const char* e = std::getenv("RTB200_PATHS");
        return e ? std::atoi(e) : -1;
    }();
    const int paths_env = ctx->paths_mode >= 0 ? ctx->paths_mode : paths_env_default;
    static const long long paths_max_rays = [] {
        const char* e = std::getenv("RTB200_PATHS_MAX_RAYS");
        return e ? std::atoll(e) : (long long)kPathsMaxPrimaryRays;
    }();
    const bool paths_legal = ctx->overlap && fp.max_level >= 1 && !fp.any_transparent && fp.glossy == 1 && fp.n_sphere == 0 && fp.n_plane == 0 && !fp.tex_on
        && !fp.tex_debug && !fp.exhaustive;
    static const int paths_from = [] {
        const char* e = std::getenv("RTB200_PATHS_FROM");
        return e ? std::max(1, std::atoi(e)) : 1;
    }();
    const bool use_paths = paths_legal && paths_env != 0 && (paths_env == 1 || (long long)batch_pixels * fp.spp <= paths_max_rays);
    // Which traversal form the levels >= 1 take: eight lanes per ray through the 8-wide tree for queues that were small in the
    // previous frame (rt_kernels.cu: k_extend_wide), one lane per ray otherwise.  The previous frame's fills are a hint, not a
    // contract: both forms are correct for any queue, the choice only decides the speed.  Frames are compared by a signature.
    static const long long wide_max_rays = [] {
        const char* e = std::getenv("RTB200_WIDE_MAX_RAYS");
        return e ? std::atoll(e) : (long long)kWideMaxRays;
    }();
    static const long long wide_max_rays_alone = [] {
        const char* e = std::getenv("RTB200_WIDE_MAX_RAYS_ALONE");
        return e ? std::atoll(e) : (long long)kWideMaxRaysAlone;
    }();
    static const long long wide_max_shadow = [] { // any-hit queries need no ordering of the children: the wide form stays ahead for larger queues
        const char* e = std::getenv("RTB200_WIDE_MAX_SHADOW");
        return e ? std::atoll(e) : (long long)kWideMaxShadowRays;
    }();
    const unsigned long long signature = ((unsigned long long)fp.W << 44) ^ ((unsigned long long)fp.H << 28) ^ ((unsigned long long)fp.spp << 20)
        ^ ((unsigned long long)fp.world << 12) ^ ((unsigned long long)fp.rank << 4) ^ (unsigned long long)fp.max_level ^ ((unsigned long long)plan.size() << 56)
        ^ ((unsigned long long)fp.n_point << 50);
    bool wide_usable = ctx->wide_mode != 0 && !ctx->counters_enabled && !fp.exhaustive && ctx->overlap && fp.max_level >= 1;
    const bool wide_hist = wide_usable && ctx->hist_valid && ctx->hist_signature == signature && plan.size() <= (size_t)lanes_wanted;
    if (wide_usable && !ctx->wide8_tried) { // the first frame that could use the 8-wide tree builds it
        bool wanted = ctx->wide_mode == 1;
        for (int l = 0; wide_hist && l < lanes_wanted && !wanted; l++)
            for (int lv = 1; lv < kLevelHistory && lv <= fp.max_level; lv++)
                wanted = wanted || (long long)ctx->hist_ext[l][lv] <= wide_max_rays;
        if (wanted) {
            int rc = ensure_wide8(ctx);
            if (rc)
                return rc;
        }
    }
    wide_usable = wide_usable && ctx->wide8_built;
    auto wide_for = [&](int lane, int level, bool shadow) {
        if (!wide_usable || level < (shadow ? 0 : 1))
            return false;
        if (shadow && (fp.any_transparent || fp.n_point == 0))
            return false;
        if (ctx->wide_mode == 1)
            return true;
        if (!wide_hist || level >= kLevelHistory)
            return false;
        if (shadow)
            return (long long)ctx->hist_sh[lane][level] <= wide_max_shadow;
        // An extend kernel runs beside the shadow kernel of the level before.  Next to a one-lane-per-ray shadow kernel, whose persistent
        // blocks fill the SMs until its queue is drained, the eight-lane form only pays for really small queues; next to an eight-lane one
        // it pays up to the size where its throughput loses (1/4 share of C3, level 2: 88 K rays each, shadow one lane per ray: 0.89 ms
        // with the extend kernel one lane per ray, 0.97 with eight).
        const bool beside_wide_shadow = fp.n_point == 0 || (!fp.any_transparent && (long long)ctx->hist_sh[lane][level - 1] <= wide_max_shadow);
        return (long long)ctx->hist_ext[lane][level] <= (beside_wide_shadow ? wide_max_rays : wide_max_rays_alone);
    };
    ctx->frame_signature = signature;
    const size_t n_batches = plan.size();
    const int n_lanes = (int)std::min<size_t>(lanes_wanted, std::max<size_t>(n_batches, 1));

    CK(ctx->accum.ensure(std::max<size_t>(n_local, 1)));
    if (want_ids) {
        CK(ctx->prim_id.ensure(std::max<size_t>(n_local, 1)));
        CK(ctx->prim_t.ensure(std::max<size_t>(n_local, 1)));
    }
    BatchDev bd[kMaxLanes];
    for (int l = 0; l < n_lanes; l++) {
        int rc = ensure_lane(ctx, ctx->lanes[l], fp, (unsigned)batch_pixels, want_ids, bd[l]);
        if (rc != RT_OK)
            return rc;
        ctx->lanes[l].used = false;
    }
    for (int l = n_lanes; l < kMaxLanes; l++)
        ctx->lanes[l].used = false;
    const SceneDev s = ctx->scene_dev();
    cudaStream_t st0 = ctx->stream;
    int launches = 0, batches = 0;
    unsigned long long traced_primary = 0;
    ctx->ev_used = 0;
    ctx->ev_stage.clear();
    CK(cudaEventRecord(ctx->ev0, st0));
    if (n_local)
        CK(cudaMemsetAsync(ctx->accum.p, 0, n_local * sizeof(float4), st0));
    CK(cudaEventRecord(ctx->ev_start, st0));

    const bool prefill = host && host->prefill && host->prefill_n;
    if (prefill) { // gather root: the other half of the framebuffer becomes background for the next frame
        CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_start, 0));
        launch_fill_background(ctx->copy, ctx->sm_count, host->prefill, host->prefill_n);
        launches++;
    }
    if (early_bg) { // rows outside the scene's projection: on their way before the first ray is traced
        CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_start, 0));
        launch_host_background(ctx->copy, ctx->sm_count, fp, 0, (unsigned)fp.n_local_tiles, nullptr, host->mapped_rgb, ctx->shared_store_gbs > 0.0 ? 0.85 * ctx->shared_store_gbs : ctx->store_gbs, ctx->shared_store_gbs > 0.0);
        launches++;
        trace_at(ctx->copy, "rows outside the scene's projection stored");
    }
    for (size_t bi = 0; bi < plan.size(); bi++, batches++) {
        const size_t first = plan[bi].first;
        const int li = batches % n_lanes;
        rt_ctx::Lane& ln = ctx->lanes[li];
        BatchDev& b = bd[li];
        cudaStream_t st = ln.main;
        if (!ln.used) {
            ln.used = true;
            CK(cudaStreamWaitEvent(st, ctx->ev_start, 0));
            CK(cudaMemsetAsync(ln.counters.p, 0, sizeof(Counters), st));
        }
        const unsigned n_lp = (unsigned)plan[bi].second;
        // primary rays of this batch: pixels of its tiles that lie inside the image, times samples per pixel
        unsigned long long n_primary = 0, n_traced = 0;
        for (size_t j = first / kTilePixels; j < (first + n_lp) / kTilePixels; j++) {
            const long long g = (long long)fp.rank + (long long)j * fp.world;
            unsigned utx, uty;
            tile_xy((unsigned)g, (unsigned)fp.tiles_x, (unsigned)fp.tile_rot, utx, uty);
            const int tx = (int)utx, ty = (int)uty;
            n_primary += (unsigned long long)std::min(kTileW, fp.W - tx * kTileW) * std::min(kTileH, fp.H - ty * kTileH);
            // the ones that walk the BVH: pixels of the tile inside the scene's projection (pixel_sees_scene)
            const int x0 = std::max(tx * kTileW, fp.vis_x0), x1 = std::min(std::min((tx + 1) * kTileW, fp.W), fp.vis_x1);
            const int y0 = std::max(ty * kTileH, fp.vis_y0), y1 = std::min(std::min((ty + 1) * kTileH, fp.H), fp.vis_y1);
            if (x1 > x0 && y1 > y0)
                n_traced += (unsigned long long)(x1 - x0) * (y1 - y0);
        }
        n_primary *= (unsigned long long)fp.spp;
        traced_primary += n_traced * (unsigned long long)fp.spp;
        for (int level = 0; level <= fp.max_level; level++) {
            if (level == paths_from && use_paths) {
                // small wavefront of an opaque scene lit by point-like lights: the remaining levels as whole paths, one launch
                // (k_paths), next to the shadow queries of the level before on the side stream
                if (ctx->overlap && level >= 2)
                    CK(cudaStreamWaitEvent(st, ln.ev_shadow[level & 1], 0));
                launch_level_reset(st, b.counters, (level & 1) ^ 1, -1, 0, level & 1);
                {
                    StageScope sc(ctx, RT_STAGE_EXTEND, st);
                    SceneDev s_ext = s;
                    s_ext.tie_by_id = fp.tie_by_id; // for the closest-hit queries; the any-hit shadow queries do not look at tie keys
                    launch_paths(st, ctx->sm_count, s_ext, ctx->root_entry, fp, b, level & 1, level, ctx->counters_enabled);
                }
                launches += 2;
                break;
            }
            const int qi = level & 1, par = level & 1;
            set_parity(ln, b, par);
            // the shadow queues of this parity were last read by the shadow kernels of level - 2
            if (ctx->overlap && level >= 2)
                CK(cudaStreamWaitEvent(st, ln.ev_shadow[par], 0));
            // level 0 has no stored ray queue: K1 generate is fused into extend / shade (rays are a function of the index)
            if (level == 0)
                launch_level_reset(st, b.counters, 1, (long long)n_lp * fp.spp, n_primary, par);
            else
                launch_level_reset(st, b.counters, qi ^ 1, -1, 0, par);
            {
                StageScope sc(ctx, RT_STAGE_EXTEND, st);
                SceneDev s_ext = s;
                s_ext.tie_by_id = fp.tie_by_id; // shadow queries keep the BVH order (shadow.cpp:42)
                if (wide_for(li, level, false))
                    launch_extend_wide(st, ctx->sm_count, s_ext, ctx->wide8_nodes, ctx->wide8_root, b, qi, level);
                else
                    launch_extend(st, ctx->sm_count, s_ext, ctx->root_entry, fp, b, qi, level, (unsigned)first, ctx->counters_enabled);
            }
            if (sparse && level == 0) { // which tile rows of this batch hold a hit: the others are not stored (k_resolve)
                launch_row_flags(st, ctx->sm_count, fp, (unsigned)first, n_lp, b.q[0].hit, ctx->row_flags.p, &b.counters->flagged_rows);
                launches++;
            }
            if (early_bg && level == 0) {
                // rows of pure background leave for the host now, on the copy stream, next to everything that follows
                launch_row_flags(st, ctx->sm_count, fp, (unsigned)first, n_lp, b.q[0].hit, ctx->row_flags.p, &b.counters->flagged_rows);
                CK(cudaEventRecord(ln.ev_packed, st));
                CK(cudaStreamWaitEvent(ctx->copy, ln.ev_packed, 0));
                trace_at(ctx->copy, "level-0 extend done, batch " + std::to_string(bi));
                launch_host_background(ctx->copy, ctx->sm_count, fp, (unsigned)(first / kTilePixels), n_lp / kTilePixels, ctx->row_flags.p, host->mapped_rgb, ctx->shared_store_gbs > 0.0 ? 0.85 * ctx->shared_store_gbs : ctx->store_gbs, ctx->shared_store_gbs > 0.0);
                trace_at(ctx->copy, "background rows stored, batch " + std::to_string(bi));
                launches += 2;
            }
            {
                StageScope sc(ctx, RT_STAGE_SHADE, st);
                launch_shade(st, ctx->sm_count, s, fp, b, qi, level, (unsigned)first);
            }
            launches += 3;
            // shadow rays of this level: on the side stream, concurrently with extend / shade of the next level
            cudaStream_t ss = ctx->overlap ? ln.side : st;
            if (ctx->overlap) {
                CK(cudaEventRecord(ln.ev_shade[par], st));
                CK(cudaStreamWaitEvent(ss, ln.ev_shade[par], 0));
            }
            if (fp.n_point > 0) {
                StageScope sc(ctx, RT_STAGE_SHADOW_POINT, ss);
                if (wide_for(li, level, true))
                    launch_shadow_point_wide(ss, ctx->sm_count, s, ctx->wide8_nodes, ctx->wide8_root, b, level);
                else
                    launch_shadow_point(ss, ctx->sm_count, s, ctx->root_entry, fp, b, level, ctx->counters_enabled);
                launches++;
            }
            if (fp.n_sphere > 0) {
                StageScope sc(ctx, RT_STAGE_SHADOW_SPHERE, ss);
                launch_shadow_sphere(ss, ctx->sm_count, s, ctx->root_entry, fp, b, ctx->counters_enabled);
                launches += 2;
            }
            if (fp.n_plane > 0) {
                StageScope sc(ctx, RT_STAGE_SHADOW_PLANE, ss);
                launch_shadow_plane(ss, ctx->sm_count, s, ctx->root_entry, fp, b, ctx->counters_enabled);
                launches += 2;
            }
            if (ctx->overlap)
                CK(cudaEventRecord(ln.ev_shadow[par], ss));
        }
        if (ctx->overlap) { // join: everything of this batch is in the accumulators before its resolve
            CK(cudaStreamWaitEvent(st, ln.ev_shadow[0], 0));
            if (fp.max_level >= 1)
                CK(cudaStreamWaitEvent(st, ln.ev_shadow[1], 0));
        }
        {
            StageScope sc(ctx, RT_STAGE_RESOLVE, st);
            launch_resolve(st, ctx->sm_count, fp, (unsigned)first, n_lp, ctx->accum.p, b.prim_id, b.prim_t, out, want_ids ? ctx->out_id.p : nullptr,
                want_ids ? ctx->out_t.p : nullptr, sparse ? ctx->row_flags.p : nullptr);
            launches++;
        }
        if (early_bg) { // the rows of this batch that were hit: straight into the host image
            trace_at(st, "resolved, batch " + std::to_string(bi));
            launch_pack_rgb_tiles(st, ctx->sm_count, fp, (unsigned)(first / kTilePixels), n_lp / kTilePixels, ctx->row_flags.p, out, host->mapped_rgb);
            launches++;
            trace_at(st, "rows with hits stored, batch " + std::to_string(bi));
        }
        if (band_download) {
            // this batch is a band of complete image rows: pack it to float3 and send it home while the other lanes render
            const int ty0 = (int)(first / kTilePixels / fp.tiles_x), ty1 = (int)((first + n_lp) / kTilePixels / fp.tiles_x);
            const size_t row_lo = (size_t)(fp.H - std::min(fp.H, ty1 * kTileH)), row_hi = (size_t)(fp.H - ty0 * kTileH);
            const size_t p0 = row_lo * fp.W, p1 = row_hi * fp.W;
            launch_pack_rgb(st, ctx->sm_count, out, ctx->rgb.p, p0, p1);
            launches++;
            CK(cudaEventRecord(ln.ev_packed, st));
            CK(cudaStreamWaitEvent(ctx->copy, ln.ev_packed, 0));
            auto trace = [&](cudaStream_t on, const char* what) {
                if (!ctx->trace_bands)
                    return;
                cudaEvent_t e;
                cudaEventCreate(&e);
                cudaEventRecord(e, on);
                ctx->trace_ev.push_back(e);
                ctx->trace_what.push_back(std::string(what) + " band " + std::to_string(bi) + " lane " + std::to_string(li) + " rows " + std::to_string(row_lo) + ".." + std::to_string(row_hi));
            };
            trace(st, "packed");
            CK(cudaMemcpyAsync(host->rgb + 3 * p0, ctx->rgb.p + 3 * p0, (p1 - p0) * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy));
            trace(ctx->copy, "copied");
        }
        CK(cudaEventRecord(ln.ev_done, st));
    }
    for (int l = 0; l < n_lanes; l++)
        if (ctx->lanes[l].used)
            CK(cudaStreamWaitEvent(st0, ctx->lanes[l].ev_done, 0));
    if (band_download) {
        CK(cudaEventRecord(ctx->ev_copied, ctx->copy));
        CK(cudaStreamWaitEvent(st0, ctx->ev_copied, 0));
    }
    if (post) {
        int rc = enqueue_postprocess(ctx, st0, out, fp.W, fp.H, ctx->post, launches);
        if (rc)
            return rc;
    }
    if (early_bg || prefill) {
        CK(cudaEventRecord(ctx->ev_copied, ctx->copy));
        CK(cudaStreamWaitEvent(st0, ctx->ev_copied, 0));
    }
    if (early_bg) {
    } else if (host && host->mapped_rgb && n_local) {
        launch_pack_rgb_tiles(st0, ctx->sm_count, fp, 0, (unsigned)fp.n_local_tiles, nullptr, out, host->mapped_rgb);
        launches++;
    }
    if (host && host->rgb && !band_download && n_local) {
        const size_t npx = (size_t)fp.W * fp.H;
        launch_pack_rgb(st0, ctx->sm_count, out, ctx->rgb.p, 0, npx);
        launches++;
        CK(cudaMemcpyAsync(host->rgb, ctx->rgb.p, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, st0));
    }
    CK(cudaEventRecord(ctx->ev1, st0));
    for (int l = 0; l < n_lanes; l++)
        if (ctx->lanes[l].used)
            CK(cudaMemcpyAsync(&ctx->h_counters[l], ctx->lanes[l].counters.p, sizeof(Counters), cudaMemcpyDeviceToHost, st0));
    CK(cudaGetLastError());
    ctx->last_launches = launches;
    ctx->last_batches = batches;
    ctx->last_traced_primary = traced_primary;
    ctx->last_gather_mode = sparse ? 2 : (host && host->foreign ? 1 : 0);
    ctx->last_local_pixels = n_local;
    ctx->frame_pending = true;
    if (ctx->trace_bands)
        std::fprintf(stderr, "[bands] host: %d launches of %d batches enqueued in %.3f ms\n", launches, batches,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    return RT_OK;
}

// Framebuffer of a frame that is not a gather frame: the context's own, except on a gather root, whose two exported halves only
// gather frames may touch (peers rely on their background fill) — such a context renders everything else into a separate buffer.
int plain_framebuffer(rt_ctx* ctx, const FrameParams& fp, float4** out)
{
    const size_t npx = (size_t)fp.W * fp.H;
    if (ctx->gather_root) {
        if (ctx->fb_plain.n < npx || ctx->fb_plain_w != fp.W || ctx->fb_plain_h != fp.H) {
            CK(ctx->fb_plain.ensure(npx));
            CK(cudaMemsetAsync(ctx->fb_plain.p, 0, npx * sizeof(float4), ctx->stream));
            ctx->fb_plain_w = fp.W;
            ctx->fb_plain_h = fp.H;
        }
        *out = ctx->fb_plain.p;
    } else {
        if (ctx->fb.n < npx || ctx->fb_w != fp.W || ctx->fb_h != fp.H) {
            CK(ctx->fb.ensure(npx));
            CK(cudaMemsetAsync(ctx->fb.p, 0, npx * sizeof(float4), ctx->stream));
            ctx->fb_w = fp.W;
            ctx->fb_h = fp.H;
        }
        *out = ctx->fb.p;
    }
    ctx->fb_last = *out;
    ctx->fb_last_w = fp.W;
    ctx->fb_last_h = fp.H;
    return RT_OK;
}

int check_ready(rt_ctx* ctx)
{
    if (ctx->n_tris <= 0)
        return fail(RT_ERR_INVALID, "no scene uploaded (rt_upload_scene)");
    if (!ctx->bvh_built)
        return fail(RT_ERR_INVALID, "BVH not built (rt_build_bvh)");
    if (ctx->n_mats <= 0)
        return fail(RT_ERR_INVALID, "no materials");
    return RT_OK;
}

// Tie keys of the triangle slots and spheres = visiting rank in the reference's own BVH over (triangles, spheres).
int refresh_tie_keys(rt_ctx* ctx)
{
    CK(ctx->d_rank.ensure((size_t)ctx->user_tris + kMaxSpheres + 1));
    const char* err = nullptr;
    if (reference_visit_rank(ctx->stream, ctx->d_pos.p, ctx->user_tris, ctx->d_spheres.p, ctx->n_spheres, ctx->d_rank.p, &err) != 0)
        return fail(RT_ERR_CUDA, std::string("reference visiting order: ") + (err ? err : "failed"));
    {
        const SceneDev sd = ctx->scene_dev();
        launch_apply_tie_keys(ctx->stream, const_cast<float4*>(sd.tri_v0), sd.tri_v2, ctx->d_rank.p, ctx->n_tris, ctx->user_tris);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

} // namespace

extern "C" {

const char* rt_last_error(void) { return g_err.c_str(); }
const char* rt_version(void) { return "rt_b200 0.1 (sm_100a)"; }

int rt_create(int device, rt_ctx** out)
{
    if (!out)
        return fail(RT_ERR_INVALID, "rt_create: null out");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(RT_ERR_CUDA, std::string("rt_create: no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    if (device < 0 || device >= count)
        return fail(RT_ERR_INVALID, "rt_create: device ordinal out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(RT_ERR_CUDA, std::string("rt_create: kernels are built for sm_100a only, found ") + prop.name);
    rt_ctx* ctx = new rt_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess
        || cudaMallocHost(&ctx->h_counters, kMaxLanes * sizeof(Counters)) != cudaSuccess) {
        delete ctx;
        return fail(RT_ERR_CUDA, "rt_create: stream / event / pinned allocation failed");
    }
    ctx->own_stream = true;
    if (const char* e = std::getenv("RTB200_BAND_ORDER")) // developer knob for A/B timing: 0 = bands top to bottom
        ctx->band_order_outer_first = e[0] != '0';
    if (const char* e = std::getenv("RTB200_ZERO_COPY"))
        ctx->zero_copy_host = e[0] != '0';
    if (const char* e = std::getenv("RTB200_CULL_PRIMARY"))
        ctx->cull_primary = e[0] != '0';
    if (const char* e = std::getenv("RTB200_TRACE_BANDS"))
        ctx->trace_bands = e[0] == '1';
    if (const char* e = std::getenv("RTB200_WIDE")) // developer knob: rt_set_wide's mode for new contexts
        ctx->wide_mode = std::max(-1, std::min(1, std::atoi(e)));
    if (const char* e = std::getenv("RTB200_QUEUE_SCALE")) {
        const double v = std::atof(e);
        if (v > 0.0 && v <= 1.0)
            ctx->queue_scale = v;
    }
    bool aux_ok = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming) == cudaSuccess
        && cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming) == cudaSuccess
        && cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) == cudaSuccess;
    for (int l = 0; l < kMaxLanes; l++) {
        rt_ctx::Lane& ln = ctx->lanes[l];
        aux_ok = aux_ok && cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking) == cudaSuccess
            && cudaStreamCreateWithFlags(&ln.side, cudaStreamNonBlocking) == cudaSuccess
            && cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming) == cudaSuccess
            && cudaEventCreateWithFlags(&ln.ev_packed, cudaEventDisableTiming) == cudaSuccess;
        for (int p = 0; p < 2; p++)
            aux_ok = aux_ok && cudaEventCreateWithFlags(&ln.ev_shade[p], cudaEventDisableTiming) == cudaSuccess
                && cudaEventCreateWithFlags(&ln.ev_shadow[p], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!aux_ok) {
        rt_destroy(ctx);
        return fail(RT_ERR_CUDA, "rt_create: side stream / event creation failed");
    }
    std::memset(ctx->h_counters, 0, kMaxLanes * sizeof(Counters));
    *out = ctx;
    return RT_OK;
}

int rt_destroy(rt_ctx* ctx)
{
    if (!ctx)
        return RT_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf<float4>* f4[] = { &ctx->d_plane, &ctx->d_v0, &ctx->d_v1, &ctx->d_v2, &ctx->d_n0, &ctx->d_n1, &ctx->d_n2, &ctx->d_nodes, &ctx->d_mats,
        &ctx->d_point, &ctx->d_sphere, &ctx->accum, &ctx->fb, &ctx->post_img, &ctx->post_a, &ctx->post_b };
    ctx->post_weights.release();
    ctx->post_rgba8.release();
    for (auto* b : f4)
        b->release();
    for (int l = 0; l < kMaxLanes; l++) {
        rt_ctx::Lane& ln = ctx->lanes[l];
        for (int k = 0; k < 2; k++) {
            DevBuf<float4>* q[] = { &ln.q_o[k], &ln.q_d[k], &ln.q_w[k], &ln.sp_p[k], &ln.sp_a[k], &ln.sp_b[k], &ln.ss_p[k], &ln.ss_a[k], &ln.ss_b[k],
                &ln.pl_p[k], &ln.pl_a[k], &ln.pl_b[k], &ln.pl_r[k], &ln.pl_acc[k] };
            for (auto* b : q)
                b->release();
            ln.q_hit[k].release();
            ln.sphere_acc[k].release();
            if (ln.ev_shade[k])
                cudaEventDestroy(ln.ev_shade[k]);
            if (ln.ev_shadow[k])
                cudaEventDestroy(ln.ev_shadow[k]);
        }
        ln.counters.release();
        if (ln.ev_done)
            cudaEventDestroy(ln.ev_done);
        if (ln.ev_packed)
            cudaEventDestroy(ln.ev_packed);
        if (ln.main)
            cudaStreamDestroy(ln.main);
        if (ln.side)
            cudaStreamDestroy(ln.side);
    }
    if (ctx->ev_start)
        cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_copied)
        cudaEventDestroy(ctx->ev_copied);
    if (ctx->copy)
        cudaStreamDestroy(ctx->copy);
    ctx->d_pos.release();
    ctx->d_nrm.release();
    ctx->d_mesh.release();
    ctx->d_perm.release();
    ctx->d_rank.release();
    ctx->prim_id.release();
    ctx->out_id.release();
    ctx->prim_t.release();
    ctx->out_t.release();
    ctx->rgb.release();
    ctx->d_spheres.release();
    ctx->d_mat_glossy.release();
    ctx->d_sphere_glossy.release();
    ctx->d_tex_texels.release();
    ctx->d_tex_table.release();
    ctx->d_mat_tex.release();
    ctx->d_uv.release();
    ctx->d_plane_lights.release();
    ctx->rays_in.release();
    ctx->flag.release();
    ctx->row_flags.release();
    ctx->fb_plain.release();
    if (ctx->wide8_nodes)
        cudaFree(ctx->wide8_nodes);
    cudaFree(ctx->build_scratch.base);
    if (ctx->build_scratch.host_word)
        cudaFreeHost(ctx->build_scratch.host_word);
    if (ctx->lbvh_nodes)
        cudaFree(ctx->lbvh_nodes);
    if (ctx->lbvh_perm)
        cudaFree(ctx->lbvh_perm);
    if (ctx->wide_nodes)
        cudaFree(ctx->wide_nodes);
    if (ctx->h_counters)
        cudaFreeHost(ctx->h_counters);
    for (cudaEvent_t e : ctx->ev_pool)
        cudaEventDestroy(e);
    if (ctx->ev0)
        cudaEventDestroy(ctx->ev0);
    if (ctx->ev1)
        cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RT_OK;
}

int rt_set_stream(rt_ctx* ctx, void* cuda_stream)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream)
        cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return RT_OK;
}

int rt_set_counters(rt_ctx* ctx, int enable)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    ctx->counters_enabled = enable != 0;
    return RT_OK;
}

int rt_set_stage_timing(rt_ctx* ctx, int enable)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    ctx->stage_timing = enable != 0;
    return RT_OK;
}

int rt_stage_times(rt_ctx* ctx, float* ms, int* launches)
{
    if (!ctx || !ms || !launches)
        return fail(RT_ERR_INVALID, "rt_stage_times: bad arguments");
    for (int k = 0; k < RT_STAGE_COUNT; k++) {
        ms[k] = ctx->stage_ms[k];
        launches[k] = ctx->stage_launches[k];
    }
    return RT_OK;
}

int rt_set_wide(rt_ctx* ctx, int mode)
{
    if (!ctx || mode < -1 || mode > 1)
        return fail(RT_ERR_INVALID, "rt_set_wide: mode must be -1 (automatic), 0 (never) or 1 (every level >= 1)");
    ctx->wide_mode = mode;
    return RT_OK;
}

int rt_set_paths(rt_ctx* ctx, int mode)
{
    if (!ctx || mode < -1 || mode > 1)
        return fail(RT_ERR_INVALID, "rt_set_paths: mode must be -1 (automatic), 0 (never) or 1 (whenever legal)");
    ctx->paths_mode = mode;
    return RT_OK;
}

int rt_measure_fp32_peak(rt_ctx* ctx, double* ginst_per_s)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (!ginst_per_s)
        return fail(RT_ERR_INVALID, "rt_measure_fp32_peak: null argument");
    CK(ctx->rays_in.ensure((size_t)ctx->sm_count * 8 * 256));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) { // first launches warm the clocks up; the best of the rest counts
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        const double n = launch_fp32_peak(ctx->stream, ctx->sm_count, ctx->rays_in.p, 8192);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep >= 2 && ms > 0.0f)
            best = std::max(best, n / (ms * 1e-3) / 1e9);
    }
    *ginst_per_s = best;
    return RT_OK;
}

int rt_set_overlap(rt_ctx* ctx, int enable)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    ctx->overlap = enable != 0;
    return RT_OK;
}

int rt_set_pipeline(rt_ctx* ctx, int lanes, int batches_per_frame, unsigned int min_batch_pixels)
{
    if (!ctx || lanes < 0 || lanes > kMaxLanes || (lanes > 0 && batches_per_frame < 1))
        return fail(RT_ERR_INVALID, "rt_set_pipeline: lanes must be 0 (automatic) or 1..4, batches_per_frame >= 1");
    ctx->n_lanes = lanes;
    ctx->batches_per_frame = batches_per_frame;
    ctx->min_batch_pixels = std::max<unsigned>(min_batch_pixels, (unsigned)kTilePixels);
    return RT_OK;
}

int rt_set_batch_rays(rt_ctx* ctx, unsigned int max_primary_rays_per_batch)
{
    if (!ctx || max_primary_rays_per_batch < (unsigned)kTilePixels)
        return fail(RT_ERR_INVALID, "batch size must be at least one tile");
    ctx->batch_rays = max_primary_rays_per_batch;
    return RT_OK;
}

int rt_upload_scene(rt_ctx* ctx, const float* pos, const float* nrm, const int* mesh_id, int64_t n_tris, const rt_material* mats, int n_mats)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_tris < 0 || (n_tris > 0 && (!pos || !nrm)))
        return fail(RT_ERR_INVALID, "rt_upload_scene: need positions and normals");
    // A scene of sphere primitives only (the reference's Spheres preset, src/scene.cpp:80-87) has no triangles: trace one
    // degenerate triangle (zero area -> NaN plane -> never hit, as in the reference) so the BVH paths stay uniform.
    static const float zero9[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    static const rt_material default_mat = { { 0.6f, 0.6f, 0.6f }, { 0, 0, 0 }, 0.0f, 1.0f };
    ctx->user_tris = n_tris;
    if (n_tris == 0) {
        pos = nrm = zero9;
        mesh_id = nullptr;
        n_tris = 1;
        if (!mats || n_mats <= 0) {
            mats = &default_mat;
            n_mats = 1;
        }
    }
    if (n_tris > (1ll << 28) - 1)
        return fail(RT_ERR_INVALID, "rt_upload_scene: at most 2^28-1 triangles");
    if (mesh_id)
        for (int64_t i = 0; i < n_tris; i++)
            if (mesh_id[i] < 0 || mesh_id[i] >= n_mats)
                return fail(RT_ERR_INVALID, "rt_upload_scene: mesh_id out of range of the material table");
    rc = upload_materials(ctx, mats, n_mats);
    if (rc)
        return rc;
    ctx->n_tris = n_tris;
    ctx->h_pos.assign(pos, pos + 9 * n_tris);
    float cm = 0.0f;
    for (float v : ctx->h_pos)
        cm = std::max(cm, std::fabs(v));
    ctx->coord_max = cm;
    for (int a = 0; a < 3; a++) {
        ctx->tri_lo[a] = FLT_MAX;
        ctx->tri_hi[a] = -FLT_MAX;
    }
    for (size_t i = 0; i < ctx->h_pos.size(); i++) { // NaN corners bound nothing (a triangle with one is never hit)
        const float v = ctx->h_pos[i];
        if (v < ctx->tri_lo[i % 3])
            ctx->tri_lo[i % 3] = v;
        if (v > ctx->tri_hi[i % 3])
            ctx->tri_hi[i % 3] = v;
    }
    CK(ctx->d_pos.ensure(9 * (size_t)n_tris));
    CK(ctx->d_nrm.ensure(9 * (size_t)n_tris));
    CK(ctx->d_mesh.ensure((size_t)n_tris));
    CK(cudaMemcpyAsync(ctx->d_pos.p, pos, 9 * (size_t)n_tris * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_nrm.p, nrm, 9 * (size_t)n_tris * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (mesh_id)
        CK(cudaMemcpyAsync(ctx->d_mesh.p, mesh_id, (size_t)n_tris * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    else
        CK(cudaMemsetAsync(ctx->d_mesh.p, 0, (size_t)n_tris * sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->bvh_built = false;
    ctx->have_uv = false; // texture coordinates and texture bindings belong to the previous scene
    ctx->n_textures = 0;
    return RT_OK;
}

int rt_build_bvh(rt_ctx* ctx, int mode)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (ctx->n_tris <= 0)
        return fail(RT_ERR_INVALID, "rt_build_bvh: upload a scene first");
    // Box padding: the reference's triangle test accepts points a few ulp outside the exact triangle (rounded
    // hit point, rounded edge functions); pad so such a point is still inside every ancestor box.
    // The slack covers rounding of the hit point, eps * (|origin| + t): 4e-5 leaves head-room up to the reference camera's
    // farthest zoom (distance 100, trackball.cpp:150) on a unit-scale scene.
    const float pad = 4e-5f * std::max(1.0f, ctx->coord_max);
    const size_t n = (size_t)ctx->n_tris;
    const bool trace_build = std::getenv("RTB200_TRACE_BUILD") != nullptr;
    const auto t_build0 = std::chrono::steady_clock::now();
    auto build_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build0).count(); };
    // a build that fails half-way must not leave the previous tree's (possibly freed) nodes behind a `built` flag
    ctx->bvh_built = false;
    ctx->nodes = nullptr;
    ctx->n_nodes = 0;
    CK(ctx->d_plane.ensure(n * kTriStride));
    CK(ctx->d_n0.ensure(n * kTriStride));
    if (!RT_TRI_AOS) {
        CK(ctx->d_v0.ensure(n));
        CK(ctx->d_v1.ensure(n));
        CK(ctx->d_v2.ensure(n));
        CK(ctx->d_n1.ensure(n));
        CK(ctx->d_n2.ensure(n));
    }
    const int* perm = nullptr;
    if (mode == RT_BVH_AUTO)
        mode = RT_BVH_PLOC_DEVICE;
    if (mode == RT_BVH_SAH_HOST) {
        HostBvh h = build_bvh_sah_host(ctx->h_pos.data(), ctx->n_tris, pad);
        if (h.depth >= kStackDepth)
            return fail(RT_ERR_INVALID, "rt_build_bvh: tree deeper than the traversal stack");
        CK(ctx->d_nodes.ensure(h.nodes.size()));
        CK(ctx->d_perm.ensure(n));
        CK(cudaMemcpyAsync(ctx->d_nodes.p, h.nodes.data(), h.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_perm.p, h.perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->nodes = ctx->d_nodes.p;
        ctx->n_nodes = (int)(h.nodes.size() / 2);
        ctx->root_entry = h.root_entry;
        ctx->bvh_depth = h.depth;
        perm = ctx->d_perm.p;
    } else if (mode == RT_BVH_LBVH_DEVICE || mode == RT_BVH_PLOC_DEVICE) {
        if (ctx->lbvh_nodes)
            cudaFree(ctx->lbvh_nodes);
        if (ctx->lbvh_perm)
            cudaFree(ctx->lbvh_perm);
        ctx->lbvh_nodes = nullptr;
        ctx->lbvh_perm = nullptr;
        DeviceBvh d;
        const char* err = nullptr;
        int built = 2;
        if (mode == RT_BVH_PLOC_DEVICE) {
            if (!ctx->build_scratch.host_word)
                CK(cudaMallocHost(&ctx->build_scratch.host_word, sizeof(unsigned long long)));
            built = build_bvh_ploc_device(ctx->stream, ctx->d_pos.p, ctx->n_tris, pad, &d, &err, &ctx->build_scratch);
            if (ctx->build_scratch.capacity > ((size_t)1 << 30)) { // a 44 M-triangle scene's 12 GB of scratch are not kept
                cudaFree(ctx->build_scratch.base);
                ctx->build_scratch.base = nullptr;
                ctx->build_scratch.capacity = 0;
            }
            if (built == 1)
                return fail(RT_ERR_CUDA, std::string("rt_build_bvh (PLOC): ") + (err ? err : "failed"));
            if (built == 0 && d.depth >= kStackDepth) { // agglomeration does not bound the depth; the Morton hierarchy does (64-bit keys)
                cudaFree(d.nodes);
                cudaFree(d.perm);
                d = DeviceBvh();
                built = 2;
            }
        }
        if (built == 2 && build_bvh_lbvh_device(ctx->stream, ctx->d_pos.p, ctx->n_tris, pad, &d, &err) != 0)
            return fail(RT_ERR_CUDA, std::string("rt_build_bvh (LBVH): ") + (err ? err : "failed"));
        if (d.depth >= kStackDepth) {
            cudaFree(d.nodes);
            cudaFree(d.perm);
            return fail(RT_ERR_INVALID, "rt_build_bvh: LBVH deeper than the traversal stack");
        }
        ctx->lbvh_nodes = d.nodes;
        ctx->lbvh_perm = d.perm;
        ctx->nodes = d.nodes;
        ctx->n_nodes = d.n_nodes;
        ctx->root_entry = d.root_entry;
        ctx->bvh_depth = d.depth;
        perm = d.perm;
    } else {
        return fail(RT_ERR_INVALID, "rt_build_bvh: unknown mode");
    }
#if RT_BVH_WIDE
    {
        if (ctx->wide_nodes)
            cudaFree(ctx->wide_nodes);
        ctx->wide_nodes = nullptr;
        WideBvh w;
        const char* err = nullptr;
        if (collapse_bvh_wide_device(ctx->stream, ctx->nodes, ctx->n_nodes, ctx->root_entry, &w, &err) != 0)
            return fail(RT_ERR_CUDA, std::string("rt_build_bvh (4-wide collapse): ") + (err ? err : "failed"));
        if (3 * w.depth + 1 >= kStackDepth) { // a node step pushes up to three entries
            cudaFree(w.nodes);
            return fail(RT_ERR_INVALID, "rt_build_bvh: 4-wide tree deeper than the traversal stack");
        }
        ctx->wide_nodes = w.nodes;
        if (w.nodes) {
            ctx->nodes = w.nodes;
            ctx->n_nodes = w.n_nodes;
            ctx->bvh_depth = w.depth;
        }
        ctx->root_entry = w.root_entry;
        // the binary tree is no longer needed
        if (ctx->lbvh_nodes) {
            cudaFree(ctx->lbvh_nodes);
            ctx->lbvh_nodes = nullptr;
        }
    }
#endif
    ctx->wide8_built = false; // the 8-wide tree of this scene is collapsed when a frame first wants it (ensure_wide8)
    ctx->wide8_tried = false;
    ctx->hist_valid = false;
    const double t_tree = build_ms();
    {
        const SceneDev sd = ctx->scene_dev();
        auto w = [](const float4* p) { return const_cast<float4*>(p); };
        launch_tri_setup(ctx->stream, ctx->d_pos.p, ctx->d_nrm.p, ctx->d_mesh.p, perm, (int)ctx->n_tris, w(sd.tri_plane), w(sd.tri_v0), w(sd.tri_v1),
            w(sd.tri_v2), w(sd.tri_n0), w(sd.tri_n1), w(sd.tri_n2));
    }
    CK(cudaGetLastError());
    rc = refresh_tie_keys(ctx);
    if (rc)
        return rc;
    ctx->bvh_built = true;
    if (trace_build)
        std::fprintf(stderr, "[build] mode %d: tree %.2f ms, triangle records + reference visiting ranks %.2f ms\n", mode, t_tree, build_ms() - t_tree);
    return RT_OK;
}

int rt_bvh_info(rt_ctx* ctx, int* n_nodes, int* depth)
{
    if (!ctx || !ctx->bvh_built)
        return fail(RT_ERR_INVALID, "rt_bvh_info: no BVH");
    if (n_nodes)
        *n_nodes = ctx->n_nodes;
    if (depth)
        *depth = ctx->bvh_depth;
    return RT_OK;
}

int rt_set_materials(rt_ctx* ctx, const rt_material* mats, int n_mats)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (ctx->n_mats > 0 && n_mats != ctx->n_mats)
        return fail(RT_ERR_INVALID, "rt_set_materials: material count differs from the uploaded scene");
    return upload_materials(ctx, mats, n_mats);
}

// Upload the point-like light table: point lights followed by spot lights, 3 x float4 each.
static int upload_point_like(rt_ctx* ctx)
{
    std::vector<float4> h = ctx->h_point;
    h.insert(h.end(), ctx->h_spot.begin(), ctx->h_spot.end());
    if (h.empty())
        h.resize(3);
    CK(ctx->d_point.ensure(h.size()));
    CK(cudaMemcpyAsync(ctx->d_point.p, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_point = ctx->n_point_user + ctx->n_spot;
    return RT_OK;
}

int rt_set_lights(rt_ctx* ctx, const rt_point_light* point, int n_point, const rt_sphere_light* sphere, int n_sphere)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_point < 0 || n_sphere < 0 || n_point > kMaxPointLights || n_sphere > kMaxSphereLights)
        return fail(RT_ERR_INVALID, "rt_set_lights: too many lights");
    if ((n_point > 0 && !point) || (n_sphere > 0 && !sphere))
        return fail(RT_ERR_INVALID, "rt_set_lights: null array");
    ctx->h_point.assign(3 * (size_t)n_point, make_float4(0, 0, 0, 0));
    for (int i = 0; i < n_point; i++) {
        ctx->h_point[3 * i] = make_float4(point[i].position[0], point[i].position[1], point[i].position[2], 0.0f);
        ctx->h_point[3 * i + 1] = make_float4(point[i].color[0], point[i].color[1], point[i].color[2], 0.0f);
    }
    ctx->n_point_user = n_point;
    std::vector<float4> hs(2 * (size_t)std::max(1, n_sphere));
    for (int i = 0; i < n_sphere; i++) {
        hs[2 * i] = make_float4(sphere[i].position[0], sphere[i].position[1], sphere[i].position[2], sphere[i].radius);
        hs[2 * i + 1] = make_float4(sphere[i].color[0], sphere[i].color[1], sphere[i].color[2], 0.0f);
    }
    CK(ctx->d_sphere.ensure(hs.size()));
    CK(cudaMemcpyAsync(ctx->d_sphere.p, hs.data(), hs.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_sphere = n_sphere;
    return upload_point_like(ctx);
}

int rt_set_spot_lights(rt_ctx* ctx, const rt_spot_light* spot, int n_spot)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_spot < 0 || n_spot > kMaxPointLights || (n_spot > 0 && !spot))
        return fail(RT_ERR_INVALID, "rt_set_spot_lights: between 0 and 16 spot lights");
    ctx->h_spot.assign(3 * (size_t)n_spot, make_float4(0, 0, 0, 0));
    for (int i = 0; i < n_spot; i++) {
        // cos(radians(angle)) with libm on the host, as the reference evaluates it (shadow.cpp:235)
        const float cut = std::cos(spot[i].angle * 0.01745329251994329576923690768489f);
        ctx->h_spot[3 * i] = make_float4(spot[i].position[0], spot[i].position[1], spot[i].position[2], 1.0f);
        ctx->h_spot[3 * i + 1] = make_float4(spot[i].color[0], spot[i].color[1], spot[i].color[2], cut);
        ctx->h_spot[3 * i + 2] = make_float4(spot[i].direction[0], spot[i].direction[1], spot[i].direction[2], 0.0f);
    }
    ctx->n_spot = n_spot;
    return upload_point_like(ctx);
}

int rt_set_plane_lights(rt_ctx* ctx, const rt_plane_light* plane, int n_plane)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_plane < 0 || n_plane > kMaxSphereLights || (n_plane > 0 && !plane))
        return fail(RT_ERR_INVALID, "rt_set_plane_lights: between 0 and 8 plane lights");
    std::vector<float4> h(4 * (size_t)std::max(1, n_plane));
    for (int i = 0; i < n_plane; i++) {
        h[4 * i] = make_float4(plane[i].position[0], plane[i].position[1], plane[i].position[2], 0.0f);
        h[4 * i + 1] = make_float4(plane[i].width[0], plane[i].width[1], plane[i].width[2], 0.0f);
        h[4 * i + 2] = make_float4(plane[i].height[0], plane[i].height[1], plane[i].height[2], 0.0f);
        h[4 * i + 3] = make_float4(plane[i].color[0], plane[i].color[1], plane[i].color[2], 0.0f);
    }
    CK(ctx->d_plane_lights.ensure(h.size()));
    CK(cudaMemcpyAsync(ctx->d_plane_lights.p, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_plane = n_plane;
    return RT_OK;
}

int rt_set_spheres(rt_ctx* ctx, const rt_sphere* spheres, int n_spheres)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_spheres < 0 || n_spheres > kMaxSpheres || (n_spheres > 0 && !spheres))
        return fail(RT_ERR_INVALID, "rt_set_spheres: between 0 and 64 spheres");
    std::vector<float4> h(3 * (size_t)std::max(1, n_spheres));
    bool any_t = false;
    for (int i = 0; i < n_spheres; i++) {
        const rt_sphere& sp = spheres[i];
        h[3 * i] = make_float4(sp.center[0], sp.center[1], sp.center[2], sp.radius);
        h[3 * i + 1] = make_float4(sp.material.kd[0], sp.material.kd[1], sp.material.kd[2], sp.material.shininess);
        h[3 * i + 2] = make_float4(sp.material.ks[0], sp.material.ks[1], sp.material.ks[2], sp.material.transparency);
        any_t |= sp.material.transparency != 1.0f;
    }
    std::vector<float> gd((size_t)std::max(1, n_spheres), 0.0f);
    for (int i = 0; i < n_spheres; i++)
        gd[i] = glossy_cone(spheres[i].material.shininess);
    CK(ctx->d_sphere_glossy.ensure(gd.size()));
    CK(cudaMemcpyAsync(ctx->d_sphere_glossy.p, gd.data(), gd.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_spheres.ensure(h.size()));
    CK(cudaMemcpyAsync(ctx->d_spheres.p, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<float> centres;
    for (int i = 0; i < n_spheres; i++)
        centres.insert(centres.end(), spheres[i].center, spheres[i].center + 3);
    std::vector<float> radii;
    for (int i = 0; i < n_spheres; i++)
        radii.push_back(spheres[i].radius);
    const bool moved = centres != ctx->h_sphere_centres || radii != ctx->h_sphere_radii; // (the rank depends on the centres only; radii for good measure)
    ctx->h_sphere_centres = centres;
    ctx->h_sphere_radii = radii;
    ctx->n_spheres = n_spheres;
    if (moved && ctx->bvh_built) { // the spheres are objects of the reference's BVH: its visiting order changes with them
        rc = refresh_tie_keys(ctx);
        if (rc)
            return rc;
    }
    ctx->spheres_transparent = any_t;
    ctx->any_transparent = ctx->mats_transparent || ctx->spheres_transparent;
    return RT_OK;
}

int rt_set_shard(rt_ctx* ctx, int rank, int world)
{
    if (!ctx || world < 1 || rank < 0 || rank >= world)
        return fail(RT_ERR_INVALID, "rt_set_shard: need 0 <= rank < world");
    ctx->rank = rank;
    ctx->world = world;
    return RT_OK;
}

int rt_render_device(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, void* d_rgba)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_ready(ctx);
    if (rc)
        return rc;
    FrameParams fp;
    rc = make_frame_params(ctx, cam, prm, fp);
    if (rc)
        return rc;
    float4* out = (float4*)d_rgba;
    const size_t npx = (size_t)fp.W * fp.H;
    HostTarget gt;
    bool gather = false;
    if (!out) {
        if (ctx->gather_root && fp.world > 1 && ctx->fb_w == fp.W && ctx->fb_h == fp.H) {
            // gather frame on the root: this frame's half, and the background fill of the other one
            const unsigned half = (unsigned)(ctx->gather_frames & 1);
            out = ctx->fb.p + half * npx;
            gt.prefill = ctx->fb.p + (half ^ 1) * npx;
            gt.prefill_n = npx;
            gather = true;
            ctx->fb_last = out;
            ctx->fb_last_w = fp.W;
            ctx->fb_last_h = fp.H;
        } else {
            rc = plain_framebuffer(ctx, fp, &out);
            if (rc)
                return rc;
        }
    } else if (d_rgba == ctx->peer_base && fp.world > 1) { // gather frame on a peer: rows with hits only, into this frame's half
        out += (ctx->gather_frames & 1) * npx;
        gt.sparse = true;
        gather = true;
    }
    if (gather)
        ctx->gather_frames++;
    gt.foreign = d_rgba != nullptr;
    return enqueue_frame(ctx, fp, out, false, ctx->batch_rays, &gt);
}

int rt_visible_rect(const rt_camera* cam, int width, int height, const float lo[3], const float hi[3], int rect[4])
{
    if (!cam || !lo || !hi || !rect || width <= 0 || height <= 0)
        return fail(RT_ERR_INVALID, "rt_visible_rect: null argument or empty image");
    FrameParams fp;
    std::memset(&fp, 0, sizeof(fp));
    fp.W = width;
    fp.H = height;
    camera_to_frame(cam, fp);
    const double l[3] = { lo[0], lo[1], lo[2] }, h[3] = { hi[0], hi[1], hi[2] };
    box_pixel_rect(l, h, fp);
    rect[0] = fp.vis_x0;
    rect[1] = fp.vis_x1;
    rect[2] = fp.vis_y0;
    rect[3] = fp.vis_y1;
    return RT_OK;
}

// Device-to-host rate of this context's link, once: a 16 MB copy into page-locked memory, best of three.
static int measure_store_rate(rt_ctx* ctx)
{
    if (ctx->store_gbs > 0.0)
        return RT_OK;
    const size_t bytes = (size_t)16 << 20;
    void* h = nullptr;
    DevBuf<unsigned char> d;
    CK(d.ensure(bytes));
    if (cudaMallocHost(&h, bytes) != cudaSuccess) {
        cudaGetLastError();
        d.release();
        ctx->store_gbs = 20.0;
        return RT_OK;
    }
    float best = 1e30f;
    for (int k = 0; k < 3; k++) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        cudaMemcpyAsync(h, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess && ms > 0.0f)
            best = std::min(best, ms);
    }
    cudaFreeHost(h);
    d.release();
    ctx->store_gbs = best < 1e29f ? 0.85 * (double)bytes / (best * 1e6) : 20.0;
    return RT_OK;
}

int rt_host_register(void* p, size_t bytes)
{
    if (!p || !bytes)
        return fail(RT_ERR_INVALID, "rt_host_register: empty range");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RT_ERR_CUDA, std::string("rt_host_register: ") + cudaGetErrorString(e));
    }
    return RT_OK;
}

int rt_host_alloc(size_t bytes, void** p)
{
    if (!p || !bytes)
        return fail(RT_ERR_INVALID, "rt_host_alloc: empty request");
    *p = nullptr;
    const cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(RT_ERR_CUDA, std::string("rt_host_alloc: ") + cudaGetErrorString(e));
    }
    return RT_OK;
}

int rt_host_free(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess)
        cudaGetLastError();
    return RT_OK;
}

int rt_host_unregister(void* p)
{
    if (p && cudaHostUnregister(p) != cudaSuccess)
        cudaGetLastError();
    return RT_OK;
}

int rt_current_device(void)
{
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        device = 0;
    }
    return device;
}

int rt_set_host_store_rate(rt_ctx* ctx, double gbs)
{
    if (!ctx || !(gbs >= 0.0))
        return fail(RT_ERR_INVALID, "rt_set_host_store_rate: need a context and a rate >= 0");
    ctx->shared_store_gbs = gbs;
    return RT_OK;
}

int rt_render_shard(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, float* rgb_host_mapped, rt_stats* stats)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_ready(ctx);
    if (rc)
        return rc;
    if (!rgb_host_mapped)
        return fail(RT_ERR_INVALID, "rt_render_shard: rgb_host_mapped is null");
    cudaPointerAttributes attr {};
    if (cudaPointerGetAttributes(&attr, rgb_host_mapped) != cudaSuccess || !attr.devicePointer
        || (attr.type != cudaMemoryTypeHost && attr.type != cudaMemoryTypeManaged && attr.type != cudaMemoryTypeDevice)) {
        cudaGetLastError();
        return fail(RT_ERR_INVALID, "rt_render_shard: the buffer must be page-locked host memory mapped into the device (cudaHostRegister / cudaHostAlloc)");
    }
    FrameParams fp;
    rc = make_frame_params(ctx, cam, prm, fp);
    if (rc)
        return rc;
    rc = measure_store_rate(ctx);
    if (rc)
        return rc;
    float4* fb = nullptr;
    rc = plain_framebuffer(ctx, fp, &fb);
    if (rc)
        return rc;
    unsigned batch = ctx->batch_rays;
    for (int attempt = 0;; attempt++) {
        HostTarget host;
        host.mapped_rgb = static_cast<float*>(attr.devicePointer);
        host.early_background = ctx->zero_copy_host && fp.spp == 1; // (a frame with several samples per pixel is stored when it is complete)
        rc = enqueue_frame(ctx, fp, fb, false, batch, &host);
        if (rc)
            return rc;
        rc = rt_sync(ctx, stats);
        // queue overflow from ray splitting: twice the head-room per primary ray, on half the batch where the batch can shrink
        if (rc == RT_ERR_OVERFLOW && ctx->last_overflow == 1 && attempt < 6) {
            ctx->headroom_shift++;
            if (batch / 2 >= (unsigned)kTilePixels * (unsigned)fp.spp)
                batch /= 2;
            continue;
        }
        ctx->headroom_shift = 0;
        return rc;
    }
}

int rt_sync(rt_ctx* ctx, rt_stats* stats)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->frame_pending) {
        ctx->frame_pending = false;
        for (int k = 0; k < RT_STAGE_COUNT; k++) {
            ctx->stage_ms[k] = 0.0f;
            ctx->stage_launches[k] = 0;
        }
        for (size_t k = 0; k + 1 < ctx->ev_used; k += 2) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ctx->ev_pool[k], ctx->ev_pool[k + 1]) == cudaSuccess) {
                ctx->stage_ms[ctx->ev_stage[k / 2]] += ms;
                ctx->stage_launches[ctx->ev_stage[k / 2]]++;
                static const bool timeline = std::getenv("RTB200_TRACE_LAUNCHES") != nullptr; // developer: when every timed launch began
                if (timeline) {
                    float t0 = 0.0f;
                    cudaEventElapsedTime(&t0, ctx->ev0, ctx->ev_pool[k]);
                    std::fprintf(stderr, "[launch] stage %d  begin %7.3f ms  took %7.3f ms\n", ctx->ev_stage[k / 2], t0, ms);
                }
            }
        }
        if (ctx->trace_bands) {
            for (size_t k = 0; k < ctx->trace_ev.size(); k++) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ctx->ev0, ctx->trace_ev[k]);
                std::fprintf(stderr, "[bands] %7.3f ms %s\n", ms, ctx->trace_what[k].c_str());
                cudaEventDestroy(ctx->trace_ev[k]);
            }
            ctx->trace_ev.clear();
            ctx->trace_what.clear();
        }
        for (int l = 0; l < kMaxLanes; l++) // the queue fills of this frame are the next frame's expectations
            if (ctx->lanes[l].used) {
                std::memcpy(ctx->hist_ext[l], ctx->h_counters[l].level_ext, sizeof(ctx->hist_ext[l]));
                std::memcpy(ctx->hist_sh[l], ctx->h_counters[l].level_sh, sizeof(ctx->hist_sh[l]));
            }
        ctx->hist_signature = ctx->frame_signature;
        ctx->hist_valid = true;
        Counters c; // totals over the lanes used by the frame
        std::memset(&c, 0, sizeof(c));
        for (int l = 0; l < kMaxLanes; l++) {
            if (!ctx->lanes[l].used)
                continue;
            const Counters& h = ctx->h_counters[l];
            c.overflow = std::max(c.overflow, h.overflow);
            c.primary_rays += h.primary_rays;
            c.shadow_queries += h.shadow_queries;
            c.secondary_rays += h.secondary_rays;
            c.flagged_rows += h.flagged_rows;
            c.node_visits += h.node_visits;
            c.tri_tests += h.tri_tests;
            c.tri_tests_full += h.tri_tests_full;
            c.ext_node_visits += h.ext_node_visits;
            c.ext_tri_tests += h.ext_tri_tests;
            c.ext_tri_tests_full += h.ext_tri_tests_full;
            c.max_ray_nodes = std::max(c.max_ray_nodes, h.max_ray_nodes);
            c.max_ray_tris = std::max(c.max_ray_tris, h.max_ray_tris);
        }
        if (ctx->counters_enabled && std::getenv("RTB200_TRACE_LAUNCHES"))
            std::fprintf(stderr, "[counters] longest query: %u boxes, %u triangles\n", c.max_ray_nodes, c.max_ray_tris);
        ctx->last_overflow = c.overflow;
        if (c.overflow == 1)
            return fail(RT_ERR_OVERFLOW, "a ray queue overflowed (transparent materials doubled the wavefront more than provisioned); lower rt_set_batch_rays");
        if (c.overflow >= 2)
            return fail(RT_ERR_OVERFLOW, "traversal stack overflow");
        if (stats) {
            std::memset(stats, 0, sizeof(*stats));
            stats->primary_rays = c.primary_rays;
            stats->shadow_queries = c.shadow_queries;
            stats->secondary_rays = c.secondary_rays;
            stats->node_visits = c.node_visits;
            stats->tri_tests = c.tri_tests;
            stats->tri_tests_full = c.tri_tests_full;
            stats->extend_node_visits = c.ext_node_visits;
            stats->extend_tri_tests = c.ext_tri_tests;
            stats->extend_tri_tests_full = c.ext_tri_tests_full;
            float ms = 0.0f;
            CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            stats->gpu_ms = ms;
            stats->kernel_launches = ctx->last_launches;
            stats->batches = ctx->last_batches;
            stats->traced_primary_rays = ctx->last_traced_primary;
            stats->gather_bytes = ctx->last_gather_mode == 2 ? (uint64_t)c.flagged_rows * kTileW * sizeof(float4)
                : ctx->last_gather_mode == 1                  ? (uint64_t)ctx->last_local_pixels * sizeof(float4)
                                                              : 0;
        }
    }
    return RT_OK;
}

int rt_framebuffer(rt_ctx* ctx, void** d_rgba, int* width, int* height)
{
    if (!ctx || !ctx->fb_last)
        return fail(RT_ERR_INVALID, "rt_framebuffer: nothing rendered yet");
    if (d_rgba)
        *d_rgba = ctx->fb_last;
    if (width)
        *width = ctx->fb_last_w;
    if (height)
        *height = ctx->fb_last_h;
    return RT_OK;
}

int rt_framebuffer_ipc_handle(rt_ctx* ctx, int width, int height, void* handle64)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (width <= 0 || height <= 0 || !handle64)
        return fail(RT_ERR_INVALID, "rt_framebuffer_ipc_handle: bad arguments");
    // two halves for alternate gather frames, both background to begin with (see rt_ctx::gather_root)
    const size_t npx = (size_t)width * height;
    if (ctx->fb.n < 2 * npx || ctx->fb_w != width || ctx->fb_h != height || !ctx->gather_root) {
        CK(ctx->fb.ensure(2 * npx));
        ctx->fb_w = width;
        ctx->fb_h = height;
    }
    launch_fill_background(ctx->stream, ctx->sm_count, ctx->fb.p, 2 * npx);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->gather_root = true;
    ctx->gather_frames = 0;
    ctx->fb_last = ctx->fb.p;
    ctx->fb_last_w = width;
    ctx->fb_last_h = height;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->fb.p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &h, 64);
    return RT_OK;
}

int rt_open_peer_framebuffer(rt_ctx* ctx, const void* handle64, void** d_rgba)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (!handle64 || !d_rgba)
        return fail(RT_ERR_INVALID, "rt_open_peer_framebuffer: bad arguments");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(d_rgba, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_base = *d_rgba;
    ctx->gather_frames = 0;
    return RT_OK;
}

int rt_set_gather_target(rt_ctx* ctx, void* d_rgba)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    ctx->peer_base = d_rgba;
    ctx->gather_frames = 0;
    return RT_OK;
}

int rt_close_peer_framebuffer(rt_ctx* ctx, void* d_rgba)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    CK(cudaIpcCloseMemHandle(d_rgba));
    if (d_rgba == ctx->peer_base)
        ctx->peer_base = nullptr;
    return RT_OK;
}

int rt_download_rgb(rt_ctx* ctx, const void* d_rgba, int width, int height, float* rgb_out)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (!d_rgba || !rgb_out || width <= 0 || height <= 0)
        return fail(RT_ERR_INVALID, "rt_download_rgb: bad arguments");
    const size_t npx = (size_t)width * height;
    CK(ctx->rgb.ensure(npx * 3 + 4));
    launch_pack_rgb(ctx->stream, ctx->sm_count, (const float4*)d_rgba, ctx->rgb.p, 0, npx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb_out, ctx->rgb.p, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_render(rt_ctx* ctx, const rt_camera* cam, const rt_params* prm, float* rgb_out, int* tri_id_out, float* t_out, rt_stats* stats)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_ready(ctx);
    if (rc)
        return rc;
    if (!rgb_out)
        return fail(RT_ERR_INVALID, "rt_render: rgb_out is null");
    FrameParams fp;
    rc = make_frame_params(ctx, cam, prm, fp);
    if (rc)
        return rc;
    const size_t npx = (size_t)fp.W * fp.H;
    const bool want_ids = tri_id_out != nullptr || t_out != nullptr;
    float4* fb = nullptr;
    rc = plain_framebuffer(ctx, fp, &fb);
    if (rc)
        return rc;
    if (want_ids) {
        CK(ctx->out_id.ensure(npx));
        CK(ctx->out_t.ensure(npx));
        CK(cudaMemsetAsync(ctx->out_id.p, 0xff, npx * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->out_t.p, 0, npx * sizeof(float), ctx->stream));
    }
    CK(ctx->rgb.ensure(npx * 3 + 4));
    unsigned batch = ctx->batch_rays;
    // Page-locked memory that this device can address (cudaHostAlloc / cudaHostRegister): the kernels store the frame into it
    // themselves and the background rows leave right after level 0 (HostTarget); otherwise bands go through the copy engine.
    // Frames with several samples per pixel take that path when they fit the batches the lanes render side by side (the rows with
    // hits are stored unpaced when a batch is resolved, which must not run next to a later batch's traversal for long).
    float* mapped = nullptr;
    const bool big = npx >= ((size_t)1 << 22);
    const bool fits = fp.spp == 1 || (ctx->n_lanes <= 0 && npx * (size_t)fp.spp <= (big ? 2 : 1) * (size_t)ctx->batch_rays);
    if (ctx->zero_copy_host && fp.world == 1 && fits && !(ctx->post_on && post_has_effect(ctx->post))) {
        cudaPointerAttributes attr {};
        if (cudaPointerGetAttributes(&attr, rgb_out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
            mapped = static_cast<float*>(attr.devicePointer);
        cudaGetLastError();
        if (mapped) {
            rc = measure_store_rate(ctx);
            if (rc)
                return rc;
        }
    }
    for (int attempt = 0;; attempt++) {
        HostTarget host;
        if (mapped) {
            host.mapped_rgb = mapped;
            host.early_background = true;
        } else
            host.rgb = rgb_out; // finished bands are packed and copied by the lanes while the rest of the frame renders
        rc = enqueue_frame(ctx, fp, fb, want_ids, batch, &host);
        if (rc)
            return rc;
        if (tri_id_out)
            CK(cudaMemcpyAsync(tri_id_out, ctx->out_id.p, npx * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (t_out)
            CK(cudaMemcpyAsync(t_out, ctx->out_t.p, npx * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        rc = rt_sync(ctx, stats);
        // queue overflow from ray splitting: twice the head-room per primary ray, on half the batch where the batch can shrink
        if (rc == RT_ERR_OVERFLOW && ctx->last_overflow == 1 && attempt < 6) {
            ctx->headroom_shift++;
            if (batch / 2 >= (unsigned)kTilePixels * (unsigned)fp.spp)
                batch /= 2;
            continue;
        }
        ctx->headroom_shift = 0;
        return rc;
    }
}

int rt_set_texcoords(rt_ctx* ctx, const float* uv)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (ctx->n_tris <= 0)
        return fail(RT_ERR_INVALID, "rt_set_texcoords: upload a scene first");
    if (!uv || ctx->user_tris == 0) {
        ctx->have_uv = false;
        return RT_OK;
    }
    CK(ctx->d_uv.ensure(3 * (size_t)ctx->user_tris));
    CK(cudaMemcpyAsync(ctx->d_uv.p, uv, 6 * (size_t)ctx->user_tris * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_uv = true;
    return RT_OK;
}

int rt_set_textures(rt_ctx* ctx, const rt_texture* textures, int n_textures, const int* material_texture, int n_materials)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (n_textures < 0 || (n_textures > 0 && (!textures || !material_texture)))
        return fail(RT_ERR_INVALID, "rt_set_textures: need textures and one texture index per material");
    if (n_textures == 0) {
        ctx->n_textures = 0;
        return RT_OK;
    }
    if (n_materials != ctx->n_mats)
        return fail(RT_ERR_INVALID, "rt_set_textures: material count differs from the uploaded scene");
    std::vector<int4> table((size_t)n_textures);
    size_t total = 0;
    for (int k = 0; k < n_textures; k++) {
        if (textures[k].width <= 0 || textures[k].height <= 0 || !textures[k].rgb)
            return fail(RT_ERR_INVALID, "rt_set_textures: empty texture");
        if (total + (size_t)textures[k].width * textures[k].height > 0x7fffffffu)
            return fail(RT_ERR_INVALID, "rt_set_textures: more than 2^31 texels");
        // a square power-of-two texture gets a mip pyramid down to 1 x 1 (Image::canUseMipmapping / initMipmap, src/image.cpp:400-430):
        // its levels follow level 0 back to back; table.w = number of levels, 0 = no pyramid
        const int tw = textures[k].width, th = textures[k].height;
        const bool pyramid = ((th & (th - 1)) == 0) && ((tw & (tw - 1)) == 0) && tw == th;
        int levels = 0;
        size_t texels_k = (size_t)tw * th;
        if (pyramid) {
            levels = 1;
            for (int w = tw; w > 1; w /= 2, levels++)
                texels_k += (size_t)(w / 2) * (w / 2);
        }
        table[k] = make_int4((int)total, tw, th, levels);
        total += texels_k;
        if (total > 0x7fffffffu)
            return fail(RT_ERR_INVALID, "rt_set_textures: more than 2^31 texels");
    }
    for (int m = 0; m < n_materials; m++)
        if (material_texture[m] < -1 || material_texture[m] >= n_textures)
            return fail(RT_ERR_INVALID, "rt_set_textures: texture index out of range");
    std::vector<float4> texels(total);
    for (int k = 0; k < n_textures; k++) {
        const size_t n = (size_t)textures[k].width * textures[k].height;
        for (size_t i = 0; i < n; i++)
            texels[table[k].x + i] = make_float4(textures[k].rgb[3 * i], textures[k].rgb[3 * i + 1], textures[k].rgb[3 * i + 2], 0.0f);
        // getReducedResolutionTexture (src/image.cpp:377-397): every texel of the next level = 0.25f * (upper left + lower left + upper right + lower right)
        size_t src = (size_t)table[k].x;
        for (int level = 1, w = textures[k].width; level < table[k].w; level++, w /= 2) {
            const size_t dst = src + (size_t)w * w;
            const int rw = w / 2;
            for (int y = 0; y < rw; y++)
                for (int x = 0; x < rw; x++) {
                    const float4 lu = texels[src + (size_t)(2 * y) * w + 2 * x], ll = texels[src + (size_t)(2 * y + 1) * w + 2 * x];
                    const float4 ru = texels[src + (size_t)(2 * y) * w + 2 * x + 1], rl = texels[src + (size_t)(2 * y + 1) * w + 2 * x + 1];
                    texels[dst + (size_t)y * rw + x] = make_float4(0.25f * (((lu.x + ll.x) + ru.x) + rl.x), 0.25f * (((lu.y + ll.y) + ru.y) + rl.y),
                        0.25f * (((lu.z + ll.z) + ru.z) + rl.z), 0.0f);
                }
            src = dst;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream)); // an earlier frame may still sample the old texels
    CK(ctx->d_tex_texels.ensure(total));
    CK(ctx->d_tex_table.ensure((size_t)n_textures));
    CK(ctx->d_mat_tex.ensure((size_t)n_materials));
    CK(cudaMemcpyAsync(ctx->d_tex_texels.p, texels.data(), total * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tex_table.p, table.data(), table.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_mat_tex.p, material_texture, (size_t)n_materials * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_textures = n_textures;
    return RT_OK;
}

int rt_set_texturing(rt_ctx* ctx, const rt_texture_params* p)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    if (!p) {
        ctx->tex_on = false;
        ctx->tex_params = rt_texture_params {}; // the reference's defaults (src/main.cpp:54-57): nearest, border, black
        return RT_OK;
    }
    if (p->filtering < RT_TEX_NEAREST || p->filtering > RT_TEX_TRILINEAR)
        return fail(RT_ERR_INVALID, "rt_set_texturing: unknown filtering method");
    if (p->out_of_bounds_x < RT_OOB_BORDER || p->out_of_bounds_x > RT_OOB_REPEAT || p->out_of_bounds_y < RT_OOB_BORDER || p->out_of_bounds_y > RT_OOB_REPEAT)
        return fail(RT_ERR_INVALID, "rt_set_texturing: unknown out-of-bounds rule");
    ctx->tex_params = *p;
    ctx->tex_on = true;
    return RT_OK;
}

int rt_set_postprocess(rt_ctx* ctx, const rt_post_params* post)
{
    if (!ctx)
        return fail(RT_ERR_INVALID, "null context");
    if (!post) {
        ctx->post_on = false;
        return RT_OK;
    }
    int rc = check_post(post);
    if (rc)
        return rc;
    rc = use_device(ctx);
    if (rc)
        return rc;
    ctx->post = normalised_post(*post);
    ctx->post_on = true;
    return ensure_post_weights(ctx, ctx->post);
}

int rt_postprocess_device(rt_ctx* ctx, const rt_post_params* post, void* d_rgba, int width, int height)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_post(post);
    if (rc)
        return rc;
    float4* img = (float4*)d_rgba;
    if (!img) {
        if (!ctx->fb_last || ctx->fb_last_w != width || ctx->fb_last_h != height)
            return fail(RT_ERR_INVALID, "rt_postprocess_device: the context holds no framebuffer of that size");
        img = ctx->fb_last;
    }
    if (width <= 0 || height <= 0)
        return fail(RT_ERR_INVALID, "rt_postprocess_device: empty image");
    int launches = 0;
    const rt_post_params p = normalised_post(*post);
    rc = ensure_post_weights(ctx, p);
    if (rc)
        return rc;
    const bool timing = ctx->stage_timing;
    ctx->stage_timing = false; // stage events belong to frames (rt_sync reads them back)
    rc = enqueue_postprocess(ctx, ctx->stream, img, width, height, p, launches);
    ctx->stage_timing = timing;
    return rc;
}

int rt_postprocess(rt_ctx* ctx, const rt_post_params* post, float* rgb, int width, int height, int via_write_bitmap, unsigned char* rgba8)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_post(post);
    if (rc)
        return rc;
    if (!rgb || width <= 0 || height <= 0)
        return fail(RT_ERR_INVALID, "rt_postprocess: need an image");
    const size_t n = (size_t)width * height;
    CK(ctx->rgb.ensure(n * 3 + 4));
    CK(ctx->post_img.ensure(n));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->rgb.p, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    launch_unpack_rgb(st, ctx->sm_count, ctx->rgb.p, ctx->post_img.p, n);
    const rt_post_params p = normalised_post(*post);
    rc = ensure_post_weights(ctx, p);
    if (rc)
        return rc;
    int launches = 0;
    const bool timing = ctx->stage_timing;
    ctx->stage_timing = false;
    if (via_write_bitmap) { // writeBitmapToFile (src/screen.cpp:40-53): bloom whatever bloom_live says, no gamma
        rc = enqueue_bloom(ctx, st, ctx->post_img.p, width, height, p, launches);
        if (rc == RT_OK && rgba8) {
            CK(ctx->post_rgba8.ensure(n * 4));
            launch_post_rgba8(st, ctx->sm_count, ctx->post_img.p, ctx->post_rgba8.p, n);
            CK(cudaMemcpyAsync(rgba8, ctx->post_rgba8.p, n * 4, cudaMemcpyDeviceToHost, st));
        }
    } else {
        rc = enqueue_postprocess(ctx, st, ctx->post_img.p, width, height, p, launches);
    }
    ctx->stage_timing = timing;
    if (rc)
        return rc;
    launch_pack_rgb(st, ctx->sm_count, ctx->post_img.p, ctx->rgb.p, 0, n);
    CK(cudaMemcpyAsync(rgb, ctx->rgb.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return RT_OK;
}

int rt_intersect(rt_ctx* ctx, const float* rays, int64_t n_rays, int use_bvh, int* tri_id, float* t)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    rc = check_ready(ctx);
    if (rc)
        return rc;
    if (n_rays < 0 || (n_rays > 0 && (!rays || !tri_id || !t)))
        return fail(RT_ERR_INVALID, "rt_intersect: bad arguments");
    if (n_rays == 0)
        return RT_OK;
    CK(ctx->rays_in.ensure(6 * (size_t)n_rays));
    CK(ctx->out_id.ensure((size_t)n_rays));
    CK(ctx->out_t.ensure((size_t)n_rays));
    CK(ctx->flag.ensure(1));
    CK(cudaMemsetAsync(ctx->flag.p, 0, sizeof(unsigned), ctx->stream));
    CK(cudaMemcpyAsync(ctx->rays_in.p, rays, 6 * (size_t)n_rays * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SceneDev s_int = ctx->scene_dev();
    s_int.tie_by_id = use_bvh ? 0 : 1;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (use_bvh == 2) { // experiment: the 8-wide tree, eight lanes per ray
        rc = ensure_wide8(ctx);
        if (rc)
            return rc;
        if (!ctx->wide8_built)
            return fail(RT_ERR_INVALID, "rt_intersect: no 8-wide tree for this scene");
#if RT_CHECKED
        s_int.n_wide_nodes = ctx->wide8_count;
#endif
        launch_intersect_wide(ctx->stream, ctx->sm_count, s_int, ctx->wide8_nodes, ctx->wide8_root, ctx->rays_in.p, (long long)n_rays, ctx->out_id.p, ctx->out_t.p);
    } else
        launch_intersect(ctx->stream, ctx->sm_count, s_int, ctx->root_entry, ctx->rays_in.p, (long long)n_rays, use_bvh, ctx->out_id.p,
            ctx->out_t.p, ctx->flag.p);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaGetLastError());
    unsigned flag = 0;
    CK(cudaMemcpyAsync(tri_id, ctx->out_id.p, (size_t)n_rays * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(t, ctx->out_t.p, (size_t)n_rays * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&flag, ctx->flag.p, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_intersect_ms, ctx->ev0, ctx->ev1);
    if (flag)
        return fail(RT_ERR_OVERFLOW, "traversal stack overflow");
    return RT_OK;
}

float rt_last_intersect_ms(rt_ctx* ctx) { return ctx ? ctx->last_intersect_ms : 0.0f; }

int rt_checked_build(void) { return RT_CHECKED ? 1 : 0; }

int rt_violations_selftest(rt_ctx* ctx)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    provoke_violation_kernels(ctx->stream);
    provoke_violation_wide8(ctx->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_violations(rt_ctx* ctx, unsigned int* counts, int n_counts)
{
    int rc = use_device(ctx);
    if (rc)
        return rc;
    if (!counts || n_counts < 0)
        return fail(RT_ERR_INVALID, "rt_violations: bad arguments");
    CK(cudaDeviceSynchronize());
    unsigned int all[kChkSites] = {};
    add_violations_kernels(all);
    add_violations_wide8(all);
    for (int k = 0; k < n_counts; k++)
        counts[k] = k < kChkSites ? all[k] : 0u;
    return RT_OK;
}

// ---- OBJ loading: host only ----
struct rt_mesh_soup {
    std::vector<float> pos, nrm, uv;
    std::vector<int> mesh_id;
    std::vector<rt_material> mats;
    std::vector<int> mat_tex;                 // texture of each mesh, -1 = none
    std::vector<std::vector<float>> texels;   // one per distinct texture file
    std::vector<rt_texture> textures;         // views of `texels`
};

int rt_load_obj(const char* path, int center_and_normalize, rt_mesh_soup** out)
{
    if (!path || !out)
        return fail(RT_ERR_INVALID, "rt_load_obj: bad arguments");
    *out = nullptr;
    try {
        std::vector<Mesh> meshes = loadMesh(path, center_and_normalize != 0);
        auto* s = new rt_mesh_soup();
        std::vector<std::filesystem::path> texture_paths;
        int mi = 0;
        for (const Mesh& mesh : meshes) {
            for (const Triangle& tri : mesh.triangles) {
                const Vertex* v[3] = { &mesh.vertices[tri.x], &mesh.vertices[tri.y], &mesh.vertices[tri.z] };
                for (int k = 0; k < 3; k++) {
                    s->pos.insert(s->pos.end(), { v[k]->p.x, v[k]->p.y, v[k]->p.z });
                    s->nrm.insert(s->nrm.end(), { v[k]->n.x, v[k]->n.y, v[k]->n.z });
                    s->uv.insert(s->uv.end(), { v[k]->texCoord.x, v[k]->texCoord.y });
                }
                s->mesh_id.push_back(mi);
            }
            const Material& m = mesh.material;
            s->mats.push_back(rt_material { { m.kd.x, m.kd.y, m.kd.z }, { m.ks.x, m.ks.y, m.ks.z }, m.shininess, m.transparency });
            int tex = -1;
            if (m.kdTexture) { // one entry per distinct file
                const Image& img = *m.kdTexture;
                for (size_t k = 0; k < texture_paths.size() && tex < 0; k++)
                    if (texture_paths[k] == img.path())
                        tex = (int)k;
                if (tex < 0) {
                    tex = (int)texture_paths.size();
                    texture_paths.push_back(img.path());
                    std::vector<float> t;
                    t.reserve(img.pixels().size() * 3);
                    for (const glm::vec3& c : img.pixels())
                        t.insert(t.end(), { c.x, c.y, c.z });
                    s->texels.push_back(std::move(t));
                    s->textures.push_back(rt_texture { img.width(), img.height(), nullptr });
                }
            }
            s->mat_tex.push_back(tex);
            mi++;
        }
        for (size_t k = 0; k < s->textures.size(); k++)
            s->textures[k].rgb = s->texels[k].data();
        *out = s;
        return RT_OK;
    } catch (const std::exception& e) {
        return fail(RT_ERR_IO, e.what());
    }
}

int64_t rt_soup_num_triangles(const rt_mesh_soup* s) { return s ? (int64_t)s->mesh_id.size() : 0; }
int rt_soup_num_meshes(const rt_mesh_soup* s) { return s ? (int)s->mats.size() : 0; }
const float* rt_soup_positions(const rt_mesh_soup* s) { return s ? s->pos.data() : nullptr; }
const float* rt_soup_normals(const rt_mesh_soup* s) { return s ? s->nrm.data() : nullptr; }
const int* rt_soup_mesh_ids(const rt_mesh_soup* s) { return s ? s->mesh_id.data() : nullptr; }
const rt_material* rt_soup_materials(const rt_mesh_soup* s) { return s ? s->mats.data() : nullptr; }
const float* rt_soup_texcoords(const rt_mesh_soup* s) { return s ? s->uv.data() : nullptr; }
int rt_soup_num_textures(const rt_mesh_soup* s) { return s ? (int)s->textures.size() : 0; }
const rt_texture* rt_soup_textures(const rt_mesh_soup* s) { return s ? s->textures.data() : nullptr; }
const int* rt_soup_material_textures(const rt_mesh_soup* s) { return s ? s->mat_tex.data() : nullptr; }
void rt_soup_free(rt_mesh_soup* s) { delete s; }

} // extern "C"
