// Launch wrappers of the wavefront kernels (definitions in rt_kernels.cu, rt_lbvh.cu).
#pragma once
#include "rt_types.h"
#include <vector>

namespace rtb {

void launch_tri_setup(cudaStream_t st, const float* pos, const float* nrm, const int* mesh_id, const int* perm, int n,
    float4* plane, float4* v0, float4* v1, float4* v2, float4* n0, float4* n1, float4* n2);
// n_current >= 0: fill of the current ray queue (level 0: primary rays are generated inside extend / shade from the
// pixel index, the queue only provides the hit slots); add_primary: primary rays of the batch (counted on the host)
void launch_level_reset(cudaStream_t st, Counters* c, int next_q, long long n_current, unsigned long long add_primary, int par);
void launch_extend(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int qi,
    int level, unsigned first_lp, bool count);
void launch_shade(cudaStream_t st, int sm_count, const SceneDev& s, const FrameParams& fp, const BatchDev& b, int qi, int level, unsigned first_lp);
void launch_shadow_point(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int level, bool count);
// The same two queues through the 8-wide tree with eight lanes per ray (small queues; extend: levels >= 1; shadow: every material opaque).
void launch_extend_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int wide_root, const BatchDev& b, int qi, int level);
void launch_shadow_point_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int wide_root, const BatchDev& b, int level);
// Bounce levels >= first_level of the rays in queue qi as whole paths (k_paths): small wavefronts of opaque scenes with point-like lights only.
void launch_paths(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, int qi, int first_level, bool count);
void launch_shadow_sphere(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count);
void launch_shadow_plane(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const FrameParams& fp, const BatchDev& b, bool count);
void launch_resolve(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned first_lp, unsigned n_lp, const float4* accum,
    const int* prim_id, const float* prim_t, float4* out, int* out_id, float* out_t, const unsigned char* row_flags = nullptr);
void launch_fill_background(cudaStream_t st, int sm_count, float4* out, size_t n);
void launch_pack_rgb(cudaStream_t st, int sm_count, const float4* in, float* out, size_t p0, size_t p1);
void launch_pack_rgb_tiles(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned tile0, unsigned n_tiles, const unsigned char* flags,
    const float4* in, float* out);
void launch_row_flags(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned first_lp, unsigned n_lp, const int2* hit, unsigned char* flags,
    unsigned* n_flagged);
void launch_host_background(cudaStream_t st, int sm_count, const FrameParams& fp, unsigned tile0, unsigned n_tiles, const unsigned char* flags, float* out,
    double gbs, bool gbs_is_shared_rate);
// FMUL / FADD issue-rate microbenchmark; returns the number of FP32 lane-instructions the launch executes.  scratch: >= sm_count * 8 * 256 floats.
double launch_fp32_peak(cudaStream_t st, int sm_count, float* scratch, int iters);
void launch_intersect(cudaStream_t st, int sm_count, const SceneDev& s, int root_entry, const float* rays, long long n, int use_bvh,
    int* tri_id, float* t_out, unsigned* overflow);

// Host binned-SAH builder (rt_bvh_host.cpp).  nodes: 2 x float4 per node in the layout of SceneDev::nodes;
// perm: BVH-order slot -> global triangle id.  Returns the tree depth.
struct HostBvh {
    std::vector<float4> nodes;
    std::vector<int> perm;
    int root_entry = 0;
    int depth = 0;
};
HostBvh build_bvh_sah_host(const float* pos, long long n_tris, float pad);

// Device LBVH builder (rt_lbvh.cu): Morton codes -> radix sort -> Karras hierarchy -> leaf collapse -> refit.
// d_pos: device copy of the triangle soup.  Outputs are device buffers allocated by the callee (cudaMalloc).
struct DeviceBvh {
    float4* nodes = nullptr;
    int* perm = nullptr;
    int n_nodes = 0;
    int root_entry = 0;
    int depth = 0;
};
int build_bvh_lbvh_device(cudaStream_t st, const float* d_pos, long long n_tris, float pad, DeviceBvh* out, const char** err);
// Device builder of SAH quality (rt_ploc.cu): Morton order -> parallel locally-ordered clustering (mutual nearest neighbours by
// merged surface area) -> depth-first leaf order -> leaf collapse.  Same outputs.  Returns 2 for scenes of at most 4 triangles
// (use the LBVH path, which emits the single leaf).
// scratch: the caller's reusable build arena (grown when too small) and one pinned host word, so that a rebuild allocates nothing but its outputs.
struct BuildScratch {
    char* base = nullptr;
    size_t capacity = 0;
    unsigned long long* host_word = nullptr; // pinned
};
int build_bvh_ploc_device(cudaStream_t st, const float* d_pos, long long n_tris, float pad, DeviceBvh* out, const char** err, BuildScratch* scratch);

// Screen post-processing (rt_post.cu): float4 images in the Screen layout.  option / gauss follow rt_b200.h's RT_FILTER_* / RT_KERNEL_*.
void launch_post_light(cudaStream_t st, int sm_count, const float4* img, float4* light, size_t n);
void launch_post_blur(cudaStream_t st, const float4* src, float4* dst, int w, int h, int f, bool gauss, const float* weights);
void launch_post_combine(cudaStream_t st, int sm_count, float4* img, const float4* light, size_t n, int option, float exposure);
void launch_post_gamma(cudaStream_t st, int sm_count, float4* img, size_t n, float e);
void launch_post_rgba8(cudaStream_t st, int sm_count, const float4* img, unsigned char* out, size_t n);
void launch_unpack_rgb(cudaStream_t st, int sm_count, const float* in, float4* out, size_t n);

// 4-wide BVH collapsed from the binary one (rt_wide.cu).  nodes: 8 x float4 per node; allocated by the callee (cudaMalloc).
struct WideBvh {
    float4* nodes = nullptr;
    int n_nodes = 0;
    int root_entry = 0; // node index, or the leaf encoding when the whole scene is one leaf
    int depth = 0;
};
int collapse_bvh_wide_device(cudaStream_t st, const float4* d_nodes2, int n_nodes2, int root_entry2, WideBvh* out, const char** err);

// rt_wide8.cu: the 8-wide tree (16 float4 per node) collapsed on the device from the binary one, and rt_intersect through it with eight
// lanes per ray.
int collapse_bvh_wide8_device(cudaStream_t st, const float4* d_nodes2, int n_nodes2, float4** nodes8, int* n_nodes8, int* root_entry8, int* depth8, const char** err);
void launch_intersect_wide(cudaStream_t st, int sm_count, const SceneDev& s, const float4* wide, int root_entry, const float* rays, long long n, int* tri_id,
    float* t_out);

// Checked build (RT_CHECKED, rt_types.h): violations counted so far by the kernels of rt_kernels.cu / rt_wide8.cu on the current
// device are ADDED to out[kChkSites].  The default build counts nothing and leaves `out` alone.
void add_violations_kernels(unsigned int* out);
void add_violations_wide8(unsigned int* out);
// one deliberate violation each (site kChkTable / kChkWideStack), for rt_violations_selftest; nothing in the default build
void provoke_violation_kernels(cudaStream_t st);
void provoke_violation_wide8(cudaStream_t st);

// Visiting rank of every object (triangles 0..n_tris-1, then spheres) in the reference's own BVH (rt_reforder.cu): the tie key
// of the traversal kernels.  d_spheres: 3 x float4 per sphere ({centre, radius} first).  d_rank: n_tris + n_spheres ints.
int reference_visit_rank(cudaStream_t st, const float* d_pos, long long n_tris, const float4* d_spheres, int n_spheres, int* d_rank, const char** err);
// tri_v0[slot].w = rank[global id] for ranked triangles (global id < n_ranked), the id itself otherwise
void launch_apply_tie_keys(cudaStream_t st, float4* v0, const float4* v2, const int* rank, long long n_slots, long long n_ranked);

} // namespace rtb
