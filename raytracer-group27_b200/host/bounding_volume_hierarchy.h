// BoundingVolumeHierarchy — same public face as the reference class (src/bounding_volume_hierarchy.h:22-33).
// Construction copies every triangle of the Scene to the GPU (the reference copies them into its own vectors,
// bounding_volume_hierarchy.cpp:80-99) and builds the device BVH; intersect() answers single-ray queries
// through the same device traversal the renderer uses.  All GPU work goes through the C ABI in
// include/rt_b200.h.
#pragma once
#include "ray_tracing.h"
#include "scene.h"

struct rt_ctx;

class BoundingVolumeHierarchy {
public:
    // bvhMode: RT_BVH_AUTO (default), RT_BVH_LBVH_DEVICE or RT_BVH_SAH_HOST; device: CUDA ordinal
    explicit BoundingVolumeHierarchy(Scene* pScene, int bvhMode = 2, int device = 0);
    ~BoundingVolumeHierarchy();
    BoundingVolumeHierarchy(const BoundingVolumeHierarchy&) = delete;
    BoundingVolumeHierarchy& operator=(const BoundingVolumeHierarchy&) = delete;

    void debugDraw(int /*level*/, bool /*showLeafNodes*/) {} // GL visual debugger: not applicable headless
    int numLevels() const { return m_numLevels; }

    // Closest hit closer than ray.t with t >= 0; updates ray.t and hitInfo.  useBVH=false loops over every
    // triangle in scene order (the reference's brute-force path); both give the same answer.
    bool intersect(Ray& ray, HitInfo& hitInfo, bool useBVH) const;

    rt_ctx* context() const { return m_ctx; }
    Scene* scene() const { return m_pScene; }

private:
    Scene* m_pScene;
    rt_ctx* m_ctx = nullptr;
    int m_numLevels = 0;
    std::vector<float> m_pos, m_nrm; // 9 floats per triangle, scene order (kept for HitInfo reconstruction)
    std::vector<int> m_meshId;
};
