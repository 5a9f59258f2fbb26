#include "render.h"
#include "rt_b200.h"
#include <stdexcept>
#include <string>

int max_reflection_level = 5;
int sphere_light_ray_count = 10;
int plane_light_1D_ray_count = 3;
int glossy_ray_count = 1;
float refraction_factor = 0.8f;
bool useBVH = false;
TextureFiltering textureFiltering { TextureFiltering::NearestNeighbor };
OutOfBoundsRule outOfBoundsRuleX { OutOfBoundsRule::Border };
OutOfBoundsRule outOfBoundsRuleY { OutOfBoundsRule::Border };
glm::vec3 textureBorderColor(0);
bool useTextures = false;
RenderTimings lastRenderTimings;

namespace {
void check(int rc, const char* what)
{
    if (rc != RT_OK)
        throw std::runtime_error(std::string(what) + ": " + rt_last_error());
}
}

void renderRayTracing(Scene& scene, const Trackball& camera, const BoundingVolumeHierarchy& bvh, Screen& screen,
    bool textureDebugging, bool anti_aliasing, bool multipleRays, int sampleSize)
{
    rt_ctx* ctx = bvh.context();

    // lights and materials are read live from the Scene every frame (src/shadow.cpp:111,141; ray_tracing.h:23-27)
    std::vector<rt_material> mats;
    for (const Mesh& mesh : scene.meshes) {
        const Material& m = mesh.material;
        mats.push_back(rt_material { { m.kd.x, m.kd.y, m.kd.z }, { m.ks.x, m.ks.y, m.ks.z }, m.shininess, m.transparency });
    }
    if (!mats.empty())
        check(rt_set_materials(ctx, mats.data(), (int)mats.size()), "rt_set_materials");
    std::vector<rt_point_light> pl;
    for (const PointLight& l : scene.pointLights)
        pl.push_back(rt_point_light { { l.position.x, l.position.y, l.position.z }, { l.color.x, l.color.y, l.color.z } });
    std::vector<rt_sphere_light> sl;
    for (const SphericalLight& l : scene.sphericalLight)
        sl.push_back(rt_sphere_light { { l.position.x, l.position.y, l.position.z }, l.radius, { l.color.x, l.color.y, l.color.z } });
    check(rt_set_lights(ctx, pl.data(), (int)pl.size(), sl.data(), (int)sl.size()), "rt_set_lights");
    std::vector<rt_spot_light> spots;
    for (const SpotLight& l : scene.spotLight)
        spots.push_back(rt_spot_light { { l.position.x, l.position.y, l.position.z }, { l.direction.x, l.direction.y, l.direction.z }, l.angle, { l.color.x, l.color.y, l.color.z } });
    check(rt_set_spot_lights(ctx, spots.data(), (int)spots.size()), "rt_set_spot_lights");
    std::vector<rt_plane_light> planes;
    for (const PlaneLight& l : scene.planeLight)
        planes.push_back(rt_plane_light { { l.position.x, l.position.y, l.position.z }, { l.width.x, l.width.y, l.width.z }, { l.height.x, l.height.y, l.height.z }, { l.color.x, l.color.y, l.color.z } });
    check(rt_set_plane_lights(ctx, planes.data(), (int)planes.size()), "rt_set_plane_lights");
    std::vector<rt_sphere> sp;
    for (const Sphere& s : scene.spheres) {
        const Material& m = s.material;
        sp.push_back(rt_sphere { { s.center.x, s.center.y, s.center.z }, s.radius,
            rt_material { { m.kd.x, m.kd.y, m.kd.z }, { m.ks.x, m.ks.y, m.ks.z }, m.shininess, m.transparency } });
    }
    check(rt_set_spheres(ctx, sp.data(), (int)sp.size()), "rt_set_spheres");

    const glm::ivec2 res = screen.resolution();
    const glm::vec3 la = camera.lookAt(), eu = camera.rotationEulerAngles();
    rt_camera cam { { la.x, la.y, la.z }, { eu.x, eu.y, eu.z }, camera.distanceFromLookAt(), camera.fovy() };
    rt_params prm {};
    prm.width = res.x;
    prm.height = res.y;
    prm.max_reflection_level = max_reflection_level;
    prm.sphere_light_ray_count = sphere_light_ray_count;
    prm.glossy_ray_count = glossy_ray_count;
    prm.refraction_factor = refraction_factor;
    prm.sample_mode = anti_aliasing ? 1 : (multipleRays ? 2 : 0);
    prm.sample_size = sampleSize;
    prm.use_bvh = useBVH ? 1 : 0; // only decides exact-t ties (rt_b200.h); the device BVH is always used
    prm.exhaustive = 0;
    prm.plane_light_ray_count_1d = plane_light_1D_ray_count;
    prm.texture_debug = textureDebugging ? 1 : 0; // src/main.cpp:355-356: one corner ray per pixel, its texel, nothing else
    rt_stats st {};
    static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "Screen pixels must be packed float3");
    // the texture branch of getFinalColor (src/main.cpp:155-171) with the knobs main.cpp sets on the Image before sampling it
    // (the texture-debug view applies the same knobs whether useTextures is set or not, src/main.cpp:81-86)
    if (useTextures || textureDebugging) {
        rt_texture_params tp { static_cast<int>(textureFiltering), static_cast<int>(outOfBoundsRuleX), static_cast<int>(outOfBoundsRuleY),
            { textureBorderColor.x, textureBorderColor.y, textureBorderColor.z } };
        check(rt_set_texturing(ctx, &tp), "rt_set_texturing");
    } else {
        check(rt_set_texturing(ctx, nullptr), "rt_set_texturing");
    }
    // screen.postprocessImage() (src/main.cpp:397-398) runs on the device at the end of the frame, before the rows come home
    check(rt_set_postprocess(ctx, &screen.postSettings()), "rt_set_postprocess");
    check(rt_render(ctx, &cam, &prm, &screen.pixels()[0].x, nullptr, nullptr, &st), "rt_render");
    lastRenderTimings.gpu_ms = st.gpu_ms;
    lastRenderTimings.rays = st.primary_rays + st.shadow_queries + st.secondary_rays;
}
