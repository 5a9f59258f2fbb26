// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Harness around the reference's own translation units.  src/ray_tracing.cpp,
// src/bounding_volume_hierarchy.cpp and src/shadow.cpp are compiled VERBATIM from /root/reference
// (see oracle/Makefile; nothing from them is copied here).  What cannot be compiled verbatim
// because it shares a translation unit with GL / GLFW / ImGui code is restated below, keeping the
// float operation order of the cited lines:
//   Trackball::position / generateRay      framework/src/trackball.cpp:65-68, 87-98
//   calcColor                              src/main.cpp:112-121
//   getFinalColor (glossy_ray_count == 1)  src/main.cpp:129-301
//   getPixelRays                           src/main.cpp:309-335
//   renderRayTracing                       src/main.cpp:340-400
//   Screen::setPixel                       src/screen.cpp:32-38
// Diffuse textures go through the reference's own Image class (src/image.cpp, also compiled verbatim), with all five
// filters; the level of detail of the mip-mapped ones comes from the reference's own ray differentials
// (src/ray_differentials.cpp, compiled verbatim too: tranfer_and_reflect_ray_differentials, transfer_ray_differentials,
// computeLevelOfDetails, called where main.cpp:83,137,168 call them).  One thing had to be DEFINED: Ray's default member
// initialisers of dD_dx / dD_dy read the members `right` and `up`, which are declared — hence constructed — after them
// (framework/include/ray.h:19-28): undefined behaviour, whatever the stack held.  defineDifferentials() below gives every
// ray the values that code evidently means: the declared right = (1,0,0), up = (0,-1,0), evaluated with the direction the
// ray had when its defaults ran — (0,0,-1) for a default-constructed ray that generateRay fills in afterwards
// (trackball.cpp:92-95), the given direction for the aggregate-initialised reflection / refraction rays (main.cpp:199,286,288).
// Child rays do not inherit differentials in the reference (they are fresh objects), so neither do they here.
// Builds oracle/_ref/libref_oracle.so, kind "reference".
#include "bounding_volume_hierarchy.h"
#include "ray_differentials.h"
#include "ray_tracing.h"
#include "shadow.h"

#include "oracle_api.h"

#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// The defined initial state of a ray's differentials (see the header comment): ray.h:19-28 with right / up as declared.
static void defineDifferentials(Ray& ray, const glm::vec3& directionWhenConstructed)
{
    ray.right = glm::vec3(1, 0, 0);
    ray.up = glm::vec3(0, -1, 0);
    const glm::vec3 direction = directionWhenConstructed;
    ray.dD_dx = (glm::dot(direction, direction) * ray.right - glm::dot(direction, ray.right) * direction) / glm::pow(glm::dot(direction, direction), 1.5f);
    ray.dD_dy = (glm::dot(direction, direction) * ray.up - glm::dot(direction, ray.up) * direction) / glm::pow(glm::dot(direction, direction), 1.5f);
    ray.dP_dx = glm::vec3(0);
    ray.dP_dy = glm::vec3(0);
}
static Ray childRay(const glm::vec3& origin, const glm::vec3& direction) // `Ray r = { origin, direction };`
{
    Ray ray = { origin, direction };
    defineDifferentials(ray, direction);
    return ray;
}

// counters bumped by the drawRay stub (oracle/ref_stubs.cpp): every iteration of the reference's
// cansee loop ends in exactly one drawRay call (src/shadow.cpp:45,49,62).
extern thread_local unsigned long long orc_drawray_calls;

namespace {

} // namespace
struct OrcRegisteredTexture { // shared with the stbi_load stand-in (ref_stubs.cpp)
    int w, h;
    std::vector<unsigned char> rgb;
};
extern std::vector<OrcRegisteredTexture> orc_registered_textures;
namespace {

std::vector<float> g_tex_uv;        // oracle_set_textures: 6 floats per triangle
std::vector<int> g_mesh_tex;
std::vector<std::string> g_tex_files; // placeholder files whose names carry the registry index
bool g_use_textures = false;        // useTextures, src/main.cpp:58
TextureFiltering g_tex_filtering = TextureFiltering::NearestNeighbor; // main.cpp:54
OutOfBoundsRule g_oob_x = OutOfBoundsRule::Border, g_oob_y = OutOfBoundsRule::Border; // main.cpp:55-56
glm::vec3 g_tex_border(0);          // main.cpp:57

std::vector<float> g_spheres;       // oracle_set_spheres
std::vector<float> g_spot, g_plane; // oracle_set_extra_lights
int g_plane_rays = 3;               // plane_light_1D_ray_count, src/main.cpp:125

struct RenderGlobals { // the file-scope knobs of src/main.cpp:58-60,123-127
    bool useBVH;
    int max_reflection_level;
    int sphere_light_ray_count;
    int glossy_ray_count;
    float refraction_factor;
    int width, height;
};

struct HeadlessTrackball { // framework/include/trackball.h:44-52 without the Window
    glm::vec3 m_lookAt;
    glm::vec3 m_rotationEulerAngles;
    float m_distanceFromLookAt;
    float m_fovy;
    float m_aspect; // Window::aspectRatio(), framework/src/window.cpp:334-337

    glm::vec3 position() const
    {
        return m_lookAt + glm::quat(m_rotationEulerAngles) * glm::vec3(0, 0, -m_distanceFromLookAt);
    }
    Ray generateRay(const glm::vec2& pixel) const
    {
        const float halfScreenPlaceHeight = std::tan(m_fovy / 2.0f);
        const float halfScreenPlaceWidth = m_aspect * halfScreenPlaceHeight;
        const glm::vec3 cameraSpaceDirection = glm::normalize(
            glm::vec3(-pixel.x * halfScreenPlaceWidth, pixel.y * halfScreenPlaceHeight, 1.0f));
        Ray ray;
        defineDifferentials(ray, glm::vec3(0.0f, 0.0f, -1.0f)); // the defaults ran on the default direction
        ray.origin = position();
        ray.direction = glm::quat(m_rotationEulerAngles) * cameraSpaceDirection;
        ray.t = std::numeric_limits<float>::max();
        return ray;
    }
};

struct Counters {
    unsigned long long primary = 0, secondary = 0;
};

glm::vec3 calcColor(const Lighting& light, const Material& material) // main.cpp:112-121
{
    glm::vec3 diffuse = material.kd * light.color * light.intensity * light.cosLightSurfaceAngle;
    glm::vec3 spec = glm::vec3(0);
    if (material.shininess > 0)
        spec = light.color * material.ks * std::pow(light.cosLightSpecAngle, material.shininess);
    return diffuse + spec;
}

glm::vec3 getFinalColorNoRayTracingJustTextures(const RenderGlobals& g, Scene& scene, const BoundingVolumeHierarchy& bvh, Ray ray) // main.cpp:75-106
{
    HitInfo hitInfo;
    if (bvh.intersect(ray, hitInfo, g.useBVH)) {
        transfer_ray_differentials(ray, hitInfo.normal); // main.cpp:83
        Material& mat = hitInfo.getMaterial(scene);
        if (mat.kdTexture) {
            Image& texture = mat.kdTexture.value();
            texture.setBorderColor(g_tex_border);
            texture.setOutOfBoundsRuleX(g_oob_x);
            texture.setOutOfBoundsRuleY(g_oob_y);
            texture.setTextureFilteringMethod(g_tex_filtering);
            const float lod = computeLevelOfDetails(ray, hitInfo); // main.cpp:99
            return texture.getPixel(hitInfo.texCoord, lod);
        }
        return glm::vec3(1);
    }
    return glm::vec3(0);
}

glm::vec3 getFinalColor(const RenderGlobals& g, Scene& scene, const BoundingVolumeHierarchy& bvh, Ray ray, int level,
    Counters& cnt) // main.cpp:129-301
{
    HitInfo hitInfo;
    if (!bvh.intersect(ray, hitInfo, g.useBVH))
        return glm::vec3(0.0f);

    tranfer_and_reflect_ray_differentials(ray, hitInfo); // main.cpp:137

    glm::vec3 color(0);
    glm::vec3 reflect = glm::reflect(glm::normalize(ray.direction), glm::normalize(hitInfo.normal));

    Material& originalMaterial = hitInfo.getMaterial(scene);
    Material matForRendering;
    matForRendering.kd = originalMaterial.kd;
    matForRendering.ks = originalMaterial.ks;
    matForRendering.shininess = originalMaterial.shininess;
    matForRendering.transparency = originalMaterial.transparency;

    if (g_use_textures && hitInfo.is_triangle && originalMaterial.kdTexture) { // main.cpp:155-171
        Image& texture = originalMaterial.kdTexture.value();
        texture.setBorderColor(g_tex_border);
        texture.setOutOfBoundsRuleX(g_oob_x);
        texture.setOutOfBoundsRuleY(g_oob_y);
        texture.setTextureFilteringMethod(g_tex_filtering);
        const float lod = computeLevelOfDetails(ray, hitInfo); // main.cpp:168
        matForRendering.kd = texture.getPixel(hitInfo.texCoord, lod);
    }

    for (const Lighting& light : getPointLights(hitInfo, reflect, scene, bvh))
        color += calcColor(light, matForRendering);
    for (const Lighting& light : getSpherelights(hitInfo, reflect, scene, bvh, g.sphere_light_ray_count))
        color += calcColor(light, matForRendering);
    for (const Lighting& light : getSpotLichts(hitInfo, reflect, scene, bvh))
        color += calcColor(light, matForRendering);
    for (const Lighting& light : getPlaneLights(hitInfo, reflect, scene, bvh, g_plane_rays))
        color += calcColor(light, matForRendering);

    if (level >= g.max_reflection_level)
        return color;

    if (matForRendering.transparency == 1.0f) {
        if (matForRendering.ks.x > 0 || matForRendering.ks.y > 0 || matForRendering.ks.z > 0) {
            glm::vec3 reflectColor = glm::vec3(0);
            Ray refRay = childRay(hitInfo.hitPoint + 0.01f * reflect, reflect);
            cnt.secondary++;
            reflectColor += matForRendering.ks * getFinalColor(g, scene, bvh, refRay, level + 1, cnt);
            if (matForRendering.shininess != 0) {
                // glossy loop `for (i = 1; i < glossy_ray_count; ...)` is empty for count == 1
                color += matForRendering.ks * reflectColor / (float)g.glossy_ray_count;
            } else {
                color += matForRendering.ks * reflectColor;
            }
        }
    } else {
        glm::vec3 l = glm::normalize(ray.direction);
        glm::vec3 n = glm::normalize(hitInfo.normal);
        float r = g.refraction_factor;
        float c = std::abs(glm::dot(l, n));
        glm::vec3 refract = r * l + (r * c - std::sqrt(1 - r * r * (1 - c * c))) * n;
        refract = glm::normalize(refract);
        float& R0 = matForRendering.transparency;
        float reflectionChance = R0 + (1 - R0) * (std::pow(1 - c, 5));
        float refractionChance = 1 - reflectionChance;
        cnt.secondary++;
        color += reflectionChance * getFinalColor(g, scene, bvh, childRay(hitInfo.hitPoint + 0.01f * reflect, reflect), level + 1, cnt);
        if (r * r * (1 - c * c) <= 1.0f) {
            cnt.secondary++;
            color += refractionChance * getFinalColor(g, scene, bvh, childRay(hitInfo.hitPoint + 0.01f * refract, refract), level + 1, cnt);
        }
    }
    return color;
}

std::vector<glm::vec2> getPixelRays(const RenderGlobals& g, glm::vec2 pixelCenter, int sampleSize) // main.cpp:309-335
{
    float offsetX = (1.0f / g.width) * (1.0f / (glm::sqrt(sampleSize) * 2));
    float offsetY = (1.0f / g.height) * (1.0f / (glm::sqrt(sampleSize) * 2));
    std::vector<glm::vec2> origins;
    const std::array<glm::vec2, 4> quadrantSigns = { glm::vec2(-1.0f, 1.0f), glm::vec2(1.0f, 1.0f),
        glm::vec2(-1.0f, -1.0f), glm::vec2(1.0f, -1.0f) };
    int moves = glm::sqrt(sampleSize) - 1;
    for (int i = 0; i < 4; i++)
        for (int x = 1; x <= moves; x = x + 2)
            for (int y = 1; y <= moves; y = y + 2)
                origins.push_back(glm::vec2(pixelCenter.x + (offsetX * quadrantSigns[i].x * x),
                    pixelCenter.y + (offsetY * quadrantSigns[i].y * y)));
    return origins;
}



void addSpheres(Scene& scene)
{
    for (size_t k = 0; k + 11 < g_spheres.size(); k += 12) {
        const float* f = &g_spheres[k];
        Material m;
        m.kd = glm::vec3(f[4], f[5], f[6]);
        m.ks = glm::vec3(f[7], f[8], f[9]);
        m.shininess = f[10];
        m.transparency = f[11];
        Sphere sp;
        sp.center = glm::vec3(f[0], f[1], f[2]);
        sp.radius = f[3];
        sp.material = m;
        scene.spheres.push_back(sp);
    }
}

Scene sceneFromSoup(const float* pos, const float* nrm, const int* mesh_id, int n_tris, const orc_material* mats, int n_mats)
{
    Scene scene;
    int n_meshes = n_mats > 0 ? n_mats : 1;
    for (int i = 0; mesh_id && i < n_tris; i++) // (oracle_closest_hit passes no materials: one mesh per id that occurs)
        n_meshes = std::max(n_meshes, mesh_id[i] + 1);
    scene.meshes.resize(n_meshes);
    for (int m = 0; m < (int)scene.meshes.size(); m++) {
        Material& mat = scene.meshes[m].material;
        if (m < n_mats && mats) {
            mat.kd = glm::vec3(mats[m].kd[0], mats[m].kd[1], mats[m].kd[2]);
            mat.ks = glm::vec3(mats[m].ks[0], mats[m].ks[1], mats[m].ks[2]);
            mat.shininess = mats[m].shininess;
            mat.transparency = mats[m].transparency;
        } else {
            mat.kd = glm::vec3(1.0f);
        }
    }
    for (int m = 0; m < (int)scene.meshes.size() && m < (int)g_mesh_tex.size(); m++)
        if (g_mesh_tex[m] >= 0 && g_mesh_tex[m] < (int)g_tex_files.size())
            scene.meshes[m].material.kdTexture = Image(g_tex_files[g_mesh_tex[m]]); // Image::Image -> stbi_load stand-in
    for (int i = 0; i < n_tris; i++) {
        Mesh& mesh = scene.meshes[mesh_id ? mesh_id[i] : 0];
        const unsigned base = (unsigned)mesh.vertices.size();
        for (int k = 0; k < 3; k++) {
            Vertex v;
            v.p = glm::vec3(pos[9 * i + 3 * k], pos[9 * i + 3 * k + 1], pos[9 * i + 3 * k + 2]);
            v.n = nrm ? glm::vec3(nrm[9 * i + 3 * k], nrm[9 * i + 3 * k + 1], nrm[9 * i + 3 * k + 2]) : glm::vec3(0, 0, 1);
            v.texCoord = g_tex_uv.size() == 6 * (size_t)n_tris ? glm::vec2(g_tex_uv[6 * i + 2 * k], g_tex_uv[6 * i + 2 * k + 1]) : glm::vec2(0.0f);
            mesh.vertices.push_back(v);
        }
        mesh.triangles.emplace_back(base, base + 1, base + 2);
    }
    return scene;
}

// Exhaustive search in the reference's brute-force order (bounding_volume_hierarchy.cpp:51-72: all triangles, then the
// spheres) using the verbatim intersection functions; the first strictly smaller t wins, so ties keep the lowest index.
int exhaustiveId(const float* pos, int n_tris, Ray ray, float& t_out, const Scene* scene = nullptr)
{
    int best = -1;
    HitInfo hi;
    for (int i = 0; i < n_tris; i++) {
        const glm::vec3 v0(pos[9 * i], pos[9 * i + 1], pos[9 * i + 2]);
        const glm::vec3 v1(pos[9 * i + 3], pos[9 * i + 4], pos[9 * i + 5]);
        const glm::vec3 v2(pos[9 * i + 6], pos[9 * i + 7], pos[9 * i + 8]);
        if (intersectRayWithTriangle(v0, v1, v2, ray, hi, 0))
            best = i;
    }
    if (scene)
        for (size_t k = 0; k < scene->spheres.size(); k++)
            if (intersectRayWithShape(scene->spheres[k], ray, hi))
                best = n_tris + (int)k;
    t_out = ray.t;
    return best;
}

} // namespace

extern "C" const char* oracle_kind(void) { return "reference"; }

extern "C" void oracle_set_extra_lights(const float* spot, int n_spot, const float* plane, int n_plane, int plane_ray_count_1d)
{
    g_spot.assign(spot, spot + (spot ? 10 * (size_t)n_spot : 0));
    g_plane.assign(plane, plane + (plane ? 12 * (size_t)n_plane : 0));
    g_plane_rays = plane_ray_count_1d;
}

extern "C" void oracle_set_spheres(const float* spheres, int n)
{
    g_spheres.assign(spheres, spheres + (spheres ? 12 * (size_t)n : 0));
}

extern "C" void oracle_set_textures(const float* tri_uv, int n_tris, const orc_texture* textures, int n_textures, const int* mesh_tex, int n_meshes,
    int use_textures, int filtering, int oob_x, int oob_y, const float* border_rgb)
{
    g_use_textures = use_textures != 0;
    g_tex_uv.assign(tri_uv, tri_uv + (tri_uv ? 6 * (size_t)n_tris : 0));
    g_mesh_tex.assign(mesh_tex, mesh_tex + (mesh_tex ? n_meshes : 0));
    orc_registered_textures.clear();
    g_tex_files.clear();
    for (int k = 0; textures && k < n_textures; k++) {
        orc_registered_textures.push_back({ textures[k].width, textures[k].height,
            std::vector<unsigned char>(textures[k].rgb, textures[k].rgb + 3 * (size_t)textures[k].width * textures[k].height) });
        // Image::Image insists on an existing file (src/image.cpp:38-41): an empty placeholder whose name carries the index
        const std::string name = (std::filesystem::temp_directory_path() / ("orc_texture_" + std::to_string(k))).string();
        std::FILE* f = std::fopen(name.c_str(), "wb");
        if (f)
            std::fclose(f);
        g_tex_files.push_back(name);
    }
    g_tex_filtering = (TextureFiltering)filtering; // 2..4: the mip-mapped filters, sampled at lod 0 (see oracle_api.h)
    g_oob_x = (OutOfBoundsRule)oob_x;
    g_oob_y = (OutOfBoundsRule)oob_y;
    g_tex_border = border_rgb ? glm::vec3(border_rgb[0], border_rgb[1], border_rgb[2]) : glm::vec3(0);
}

extern "C" int oracle_render(const float* pos, const float* nrm, const int* mesh_id, int n_tris,
    const orc_material* mats, int n_mats,
    const float* point_lights, int n_point, const float* sphere_lights, int n_sphere,
    const orc_camera* cam, const orc_params* prm,
    float* rgb, int* tri_id, float* t_hit, orc_stats* stats)
{
    if ((!pos && n_tris > 0) || !cam || !prm || prm->width <= 0 || prm->height <= 0 || prm->glossy_ray_count != 1)
        return 1;
    Scene scene = sceneFromSoup(pos, nrm, mesh_id, n_tris, mats, n_mats);
    addSpheres(scene);
    for (int i = 0; i < n_point; i++)
        scene.pointLights.push_back(PointLight { glm::vec3(point_lights[6 * i], point_lights[6 * i + 1], point_lights[6 * i + 2]),
            glm::vec3(point_lights[6 * i + 3], point_lights[6 * i + 4], point_lights[6 * i + 5]) });
    for (int i = 0; i < n_sphere; i++)
        scene.sphericalLight.push_back(SphericalLight { glm::vec3(sphere_lights[7 * i], sphere_lights[7 * i + 1], sphere_lights[7 * i + 2]),
            sphere_lights[7 * i + 3], glm::vec3(sphere_lights[7 * i + 4], sphere_lights[7 * i + 5], sphere_lights[7 * i + 6]) });
    for (size_t k = 0; k + 9 < g_spot.size(); k += 10) {
        const float* f = &g_spot[k];
        scene.spotLight.push_back(SpotLight { glm::vec3(f[0], f[1], f[2]), glm::vec3(f[3], f[4], f[5]), f[6], glm::vec3(f[7], f[8], f[9]) });
    }
    for (size_t k = 0; k + 11 < g_plane.size(); k += 12) {
        const float* f = &g_plane[k];
        scene.planeLight.push_back(PlaneLight { glm::vec3(f[0], f[1], f[2]), glm::vec3(f[3], f[4], f[5]), glm::vec3(f[6], f[7], f[8]), glm::vec3(f[9], f[10], f[11]) });
    }
    BoundingVolumeHierarchy bvh(&scene);

    RenderGlobals g;
    g.useBVH = prm->use_bvh != 0;
    g.max_reflection_level = prm->max_reflection_level;
    g.sphere_light_ray_count = prm->sphere_light_ray_count;
    g.glossy_ray_count = prm->glossy_ray_count;
    g.refraction_factor = prm->refraction_factor;
    g.width = prm->width;
    g.height = prm->height;

    HeadlessTrackball camera;
    camera.m_lookAt = glm::vec3(cam->look_at[0], cam->look_at[1], cam->look_at[2]);
    camera.m_rotationEulerAngles = glm::vec3(cam->euler[0], cam->euler[1], cam->euler[2]);
    camera.m_distanceFromLookAt = cam->dist;
    camera.m_fovy = cam->fovy;
    camera.m_aspect = float(prm->width) / float(prm->height);

    const int W = prm->width, H = prm->height;
    const int xs = prm->x_step > 0 ? prm->x_step : 1, ys = prm->y_step > 0 ? prm->y_step : 1;
    const int nyj = (H - prm->y0 + ys - 1) / ys;
#ifdef _OPENMP
    const int nthreads = prm->num_threads > 0 ? prm->num_threads : omp_get_max_threads();
#else
    const int nthreads = 1;
#endif
    unsigned long long n_primary = 0, n_secondary = 0, n_shadow = 0;
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : n_primary, n_secondary, n_shadow)
    for (int j = 0; j < nyj; j++) {
        const int y = prm->y0 + j * ys;
        Counters cnt;
        const unsigned long long draw0 = orc_drawray_calls;
        for (int x = prm->x0; x < W; x += xs) {
            // main.cpp:350-353
            const glm::vec2 normalizedPixelPos { float(x) / W * 2.0f - 1.0f, float(y) / H * 2.0f - 1.0f };
            const Ray cameraRay = camera.generateRay(normalizedPixelPos);
            glm::vec3 out(0);
            Ray firstRay = cameraRay; // the ray whose closest hit is reported in tri_id / t_hit: first sample of the pixel
            if (prm->texture_debug) { // main.cpp:355-356
                cnt.primary++;
                out = getFinalColorNoRayTracingJustTextures(g, scene, bvh, cameraRay);
            } else if (prm->sample_mode == 1) { // main.cpp:358-375
                float offsetX = 1.0f / W * 0.25f;
                float offsetY = 1.0f / H * 0.25f;
                std::array<glm::vec2, 4> offsets;
                offsets[0] = glm::vec2(normalizedPixelPos.x - offsetX, normalizedPixelPos.y + offsetY);
                offsets[1] = glm::vec2(normalizedPixelPos.x + offsetX, normalizedPixelPos.y + offsetY);
                offsets[2] = glm::vec2(normalizedPixelPos.x - offsetX, normalizedPixelPos.y - offsetY);
                offsets[3] = glm::vec2(normalizedPixelPos.x + offsetX, normalizedPixelPos.y - offsetY);
                glm::vec3 avgColor(0);
                firstRay = camera.generateRay(offsets[0]);
                for (int i = 0; i < 4; i++) {
                    Ray ray = camera.generateRay(offsets[i]);
                    cnt.primary++;
                    avgColor += getFinalColor(g, scene, bvh, ray, 0, cnt);
                }
                avgColor *= 0.25;
                out = avgColor;
            } else if (prm->sample_mode == 2) { // main.cpp:377-385
                std::vector<glm::vec2> rayOrigins = getPixelRays(g, normalizedPixelPos, prm->sample_size);
                glm::vec3 avgColor(0);
                if (!rayOrigins.empty())
                    firstRay = camera.generateRay(rayOrigins[0]);
                for (auto& rayOrigin : rayOrigins) {
                    Ray ray = camera.generateRay(rayOrigin);
                    cnt.primary++;
                    avgColor += getFinalColor(g, scene, bvh, ray, 0, cnt);
                }
                avgColor = avgColor * (float)(1.0f / prm->sample_size);
                out = avgColor;
            } else {
                cnt.primary++;
                out = getFinalColor(g, scene, bvh, cameraRay, 0, cnt);
            }
            const size_t i = (size_t)(H - 1 - y) * W + x; // Screen::setPixel, screen.cpp:36
            if (rgb) {
                rgb[3 * i] = out.x;
                rgb[3 * i + 1] = out.y;
                rgb[3 * i + 2] = out.z;
            }
            if (tri_id || t_hit) {
                float t;
                const int id = exhaustiveId(pos, n_tris, firstRay, t, &scene);
                if (tri_id)
                    tri_id[i] = id;
                if (t_hit)
                    t_hit[i] = t;
            }
        }
        n_primary += cnt.primary;
        n_secondary += cnt.secondary;
        n_shadow += orc_drawray_calls - draw0;
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->primary_rays = n_primary;
        stats->secondary_rays = n_secondary;
        stats->shadow_queries = n_shadow;
        stats->seconds = std::chrono::duration<double>(t1 - t0).count();
        stats->threads = nthreads;
    }
    return 0;
}

extern "C" int oracle_closest_hit(const float* pos, const float* nrm, const int* mesh_id, int n_tris,
    const float* rays, int n_rays, int use_bvh, int* tri_id, float* t_hit)
{
    if ((!pos && n_tris > 0) || !rays)
        return 1;
    Scene scene = sceneFromSoup(pos, nrm, mesh_id, n_tris, nullptr, 0);
    addSpheres(scene);
    BoundingVolumeHierarchy bvh(&scene);
#pragma omp parallel for schedule(dynamic, 64)
    for (int r = 0; r < n_rays; r++) {
        Ray ray;
        ray.origin = glm::vec3(rays[6 * r], rays[6 * r + 1], rays[6 * r + 2]);
        ray.direction = glm::vec3(rays[6 * r + 3], rays[6 * r + 4], rays[6 * r + 5]);
        ray.t = std::numeric_limits<float>::max();
        float t;
        int id = exhaustiveId(pos, n_tris, ray, t, &scene);
        if (use_bvh) {
            Ray rb = ray;
            HitInfo hi;
            const bool hit = bvh.intersect(rb, hi, true);
            // report the BVH traversal's t; the id is the lowest-index triangle reproducing that t
            if (!hit) {
                id = -1;
                t = std::numeric_limits<float>::max();
            } else if (rb.t != t) {
                t = rb.t;
                id = -2; // BVH culled the exhaustive winner (AABB slab rounding); id unknown
                HitInfo h2;
                for (int i = 0; i < n_tris; i++) {
                    Ray rr = ray;
                    const glm::vec3 v0(pos[9 * i], pos[9 * i + 1], pos[9 * i + 2]);
                    const glm::vec3 v1(pos[9 * i + 3], pos[9 * i + 4], pos[9 * i + 5]);
                    const glm::vec3 v2(pos[9 * i + 6], pos[9 * i + 7], pos[9 * i + 8]);
                    if (intersectRayWithTriangle(v0, v1, v2, rr, h2, 0) && rr.t == rb.t) {
                        id = i;
                        break;
                    }
                }
            }
        }
        if (tri_id)
            tri_id[r] = id;
        if (t_hit)
            t_hit[r] = t;
    }
    return 0;
}
