// loadScene — the reference's presets (src/scene.cpp:4-150) as a table: which OBJ, whether it is centred and
// scaled, and where the lights sit.  Spot / plane lights and sphere primitives are stored in the Scene for API
// fidelity; the device path traces triangles with point and spherical lights.
#include "scene.h"

namespace {

struct Preset {
    const char* obj;   // nullptr: no mesh
    bool normalize;
    bool whiteKd;      // subMeshes[0].material.kd = 1 (scene.cpp:13,102)
};

void addMeshes(Scene& scene, const std::filesystem::path& dataDir, const Preset& p)
{
    if (!p.obj)
        return;
    std::vector<Mesh> sub = loadMesh(dataDir / p.obj, p.normalize);
    if (p.whiteKd && !sub.empty())
        sub[0].material.kd = glm::vec3(1.0f);
    for (Mesh& m : sub)
        scene.meshes.push_back(std::move(m));
}

Material sphereMaterial(glm::vec3 kd, glm::vec3 ks = glm::vec3(0.0f), float shininess = 1.0f, float transparency = 1.0f)
{
    Material m;
    m.kd = kd;
    m.ks = ks;
    m.shininess = shininess;
    m.transparency = transparency;
    return m;
}

} // namespace

Scene loadScene(SceneType type, const std::filesystem::path& dataDir)
{
    Scene scene;
    const glm::vec3 white(1.0f);
    const PointLight keyLight { glm::vec3(-1, 1, -1), white };
    switch (type) {
    case SingleTriangle:
        addMeshes(scene, dataDir, { "tr_def.obj", false, true });
        scene.pointLights.push_back(keyLight);
        scene.sphericalLight.push_back(SphericalLight { glm::vec3(-2.1f, 1.24f, -0.51f), 0.5f, glm::vec3(1.0f, 0.0f, 1.0f) });
        break;
    case Cube:
        addMeshes(scene, dataDir, { "cube.obj", false, false });
        scene.pointLights.push_back(keyLight);
        scene.spotLight.push_back(SpotLight { glm::vec3(-1.2f, -1.0f, -1.0f), glm::vec3(1.0f, 1.2f, 1.0f), 10.0f, white });
        break;
    case CornellBox:
        addMeshes(scene, dataDir, { "CornellBox-Mirror-Rotated.obj", true, false });
        scene.spheres.push_back(Sphere { glm::vec3(-0.2f, 0.15f, -0.25f), 0.2f, sphereMaterial(glm::vec3(0.0f), glm::vec3(0.0f), 1.0f, 0.0f) });
        scene.pointLights.push_back(PointLight { glm::vec3(0.0f, 0.58f, 0.0f), white });
        break;
    case CornellBoxSphericalLight:
        addMeshes(scene, dataDir, { "CornellBox-Mirror-Rotated.obj", true, false });
        scene.spheres.push_back(Sphere { glm::vec3(-0.2f, 0.15f, -0.25f), 0.2f, sphereMaterial(glm::vec3(0.0f), glm::vec3(0.0f), 1.0f, 0.0f) });
        scene.sphericalLight.push_back(SphericalLight { glm::vec3(0.0f, 0.45f, 0.0f), 0.1f, white });
        break;
    case CornellBoxPlaneLight:
        addMeshes(scene, dataDir, { "CornellBox-Mirror-Rotated.obj", true, false });
        scene.planeLight.push_back(PlaneLight { glm::vec3(-0.1f, 0.63f, -0.1f), glm::vec3(0.15f, -0.05f, 0.0f), glm::vec3(0.0f, 0.0f, 0.2f), white });
        break;
    case Monkey:
        addMeshes(scene, dataDir, { "monkey-rotated.obj", true, false });
        scene.pointLights.push_back(keyLight);
        scene.pointLights.push_back(PointLight { glm::vec3(1, -1, -1), white });
        break;
    case Teapot:
        addMeshes(scene, dataDir, { "teapot.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    case Dragon:
        addMeshes(scene, dataDir, { "dragon.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    case Spheres:
        scene.spheres.push_back(Sphere { glm::vec3(3.0f, -2.0f, 10.2f), 1.0f, sphereMaterial(glm::vec3(0.8f, 0.2f, 0.2f)) });
        scene.spheres.push_back(Sphere { glm::vec3(-2.0f, 2.0f, 4.0f), 2.0f, sphereMaterial(glm::vec3(0.6f, 0.8f, 0.2f)) });
        scene.spheres.push_back(Sphere { glm::vec3(0.0f, 0.0f, 6.0f), 0.75f, sphereMaterial(glm::vec3(0.2f, 0.2f, 0.8f)) });
        scene.pointLights.push_back(PointLight { glm::vec3(3, 0, 3), glm::vec3(15.0f) });
        break;
    case Custom:
        addMeshes(scene, dataDir, { "custom.obj", false, false });
        scene.pointLights.push_back(keyLight);
        break;
    case ChessBoard:
        addMeshes(scene, dataDir, { "checker.obj", false, true });
        scene.sphericalLight.push_back(SphericalLight { glm::vec3(-1, 100, -25), 10.0f, white });
        break;
    case AndreasScene:
        addMeshes(scene, dataDir, { "AndreasScene.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    case CatalinScene:
        addMeshes(scene, dataDir, { "CatalinScene.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    case MikeScene:
        addMeshes(scene, dataDir, { "MikeScene.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    case MikeScene2:
        addMeshes(scene, dataDir, { "MikeScene2.obj", true, false });
        scene.pointLights.push_back(PointLight { glm::vec3(-2, 1, -2), white });
        break;
    case Bookeshelf:
        addMeshes(scene, dataDir, { "bookshelf.obj", true, false });
        scene.pointLights.push_back(keyLight);
        break;
    }
    return scene;
}
