#include "trackball.h"
#include <cmath>

Trackball::Trackball(Window* pWindow, float fovy, float distanceFromLookAt, float rotationX, float rotationY)
    : Trackball(pWindow, fovy, glm::vec3(0.0f), distanceFromLookAt, rotationX, rotationY)
{
}

Trackball::Trackball(Window* pWindow, float fovy, const glm::vec3& lookAt, float distanceFromLookAt, float rotationX, float rotationY)
    : m_pWindow(pWindow), m_fovy(fovy), m_lookAt(lookAt), m_distanceFromLookAt(distanceFromLookAt), m_rotationEulerAngles(rotationX, rotationY, 0.0f)
{
}

void Trackball::setCamera(const glm::vec3 lookAt, const glm::vec3 rotations, const float dist)
{
    m_lookAt = lookAt;
    m_rotationEulerAngles = rotations;
    m_distanceFromLookAt = dist;
}

// framework/src/trackball.cpp:65-68
glm::vec3 Trackball::position() const
{
    return m_lookAt + glm::quat(m_rotationEulerAngles) * glm::vec3(0.0f, 0.0f, -m_distanceFromLookAt);
}

glm::vec3 Trackball::forward() const { return glm::quat(m_rotationEulerAngles) * glm::vec3(0.0f, 0.0f, 1.0f); }
glm::vec3 Trackball::up() const { return glm::quat(m_rotationEulerAngles) * glm::vec3(0.0f, 1.0f, 0.0f); }
glm::vec3 Trackball::left() const { return glm::quat(m_rotationEulerAngles) * glm::vec3(1.0f, 0.0f, 0.0f); }

// framework/src/trackball.cpp:87-98 — host twin of the device `generate` kernel, used for single debug rays.
Ray Trackball::generateRay(const glm::vec2& pixel) const
{
    const float halfH = std::tan(m_fovy / 2.0f);
    const float halfW = m_pWindow->aspectRatio() * halfH;
    const glm::vec3 camDir = glm::normalize(glm::vec3(-pixel.x * halfW, pixel.y * halfH, 1.0f));
    Ray ray;
    ray.origin = position();
    ray.direction = glm::quat(m_rotationEulerAngles) * camDir;
    ray.t = std::numeric_limits<float>::max();
    return ray;
}
