// Trackball — the reference's orbit camera (framework/include/trackball.h:14-55) without the mouse callbacks:
// same constructor shapes, setCamera / position / lookAt / generateRay.
#pragma once
#include "ray.h"
#include "window.h"
#include <glm/gtc/quaternion.hpp>
#include <glm/vec2.hpp>
#include <glm/vec3.hpp>

class Trackball {
public:
    // fovy in radians
    Trackball(Window* pWindow, float fovy, float distanceFromLookAt = 4.0f, float rotationX = 0.0f, float rotationY = 0.0f);
    Trackball(Window* pWindow, float fovy, const glm::vec3& lookAt, float distanceFromLookAt = 4.0f, float rotationX = 0.0f, float rotationY = 0.0f);

    [[nodiscard]] glm::vec3 left() const;
    [[nodiscard]] glm::vec3 up() const;
    [[nodiscard]] glm::vec3 forward() const;
    [[nodiscard]] glm::vec3 position() const;
    [[nodiscard]] glm::vec3 lookAt() const { return m_lookAt; }
    void setCamera(const glm::vec3 lookAt, const glm::vec3 rotations, const float dist);
    // pixel in NDC: (-1,-1) bottom left, (+1,+1) top right
    [[nodiscard]] Ray generateRay(const glm::vec2& pixel) const;

    // accessors the device renderer needs to rebuild the same camera
    [[nodiscard]] float fovy() const { return m_fovy; }
    [[nodiscard]] float distanceFromLookAt() const { return m_distanceFromLookAt; }
    [[nodiscard]] glm::vec3 rotationEulerAngles() const { return m_rotationEulerAngles; }
    [[nodiscard]] const Window* window() const { return m_pWindow; }

private:
    const Window* m_pWindow;
    float m_fovy;
    glm::vec3 m_lookAt { 0.0f };
    float m_distanceFromLookAt;
    glm::vec3 m_rotationEulerAngles { 0.0f };
};
