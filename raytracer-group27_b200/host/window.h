// Headless stand-in for the reference's GLFW window wrapper (framework/include/window.h): the render path only
// ever asks it for the aspect ratio (framework/src/window.cpp:334-337, used by Trackball::generateRay).
#pragma once
#include <glm/vec2.hpp>
#include <string_view>

enum class OpenGLVersion { GL2, GL3, GL45 };

class Window {
public:
    Window(std::string_view /*title*/, const glm::ivec2& windowSize, OpenGLVersion = OpenGLVersion::GL2) : m_windowSize(windowSize) {}
    explicit Window(const glm::ivec2& windowSize) : m_windowSize(windowSize) {}
    [[nodiscard]] glm::ivec2 windowSize() const { return m_windowSize; }
    [[nodiscard]] float aspectRatio() const { return float(m_windowSize.x) / float(m_windowSize.y); }

private:
    glm::ivec2 m_windowSize;
};
