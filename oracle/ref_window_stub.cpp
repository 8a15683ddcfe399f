// ref_window_stub.cpp — TEST INFRASTRUCTURE (oracle).  Headless stand-in for the reference's
// Src/Linux/RenderWindow_X11.cpp, which needs X11 headers this image does not have.  It defines
// the eight members declared in Src/Linux/RenderWindow_X11.h:14-26 as no-ops so that the
// unmodified RayTracerProgram.cpp links; nothing on the ray/scene hot path touches the window.
#include "Linux/RenderWindow_X11.h"

struct RenderWindow::X11WindowContext {};

RenderWindow::RenderWindow() : Context(nullptr) {}
RenderWindow::~RenderWindow() {}
bool RenderWindow::Create(int, int, bool, int) { return true; }
void RenderWindow::Destroy() {}
void RenderWindow::SetRenderBufferParameters(int, int, void*) {}
void RenderWindow::RunWindowLoop(RayTracerProgram*) {}
void RenderWindow::SetTitle(const char*) {}
void RenderWindow::PresentRenderBuffer() {}
