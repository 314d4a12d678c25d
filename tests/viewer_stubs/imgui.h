// Declaration-only stand-in for Dear ImGui 1.89.x's imgui.h: just what host/rt_viewer.cpp uses. Never linked. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <stddef.h>
#define IMGUI_CHECKVERSION() ((void)0)
struct ImVec2 { float x, y; };
struct ImDrawData;
struct ImGuiIO { bool ConfigDragClickToInputText, WantCaptureMouse, WantCaptureKeyboard; float Framerate; ImVec2 DisplayFramebufferScale; };
namespace ImGui {
void* CreateContext(); void DestroyContext(); ImGuiIO& GetIO(); void StyleColorsDark(); void NewFrame(); void Render(); ImDrawData* GetDrawData();
bool Begin(const char*); void End(); bool BeginMenu(const char*); void EndMenu(); bool MenuItem(const char*, const char* shortcut = nullptr);
bool CollapsingHeader(const char*); bool Button(const char*); void NewLine(); void Text(const char*, ...);
bool InputText(const char*, char*, size_t); bool InputInt(const char*, int*); bool InputFloat(const char*, float*); bool InputFloat3(const char*, float[3]);
bool DragFloat3(const char*, float[3], float); bool SliderInt(const char*, int*, int, int); bool SliderFloat(const char*, float*, float, float);
bool ColorPicker3(const char*, float[3]);
}
