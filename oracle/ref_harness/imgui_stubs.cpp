// Link-only stand-ins for the eight Dear ImGui widgets Object.hpp:44-78,148-152,207-217 call
// inline from OnGUI(). The oracle never opens a GUI; the real widgets are the viewer's
// business (out of scope). Signatures come from the reference's own imgui.h.
#include "imgui.h"
namespace ImGui {
bool InputText(const char*, char*, size_t, ImGuiInputTextFlags, ImGuiInputTextCallback, void*) { return false; }
bool DragFloat3(const char*, float[3], float, float, float, const char*, ImGuiSliderFlags) { return false; }
bool CollapsingHeader(const char*, ImGuiTreeNodeFlags) { return false; }
bool ColorPicker3(const char*, float[3], ImGuiColorEditFlags) { return false; }
bool InputFloat3(const char*, float[3], const char*, ImGuiInputTextFlags) { return false; }
bool InputFloat(const char*, float*, float, float, const char*, ImGuiInputTextFlags) { return false; }
bool SliderFloat(const char*, float*, float, float, const char*, ImGuiSliderFlags) { return false; }
void NewLine() {}
}
