// A SCRIPTED, headless double of the few SDL2 / Dear ImGui entry points host/rt_viewer.cpp uses (declared in this directory's
// SDL.h / imgui*.h): no window, no GPU API of its own. Linked with the viewer's real source it turns `rt_viewer` into a program
// whose main loop - event handling, fly camera, inspector edits, picking, the frame state machine, rt_render_frame into the
// streamed surface - runs against librt_b200 on the GPU under test (tests/test_gpu_round2.py), driven by a script instead of
// a user. TEST INFRASTRUCTURE ONLY: the real front end is built with `make viewer` where SDL2 and Dear ImGui exist.
//
// Script (env RT_VIEWER_SCRIPT), one command per line, `<frame> <command> [args]`; frame = number of frames presented so far:
//   quit | rmb down|up | motion DX DY | click X Y | key P|DELETE | hold W|A|S|D|Q|E|LSHIFT on|off
//   menu <label> | button <label> | setint <label>=V | setfloat <label>=V | settext <label>=TEXT | dump FILE.ppm
// Log (env RT_VIEWER_LOG): `frame N title <window title>`, `frame N name <inspector name field>`, `frame N surface <fnv1a64>`.
#include <SDL.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "imgui.h"
#include "imgui_impl_sdl2.h"
#include "imgui_impl_sdlrenderer.h"

namespace {

struct Cmd { int frame; std::string what, arg; };
struct Double {
    std::vector<Cmd> script;
    std::deque<SDL_Event> events;
    int frame = 0, queued_for = -1;
    FILE* log = nullptr;
    Uint8 keys[512] = {0};
    int mouse_x = 0, mouse_y = 0;
    int tex_w = 0, tex_h = 0;
    std::vector<uint32_t> texture;
    std::string title;
    ImGuiIO io{};
    bool loaded = false, pending = false;                     // pending: a frame has begun and its end is not logged yet

    void load() {
        if (loaded) return;
        loaded = true;
        io.Framerate = 60.f; io.DisplayFramebufferScale = ImVec2{1.f, 1.f};
        if (const char* lp = getenv("RT_VIEWER_LOG")) log = fopen(lp, "w");
        const char* sp = getenv("RT_VIEWER_SCRIPT");
        FILE* f = sp ? fopen(sp, "r") : nullptr;
        if (!f) { script.push_back({0, "quit", ""}); return; }      // no script: one frame and out
        char line[1024];
        while (fgets(line, sizeof line, f)) {
            int fr = 0, used = 0;
            char what[64];
            if (line[0] == '#' || sscanf(line, "%d %63s %n", &fr, what, &used) < 2) continue;
            std::string arg = line + used;
            while (!arg.empty() && (arg.back() == '\n' || arg.back() == '\r' || arg.back() == ' ')) arg.pop_back();
            script.push_back({fr, what, arg});
        }
        fclose(f);
    }
    void say(const char* fmt, ...) {
        if (!log) return;
        fprintf(log, "frame %d ", frame);
        va_list ap; va_start(ap, fmt); vfprintf(log, fmt, ap); va_end(ap);
        fputc('\n', log); fflush(log);
    }
    static SDL_Scancode scancode(const std::string& k) {
        static const struct { const char* name; SDL_Scancode code; } table[] = {
            {"W", SDL_SCANCODE_W}, {"A", SDL_SCANCODE_A}, {"S", SDL_SCANCODE_S}, {"D", SDL_SCANCODE_D}, {"Q", SDL_SCANCODE_Q},
            {"E", SDL_SCANCODE_E}, {"P", SDL_SCANCODE_P}, {"DELETE", SDL_SCANCODE_DELETE}, {"LSHIFT", SDL_SCANCODE_LSHIFT}};
        for (const auto& t : table) if (k == t.name) return t.code;
        return SDL_SCANCODE_LSHIFT;
    }
    // the frame's input events become available at its first SDL_PollEvent
    void queue_frame_events() {
        if (queued_for == frame) return;
        queued_for = frame;
        for (const Cmd& c : script) {
            if (c.frame != frame) continue;
            SDL_Event e; memset(&e, 0, sizeof e);
            if (c.what == "quit") { e.type = SDL_QUIT; events.push_back(e); }
            else if (c.what == "rmb") { e.type = c.arg == "down" ? SDL_MOUSEBUTTONDOWN : SDL_MOUSEBUTTONUP; e.button.button = SDL_BUTTON_RIGHT; events.push_back(e); }
            else if (c.what == "motion") { int dx = 0, dy = 0; sscanf(c.arg.c_str(), "%d %d", &dx, &dy); e.type = SDL_MOUSEMOTION; e.motion.xrel = dx; e.motion.yrel = dy; events.push_back(e); }
            else if (c.what == "click") { sscanf(c.arg.c_str(), "%d %d", &mouse_x, &mouse_y); e.type = SDL_MOUSEBUTTONDOWN; e.button.button = SDL_BUTTON_LEFT; events.push_back(e); }
            else if (c.what == "key") { e.type = SDL_KEYDOWN; e.key.repeat = 0; e.key.keysym.scancode = scancode(c.arg); events.push_back(e); }
            else if (c.what == "hold") {
                const size_t sp = c.arg.find(' ');
                keys[scancode(c.arg.substr(0, sp))] = (sp != std::string::npos && c.arg.substr(sp + 1) == "on") ? 1 : 0;
            }
        }
    }
    // a widget command of this frame: `<what> <label>` or `<what> <label>=<value>`
    const Cmd* widget(const char* what, const char* label, std::string* value = nullptr) const {
        for (const Cmd& c : script) {
            if (c.frame != frame || c.what != what) continue;
            const size_t eq = c.arg.find('=');
            if (c.arg.substr(0, eq) != label) continue;
            if (value) *value = eq == std::string::npos ? "" : c.arg.substr(eq + 1);
            return &c;
        }
        return nullptr;
    }
    void end_of_frame() {
        if (!texture.empty()) {
            unsigned long long h = 1469598103934665603ull;
            for (uint32_t v : texture) for (int b = 0; b < 4; ++b) { h ^= (v >> (8 * b)) & 0xffu; h *= 1099511628211ull; }
            say("surface %016llx", h);
        }
        say("title %s", title.c_str());
        for (const Cmd& c : script) {
            if (c.frame != frame || c.what != "dump" || texture.empty()) continue;
            if (FILE* f = fopen(c.arg.c_str(), "wb")) {
                fprintf(f, "P6\n%d %d\n255\n", tex_w, tex_h);
                for (uint32_t v : texture) { const unsigned char rgb[3] = {(unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v}; fwrite(rgb, 1, 3, f); }
                fclose(f);
            }
        }
        ++frame;
    }
} g;

int g_handles[3];

}  // namespace

// ---- SDL ------------------------------------------------------------------------------------------------------------
int SDL_Init(Uint32) { g.load(); return 0; }
void SDL_Quit() {
    if (g.pending) { g.end_of_frame(); g.pending = false; }  // the loop's last frame (the one that saw SDL_QUIT) has no next poll
    if (g.log) { fclose(g.log); g.log = nullptr; }
}
const char* SDL_GetError() { return "scripted SDL double"; }
SDL_Window* SDL_CreateWindow(const char*, int, int, int, int, Uint32) { return reinterpret_cast<SDL_Window*>(&g_handles[0]); }
void SDL_DestroyWindow(SDL_Window*) {}
void SDL_SetWindowTitle(SDL_Window*, const char* t) { g.title = t; }
SDL_Renderer* SDL_CreateRenderer(SDL_Window*, int, Uint32) { return reinterpret_cast<SDL_Renderer*>(&g_handles[1]); }
void SDL_DestroyRenderer(SDL_Renderer*) {}
SDL_Texture* SDL_CreateTexture(SDL_Renderer*, Uint32, int, int w, int h) {
    g.tex_w = w; g.tex_h = h;
    return reinterpret_cast<SDL_Texture*>(&g_handles[2]);
}
void SDL_DestroyTexture(SDL_Texture*) {}
int SDL_UpdateTexture(SDL_Texture*, const SDL_Rect*, const void* pixels, int pitch) {
    g.texture.resize((size_t)g.tex_w * g.tex_h);
    for (int y = 0; y < g.tex_h; ++y) memcpy(g.texture.data() + (size_t)y * g.tex_w, static_cast<const char*>(pixels) + (size_t)y * pitch, (size_t)g.tex_w * 4);
    return 0;
}
int SDL_RenderSetScale(SDL_Renderer*, float, float) { return 0; }
int SDL_RenderClear(SDL_Renderer*) { return 0; }
int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*) { return 0; }
// the title of a frame is set AFTER its present (Raytracer.cpp:552-567): the frame ends at the next loop's first poll
void SDL_RenderPresent(SDL_Renderer*) {}
int SDL_PollEvent(SDL_Event* e) {
    static bool in_frame = false;
    if (!in_frame) {                                          // first poll of a loop iteration: the previous frame is complete
        if (g.pending) g.end_of_frame();
        g.pending = true;
        in_frame = true;
        g.queue_frame_events();
    }
    if (g.events.empty()) { in_frame = false; return 0; }
    *e = g.events.front(); g.events.pop_front();
    return 1;
}
const Uint8* SDL_GetKeyboardState(int*) { return g.keys; }
Uint32 SDL_GetMouseState(int* x, int* y) { if (x) *x = g.mouse_x; if (y) *y = g.mouse_y; return 0; }

// ---- Dear ImGui -----------------------------------------------------------------------------------------------------
namespace ImGui {
void* CreateContext() { g.load(); return &g; }
void DestroyContext() {}
ImGuiIO& GetIO() { g.load(); return g.io; }
void StyleColorsDark() {}
void NewFrame() {}
void Render() {}
ImDrawData* GetDrawData() { return nullptr; }
bool Begin(const char*) { return true; }
void End() {}
bool BeginMenu(const char*) { return true; }                 // every menu is open: its items are evaluated every frame
void EndMenu() {}
bool MenuItem(const char* label, const char*) { return g.widget("menu", label) != nullptr; }
bool CollapsingHeader(const char*) { return true; }
bool Button(const char* label) { return g.widget("button", label) != nullptr; }
void NewLine() {}
void Text(const char*, ...) {}
bool InputText(const char* label, char* buf, size_t n) {
    if (!strcmp(label, "Name")) g.say("name %s", buf);
    std::string v;
    if (!g.widget("settext", label, &v)) return false;
    snprintf(buf, n, "%s", v.c_str());
    return true;
}
static bool set_int(const char* label, int* p) { std::string v; if (!g.widget("setint", label, &v)) return false; *p = atoi(v.c_str()); return true; }
static bool set_floats(const char* label, float* p, int n) {
    std::string v;
    if (!g.widget("setfloat", label, &v)) return false;
    const char* s = v.c_str();
    for (int i = 0; i < n; ++i) { char* end = nullptr; p[i] = strtof(s, &end); if (end == s) break; s = end; }
    return true;
}
bool InputInt(const char* label, int* v) { return set_int(label, v); }
bool SliderInt(const char* label, int* v, int, int) { return set_int(label, v); }
bool InputFloat(const char* label, float* v) { return set_floats(label, v, 1); }
bool SliderFloat(const char* label, float* v, float, float) { return set_floats(label, v, 1); }
bool InputFloat3(const char* label, float v[3]) { return set_floats(label, v, 3); }
bool DragFloat3(const char* label, float v[3], float) { return set_floats(label, v, 3); }
bool ColorPicker3(const char* label, float v[3]) { return set_floats(label, v, 3); }
}  // namespace ImGui

bool ImGui_ImplSDL2_InitForSDLRenderer(SDL_Window*, SDL_Renderer*) { return true; }
void ImGui_ImplSDL2_Shutdown() {}
void ImGui_ImplSDL2_NewFrame() {}
bool ImGui_ImplSDL2_ProcessEvent(const SDL_Event*) { return false; }
bool ImGui_ImplSDLRenderer_Init(SDL_Renderer*) { return true; }
void ImGui_ImplSDLRenderer_Shutdown() {}
void ImGui_ImplSDLRenderer_NewFrame() {}
void ImGui_ImplSDLRenderer_RenderDrawData(ImDrawData*) {}
