// Declaration-only stand-in (Dear ImGui 1.89.x SDL_Renderer back end). TEST INFRASTRUCTURE ONLY.
#pragma once
struct SDL_Renderer; struct ImDrawData;
bool ImGui_ImplSDLRenderer_Init(SDL_Renderer*); void ImGui_ImplSDLRenderer_Shutdown(); void ImGui_ImplSDLRenderer_NewFrame(); void ImGui_ImplSDLRenderer_RenderDrawData(ImDrawData*);
