// Declaration-only stand-in (Dear ImGui 1.89.x SDL2 platform back end). TEST INFRASTRUCTURE ONLY.
#pragma once
struct SDL_Window; struct SDL_Renderer; union SDL_Event;
bool ImGui_ImplSDL2_InitForSDLRenderer(SDL_Window*, SDL_Renderer*); void ImGui_ImplSDL2_Shutdown(); void ImGui_ImplSDL2_NewFrame(); bool ImGui_ImplSDL2_ProcessEvent(const SDL_Event*);
