// Declaration-only stand-in for <SDL.h> (SDL2 is not in this image): just what host/rt_viewer.cpp uses, so that the
// viewer source can be syntax-checked (tests/test_capi_cpu.py). Never linked. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <stdint.h>
typedef uint8_t Uint8; typedef uint32_t Uint32; typedef int32_t Sint32;
struct SDL_Window; struct SDL_Renderer; struct SDL_Texture; struct SDL_Rect;
enum { SDL_INIT_VIDEO = 0x20, SDL_WINDOWPOS_CENTERED = 0x2FFF0000, SDL_RENDERER_ACCELERATED = 2, SDL_PIXELFORMAT_ARGB8888 = 0x16362004, SDL_TEXTUREACCESS_STREAMING = 1 };
enum { SDL_QUIT = 0x100, SDL_WINDOWEVENT = 0x200, SDL_KEYDOWN = 0x300, SDL_MOUSEMOTION = 0x400, SDL_MOUSEBUTTONDOWN, SDL_MOUSEBUTTONUP };
enum { SDL_WINDOWEVENT_CLOSE = 14, SDL_BUTTON_LEFT = 1, SDL_BUTTON_RIGHT = 3 };
enum SDL_Scancode { SDL_SCANCODE_A = 4, SDL_SCANCODE_D = 7, SDL_SCANCODE_E = 8, SDL_SCANCODE_P = 19, SDL_SCANCODE_Q = 20, SDL_SCANCODE_S = 22, SDL_SCANCODE_W = 26,
                    SDL_SCANCODE_DELETE = 76, SDL_SCANCODE_LSHIFT = 225 };
struct SDL_Keysym { SDL_Scancode scancode; };
struct SDL_KeyboardEvent { Uint32 type; Uint8 repeat; SDL_Keysym keysym; };
struct SDL_MouseMotionEvent { Uint32 type; Sint32 xrel, yrel; };
struct SDL_MouseButtonEvent { Uint32 type; Uint8 button; };
struct SDL_WindowEvent { Uint32 type; Uint8 event; };
union SDL_Event { Uint32 type; SDL_KeyboardEvent key; SDL_MouseMotionEvent motion; SDL_MouseButtonEvent button; SDL_WindowEvent window; };
int SDL_Init(Uint32); void SDL_Quit(); const char* SDL_GetError();
SDL_Window* SDL_CreateWindow(const char*, int, int, int, int, Uint32); void SDL_DestroyWindow(SDL_Window*); void SDL_SetWindowTitle(SDL_Window*, const char*);
SDL_Renderer* SDL_CreateRenderer(SDL_Window*, int, Uint32); void SDL_DestroyRenderer(SDL_Renderer*);
SDL_Texture* SDL_CreateTexture(SDL_Renderer*, Uint32, int, int, int); void SDL_DestroyTexture(SDL_Texture*);
int SDL_UpdateTexture(SDL_Texture*, const SDL_Rect*, const void*, int);
int SDL_RenderSetScale(SDL_Renderer*, float, float); int SDL_RenderClear(SDL_Renderer*); int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, const SDL_Rect*, const SDL_Rect*);
void SDL_RenderPresent(SDL_Renderer*);
int SDL_PollEvent(SDL_Event*); const Uint8* SDL_GetKeyboardState(int*); Uint32 SDL_GetMouseState(int*, int*);
