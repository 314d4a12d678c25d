/* Force-included (-include) before the reference's own headers when building
 * oracle/_ref from the sources where they lie under $REF_DIR.
 * TEST INFRASTRUCTURE ONLY - nothing here ships in the product library.
 *
 * <math.h>/<stdlib.h> must come first so ::abs(float) is visible when
 * Common.hpp:330 is parsed (the real TU gets that from Raytracer.cpp:7);
 * otherwise abs() binds to int abs(int) and every Box silently misses. */
#pragma once
#include <math.h>
#include <stdlib.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <iostream>
#include <cmath>
#include <cfloat>
#include <climits>
#include <limits>
#include <string>
#include <chrono>
#include <thread>
#include <random>

/* SDL types the hot path touches (Raytracer.cpp:50-52,64; Common.hpp:189-208). */
typedef uint32_t Uint32;
typedef uint8_t Uint8;
struct SDL_PixelFormat { Uint8 BytesPerPixel; };
struct SDL_Surface { void* pixels; int pitch; SDL_PixelFormat* format; };
struct SDL_Texture;
struct SDL_Renderer;

/* MSVC-only keywords / CRT used by Object.hpp:11,15,46 and Scene.hpp:12. */
#define abstract
#define sealed final
template <size_t N> inline void strcpy_s(char (&dst)[N], const char* src) {
    strncpy(dst, src, N - 1); dst[N - 1] = 0;
}
static inline void Sleep(int ms) { std::this_thread::sleep_for(std::chrono::milliseconds(ms)); }

/* Resolution is a compile-time #define in the reference (Raytracer.cpp:26-27); the harness
 * makes it a run-time global so one build serves every golden resolution. */
extern int g_ref_w, g_ref_h;
#define SCREEN_WIDTH g_ref_w
#define SCREEN_HEIGHT g_ref_h
#define THREADS 16

/* rand(): the reference draws from the C library (Raytracer.cpp:93-95,165,182). Its native
 * platform is MSVC (RAND_MAX 32767, per-thread state). The harness supplies the stream:
 * either an MSVC-style thread-local LCG or the counter-based Philox stream the GPU uses. */
extern "C" int ref_rand(void);
#define rand ref_rand
#undef RAND_MAX
#define RAND_MAX 32767
