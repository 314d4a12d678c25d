// Host shim so tests can compile csrc/rt_device.cuh with g++ and run the DEVICE functions'
// logic on the CPU (test infrastructure only - the product never runs this).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct int2 { int x, y; };
static inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
// PRMT: nibble k of s selects result byte k from the 8 bytes of (y:x); bit 3 of the nibble replicates the byte's sign bit
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t sel = (s >> (4 * k)) & 0xf;
        uint32_t b = (uint32_t)(src >> (8 * (sel & 7))) & 0xff;
        // (the CUDA intrinsic ignores the replicate-sign flag of PTX prmt: only selector bits 2..0 count)
        r |= b << (8 * k);
    }
    return r;
}
