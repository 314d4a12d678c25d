// rand() provider for the oracle/_ref build (see shim.h). TEST INFRASTRUCTURE ONLY.
// Compiled WITHOUT the force-include shim so it can reach glibc's real rand().
#include <stdint.h>
#include <stdlib.h>
#include "../philox.h"

namespace {
struct RefRng {
    int mode = 0;             // 0 MSVC LCG, 1 Philox stream, 2 glibc rand()
    uint32_t lcg = 1u;
    uint32_t pixel = 0, sample = 0, block = 0, widx = 0;
    bool have = false;
    uint32_t buf[4] = {0, 0, 0, 0};
};
thread_local RefRng t_rng;
uint32_t g_key[2] = {0, 0};
int g_default_mode = 0;
}

extern "C" {
void ref_rng_mode(int mode) { t_rng.mode = mode; g_default_mode = mode; }
void ref_rng_seed_lcg(uint32_t s) { t_rng.lcg = s; }
void ref_rng_key(uint32_t k0, uint32_t k1) { g_key[0] = k0; g_key[1] = k1; }
void ref_rng_begin_path(uint32_t pixel, uint32_t sample) {
    t_rng.pixel = pixel; t_rng.sample = sample; t_rng.block = 0; t_rng.widx = 0; t_rng.have = false;
}
int ref_rand(void) {
    RefRng& r = t_rng;
    if (r.mode == 0) {        // MSVC CRT: state*214013+2531011, return bits 16..30
        r.lcg = r.lcg * 214013u + 2531011u;
        return (int)((r.lcg >> 16) & 0x7fffu);
    }
    if (r.mode == 2) return ::rand() >> 16;   // glibc: one global, locked state (31 bits -> 15)
    if (!r.have) {
        const uint32_t ctr[4] = {r.pixel, r.sample, r.block, 0u};
        philox4x32_10(ctr, g_key, r.buf);
        r.have = true;
    }
    uint32_t w = r.buf[r.widx];
    if (r.widx == 3) { r.block++; r.widx = 0; r.have = false; }   // consecutive words, 4 per block
    else r.widx++;
    return (int)(w >> 17);
}
}
