/* Philox4x32-10 counter-based RNG (Salmon et al., SC'11), restated for the oracle.
 * TEST INFRASTRUCTURE ONLY: used by oracle/pt_oracle.c and oracle/ref_harness (which feeds
 * the reference's own rand() call sites, Raytracer.cpp:93-95,165,182, from this stream).
 * The product has its own device implementation (csrc/rt_rng.cuh); tests pin the two
 * against each other and against the Random123 known-answer vectors.
 *
 * Stream layout: the reference's rand() calls of ONE path, in call order, read consecutive words
 * of one Philox stream: counter = (pixel = x + y*W [y-up], sample index, block, 0), key =
 * (seed_lo, seed_hi), draw j = word (j & 3) of block (j >> 2). With the reference's call order
 * (Raytracer.cpp:165, then per bounce :93-95 and :182) that is, for the hit at depth k:
 *   block k: word0 -> specular coin drawn at that hit, word1..3 -> direction x,y,z of the
 *            scatter that leaves it.
 * So one 128-bit block serves one hit; a path never shares words with another path.
 * A word becomes the reference's `rand()` value as word >> 17, i.e. 15 bits in
 * [0, RAND_MAX=32767] - the MSVC C runtime the reference ships on. */
#ifndef ORACLE_PHILOX_H
#define ORACLE_PHILOX_H
#include <stdint.h>
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif
