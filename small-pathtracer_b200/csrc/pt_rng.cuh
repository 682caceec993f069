// pt_rng.cuh — Philox4x32-10 (Salmon, Moraes, Dror, Shaw; SC'11), the counter-based generator of
// the production engine.  Stream layout: ctr = (pixel, sample, vertex, purpose), key = seed; the block of a path's
// first vertex also supplies the sub-pixel jitter of its camera ray.
// Replaces the reference's libc rand() jitter/light draws (src/smallpt.cpp:365-366,533-534) and the
// per-row erand48 stream (:530) with draws that depend only on (pixel, sample, vertex): any sharding
// of the image gives bit-identical pixels.
#ifndef PT_RNG_CUH
#define PT_RNG_CUH

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#define PT_PHILOX_M0 0xD2511F53u
#define PT_PHILOX_M1 0xCD9E8D57u
#define PT_PHILOX_W0 0x9E3779B9u
#define PT_PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(PT_PHILOX_M0, c0), lo0 = PT_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(PT_PHILOX_M1, c2), lo1 = PT_PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += PT_PHILOX_W0;
        k1 += PT_PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// round keys of a Philox key: k + r * W (host: once per render, into KParams::philox_rk)
#ifdef __CUDACC_RTC__
__device__
#else
__host__ __device__
#endif
inline void philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk)[10][2])
{
    for (uint32_t r = 0; r < 10; r++) { rk[r][0] = k0 + r * PT_PHILOX_W0; rk[r][1] = k1 + r * PT_PHILOX_W1; }
}

// The same generator with the ten round keys (k0 + r W0, k1 + r W1) taken from a table the host filled: the key is
// a per-render constant, so its schedule needs no instruction in the kernel.
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[10][2])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(PT_PHILOX_M0, c0), lo0 = PT_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(PT_PHILOX_M1, c2), lo1 = PT_PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ rk[r][0];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[r][1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// [0,1) with 24 random bits (exact in FP32; never 1.0f)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

enum { PT_DRAW_A = 0, PT_DRAW_B = 1, PT_DRAW_LIGHT0 = 2 };

#endif
