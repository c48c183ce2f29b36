// sops_rng.cuh — device-side random streams of the SOPS chain kernels (sm_100a).
//
//  * ChaCha12 + PCG32 seed expansion + widening-multiply Uniform: the reference's *init* stream
//    (rand 0.8 StdRng::seed_from_u64 / Uniform::new(0,G), src/SOPSCore/mod.rs:56,113,126-127).
//  * WyRand double chain: the reference's *move* stream (fastrand 1.9 Rng::new() per draw,
//    mod.rs:59-62,274,278,285) made reproducible by starting the instance's thread-local generator at
//    move_seed.  The k-th Rng::new() is a pure function of (move_seed + k*INC), i.e. counter based.
//  * Philox2x32-10: fast mode, one block per chain step.
#pragma once
#include <cstdint>

namespace sops {

// ------------------------------------------------------------------------------------ ChaCha12 --
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int n) { return __funnelshift_l(v, v, n); }

#define SOPS_QR(a, b, c, d)                                  \
    a += b; d ^= a; d = rotl32(d, 16);                       \
    c += d; b ^= c; b = rotl32(b, 12);                       \
    a += b; d ^= a; d = rotl32(d, 8);                        \
    c += d; b ^= c; b = rotl32(b, 7);

// rand_core 0.6 SeedableRng::seed_from_u64: PCG32 output function over an LCG, 8 key words.
__device__ __forceinline__ void stdrng_key(uint64_t state, uint32_t key[8]) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        state = state * MUL + INC;
        uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        key[k] = __funnelshift_r(xs, xs, rot);
    }
}

// One ChaCha12 block (6 double rounds), 64-bit block counter, stream id 0 (rand_chacha layout).
__device__ __forceinline__ void chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3];
    uint32_t x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = 0, x15 = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        SOPS_QR(x0, x4, x8, x12) SOPS_QR(x1, x5, x9, x13) SOPS_QR(x2, x6, x10, x14) SOPS_QR(x3, x7, x11, x15)
        SOPS_QR(x0, x5, x10, x15) SOPS_QR(x1, x6, x11, x12) SOPS_QR(x2, x7, x8, x13) SOPS_QR(x3, x4, x9, x14)
    }
    out[0] = x0 + 0x61707865u; out[1] = x1 + 0x3320646eu; out[2] = x2 + 0x79622d32u; out[3] = x3 + 0x6b206574u;
    out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
    out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
    out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32); out[14] = x14; out[15] = x15;
}

// ------------------------------------------------------------------------------------- WyRand --
#define SOPS_WY_INC 0xA0761D6478BD642FULL
#define SOPS_WY_XOR 0xE7037ED1A0B428DBULL

__device__ __forceinline__ uint64_t wy_mix(uint64_t s) {
    uint64_t b = s ^ SOPS_WY_XOR;
    return (s * b) ^ __umul64hi(s, b);
}
// value of gen_u64() of the generator created by the k-th (1-based) fastrand::Rng::new() of an
// instance whose thread-local generator started at `move_seed`
__device__ __forceinline__ uint64_t wy_draw(uint64_t move_seed, uint64_t k) {
    uint64_t tl = move_seed + k * SOPS_WY_INC;      // thread-local state after k draws
    uint64_t fresh = wy_mix(tl);                    // = seed of the fresh Rng
    return wy_mix(fresh + SOPS_WY_INC);             // its first gen_u64
}
// j-th redraw (j >= 1) of that same fresh generator — only reached inside Lemire's rejection loop
__device__ __noinline__ uint64_t wy_redraw(uint64_t move_seed, uint64_t k, uint32_t j) {
    uint64_t tl = move_seed + k * SOPS_WY_INC;
    uint64_t fresh = wy_mix(tl);
    return wy_mix(fresh + (uint64_t)(j + 1) * SOPS_WY_INC);
}
// fastrand gen_mod_u64(n) given the first raw draw r (draw number k): Lemire with the rejection loop
__device__ __forceinline__ uint32_t wy_mod_u64(uint64_t r, uint32_t n, uint64_t move_seed, uint64_t k) {
    uint64_t hi = __umul64hi(r, (uint64_t)n);
    uint64_t lo = r * (uint64_t)n;
    if (lo < n) {                                   // p = n / 2^64
        uint64_t t = (0 - (uint64_t)n) % n;
        uint32_t j = 1;
        while (lo < t) {
            r = wy_redraw(move_seed, k, j++);
            hi = __umul64hi(r, (uint64_t)n);
            lo = r * (uint64_t)n;
        }
    }
    return (uint32_t)hi;
}
// fastrand gen_mod_u32(1000) on gen_u32() = low word of gen_u64(); returns 0..999 (caller adds 1)
__device__ __forceinline__ uint32_t wy_mod_1000(uint32_t r32, uint64_t move_seed, uint64_t k) {
    uint64_t m = (uint64_t)r32 * 1000u;
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    if (lo < 1000u) {                               // p = 1000 / 2^32
        const uint32_t t = (0u - 1000u) % 1000u;    // 296
        uint32_t j = 1;
        while (lo < t) {
            uint32_t r = (uint32_t)wy_redraw(move_seed, k, j++);
            m = (uint64_t)r * 1000u;
            hi = (uint32_t)(m >> 32); lo = (uint32_t)m;
        }
    }
    return hi;
}

// ------------------------------------------------------------------------------ Philox2x32-10 --
__device__ __forceinline__ void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t& o0, uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(c0), "r"(0xD256D193u));   // one IMAD.WIDE: hi and lo together
        const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    o0 = c0; o1 = c1;
}

}  // namespace sops
