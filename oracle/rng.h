// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the random-number stacks the reference's hot path draws from.  None of these
// crates is vendored in /root/reference (no Cargo.lock; caret ranges at Cargo.toml:9-15), so each is
// restated from its published algorithm:
//   * rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4:  StdRng::seed_from_u64 + Uniform<usize>
//       call sites: src/SOPSCore/mod.rs:56,113,126-127 (and the same lines in the other three envs).
//       PINNED: reproduces the reference's own step-0 snapshot grids 729/729 cells
//       (tests/golden/) and the RFC 8439 ChaCha20 block vector.
//   * fastrand 1.9.0 (WyRand): call sites src/SOPSCore/mod.rs:61,274,278,285 etc.
//       PARITY UNPINNED for the move stream: the reference seeds it from time/thread-id
//       (unreproducible by construction), so the oracle DEFINES the reproducible stream:
//       the thread-local generator of an instance starts at `move_seed`.
//   * Philox2x32-10 (Salmon et al. 2011, Random123): the fast-mode counter RNG (ours, not the
//       reference's); pinned by the Random123 known-answer vectors in tests.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>

namespace oracle {

// ---------------------------------------------------------------- ChaCha (rand_chacha 0.3.1) ----
static inline uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
static inline uint32_t rotr32(uint32_t v, int n) { return (v >> n) | (v << (32 - n)); }

static inline void chacha_qr(uint32_t* x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 16);
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 12);
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 8);
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 7);
}

// One ChaCha block with `double_rounds` double rounds (6 -> ChaCha12, 10 -> ChaCha20).
// Layout as rand_chacha: constants | key[8] | 64-bit block counter | 64-bit stream id (0).
static inline void chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream,
                                int double_rounds, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32),
                       (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
    memcpy(x, in, sizeof(x));
    for (int r = 0; r < double_rounds; ++r) {
        chacha_qr(x, 0, 4, 8, 12); chacha_qr(x, 1, 5, 9, 13);
        chacha_qr(x, 2, 6, 10, 14); chacha_qr(x, 3, 7, 11, 15);
        chacha_qr(x, 0, 5, 10, 15); chacha_qr(x, 1, 6, 11, 12);
        chacha_qr(x, 2, 7, 8, 13); chacha_qr(x, 3, 4, 9, 14);
    }
    for (int k = 0; k < 16; ++k) out[k] = x[k] + in[k];
}

// rand 0.8 StdRng == ChaCha12Rng.  seed_from_u64 is rand_core 0.6's PCG32 expansion.
struct StdRng {
    uint32_t key[8];
    uint64_t block;     // next block counter
    uint32_t buf[16];
    int pos;            // next unread word in buf (16 = empty)

    explicit StdRng(uint64_t seed) { seed_from_u64(seed); }

    void seed_from_u64(uint64_t state) {
        const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
        for (int k = 0; k < 8; ++k) {
            state = state * MUL + INC;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            key[k] = rotr32(xorshifted, (int)rot);
        }
        block = 0;
        pos = 16;
    }
    uint32_t next_u32() {
        if (pos >= 16) { chacha_block(key, block++, 0, 6, buf); pos = 0; }
        return buf[pos++];
    }
    // BlockRng::next_u64: two consecutive words, low first.  The init path only ever calls
    // next_u64, so the index stays even and the buffer-straddle case cannot occur.
    uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return lo | (hi << 32);
    }
    // rand 0.8 UniformInt<usize>::new(0, range).sample  (widening multiply + zone rejection)
    uint64_t uniform_below(uint64_t range) {
        uint64_t ints_to_reject = (0 - range) % range;   // (2^64 - range) mod range
        uint64_t zone = ~(uint64_t)0 - ints_to_reject;
        for (;;) {
            uint64_t v = next_u64();
            unsigned __int128 m = (unsigned __int128)v * range;
            uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
            if (lo <= zone) return hi;
        }
    }
};

// ---------------------------------------------------------------- fastrand 1.9.0 (WyRand) -------
// test instrumentation: how often Lemire's rejection loop actually redrew (u32 and u64 paths)
static std::atomic<unsigned long long> g_lemire_redraws{0};

struct WyRand {
    uint64_t s;
    explicit WyRand(uint64_t seed) : s(seed) {}
    uint64_t gen_u64() {
        s += 0xA0761D6478BD642FULL;
        unsigned __int128 t = (unsigned __int128)s * (unsigned __int128)(s ^ 0xE7037ED1A0B428DBULL);
        return (uint64_t)t ^ (uint64_t)(t >> 64);
    }
    uint32_t gen_u32() { return (uint32_t)gen_u64(); }
    uint64_t gen_mod_u64(uint64_t n) {              // Lemire, 64-bit
        uint64_t r = gen_u64();
        unsigned __int128 m = (unsigned __int128)r * n;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo < n) {
            uint64_t t = (0 - n) % n;
            while (lo < t) {
                g_lemire_redraws.fetch_add(1, std::memory_order_relaxed);
                r = gen_u64();
                m = (unsigned __int128)r * n;
                hi = (uint64_t)(m >> 64); lo = (uint64_t)m;
            }
        }
        return hi;
    }
    uint32_t gen_mod_u32(uint32_t n) {              // Lemire, 32-bit
        uint32_t r = gen_u32();
        uint64_t m = (uint64_t)r * n;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo < n) {
            uint32_t t = (0u - n) % n;
            while (lo < t) {
                g_lemire_redraws.fetch_add(1, std::memory_order_relaxed);
                r = gen_u32();
                m = (uint64_t)r * n;
                hi = (uint32_t)(m >> 32); lo = (uint32_t)m;
            }
        }
        return hi;
    }
    uint64_t usize_below(uint64_t n) { return gen_mod_u64(n); }                  // rng.usize(..n)
    uint16_t u16_1_to_1000() { return (uint16_t)(1 + gen_mod_u32(1000)); }       // rng.u16(1..=1000)
};

// ---------------------------------------------------------------- Philox2x32-10 ----------------
static inline void philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p = (uint64_t)0xD256D193u * c0;
        uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    out[0] = c0; out[1] = c1;
}

// ---------------------------------------------------------------- per-step draw sources ---------
// One move_particles call makes 2 draws (particle, direction) and a 3rd (probability) only when the
// move is possible.  Both modes expose the same three calls, in the reference's call order.
struct MoveRng {
    int mode;            // 0 = verification (fastrand thread-local chain), 1 = fast (Philox counter)
    WyRand tl;           // mode 0: the instance's "thread-local" generator (fastrand::Rng::new() seeds from it)
    uint64_t move_seed;  // mode 1
    uint64_t step;       // mode 1: index of the current step
    uint32_t r1_lo;      // mode 1: remainder word carried from the direction draw to the probability draw

    MoveRng(int mode_, uint64_t seed) : mode(mode_), tl(seed), move_seed(seed), step(0), r1_lo(0) {}

    // mode 1 derives all three draws of step k from ONE Philox2x32-10 block:
    //   (r0, r1) = philox(ctr = {lo32(k), hi32(k) ^ hi32(move_seed)}, key = lo32(move_seed))
    //   particle = mulhi32(r0, n);  m = r1 * 6: direction = hi32(m);  prob = 1 + mulhi32(lo32(m), 1000)
    void begin_step(uint64_t k) { step = k; }
    uint32_t particle(uint32_t n) {
        if (mode == 0) { WyRand r(tl.gen_u64()); return (uint32_t)r.usize_below(n); }
        uint32_t o[2];
        philox2x32_10((uint32_t)step, (uint32_t)(step >> 32) ^ (uint32_t)(move_seed >> 32),
                      (uint32_t)move_seed, o);
        uint64_t m = (uint64_t)o[1] * 6u;
        dir_ = (uint32_t)(m >> 32);
        r1_lo = (uint32_t)m;
        return (uint32_t)(((uint64_t)o[0] * n) >> 32);
    }
    uint32_t direction() {
        if (mode == 0) { WyRand r(tl.gen_u64()); return (uint32_t)r.usize_below(6); }
        return dir_;
    }
    uint32_t prob_1_1000() {
        if (mode == 0) { WyRand r(tl.gen_u64()); return r.u16_1_to_1000(); }
        return 1u + (uint32_t)(((uint64_t)r1_lo * 1000u) >> 32);
    }
private:
    uint32_t dir_ = 0;
};

}  // namespace oracle
