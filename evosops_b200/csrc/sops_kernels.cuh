// sops_kernels.cuh — the SOPS chain kernels (sm_100a): one warp per (genome, size, seed) instance.
//
// Replaces, per instance, the reference's  init_sops_env → simulate(false) → evaluate_fitness  path:
//   aggregation  src/SOPSCore/mod.rs:106-151, 271-290, 296-318
//   separation   src/SOPSCore/separation.rs:100-240, 649-805, 811-843
//   coating      src/SOPSCore/coating.rs:229-338, 439-555, 561-592
//   locomotion   src/SOPSCore/locomotion.rs:98-255, 260-290, 344-532, 551-654, 661-706
//
// Design:
//   * The lattice lives in shared memory as bit planes (1, 2, 4 or 8 32-bit words per row of the grid padded by two
//     cells per side), the particle list as packed u16 (row+2 | (col+2) << 8), the genome as u16 acceptance thresholds.
//   * The Markov chain is strictly sequential, but only ~0.3-15 % of proposals change the state.  The 32
//     lanes therefore evaluate the next 32 proposals SPECULATIVELY against the current lattice and keep
//     what they read — the 5x5 bit window around the source, which contains every cell adjacent to source
//     or target — in a register.  What the sequential chain would have shown lane l is that window XOR the moves of
//     the accepted lanes before l.
//   * Parallel commit rounds (resolve_batch_jacobi, ..._sep): that is a fixed point over the lanes.  The accepted moves
//     are compacted into a per-warp list, every later lane fetches each move's exact patch of its window from a small
//     look-up table, re-derives, the accepted set is voted again, changed lanes are toggled in or out, and at the fixed
//     point all accepted lanes commit at once (shared-memory atomics).  A lane that proposed a particle an earlier lane
//     moves cuts the batch (the next batch starts there).  Serial rounds (resolve_batch): the lowest accepted lane
//     commits, later lanes patch their window arithmetically and only touched lanes re-derive — fewer instructions
//     when at most one lane of a batch accepts.  Adaptive rounds (fast mode; chains with a warp scheduler to themselves):
//     a chain that rarely commits evaluates THREE batches side by side against the same lattice — the independent work
//     that fills a lone warp's idle issue slots — and resolves them in chain order by serial rounds (resolve_batch_wide);
//     every chain switches between the forms on its own commit counts (run_instance).  ROUNDS is a template parameter;
//     every form gives the sequential chain's bits.
//   * Aggregation decides through a perfect hash of the 8 neighbourhood cells into a per-instance threshold table.
//   * Both RNG modes are counter based: fast = Philox2x32-10 keyed per step; verify = the WyRand
//     double chain, whose draw count per step (2 or 3) depends on the state — resolved per batch by a
//     warp scan over the lanes' alignment-automaton transition maps; draws are reduced once to packed words and
//     prefetched one batch ahead.
//   * Hot loops are written without per-lane branches around loads: ptxas makes those divergent regions, and one costs
//     a lone warp ~60 cycles.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sops_rng.cuh"

// ROUNDS_ADAPTIVE: batches evaluated side by side while a chain is calm, and the switching thresholds (see run_instance).
// `ema` is the running mean of commits per 32-step batch x 64.
#ifndef SOPS_WIDE_N
#define SOPS_WIDE_N 3
#endif
#ifndef SOPS_WIDE_HOT1
#define SOPS_WIDE_HOT1 6            // aggregation, calm -> hot: commits in one wide iteration (32 N steps) ...
#endif
#ifndef SOPS_WIDE_HOT2
#define SOPS_WIDE_HOT2 9            // ... or in two consecutive wide iterations
#endif
#ifndef SOPS_WIDE_CALM
#define SOPS_WIDE_CALM 64           // aggregation, hot -> calm: ema below this (one commit per batch)
#endif
// The other behaviours re-derive through divergent regions when a commit touches a later proposal, so a commit of the
// side-by-side form costs them more and they leave it earlier; and their parallel rounds carry a higher fixed cost per
// batch than aggregation's, so between calm and hot they go one batch at a time through the serial rounds.
#ifndef SOPS_WIDE_HOT1_O
#define SOPS_WIDE_HOT1_O 3
#endif
#ifndef SOPS_WIDE_HOT2_O
#define SOPS_WIDE_HOT2_O 4
#endif
#ifndef SOPS_WIDE_CALM_O
#define SOPS_WIDE_CALM_O 19         // -> side by side: fewer than 0.3 commits per batch
#endif
#ifndef SOPS_WIDE_MID_O
#define SOPS_WIDE_MID_O 112         // serial <-> parallel rounds: 1.75 commits per batch
#endif
// Separation has the two single-batch phases only (three batches side by side gained 3 % on its calm chains: every one of its
// steps is "possible" and its decision is ~60 instructions); its serial rounds find a swap partner through the
// cell -> particle map the parallel rounds keep, and keep it exact
#ifndef SOPS_WIDE_MID_SEP
#define SOPS_WIDE_MID_SEP 80          // serial <-> parallel rounds: 1.25 commits per batch
#endif

namespace sops {

enum { AGG = 0, SEP = 1, COAT = 2, LOCO = 3 };
enum { MODE_VERIFY = 0, MODE_FAST = 1 };
// commit-round form of the warp kernel: all accepted lanes at once (fixed-point patch rounds) or
// one commit per round with re-derivation of the lanes whose window changed (every behaviour)
// ROUNDS_ADAPTIVE (fast mode, chains that have a warp scheduler to themselves): every chain switches, on its own running
// acceptance rate, between the parallel rounds and serial rounds over SOPS_WIDE_N batches evaluated side by side
enum { ROUNDS_PARALLEL = 0, ROUNDS_SERIAL = 1, ROUNDS_ADAPTIVE = 2 };

// --- the adaptive rounds' controller (run_instance): which form the next batch of a chain takes ----------------------
// `ema` = running mean of commits per 32-step batch in 1/64 (decay 1/8 per batch; it bottoms out at 7/64 when nothing
// commits, below every threshold).  Aggregation has no serial single-batch phase, separation no side-by-side phase.
enum { FORM_WIDE = 0, FORM_ONE_SERIAL = 1, FORM_ONE_PARALLEL = 2 };
__host__ __device__ constexpr uint32_t adapt_ema_next(uint32_t ema, uint32_t commits) { return ema - (ema >> 3) + 8u * commits; }
// side by side -> single batches: commits in one wide iteration, or in two consecutive ones
__host__ __device__ constexpr uint32_t adapt_hot1(int beh) { return beh == 0 ? SOPS_WIDE_HOT1 : SOPS_WIDE_HOT1_O; }
__host__ __device__ constexpr uint32_t adapt_hot2(int beh) { return beh == 0 ? SOPS_WIDE_HOT2 : SOPS_WIDE_HOT2_O; }
// ema below calm: side by side; below mid: single batches through the serial rounds; else the parallel rounds
__host__ __device__ constexpr uint32_t adapt_calm(int beh) { return beh == 0 ? SOPS_WIDE_CALM : beh == 1 ? 0u : SOPS_WIDE_CALM_O; }
__host__ __device__ constexpr uint32_t adapt_mid(int beh) { return beh == 0 ? 0u : beh == 1 ? SOPS_WIDE_MID_SEP : SOPS_WIDE_MID_O; }
template <int BEH>
__host__ __device__ constexpr int adapt_form(uint32_t ema) {
    if constexpr (adapt_calm(BEH) != 0u) { if (ema < adapt_calm(BEH)) return FORM_WIDE; }
    if constexpr (adapt_mid(BEH) != 0u) { if (ema < adapt_mid(BEH)) return FORM_ONE_SERIAL; }
    return FORM_ONE_PARALLEL;
}

struct Instance {                   // 40 bytes, built on the host, sorted by cost (longest first)
    uint64_t seed;                  // init seed (StdRng::seed_from_u64)
    uint64_t move_seed;
    uint64_t steps;
    uint32_t genome;                // genome index
    uint32_t slot;                  // output slot on this device
    uint32_t aux;                   // coat: element offset of this size's distance table
    uint8_t arena, layers, coat, orient;
};

struct DevResult { uint32_t i0, i1, i2, status; };

struct KParams {
    const Instance* inst;
    uint32_t n_inst;
    const uint8_t* genomes;
    uint32_t genome_len;
    uint16_t prob[16];              // gene -> per-mille acceptance threshold
    uint32_t* work_counter;
    DevResult* out;
    unsigned long long* counters;   // [0] attempts [1] possible [2] committed
    uint32_t* dump;                 // optional final-state dump (words), dump_stride words per slot
    uint32_t dump_stride;
    uint32_t dump_slot0;            // first slot of this launch (dump rows are launch-relative)
    const uint16_t* coat_dist;      // coat: per-size G*G distance tables (0xFFFF = none)
    // per-warp shared-memory layout, in 32-bit words.  The grid is padded by TWO cells on every side
    // (PG = G + 4) so the 5x5 window of any particle stays inside the plane.
    uint32_t words_per_warp;
    uint32_t off_p0, off_p1, off_p2, off_pos, off_thr, off_aux;
    uint32_t off_tab;               // AGG: hashed-window threshold table, SOPS_AGG_TAB_WORDS words.  SEP (parallel rounds): inverse map
    uint32_t inv_words;             // SEP: words of the cell -> particle map (0: not kept)
    uint32_t plane_words, pos_words;
};

#define SOPS_LIGHT_WORDS 64u        // 128 u16 beams: loco is defined for arena_layers <= 63 (u8 arithmetic upstream), 2a+1 <= 127
#define SOPS_LOCO_AUX_WORDS (SOPS_LIGHT_WORDS + 128u)   // + one u32 per beam of init scratch

// ---------------------------------------------------------------------- direction tables --------
// Directions in the reference's order (mod.rs:80-82): (-1,0) (1,0) (0,-1) (0,1) (-1,-1) (1,1).
// A 9-bit window around a cell holds rows (r-1, r, r+1) x cols (c-1, c, c+1): bit = 3*row + col.
// For a move s -> t = s + D[d]:  MID = common neighbours, BACK = nbrs(s) \ {t, MID},
// FRONT = nbrs(t) \ {s, MID}.  Table word = MID(s-window) | BACK(s-window) << 9 | FRONT(t-window) << 18.
__host__ __device__ constexpr int dir_bit(int d) {
    return d == 0 ? 1 : d == 1 ? 7 : d == 2 ? 3 : d == 3 ? 5 : d == 4 ? 0 : 8;
}
__host__ __device__ constexpr int dir_adj(int d, int k) {   // the two directions 60 degrees either side of d
    return d == 0 ? (k ? 4 : 3) : d == 4 ? (k ? 2 : 0) : d == 2 ? (k ? 1 : 4)
         : d == 1 ? (k ? 5 : 2) : d == 5 ? (k ? 3 : 1) : (k ? 0 : 5);
}
constexpr uint32_t NB_MASK = 0x1ABu;                        // the six neighbours inside the window
__host__ __device__ constexpr uint32_t dir_mid(int d) { return (1u << dir_bit(dir_adj(d, 0))) | (1u << dir_bit(dir_adj(d, 1))); }
__host__ __device__ constexpr uint32_t dir_word(int d) {
    return dir_mid(d) | ((NB_MASK & ~dir_mid(d) & ~(1u << dir_bit(d))) << 9) |
           ((NB_MASK & ~dir_mid(d ^ 1) & ~(1u << dir_bit(d ^ 1))) << 18);
}
// packed position delta of direction d on (row | col << 8):  -1, +1, -256, +256, -257, +257
__host__ __device__ constexpr int dir_dr(int d) { return d == 0 || d == 4 ? -1 : d == 1 || d == 5 ? 1 : 0; }
__host__ __device__ constexpr int dir_dc(int d) { return d == 2 || d == 4 ? -1 : d == 3 || d == 5 ? 1 : 0; }
__host__ __device__ constexpr int dir_delta(int d) { return dir_dr(d) + 256 * dir_dc(d); }
// The 5x5 window U around the source s (bit = WS*(dr+2) + (dc+2), row stride WS = 6, s at bit 14) contains both
// 3x3 windows: 3x3 bit (r, c) of the window centred at s + (or, oc) sits at U bit WS*(r+1+or) + (c+1+oc).
// (Stride 6, not 5: with it every direction's 8 neighbourhood cells have a perfect multiplicative hash, below.)
constexpr int WS = 6;
__host__ __device__ constexpr uint32_t lift9(uint32_t m9, int orow, int ocol) {
    uint32_t m = 0;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            if ((m9 >> (3 * r + c)) & 1u) m |= 1u << (WS * (r + 1 + orow) + (c + 1 + ocol));
    return m;
}
__host__ __device__ constexpr uint32_t dir_back5(int d) { return lift9((dir_word(d) >> 9) & 0x1ff, 0, 0); }
__host__ __device__ constexpr uint32_t dir_mid5(int d) { return lift9(dir_word(d) & 0x1ff, 0, 0); }
__host__ __device__ constexpr uint32_t dir_front5(int d) { return lift9(dir_word(d) >> 18, dir_dr(d), dir_dc(d)); }
__host__ __device__ constexpr uint32_t dir_tbit(int d) { return (uint32_t)(2 * WS + 2 + WS * dir_dr(d) + dir_dc(d)); }
// Aggregation decides from the occupancy of the 8 cells around (s, t) only.  ((U & dir_all5) * dir_magic) >> 24 is a
// perfect hash of those 8 bits (a bijection onto 0..255, checked at compile time), so the whole
// back/mid/front popcount -> gene -> threshold chain (mod.rs:186-232, 284) becomes ONE multiply and ONE load from a
// per-instance table [6][256] of thresholds that init_instance fills from the genome.
__host__ __device__ constexpr uint32_t dir_all5(int d) { return dir_back5(d) | dir_mid5(d) | dir_front5(d); }
__host__ __device__ constexpr uint32_t dir_magic(int d) {
    return d == 0 ? 0x20102010u : d == 1 ? 0x00801024u : d == 2 ? 0x00080200u : d == 3 ? 0x00040100u : d == 4 ? 0x40202010u : 0x00400809u;
}
// k-th subset of a mask's set bits (software pdep)
__host__ __device__ constexpr uint32_t deposit8(uint32_t k, uint32_t mask) {
    uint32_t x = 0;
    for (int b = 0; b < 8; ++b) {
        const uint32_t low = mask & (0u - mask);
        if ((k >> b) & 1u) x |= low;
        mask ^= low;
    }
    return x;
}
__host__ __device__ constexpr bool magic_is_perfect(int d) {
    uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t k = 0; k < 256; ++k) {
        const uint32_t h = (deposit8(k, dir_all5(d)) * dir_magic(d)) >> 24;
        if ((seen[h >> 5] >> (h & 31u)) & 1u) return false;
        seen[h >> 5] |= 1u << (h & 31u);
    }
    return true;
}
static_assert(magic_is_perfect(0) && magic_is_perfect(1) && magic_is_perfect(2) && magic_is_perfect(3) &&
              magic_is_perfect(4) && magic_is_perfect(5), "window hash must be a bijection for every direction");
#define SOPS_AGG_TAB_WORDS (6u * 256u / 2u)                  // u16 thresholds, [direction][hash]

__device__ __forceinline__ uint32_t lowmask(uint32_t r) { return r >= 32 ? 0xffffffffu : ((1u << r) - 1u); }

// ------------------------------------------------------------------------------ bit planes ------
// RW 32-bit words per row: 1 and 2 are specialised (PG <= 32 / 64), the generic form (4, 8) serves PG <= 128 / 256
template <int RW> struct Plane {
    static __device__ __forceinline__ uint32_t seg(const uint32_t* p, uint32_t r, uint32_t c0, uint32_t m) {
        const uint32_t* row = p + RW * r;
        const uint32_t k = c0 >> 5;
        // a 5-bit segment never crosses the end of the row, so the clamped `hi` word is only read, never used, there
        return __funnelshift_r(row[k], row[k + 1 < RW ? k + 1 : RW - 1], c0 & 31u) & m;
    }
    static __device__ __forceinline__ uint32_t bit(const uint32_t* p, uint32_t r, uint32_t c) { return (p[RW * r + (c >> 5)] >> (c & 31)) & 1u; }
    static __device__ __forceinline__ void atom_or(uint32_t* p, uint32_t r, uint32_t c) { atomicOr(&p[RW * r + (c >> 5)], 1u << (c & 31)); }
    static __device__ __forceinline__ void atom_xor(uint32_t* p, uint32_t r, uint32_t c) { atomicXor(&p[RW * r + (c >> 5)], 1u << (c & 31)); }
    static __device__ __forceinline__ const uint32_t* word(const uint32_t* p, uint32_t r, uint32_t c) { return p + RW * r + (c >> 5); }
};
template <> struct Plane<1> {       // padded grid side <= 32: one u32 per row
    static __device__ __forceinline__ uint32_t seg(const uint32_t* p, uint32_t r, uint32_t c0, uint32_t m) { return (p[r] >> c0) & m; }
    static __device__ __forceinline__ uint32_t bit(const uint32_t* p, uint32_t r, uint32_t c) { return (p[r] >> c) & 1u; }
    static __device__ __forceinline__ void atom_or(uint32_t* p, uint32_t r, uint32_t c) { atomicOr(&p[r], 1u << c); }
    static __device__ __forceinline__ void atom_xor(uint32_t* p, uint32_t r, uint32_t c) { atomicXor(&p[r], 1u << c); }
    static __device__ __forceinline__ const uint32_t* word(const uint32_t* p, uint32_t r, uint32_t c) { return p + r; }
};
template <> struct Plane<2> {       // padded grid side <= 64: one u64 per row
    static __device__ __forceinline__ uint32_t seg(const uint32_t* p, uint32_t r, uint32_t c0, uint32_t m) {
        uint64_t w = *reinterpret_cast<const uint64_t*>(p + 2 * r);
        return (uint32_t)(w >> c0) & m;
    }
    static __device__ __forceinline__ uint32_t bit(const uint32_t* p, uint32_t r, uint32_t c) { return (p[2 * r + (c >> 5)] >> (c & 31)) & 1u; }
    static __device__ __forceinline__ void atom_or(uint32_t* p, uint32_t r, uint32_t c) { atomicOr(&p[2 * r + (c >> 5)], 1u << (c & 31)); }
    static __device__ __forceinline__ void atom_xor(uint32_t* p, uint32_t r, uint32_t c) { atomicXor(&p[2 * r + (c >> 5)], 1u << (c & 31)); }
    static __device__ __forceinline__ const uint32_t* word(const uint32_t* p, uint32_t r, uint32_t c) { return p + 2 * r + (c >> 5); }
};

// SEP, parallel rounds: index of a packed cell in the cell -> particle map (u16 entries, 32 * RW per row)
template <int RW>
__device__ __forceinline__ uint32_t inv_index(uint32_t cell) { return (cell & 0xffu) * (32u * RW) + (cell >> 8); }

// 3x3 window around padded cell (r, c): bit = 3*row + col (fitness sums)
template <int RW>
__device__ __forceinline__ uint32_t win9(const uint32_t* p, uint32_t r, uint32_t c) {
    return Plane<RW>::seg(p, r - 1, c - 1, 7u) | (Plane<RW>::seg(p, r, c - 1, 7u) << 3) | (Plane<RW>::seg(p, r + 1, c - 1, 7u) << 6);
}
// 5x5 window around padded cell (r, c): bit = WS*row + col.  r, c >= 2 always (two padding cells).
template <int RW>
__device__ __forceinline__ uint32_t win25(const uint32_t* p, uint32_t r, uint32_t c) {
    const uint32_t c0 = c - 2;
    return Plane<RW>::seg(p, r - 2, c0, 31u) | (Plane<RW>::seg(p, r - 1, c0, 31u) << WS) | (Plane<RW>::seg(p, r, c0, 31u) << (2 * WS)) |
           (Plane<RW>::seg(p, r + 1, c0, 31u) << (3 * WS)) | (Plane<RW>::seg(p, r + 2, c0, 31u) << (4 * WS));
}
// XOR mask of the two flipped cells x and x + delta (a committed move) in the 5x5 window whose top-left
// cell is `org` (= centre - 0x0202).  Cells outside the window contribute 0.
__host__ __device__ __forceinline__ uint32_t win25_patch(uint32_t x, uint32_t delta, uint32_t org) {
    const uint32_t v = x - org;                              // (dr + 2) | (dc + 2) << 8 when inside
    const uint32_t w = v + delta;
    const bool vin = ((v | (v + 0x0303u)) & 0xFFFFF8F8u) == 0u;   // both fields in 0..4 (a borrow sets high bits)
    const bool win_ = ((w | (w + 0x0303u)) & 0xFFFFF8F8u) == 0u;
    // bit index WS*row + col of a valid cell: one byte-wise dot product of (row, col, 0, 0) with (WS, 1, 0, 0)
#ifdef __CUDA_ARCH__
    const uint32_t iv = __dp4a(v, (uint32_t)WS | 0x0100u, 0u), iw = __dp4a(w, (uint32_t)WS | 0x0100u, 0u);
#else
    const uint32_t iv = (v & 0xffu) * (uint32_t)WS + ((v >> 8) & 0xffu), iw = (w & 0xffu) * (uint32_t)WS + ((w >> 8) & 0xffu);
#endif
    const uint32_t mv = 1u << (iv & 31u), mw = 1u << (iw & 31u);
    return (vin ? mv : 0u) ^ (win_ ? mw : 0u);
}

// Patch look-up table for the parallel commit rounds (resolve_batch_jacobi): what a committed move s_f -> s_f + D[d]
// flips in the 5x5 window centred at ANOTHER lane's source s_l, by [d][s_f.row - s_l.row + 3][s_f.col - s_l.col + 3]
// (offsets -3..3: the target can lie inside the window when the source is one cell outside).  Bit 31: s_f == s_l (the
// same particle was proposed again: that lane is stale once f commits); bit 30: the move's target is s_l (SEP: lane l
// proposed the swap partner).  Window bits are 0..28, so the flags never reach a decision mask.
__host__ __device__ constexpr uint32_t plut_entry(int d, int r3, int c3) {
    const int r = r3 - 3, c = c3 - 3, tr = r + dir_dr(d), tc = c + dir_dc(d);
    uint32_t m = 0;
    if (r >= -2 && r <= 2 && c >= -2 && c <= 2) m ^= 1u << (WS * (r + 2) + (c + 2));
    if (tr >= -2 && tr <= 2 && tc >= -2 && tc <= 2) m ^= 1u << (WS * (tr + 2) + (tc + 2));
    if (r == 0 && c == 0) m |= 0x80000000u;
    if (tr == 0 && tc == 0) m |= 0x40000000u;
    return m;
}
#define SOPS_PLUT_WORDS (6u * 49u)
// byte offset of entry [d][r3][c3] as ONE dp4a of the packed (r3 | c3 << 8 | d << 16) with these byte weights
#define SOPS_PLUT_DP4A 0x00C4041Cu                          // 4 * (7, 1, 49, 0)

// group index tables --------------------------------------------------------------------------------
// separation.rs:205-222: (occupied c, same colour k) -> c(c+1)/2 + (c-k)
__device__ __forceinline__ uint32_t sep_idx(uint32_t c, uint32_t k) { return ((c * (c + 1)) >> 1) + (c - k); }
// coating.rs:302-320: nibble tables indexed by particles*4 + objects
__device__ __forceinline__ uint32_t coat_idx3(uint32_t p, uint32_t o) { return (uint32_t)(0x0009004806573210ULL >> ((p * 4 + o) * 4)) & 15u; }
__device__ __forceinline__ uint32_t coat_idx2(uint32_t p, uint32_t o) { return (uint32_t)(0x0000000500340210ULL >> ((p * 4 + o) * 4)) & 15u; }

struct Geo {
    uint32_t a, G, n;
    uint32_t n3;                    // sep: particles per colour
    uint32_t orient;                // loco
};

// Planes:  AGG / LOCO: p0 = particles, p1 = static not-enterable (boundary + padding)
//          COAT:       p0 = particles, p1 = static not-enterable (boundary + padding + object), p2 = object
//          SEP:        p0, p1 = colour bits, p2 = static not-enterable (boundary + padding)
struct Lat {
    uint32_t* p0; uint32_t* p1; uint32_t* p2;
    uint16_t* pos; uint16_t* thr; uint16_t* light;
    uint16_t* tab;                  // AGG: thresholds by [direction][window hash].  SEP: cell -> particle map (when kept)
    bool inv;                       // SEP: the cell -> particle map is kept (parallel / adaptive rounds)
    uint32_t plut;                  // shared-window address of the patch look-up table (per CTA)
    uint4* cmp;                     // this warp's scratch list of accepted moves (32 entries)
    const uint4* dirtab;            // per direction: back5, mid5, front5 masks, delta | tbit << 16
                                    // (AGG: all-8-cells mask, hash multiplier, table offset, delta | tbit << 16)
};
template <int BEH>
__device__ __forceinline__ uint4 dirtab_entry(int d) {
    const uint32_t w = ((uint32_t)dir_delta(d) & 0xffffu) | (dir_tbit(d) << 16);
    if (BEH == AGG) return make_uint4(dir_all5(d), dir_magic(d), (uint32_t)d << 8, w);
    return make_uint4(dir_back5(d), dir_mid5(d), dir_front5(d), w);
}

// locomotion.rs:260-290 on unpadded coordinates; light[b] = x | y << 8.  Beam id or -1 when the cell is on no beam.
// Branch-free: the beam through (x, y) is cx*x + cy*y + c0 with (2,-1,0), (-1,2,0), (1,1,-a) for orientation 1, 2, 3, and a
// cell lies on a beam iff that value is in 0..2a (for orientation 3 that is the reference's  a <= x + y <= 3a).
__device__ __forceinline__ int loco_beam(uint32_t o, int a, int x, int y) {
    const int cx = o == 1 ? 2 : o == 2 ? -1 : 1, cy = o == 1 ? -1 : o == 2 ? 2 : 1, c0 = o == 3 ? -a : 0;
    const int b = cx * x + cy * y + c0;
    return (b >= 0 && b <= 2 * a) ? b : -1;
}
// sensing against given beam fronts (light[b] = x | y << 8), packed like Prop::aux: orientation 1 compares y, the others x
__device__ __forceinline__ uint32_t loco_aux(uint32_t o, int bs, int bt, uint32_t fs, uint32_t ft, int sx, int sy, int tx, int ty) {
    const int cs = o == 1 ? sy : sx, ct = o == 1 ? ty : tx;
    const int ls = (int)(o == 1 ? fs >> 8 : fs & 0xffu), lt = (int)(o == 1 ? ft >> 8 : ft & 0xffu);
    const uint32_t cur = (bs >= 0 && cs >= ls) ? 1u : 0u;
    const uint32_t fut = (bt >= 0 && ct >= lt) ? 1u : 0u;
    const uint32_t li = (cur == fut) ? 2u : cur;             // (0,1)->0 (1,0)->1 else 2   (locomotion.rs:230-235)
    return li * 48u | ((uint32_t)(bs + 1) << 8) | ((uint32_t)(bt + 1) << 16) | (cur << 24) | (fut << 25);
}

__device__ __forceinline__ uint32_t sep_colour(const Geo& g, uint32_t idx) {   // hand-out order, separation.rs:126-146
    return 1u + (idx >= g.n3) + (idx >= 2 * g.n3);
}
// 9-bit mask of window cells whose 2-bit colour equals col (col != 0 so empty cells never match)
__device__ __forceinline__ uint32_t same_mask(uint32_t w0, uint32_t w1, uint32_t col) {
    uint32_t m0 = (col & 1) ? 0x1ffu : 0u, m1 = (col & 2) ? 0x1ffu : 0u;
    return ~((w0 ^ m0) | (w1 ^ m1)) & 0x1ffu;
}

// --- one speculative proposal ---------------------------------------------------------------------------
struct Prop {
    uint32_t idx, d, pd;            // particle, direction, probability draw (0..999): fixed for the step
    uint32_t s, t;                  // packed padded source / target
    uint32_t tt;                    // t | tbit << 16 | SEP: flipped colour bits << 28 (what an accepted lane broadcasts)
    uint32_t mB, mM, mF, tb;        // BACK / MID / FRONT masks in the 5x5 window; tb: SEP bit index of the target cell, else a
                                    // one-hot mask (target cell, or the source cell when the target is statically blocked)
                                    // (AGG: mB = mask of all 8 cells, mM = hash multiplier, mF = table offset)
    uint32_t U0;                    // 5x5 window of plane p0 around s
    uint32_t U1;                    // SEP: 5x5 window of plane p1.  COAT: object window (static)
    uint32_t aux;                   // LOCO: light_idx*48 | beam(s)+1 << 8 | beam(t)+1 << 16 | senses(s) << 24 | senses(t) << 25.  SEP: mover colour
    uint32_t kind;                  // 0 blocked, 1 move into empty, 2 swap (SEP)
    uint32_t lfront;                // LOCO: the beam fronts this proposal sensed, light[beam(s)] | light[beam(t)] << 16
    bool bstat;                     // target statically not enterable
    bool poss, eff;                 // target enterable now; would change the state
};

// 25-bit mask of window cells whose 2-bit colour equals col (col != 0 so empty cells never match)
__device__ __forceinline__ uint32_t same_mask25(uint32_t w0, uint32_t w1, uint32_t col) {
    uint32_t m0 = (col & 1) ? 0xffffffffu : 0u, m1 = (col & 2) ? 0xffffffffu : 0u;
    return ~((w0 ^ m0) | (w1 ^ m1));
}

// decision from the registers (window, static flag, draws): no lattice access, one or two threshold loads
template <int BEH>
__device__ __forceinline__ void derive(const Lat& L, const Geo& g, Prop& p) {
    if (BEH == AGG) {
        // hashed window -> threshold: one multiply, one load (table built by init_instance)
        // (integer arithmetic up to ONE final compare: chained predicates cost ~13 cycles each on the commit path,
        // and the load must not wait for the possible-ness test)
        const uint32_t thr = L.tab[p.mF + (((p.U0 & p.mB) * p.mM) >> 24)];
        // p.tb is a one-hot MASK here: the target cell — or, when the target is statically not enterable, the source
        // cell, which always holds this lane's own particle — so "possible" is one AND (mod.rs:235-251).  That
        // predicate is ready long before the threshold arrives; the compare that feeds the vote just ANDs it in.
        p.poss = (p.U0 & p.tb) == 0u;
        p.kind = p.poss ? 1u : 0u;
        p.eff = p.poss && p.pd < thr;                        // mod.rs:284-285: possible and draw below the threshold
    } else if (BEH == LOCO) {
        uint32_t gi = (__popc(p.U0 & p.mB) * 3 + __popc(p.U0 & p.mM)) * 4 + __popc(p.U0 & p.mF);
        gi += p.aux & 0xffu;
        uint32_t thr = L.thr[gi];
        asm volatile("" : "+r"(thr));                        // (keeps the count / load chain out of a divergent "if possible" region)
        p.poss = (p.U0 & p.tb) == 0u;                        // locomotion.rs:551-568 (p.tb: the one-hot mask described above)
        p.kind = p.poss ? 1u : 0u;
        p.eff = p.poss && p.pd < thr;                        // mod.rs:284-285
    } else if (BEH == COAT) {
        const uint32_t b = coat_idx3(__popc(p.U0 & p.mB), __popc(p.U1 & p.mB));
        const uint32_t m = coat_idx2(__popc(p.U0 & p.mM), __popc(p.U1 & p.mM));
        const uint32_t f = coat_idx3(__popc(p.U0 & p.mF), __popc(p.U1 & p.mF));
        uint32_t thr = L.thr[(b * 6 + m) * 10 + f];
        asm volatile("" : "+r"(thr));                        // (keeps the count / load chain out of a divergent "if possible" region)
        p.poss = (p.U0 & p.tb) == 0u;                        // coating.rs:515-530 (object cells are in bstat, hence in the mask)
        p.kind = p.poss ? 1u : 0u;
        p.eff = p.poss && p.pd < thr;
    } else {
        const uint32_t occ = p.U0 | p.U1;
        const uint32_t col = p.aux;
        const uint32_t sm = same_mask25(p.U0, p.U1, col);
        const uint32_t oB = __popc(occ & p.mB), oM = __popc(occ & p.mM), oF = __popc(occ & p.mF);
        const uint32_t b = sep_idx(oB, __popc(sm & p.mB));
        const uint32_t m = sep_idx(oM, __popc(sm & p.mM));
        const uint32_t f = sep_idx(oF, __popc(sm & p.mF));
        uint32_t thr = L.thr[(b * 6 + m) * 10 + f];
        // the partner (at t) evaluated for the flipped move t -> s (separation.rs:768-778): its BACK is our
        // FRONT, its FRONT our BACK, "same colour" relative to ITS colour.  Computed unconditionally
        // (garbage when t is empty, then unused).
        const uint32_t col2 = ((p.U0 >> (p.tb & 31u)) & 1u) | (((p.U1 >> (p.tb & 31u)) & 1u) << 1);
        const uint32_t pm = same_mask25(p.U0, p.U1, col2);
        const uint32_t b2 = sep_idx(oF, __popc(pm & p.mF));
        const uint32_t m2 = sep_idx(oM, __popc(pm & p.mM));
        const uint32_t f2 = sep_idx(oB, __popc(pm & p.mB));
        const uint32_t thr2 = L.thr[(b2 * 6 + m2) * 10 + f2];
        p.kind = p.bstat ? 0u : (col2 ? 2u : 1u);            // separation.rs:720-740
        p.poss = p.kind != 0;
        if (p.kind == 2) thr = min(thr, thr2);               // max gene == min probability (tables are non-increasing)
        // a same-colour swap draws and is accepted but nothing moves (separation.rs:287-288)
        p.eff = p.poss && p.pd < thr && !(p.kind == 2 && col2 == col);
        // a move flips col (col2 == 0), a swap col ^ col2; bit 31 says "swap" to the lanes the move is broadcast to
        p.tt = (p.tt & 0x0fffffffu) | ((col ^ col2) << 28) | (p.kind == 2 ? 0x80000000u : 0u);
    }
}

// what a proposal reads from the current lattice: position, window(s), static flag
template <int BEH, int RW>
__device__ __forceinline__ void load_proposal(const Lat& L, const Geo& g, Prop& p) {
    p.s = L.pos[p.idx];
    const uint4 dt = L.dirtab[p.d];
    p.mB = dt.x; p.mM = dt.y; p.mF = dt.z; p.tb = dt.w >> 16;   // (all but SEP: turned into a mask once bstat is known, below)
    p.t = (p.s + dt.w) & 0xffffu;
    p.tt = p.t | (dt.w & 0xffff0000u);
    const uint32_t sr = p.s & 0xff, sc = p.s >> 8, tr = p.t & 0xff, tc = p.t >> 8;
    p.U0 = win25<RW>(L.p0, sr, sc);
    p.U1 = 0;
    p.aux = 0;
    if (BEH == SEP) {
        p.U1 = win25<RW>(L.p1, sr, sc);
        p.bstat = Plane<RW>::bit(L.p2, tr, tc);
        p.aux = sep_colour(g, p.idx);
    } else {
        p.bstat = Plane<RW>::bit(L.p1, tr, tc);
        p.tb = 1u << (p.bstat ? (uint32_t)(2 * WS + 2) : (p.tb & 31u));   // AGG / COAT / LOCO: see derive
        if (BEH == COAT) p.U1 = win25<RW>(L.p2, sr, sc);
        if (BEH == LOCO) {                                   // locomotion.rs:230-235, 596-598
            const int a = (int)g.a, sx = (int)sr - 2, sy = (int)sc - 2, tx = (int)tr - 2, ty = (int)tc - 2;
            const int bs = loco_beam(g.orient, a, sx, sy), bt = loco_beam(g.orient, a, tx, ty);
            const uint32_t fs = L.light[bs < 0 ? 0 : bs], ft = L.light[bt < 0 ? 0 : bt];   // (unused when on no beam)
            p.lfront = fs | (ft << 16);
            p.aux = loco_aux(g.orient, bs, bt, fs, ft, sx, sy, tx, ty);
        }
    }
}
// full evaluation of a proposal against the current lattice: position, window, static flag, decision
template <int BEH, int RW>
__device__ __forceinline__ void evaluate(const Lat& L, const Geo& g, Prop& p) {
    load_proposal<BEH, RW>(L, g, p);
    derive<BEH>(L, g, p);
}

// cheap "is the target enterable" for the verification-mode alignment walk
template <int BEH, int RW>
__device__ __forceinline__ bool probe_possible(const Lat& L, uint32_t idx, uint32_t d) {
    const uint32_t s = L.pos[idx];
    const uint32_t t = (s + L.dirtab[d].w) & 0xffffu;
    const uint32_t tr = t & 0xff, tc = t >> 8;
    if (BEH == SEP) return !Plane<RW>::bit(L.p2, tr, tc);
    return !(Plane<RW>::bit(L.p1, tr, tc) | Plane<RW>::bit(L.p0, tr, tc));
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// predicated 16-bit shared-memory store (no divergent region)
__device__ __forceinline__ void st16_if(bool on, const void* ptr, uint32_t v) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q st.shared.u16 [%1], %2;\n\t}"
        :: "r"((uint32_t)on), "r"(smem_addr(ptr)), "h"((uint16_t)v) : "memory");
}
// one lane (on != 0) flips one bit in each of two shared-memory words
__device__ __forceinline__ void own_flip(bool on, uint32_t a0, uint32_t a1, uint32_t m0, uint32_t m1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 u, v;\n\t"
        "setp.ne.u32 q, %0, 0;\n\t"
        "@q ld.shared.b32 u, [%1];\n\t"
        "@q ld.shared.b32 v, [%2];\n\t"
        "@q xor.b32 u, u, %3;\n\t"
        "@q xor.b32 v, v, %4;\n\t"
        "@q st.shared.b32 [%1], u;\n\t"
        "@q st.shared.b32 [%2], v;\n\t}"
        :: "r"((uint32_t)on), "r"(a0), "r"(a1), "r"(m0), "r"(m1) : "memory");
}

// --- commit of lane f's accepted proposal + incremental repair of every later lane --------------------
// Lane f flips its two cells (fire-and-forget shared reductions) and rewrites its position.  Every lane then
// patches its register window with the flipped cells; a lane whose window really changed re-derives, a
// lane that proposed the moved particle itself (or, LOCO, that senses a rewritten beam) re-reads the lattice.
// NLATER: later[0 .. NLATER-1] are further proposals of every lane that come after ALL of p's in chain order (the batches
// evaluated side by side with this one, resolve_batch_wide): each is repaired like a later lane's.
template <int BEH, int RW, int NLATER>
__device__ __forceinline__ void commit_and_repair(const Lat& L, const Geo& g, uint32_t lane, uint32_t f, uint32_t sf, uint32_t ttf,
                                                  bool active, Prop& p, Prop* later) {
    const uint32_t tf = ttf & 0xffffu;
    uint32_t x = 1;                                          // plane bits flipped at both cells
    uint32_t wrote = 0;                                      // LOCO: rewritten beams (+1), 0 = none
    const uint32_t sfr = sf & 0xff, sfc = sf >> 8, tfr = tf & 0xff, tfc = tf >> 8;
    const bool mine = lane == f;
    if (BEH == SEP) {
        // colour bits flipped at both cells: a move carries the mover's colour, a swap exchanges two colours
        x = (ttf >> 28) & 3u;
        const bool swap = (ttf >> 31) != 0u;                 // warp-uniform (broadcast)
        uint32_t partner = 0xffffffffu;
        if (swap) {
            // the partner's index: the reference's participants.iter().position scan (separation.rs:302); from the
            // cell -> particle map when the kernel keeps one (adaptive rounds), else a warp-wide scan of the position list
            // (every lane scans its strided share with independent loads, one warp reduction at the end)
            if (L.inv) {
                partner = L.tab[inv_index<RW>(tf)];
                __syncwarp();                                // every lane has read the map before lane f rewrites it
            } else {
                uint32_t found = 0xffffffffu;
#pragma unroll 4
                for (uint32_t i = lane; i < g.n; i += 32) if (L.pos[i] == tf) found = i;
                partner = __reduce_min_sync(0xffffffffu, found);
            }
        }
        // the committing lane updates the two colour planes itself, from its OWN source / target (predicated loads
        // and stores: no broadcast-derived address arithmetic, no "which plane" / "same word" branches)
        {
            const uint32_t osr = p.s & 0xff, osc = p.s >> 8, otr = p.t & 0xff, otc = p.t >> 8;
            const uint32_t a_ws = smem_addr(Plane<RW>::word(L.p0, osr, osc)), a_wt = smem_addr(Plane<RW>::word(L.p0, otr, otc));
            const uint32_t plane1 = smem_addr(L.p1) - smem_addr(L.p0);
            uint32_t m_s = 1u << (osc & 31u), m_t = 1u << (otc & 31u);
            if (a_ws == a_wt) m_s = m_t = m_s ^ m_t;         // both cells in one word: the two stores write the same value
            own_flip(mine && (x & 1u), a_ws, a_wt, m_s, m_t);
            own_flip(mine && (x & 2u), a_ws + plane1, a_wt + plane1, m_s, m_t);
        }
        if (mine) L.pos[p.idx] = (uint16_t)tf;
        if (swap) L.pos[partner] = (uint16_t)sf;
        if (L.inv && mine) {                                 // keep the map exact for the batches that go through the parallel rounds
            L.tab[inv_index<RW>(tf)] = (uint16_t)p.idx;
            if (swap) L.tab[inv_index<RW>(sf)] = (uint16_t)partner;
        }
    } else {
        if (BEH == LOCO) {                                   // locomotion.rs:344-519, net effect
            const uint32_t auxf = __shfl_sync(0xffffffffu, p.aux, f);
            const uint32_t b1 = ((auxf >> 24) & 1u) ? ((auxf >> 8) & 0xff) : 0u;      // beam fronts lane f rewrites
            const uint32_t b2 = ((auxf >> 25) & 1u) ? ((auxf >> 16) & 0xff) : 0u;
            if (b1) L.light[b1 - 1] = (uint16_t)((sfr - 2) | ((sfc - 2) << 8));        // warp-uniform branches and stores
            if (b2) L.light[b2 - 1] = (uint16_t)((tfr - 2) | ((tfc - 2) << 8));
            wrote = b1 | (b2 << 8);
        }
        if (BEH != LOCO) {
            // AGG / COAT lattice update by the committing lane itself, from its OWN source / target: predicated loads and
            // stores, no "same word" branch (coating generation -13 %; locomotion measured 1-3 % slower this way and
            // keeps the warp-uniform form below)
            const uint32_t osr = p.s & 0xff, osc = p.s >> 8, otr = p.t & 0xff, otc = p.t >> 8;
            const uint32_t a_ws = smem_addr(Plane<RW>::word(L.p0, osr, osc)), a_wt = smem_addr(Plane<RW>::word(L.p0, otr, otc));
            uint32_t m_s = 1u << (osc & 31u), m_t = 1u << (otc & 31u);
            if (a_ws == a_wt) m_s = m_t = m_s ^ m_t;         // both cells in one word: the two stores write the same value
            own_flip(mine, a_ws, a_wt, m_s, m_t);
        }
    }
    // LOCO lattice update: sf and tf were broadcast, so the two lattice words, their masks and the new values are
    // warp-uniform.  Every lane performs the same read-modify-write (same address, same value: one wavefront, no
    // divergence, no atomics): the loads are issued here, the stores after the window patch below.
    uint32_t* ws = nullptr; uint32_t* wt = nullptr;
    uint32_t wa = 0, wb = 0;
    if (BEH == LOCO) {
        ws = const_cast<uint32_t*>(Plane<RW>::word(L.p0, sfr, sfc));
        wt = const_cast<uint32_t*>(Plane<RW>::word(L.p0, tfr, tfc));
        wa = *ws; wb = *wt;
    }
    // ---- repair ----
    const uint32_t dU = win25_patch(sf, tf - sf, p.s - 0x0202u);
    bool reread = (p.s == sf);                               // the moved particle itself was proposed again
    if (BEH == SEP) {
        __syncwarp();
        if (x & 1) p.U0 ^= dU;
        if (x & 2) p.U1 ^= dU;
        reread |= (p.s == tf);                               // the swap partner was proposed
    } else {
        p.U0 ^= dU;
        if (BEH == LOCO && wrote) {
            const uint32_t b1 = wrote & 0xff, b2 = wrote >> 8;
            const uint32_t bs = (p.aux >> 8) & 0xff, bt = (p.aux >> 16) & 0xff;
            reread |= (b1 && (bs == b1 || bt == b1)) || (b2 && (bs == b2 || bt == b2));
        }
        if (BEH == LOCO) {
            const uint32_t ms = 1u << (sfc & 31u), mt = 1u << (tfc & 31u);
            if (ws == wt) {                                  // warp-uniform: both cells in one word
                *ws = wa ^ ms ^ mt;
            } else {
                *ws = wa ^ ms;
                *wt = wb ^ mt;
            }
        }
        if (mine) L.pos[p.idx] = (uint16_t)tf;
        __syncwarp();
    }
    if (NLATER == 0 || BEH != AGG) {
        if (active && lane > f && (dU != 0u || reread)) {
            if (reread) evaluate<BEH, RW>(L, g, p);
            else derive<BEH>(L, g, p);
        }
    }
    bool rr[NLATER > 0 ? NLATER : 1];
    bool any_rr = reread && lane > f;
#pragma unroll
    for (int j = 0; j < NLATER; ++j) {
        Prop& q = later[j];
        const uint32_t dU2 = win25_patch(sf, tf - sf, q.s - 0x0202u);
        bool reread2 = (q.s == sf);
        if (BEH == SEP) {
            if (x & 1) q.U0 ^= dU2;
            if (x & 2) q.U1 ^= dU2;
            reread2 |= (q.s == tf);
        } else {
            q.U0 ^= dU2;
            if (BEH == LOCO && wrote) {
                const uint32_t b1 = wrote & 0xff, b2 = wrote >> 8;
                const uint32_t bs = (q.aux >> 8) & 0xff, bt = (q.aux >> 16) & 0xff;
                reread2 |= (b1 && (bs == b1 || bt == b1)) || (b2 && (bs == b2 || bt == b2));
            }
        }
        rr[j] = reread2;
        any_rr = any_rr || reread2;
        if (BEH != AGG && (dU2 != 0u || reread2)) {
            if (reread2) evaluate<BEH, RW>(L, g, q);
            else derive<BEH>(L, g, q);
        }
    }
    if (NLATER > 0 && BEH == AGG) {
        // Aggregation's decision is one multiply and one load: everything later in chain order is re-derived without a
        // branch (a divergent region per proposal costs more than the loads); only a lane that proposed the moved particle
        // again re-reads the lattice.
        {
            Prop t = p;
            derive<BEH>(L, g, t);
            if (lane > f) { p.poss = t.poss; p.kind = t.kind; p.eff = t.eff; }
        }
#pragma unroll
        for (int j = 0; j < NLATER; ++j) derive<BEH>(L, g, later[j]);
        if (any_rr) {
            if (reread && lane > f) evaluate<BEH, RW>(L, g, p);
#pragma unroll
            for (int j = 0; j < NLATER; ++j)
                if (rr[j]) evaluate<BEH, RW>(L, g, later[j]);
        }
    }
}
template <int BEH, int RW>
__device__ __forceinline__ void commit_and_repair(const Lat& L, const Geo& g, uint32_t lane, uint32_t f, uint32_t sf, uint32_t ttf,
                                                  bool active, Prop& p) {
    commit_and_repair<BEH, RW, 0>(L, g, lane, f, sf, ttf, active, p, nullptr);
}

// --- parallel commit rounds (aggregation, coating, locomotion; separation below) ----------------------------------
// The sequential chain asks, for lane l:  decision_l = g(window_l(S0) ^ XOR of the moves of the ACCEPTED lanes f < l).
// That is a fixed point of a strictly lower-triangular map over the 32 lanes, solved by iteration instead of one
// commit after the other:
//   1. every lane XORs into P the exact patch of every currently accepted earlier lane (a warp-uniform loop over the
//      accepted set; one shuffle + one table load per accepted lane, all independent of each other);
//   2. every lane re-derives from W0 ^ P; the accepted set is voted again; lanes whose decision changed are toggled
//      in (their patch is XORed in or out of the later lanes) and step 2 repeats — rarely: a patched window seldom
//      flips a decision;
//   3. all accepted lanes commit AT ONCE: shared-memory atomics on the lattice words, one position store each.
// A lane that proposed the particle an earlier accepted lane moves holds the wrong window altogether (stale): the batch
// is cut at the lowest such lane (returned bound R, as for VERIFY's alignment cuts; the caller restarts there).
// Exactness: at the fixed point every lane below R decided on exactly the window the sequential chain shows it.
// m = s_f | d_f << 16 | f << 24 (what lane f publishes), bias = 0x0303 - s_l, mykey = l << 24: the patch of f's move
// in lane l's window if f < l, else 0.  (Sentinel m = 0xffffffff: never earlier than anybody.)
__device__ __forceinline__ uint32_t plut_fetch(uint32_t m, uint32_t bias, uint32_t lut, uint32_t mykey) {
    const uint32_t v = m + bias;                             // (s_f - s_l + 3) per byte | d << 16 | f << 24 (garbage above a byte that borrows)
    const uint32_t z = (v | (v + 0x0101u)) & 0xF8F8u;        // both offsets in 0..6?
    uint32_t dU = 0;
#ifdef __CUDA_ARCH__
    const uint32_t addr = __dp4a(v, SOPS_PLUT_DP4A, lut);
    asm("{\n\t.reg .pred q, r;\n\t"
        "setp.lt.u32 r, %2, %3;\n\t"
        "setp.eq.and.u32 q, %1, 0, r;\n\t"
        "@q ld.shared.b32 %0, [%4];\n\t}"
        : "+r"(dU) : "r"(z), "r"(m), "r"(mykey), "r"(addr));
#endif
    return dU;
}
// XOR the patches of the lanes in X (warp-uniform mask) into P, for the lanes after each of them; two per trip
// (the rare re-vote path: the first pass reads the compacted list instead)
__device__ __forceinline__ void toggle_patches(uint32_t X, uint32_t desc, uint32_t bias, uint32_t lut, uint32_t mykey, uint32_t& P) {
    while (X) {
        const uint32_t f1 = (uint32_t)__ffs(X) - 1u;
        X &= X - 1u;
        const uint32_t f2 = X ? (uint32_t)__ffs(X) - 1u : 0xffffffffu;
        X &= X - 1u;
        const uint32_t m1 = __shfl_sync(0xffffffffu, desc, f1);
        uint32_t m2 = __shfl_sync(0xffffffffu, desc, f2);
        m2 = f2 == 0xffffffffu ? 0xffffffffu : m2;
        P ^= plut_fetch(m1, bias, lut, mykey) ^ plut_fetch(m2, bias, lut, mykey);
    }
}
// Locomotion adds the light table to this: an accepted move rewrites the front of the beam through its source (if the
// mover senses light there) and of the beam through its target (if it will sense light there) (locomotion.rs:344-519, net
// effect), and a later lane on one of those beams senses against the NEW front.  Writers are rare (a particle senses
// light only at the front of its beam), so every iteration walks the currently accepted writers in lane order with
// shuffles, gives each later lane the front of the latest earlier writer of its two beams, senses again and decides; a
// lane whose written beams changed counts as changed.  The commit stores the fronts in lane order.
__device__ __forceinline__ uint32_t loco_writes(uint32_t aux) {   // beams (+1) a move with this sensing rewrites: s-beam | t-beam << 8
    return (((aux >> 24) & 1u) ? ((aux >> 8) & 0xffu) : 0u) | ((((aux >> 25) & 1u) ? ((aux >> 16) & 0xffu) : 0u) << 8);
}
template <int BEH, int RW, bool VERIFY>
__device__ __forceinline__ uint32_t resolve_batch_jacobi(const Lat& L, const Geo& g, uint32_t lane, bool active, Prop& p, uint32_t& ncommit) {
    const bool any = __any_sync(0xffffffffu, p.eff);         // (both votes are issued before the branch: it waits for the first only)
    uint32_t Acur = __ballot_sync(0xffffffffu, p.eff);       // the lanes whose moves the later lanes' P holds
    if (!any) return 32;                                     // most batches of a low-acceptance chain end here
    const bool poss0 = p.poss;                               // VERIFY: the possible-ness this lane's draws were aligned with
    const uint32_t W0 = p.U0;
    const uint32_t mykey = lane << 24;
    const uint32_t desc = p.s | (p.d << 16) | mykey;
    const uint32_t bias = 0x0303u - p.s;
    // the accepted lanes' moves, compacted in lane order into the warp's scratch list (every lane writes one entry:
    // the others fill the rest with sentinels), then read back four at a time by broadcast loads
    const uint32_t below = lowmask(lane);
    const uint32_t K = (uint32_t)__popc(Acur);
    {
        const uint32_t slot = (uint32_t)__popc((p.eff ? Acur : ~Acur) & below) + (p.eff ? 0u : K);
        reinterpret_cast<uint32_t*>(L.cmp)[slot] = p.eff ? desc : 0xffffffffu;
    }
    __syncwarp();
    auto block = [&](uint32_t b) -> uint32_t {
        const uint4 m = L.cmp[b];
        return plut_fetch(m.x, bias, L.plut, mykey) ^ plut_fetch(m.y, bias, L.plut, mykey) ^
               plut_fetch(m.z, bias, L.plut, mykey) ^ plut_fetch(m.w, bias, L.plut, mykey);
    };
    uint32_t P = block(0) ^ block(1);                        // eight entries up front (sentinels past the K-th), the rest rarely
    if (K > 8) {
#pragma unroll 1
        for (uint32_t b = 2; 4 * b < K; ++b) P ^= block(b);
    }
    // LOCO: this lane's beams (+1, 0 = none), cells in the light table's coordinates
    const uint32_t bs1 = (p.aux >> 8) & 0xffu, bt1 = (p.aux >> 16) & 0xffu;
    const int sx = (int)(p.s & 0xff) - 2, sy = (int)(p.s >> 8) - 2, tx = (int)(p.t & 0xff) - 2, ty = (int)(p.t >> 8) - 2;
    const uint32_t su = (uint32_t)sx | ((uint32_t)sy << 8), tu = (uint32_t)tx | ((uint32_t)ty << 8);
    bool acc = p.eff;
    uint32_t wr = 0;                                         // LOCO: the beams this lane's move is currently taken to rewrite
    if (BEH == LOCO) wr = acc ? loco_writes(p.aux) : 0u;
    uint32_t S;
    for (;;) {
        // stale lanes proposed a particle that an earlier accepted lane moves: the batch is cut at the lowest one (the
        // flag's parity is exact there: its first earlier mover is the only one below it that is not stale itself)
        S = __ballot_sync(0xffffffffu, active && (P >> 31) != 0u);
        p.U0 = W0 ^ P;
        if (BEH == LOCO) {
            uint32_t Wb = __ballot_sync(0xffffffffu, wr != 0u);
            uint32_t fs = p.lfront & 0xffffu, ft = p.lfront >> 16;
            const uint32_t wa = wr | (su << 16);
            while (Wb) {                                     // rare: some accepted move rewrites a beam front
                const uint32_t f = (uint32_t)__ffs(Wb) - 1u;
                Wb &= Wb - 1u;
                const uint32_t w = __shfl_sync(0xffffffffu, wa, f), v2 = __shfl_sync(0xffffffffu, tu, f);
                const uint32_t b1 = w & 0xffu, b2 = (w >> 8) & 0xffu, v1 = w >> 16;
                const bool later = lane > f;
                if (later && b1 != 0u) { fs = b1 == bs1 ? v1 : fs; ft = b1 == bt1 ? v1 : ft; }
                if (later && b2 != 0u) { fs = b2 == bs1 ? v2 : fs; ft = b2 == bt1 ? v2 : ft; }
            }
            p.aux = loco_aux(g.orient, (int)bs1 - 1, (int)bt1 - 1, fs, ft, sx, sy, tx, ty);
        }
        derive<BEH>(L, g, p);
        bool changed = p.eff != acc;
        uint32_t wr_new = 0;
        if (BEH == LOCO) {
            wr_new = p.eff ? loco_writes(p.aux) : 0u;
            changed = changed || wr_new != wr;
        }
        if (!__any_sync(0xffffffffu, changed)) break;        // fixed point: every lane decided on the window it is shown
        const uint32_t X = __ballot_sync(0xffffffffu, p.eff != acc);
        toggle_patches(X, desc, bias, L.plut, mykey, P);
        Acur ^= X;
        acc = p.eff;
        wr = wr_new;
    }
    if (VERIFY) S |= __ballot_sync(0xffffffffu, active && p.poss != poss0);   // a flipped possible-ness shifts every later draw: cut there too
    const uint32_t keep = ~S & (S - 1u);                     // the lanes below the lowest cut lane (all of them when S == 0)
    const uint32_t C = Acur & keep;
    const bool mine = (C >> lane) & 1u;
    {
        // every lane issues the two atomics: no divergent region.  A lane that does not commit XORs zero into its OWN
        // word of the warp's scratch list, so the idle lanes never pile up on the few lattice rows (bank conflicts)
        uint32_t* idle = reinterpret_cast<uint32_t*>(L.cmp) + lane;
        uint32_t* ws = mine ? const_cast<uint32_t*>(Plane<RW>::word(L.p0, p.s & 0xff, p.s >> 8)) : idle;
        uint32_t* wt = mine ? const_cast<uint32_t*>(Plane<RW>::word(L.p0, p.t & 0xff, p.t >> 8)) : idle;
        atomicXor(ws, mine ? 1u << ((p.s >> 8) & 31u) : 0u);
        atomicXor(wt, mine ? 1u << ((p.t >> 8) & 31u) : 0u);
        if (mine) L.pos[p.idx] = (uint16_t)p.t;
    }
    if (BEH == LOCO) {                                       // beam fronts, in lane order: the last writer of a beam wins
        uint32_t Wm = __ballot_sync(0xffffffffu, mine && wr != 0u);
        while (Wm) {
            const uint32_t f = (uint32_t)__ffs(Wm) - 1u;
            Wm &= Wm - 1u;
            if (lane == f) {
                if (wr & 0xffu) L.light[(wr & 0xffu) - 1u] = (uint16_t)su;
                if (wr >> 8) L.light[(wr >> 8) - 1u] = (uint16_t)tu;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    ncommit += (uint32_t)__popc(C);
    return (uint32_t)__popc(keep);
}

// --- parallel commit rounds, separation ----------------------------------------------------------------------------
// Same fixed point as above with three differences.  (1) What a lane's move flips depends on its own decision: a move
// carries the mover's colour, a swap exchanges two colours (x = colour bits flipped at both cells) — so a lane's
// published descriptor can change from one iteration to the next even while it stays accepted, and every iteration
// rebuilds the compacted list and the patches from scratch (iterations past the first are rare).  (2) A swap moves a
// SECOND particle: a lane that proposed the swap partner is stale too (table bit 30, counted for swaps only), and a
// swapping lane whose target cell an earlier accepted lane changes cannot know its partner's identity from the S0
// inverse map — both cut the batch like the mover-stale lanes do.  Flags are OR-accumulated here (two swaps can hit the
// same lane).  (3) The partner's index comes from an inverse map cell -> particle kept beside the colour planes
// (the reference scans `participants` for it, separation.rs:302).
template <int RW, bool VERIFY>
__device__ __forceinline__ uint32_t resolve_batch_jacobi_sep(const Lat& L, const Geo& g, uint32_t lane, bool active, Prop& p, uint32_t& ncommit) {
    if (!__any_sync(0xffffffffu, p.eff)) return 32;
    const uint32_t W00 = p.U0, W01 = p.U1;
    const uint32_t mykey = lane << 27;
    const uint32_t base_desc = p.s | (p.d << 16) | mykey;
    const uint32_t bias = 0x0303u - p.s;
    const uint32_t below = lowmask(lane);
    const uint32_t tmask = 1u << (p.tb & 31u);               // the target cell in this lane's window
    uint32_t A, S;
    for (;;) {
        A = __ballot_sync(0xffffffffu, p.eff);
        const uint32_t K = (uint32_t)__popc(A);
        {   // descriptor of the current decision: flipped colour bits at 24..25, "swap" at 26
            const uint32_t desc = base_desc | (((p.tt >> 28) & 3u) << 24) | ((p.tt >> 31) << 26);
            const uint32_t slot = (uint32_t)__popc((p.eff ? A : ~A) & below) + (p.eff ? 0u : K);
            reinterpret_cast<uint32_t*>(L.cmp)[slot] = p.eff ? desc : 0xffffffffu;
        }
        __syncwarp();
        uint32_t P0 = 0, P1 = 0;
        S = 0;
        auto block = [&](uint32_t b) {
            const uint4 m4 = L.cmp[b];
            const uint32_t mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t m = mm[i];
                uint32_t dU = plut_fetch(m, bias, L.plut, mykey);
                dU &= (m & (1u << 26)) ? 0xffffffffu : 0xbfffffffu;      // the partner flag counts for swaps only
                P0 ^= (m & (1u << 24)) ? dU : 0u;
                P1 ^= (m & (1u << 25)) ? dU : 0u;
                S |= dU;
            }
        };
        block(0);                                            // (sentinels past the K-th entry)
        if (K > 4) {
            block(1);
#pragma unroll 1
            for (uint32_t b = 2; 4 * b < K; ++b) block(b);
        }
        p.U0 = W00 ^ (P0 & 0x1fffffffu);
        p.U1 = W01 ^ (P1 & 0x1fffffffu);
        const bool eff0 = p.eff;
        const uint32_t tt0 = p.tt;
        derive<SEP>(L, g, p);
        const bool changed = (p.eff != eff0) || (p.eff && ((p.tt ^ tt0) >> 28) != 0u);
        __syncwarp();                                        // the list is rewritten by the next iteration
        if (!__any_sync(0xffffffffu, changed)) break;
    }
    const bool cut = (S >> 30) != 0u || ((S & tmask) != 0u && p.kind == 2u);
    const uint32_t Sb = __ballot_sync(0xffffffffu, active && cut);
    const uint32_t keep = ~Sb & (Sb - 1u);
    const uint32_t C = A & keep;
    {
        // every lane issues the four atomics (zero masks for the lanes and planes that do not flip): no divergent region
        const bool mine = (C >> lane) & 1u;
        const uint32_t x = mine ? (p.tt >> 28) & 3u : 0u;
        const bool swap = mine && (p.tt >> 31) != 0u;
        // (a lane / plane that does not flip XORs zero into the lane's own word of the scratch list: no pile-up on the lattice rows)
        uint32_t* idle = reinterpret_cast<uint32_t*>(L.cmp) + lane;
        uint32_t* ws = const_cast<uint32_t*>(Plane<RW>::word(L.p0, p.s & 0xff, p.s >> 8));
        uint32_t* wt = const_cast<uint32_t*>(Plane<RW>::word(L.p0, p.t & 0xff, p.t >> 8));
        const uint32_t plane1 = (uint32_t)(L.p1 - L.p0);
        const uint32_t ms_ = 1u << ((p.s >> 8) & 31u), mt_ = 1u << ((p.t >> 8) & 31u);
        atomicXor((x & 1u) ? ws : idle, (x & 1u) ? ms_ : 0u);
        atomicXor((x & 1u) ? wt : idle, (x & 1u) ? mt_ : 0u);
        atomicXor((x & 2u) ? ws + plane1 : idle, (x & 2u) ? ms_ : 0u);
        atomicXor((x & 2u) ? wt + plane1 : idle, (x & 2u) ? mt_ : 0u);
        const uint32_t is = inv_index<RW>(p.s), it = inv_index<RW>(p.t);
        const uint32_t partner = L.tab[it];                  // (garbage when the target is empty: used by swaps only)
        __syncwarp();                                        // every lane has read the map before any lane rewrites it
        st16_if(swap, L.pos + (swap ? partner : 0u), p.s);
        st16_if(swap, L.tab + is, partner);
        st16_if(mine, L.pos + p.idx, p.t);
        st16_if(mine, L.tab + it, p.idx);
    }
    __syncwarp();
    ncommit += (uint32_t)__popc(C);
    return (uint32_t)__popc(keep);
}

// --- commit rounds of one speculative batch ---------------------------------------------------------------
// On entry every active lane holds a proposal evaluated against the current lattice; lane order = chain
// order.  VERIFY: a repair that flips a lane's possible-ness changes how many draws that step consumes, so
// the batch is cut at that lane (returned bound R; lanes >= R are not retired).  ncommit is warp-uniform.
template <int BEH, int RW, bool VERIFY, int ROUNDS>
__device__ __forceinline__ uint32_t resolve_batch(const Lat& L, const Geo& g, uint32_t lane, bool active, Prop& p, uint32_t& ncommit) {
    if (ROUNDS != ROUNDS_SERIAL && BEH == SEP) return resolve_batch_jacobi_sep<RW, VERIFY>(L, g, lane, active, p, ncommit);
    if (ROUNDS != ROUNDS_SERIAL) return resolve_batch_jacobi<BEH, RW, VERIFY>(L, g, lane, active, p, ncommit);
    uint32_t R = 32;
    for (;;) {
        uint32_t A = __ballot_sync(0xffffffffu, p.eff);
        if (VERIFY) A &= lowmask(R);
        if (!A) break;
        const uint32_t f = __ffs(A) - 1;
        const uint32_t sf = __shfl_sync(0xffffffffu, p.s, f);
        const uint32_t ttf = __shfl_sync(0xffffffffu, p.tt, f);
        const bool was = p.poss;
        commit_and_repair<BEH, RW>(L, g, lane, f, sf, ttf, active, p);
        p.eff = p.eff && lane != f;
        ++ncommit;
        if (VERIFY) {
            const uint32_t X = __ballot_sync(0xffffffffu, active && lane > f && p.poss != was) & lowmask(R);
            if (X) R = (uint32_t)__ffs(X) - 1;
        }
    }
    return R;
}

// Serial rounds over N batches evaluated side by side against the same lattice (fast mode, full batches): chain order is
// batch 0's lanes, then batch 1's, ...  A commit of batch i repairs the proposals of the batches after it too (every lane of
// those is later).  N times the independent work per warp instruction slot for a chain that rarely commits.
template <int BEH, int RW, int N, int I = 0>
__device__ __forceinline__ void resolve_batch_wide(const Lat& L, const Geo& g, uint32_t lane, Prop* b, uint32_t& ncommit) {
    if constexpr (I < N) {
        for (;;) {
            const uint32_t A = __ballot_sync(0xffffffffu, b[I].eff);
            if (!A) break;
            const uint32_t f = __ffs(A) - 1;
            const uint32_t sf = __shfl_sync(0xffffffffu, b[I].s, f);
            const uint32_t ttf = __shfl_sync(0xffffffffu, b[I].tt, f);
            commit_and_repair<BEH, RW, N - 1 - I>(L, g, lane, f, sf, ttf, true, b[I], b + I + 1);
            b[I].eff = b[I].eff && lane != f;
            ++ncommit;
        }
        resolve_batch_wide<BEH, RW, N, I + 1>(L, g, lane, b, ncommit);
    }
}

// --- initial configuration ---------------------------------------------------------------------------
template <int BEH, int RW>
__device__ __forceinline__ uint32_t init_instance(const KParams& P, const Lat& L, const Geo& g, const Instance& in, uint32_t lane) {
    const uint32_t G = g.G, a = g.a, PG = G + 4;
    uint32_t* pstat = (BEH == SEP) ? L.p2 : L.p1;
    for (uint32_t w = lane; w < P.plane_words; w += 32) {
        L.p0[w] = 0; L.p1[w] = 0;
        if (BEH == SEP || BEH == COAT) L.p2[w] = 0;
    }
    __syncwarp();
    // static "not enterable" mask: everything except arena cells |i - j| <= a (mod.rs:115-122)
    for (uint32_t r = lane; r < PG; r += 32) {
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            uint32_t m = 0xffffffffu;
            if (r >= 2 && r <= G + 1) {
                int lo = (int)r - (int)a; if (lo < 2) lo = 2;
                int hi = (int)r + (int)a; if (hi > (int)G + 1) hi = (int)G + 1;
                lo -= 32 * w; hi -= 32 * w;
                if (lo < 0) lo = 0;
                if (hi > 31) hi = 31;
                if (lo <= hi) {
                    uint32_t v = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
                    m = ~v;
                }
            }
            pstat[r * RW + w] = m;
        }
    }
    __syncwarp();
    if (BEH == COAT) {                                       // object hexagon, coating.rs:261-284
        uint32_t o = in.layers, c = in.coat, side = 2 * o + 1, base = a - o + c + 2;   // +2: padded
        for (uint32_t k = lane; k < side * side; k += 32) {
            uint32_t u = k / side, v = k % side;
            int dd = (int)u - (int)v;
            if (dd <= (int)o && dd >= -(int)o) {
                Plane<RW>::atom_or(L.p2, base + u, base + v);
                Plane<RW>::atom_or(L.p1, base + u, base + v);
            }
        }
        __syncwarp();
    }
    // seeded rejection placement (mod.rs:125-137): lane l expands ChaCha12 block (base + l) = 4 candidate
    // cells; candidates are then consumed 32 at a time in stream order.
    uint32_t key[8];
    stdrng_key(in.seed, key);
    const uint64_t zone = ~(uint64_t)0 - ((0 - (uint64_t)G) % G);
    uint32_t placed = 0, status = 0;
    for (uint64_t blk = 0; placed < g.n; blk += 32) {
        uint32_t w[16];
        chacha12_block(key, blk + lane, w);
        uint32_t cand[4];
        bool rej = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t vi = (uint64_t)w[4 * k] | ((uint64_t)w[4 * k + 1] << 32);
            uint64_t vj = (uint64_t)w[4 * k + 2] | ((uint64_t)w[4 * k + 3] << 32);
            rej |= (vi * G > zone) | (vj * G > zone);
            uint32_t i = (uint32_t)__umul64hi(vi, (uint64_t)G), j = (uint32_t)__umul64hi(vj, (uint64_t)G);
            cand[k] = (i + 2) | ((j + 2) << 8);
        }
        if (__any_sync(0xffffffffu, rej)) status = 1;        // Uniform zone rejection (p ~ G/2^64): not replayed
        uint32_t c01 = cand[0] | (cand[1] << 16), c23 = cand[2] | (cand[3] << 16);
        for (uint32_t sub = 0; sub < 4 && placed < g.n; ++sub) {
            uint32_t src = sub * 8 + (lane >> 2), slot = lane & 3;
            uint32_t a01 = __shfl_sync(0xffffffffu, c01, src), a23 = __shfl_sync(0xffffffffu, c23, src);
            uint32_t cell = (((slot & 2) ? a23 : a01) >> ((slot & 1) * 16)) & 0xffffu;
            uint32_t pr = cell & 0xff, pc = cell >> 8;
            bool free_ = !(Plane<RW>::bit(L.p0, pr, pc) | Plane<RW>::bit(L.p1, pr, pc));
            if (BEH == SEP) free_ = free_ && !Plane<RW>::bit(L.p2, pr, pc);
            uint32_t same = __match_any_sync(0xffffffffu, cell);
            bool acc = free_ && (lane == (uint32_t)(__ffs(same) - 1));
            uint32_t am = __ballot_sync(0xffffffffu, acc);
            uint32_t ord = placed + __popc(am & lowmask(lane));
            if (acc && ord < g.n) {
                L.pos[ord] = (uint16_t)cell;
                if (BEH == SEP) {
                    if (P.inv_words) L.tab[inv_index<RW>(cell)] = (uint16_t)ord;   // parallel rounds: cell -> particle
                    uint32_t col = sep_colour(g, ord);
                    if (col & 1) Plane<RW>::atom_or(L.p0, pr, pc);
                    if (col & 2) Plane<RW>::atom_or(L.p1, pr, pc);
                } else {
                    Plane<RW>::atom_or(L.p0, pr, pc);
                }
            }
            placed += __popc(am);
            __syncwarp();
        }
    }
    if (BEH == LOCO) {
        // light table: zig-zag initial fronts (locomotion.rs:116-168) raised to the front-most
        // particle of each beam during placement (199-225).  The sequential "replace if >=" rule is
        // order independent (a running maximum), so it is rebuilt here with shared-memory atomicMax.
        uint32_t* tmp = reinterpret_cast<uint32_t*>(L.light) + SOPS_LIGHT_WORDS;   // init scratch behind the light table
        const uint32_t o = g.orient, nb = 2 * a + 1;
        for (uint32_t b = lane; b < nb; b += 32) {
            uint32_t x, y;
            if (o == 3) { x = (b + 1) / 2; y = a + b / 2; }
            else { x = (b + 1) / 2; y = b & 1; }
            L.light[b] = (uint16_t)(x | (y << 8));
            tmp[b] = 0;
        }
        __syncwarp();
        for (uint32_t k = lane; k < g.n; k += 32) {
            uint32_t cell = L.pos[k];
            int x = (int)(cell & 0xff) - 2, y = (int)(cell >> 8) - 2;
            int b; uint32_t val, keyv;
            if (o == 1) { b = 2 * x - y; val = (uint32_t)y; keyv = ((uint32_t)y << 8) | (uint32_t)x; }
            else if (o == 2) { b = 2 * y - x; val = (uint32_t)x; keyv = ((uint32_t)x << 8) | (uint32_t)y; }
            else { b = x + y - (int)a; val = (uint32_t)x; keyv = ((uint32_t)x << 8) | (uint32_t)y; }
            bool ok = (o == 3) ? (x + y >= (int)a && x + y < 3 * (int)a) : (b >= 0 && b <= 2 * (int)a);
            if (ok) {
                uint32_t init = L.light[b];
                uint32_t iv = (o == 1) ? (init >> 8) : (init & 0xff);
                if (val >= iv) atomicMax(&tmp[b], keyv + 1);
            }
        }
        __syncwarp();
        for (uint32_t b = lane; b < nb; b += 32) {
            uint32_t kv = tmp[b];
            if (kv) {
                kv -= 1;
                uint32_t hi = kv >> 8, lo = kv & 0xff;
                uint32_t x = (o == 1) ? lo : hi, y = (o == 1) ? hi : lo;
                L.light[b] = (uint16_t)(x | (y << 8));
            }
        }
        __syncwarp();
    }
    // genome -> acceptance thresholds (gene_probability(), mod.rs:85-87)
    const uint8_t* gen = P.genomes + (size_t)in.genome * P.genome_len;
    for (uint32_t k = lane; k < P.genome_len; k += 32) L.thr[k] = P.prob[gen[k] & 15];
    __syncwarp();
    if (BEH == AGG) {
        // threshold of every occupancy pattern of the 8 cells around (s, t), per direction, stored at the pattern's
        // hash: gene index [back][mid][front] as in mod.rs:186-232
        for (uint32_t k = lane; k < 6u * 256u; k += 32) {
            const uint32_t d = k >> 8;
            const uint4 dt = L.dirtab[d];
            const uint32_t mB = d == 0 ? dir_back5(0) : d == 1 ? dir_back5(1) : d == 2 ? dir_back5(2) : d == 3 ? dir_back5(3) : d == 4 ? dir_back5(4) : dir_back5(5);
            const uint32_t mM = d == 0 ? dir_mid5(0) : d == 1 ? dir_mid5(1) : d == 2 ? dir_mid5(2) : d == 3 ? dir_mid5(3) : d == 4 ? dir_mid5(4) : dir_mid5(5);
            const uint32_t x = deposit8(k & 255u, dt.x);
            const uint32_t nb = __popc(x & mB), nm = __popc(x & mM), nf = __popc(x & ~(mB | mM));
            L.tab[dt.z + ((x * dt.y) >> 24)] = L.thr[(nb * 3 + nm) * 4 + nf];
        }
        __syncwarp();
    }
    return status;
}

// --- fitness integers ---------------------------------------------------------------------------------
template <int BEH, int RW>
__device__ __forceinline__ void fitness(const KParams& P, const Lat& L, const Geo& g, const Instance& in, uint32_t lane,
                                        uint32_t& i0, uint32_t& i1) {
    uint32_t s0 = 0, s1 = 0;
    for (uint32_t k = lane; k < g.n; k += 32) {
        uint32_t cell = L.pos[k], r = cell & 0xff, c = cell >> 8;
        if (BEH == AGG || BEH == LOCO) {
            s0 += __popc(win9<RW>(L.p0, r, c) & NB_MASK);                       // mod.rs:166-178
            if (BEH == LOCO) {
                int x = (int)r - 2, y = (int)c - 2;
                s1 += g.orient == 1 ? (uint32_t)y : g.orient == 2 ? (uint32_t)x : (uint32_t)(x - y + (int)g.a);
            }
        } else if (BEH == SEP) {
            uint32_t w0 = win9<RW>(L.p0, r, c), w1 = win9<RW>(L.p1, r, c);
            s0 += __popc((w0 | w1) & NB_MASK);                                  // separation.rs:255-275
            s1 += __popc(same_mask(w0, w1, sep_colour(g, k)) & NB_MASK);
        } else {
            uint32_t dv = P.coat_dist[in.aux + (r - 2) * g.G + (c - 2)];        // coating.rs:561-574
            if (dv != 0xFFFFu) s1 += dv;
        }
    }
    i0 = __reduce_add_sync(0xffffffffu, s0);
    i1 = __reduce_add_sync(0xffffffffu, s1);
}

// --- one instance, start to finish ------------------------------------------------------------------
template <int BEH, int RW, int MODE, int ROUNDS>
__device__ __forceinline__ void run_instance(const KParams& P, const Lat& L, const Instance& in, uint32_t lane) {
    Geo g;
    g.a = in.arena; g.G = 2 * g.a + 1; g.orient = in.orient;
    {
        uint32_t p = in.layers, t = p + in.coat;
        g.n = BEH == AGG ? 3 * p * (p + 1) + 1 : BEH == COAT ? 3 * t * (t + 1) - 3 * p * (p + 1) : 3 * p * (p + 1);
    }
    g.n3 = g.n / 3;
    uint32_t status = init_instance<BEH, RW>(P, L, g, in, lane);

    const uint64_t steps = in.steps, ms = in.move_seed;
    const uint32_t n = g.n;
    uint64_t nposs = 0;                                      // per lane
    uint32_t ncommit = 0;                                    // warp-uniform

    if (MODE == MODE_FAST) {
        // Every batch retires 32 consecutive steps; the random numbers of batch b+1 do not depend on the
        // lattice, so they are drawn while batch b's shared-memory loads are in flight.
        const uint32_t key = (uint32_t)ms, khi = (uint32_t)(ms >> 32);
        // one step's draws packed in a word: particle | direction << 16 | probability draw << 19
        auto draw = [&](uint64_t st) -> uint32_t {
            uint32_t r0, r1;
            philox2x32_10((uint32_t)st, (uint32_t)(st >> 32) ^ khi, key, r0, r1);
            const uint64_t m6 = (uint64_t)r1 * 6u;
            return __umulhi(r0, n) | ((uint32_t)(m6 >> 32) << 16) | (__umulhi((uint32_t)m6, 1000u) << 19);
        };
        auto unpack = [&](uint32_t w, Prop& q) { q.idx = w & 0xffffu; q.d = (w >> 16) & 7u; q.pd = w >> 19; };
        // A batch normally retires its 32 steps; the parallel commit rounds may cut it at a lane R < 32 (a lane that
        // proposed a particle an earlier lane of the batch moved): the next batch then starts at that step, and the
        // draws move down by R lanes (this batch's upper lanes, the prefetched batch's lower ones).
        auto realign = [&](uint32_t curw, uint32_t nxtw, uint32_t R) -> uint32_t {
            const uint32_t src = lane + R;
            const uint32_t lo = __shfl_sync(0xffffffffu, curw, src), hi = __shfl_sync(0xffffffffu, nxtw, src);
            return src < 32 ? lo : hi;
        };
        Prop cur;
        uint32_t curw = draw((uint64_t)lane);
        uint64_t k = 0;                                      // steps retired
        // ROUNDS_ADAPTIVE: a chain that rarely commits gains nothing from the parallel rounds, and a lone warp leaves most
        // of its scheduler's issue slots empty waiting on one batch's loads.  While the chain is calm, N batches (32 N steps)
        // are evaluated side by side against the same lattice — N times the independent work — and resolved in chain order
        // by serial rounds (resolve_batch_wide: a commit repairs every later proposal of all N batches).  The chain
        // switches on its own commit counts: calm -> hot when a wide iteration commits >= SOPS_WIDE_HOT1 times (or two in
        // a row >= SOPS_WIDE_HOT2), hot -> calm when the running mean of commits per batch falls below SOPS_WIDE_CALM / 64.
        // Either form retires the sequential chain's steps, so the bits do not depend on the switching.
        constexpr bool ADAPT = ROUNDS == ROUNDS_ADAPTIVE;
        constexpr int N = (ADAPT && BEH != SEP) ? SOPS_WIDE_N : 1;      // (separation: no side-by-side phase)
        // `room` steps must remain after the batch for the caller's loop to go on: for an uncut batch (the usual case) that test
        // is made before the batch, off the path from the commit round to the loop branch
        auto narrow_batch = [&](uint32_t room = 32) -> bool {
            const bool more_if_full = k + 32 + room <= steps;
            const uint32_t nxtw = draw(k + 32 + lane);
            unpack(curw, cur);
            evaluate<BEH, RW>(L, g, cur);
            const uint32_t R = resolve_batch<BEH, RW, false, ROUNDS>(L, g, lane, true, cur, ncommit);
            nposs += (cur.poss && lane < R) ? 1u : 0u;
            if (R == 32) { curw = nxtw; k += 32; return more_if_full; }
            curw = realign(curw, nxtw, R);
            k += R;
            return k + room <= steps;
        };
        if (ADAPT) {
            constexpr int WIDE = FORM_WIDE, ONE_SERIAL = FORM_ONE_SERIAL, ONE_PARALLEL = FORM_ONE_PARALLEL;
            constexpr uint32_t HOT1 = adapt_hot1(BEH), HOT2 = adapt_hot2(BEH), CALM = adapt_calm(BEH), MID = adapt_mid(BEH);
            auto form_of = [&](uint32_t e) -> int { return adapt_form<BEH>(e); };
            int form = ONE_PARALLEL, form_next = ONE_PARALLEL;
            uint32_t ema = 2u * (MID > CALM ? MID : CALM);
            while (k + 32 * N <= steps) {                    // (the chain's last steps go through the plain loop below)
                if (form == WIDE) {
                    Prop b[N];
                    uint32_t w[N];
                    w[0] = curw;
#pragma unroll
                    for (int i = 1; i < N; ++i) w[i] = draw(k + 32 * i + lane);
                    uint32_t prev = 0;
                    do {
                        uint32_t nw[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) nw[i] = draw(k + 32 * (N + i) + lane);
#pragma unroll
                        for (int i = 0; i < N; ++i) unpack(w[i], b[i]);
#pragma unroll
                        for (int i = 0; i < N; ++i) evaluate<BEH, RW>(L, g, b[i]);
                        bool anyeff = false;
#pragma unroll
                        for (int i = 0; i < N; ++i) anyeff = anyeff || b[i].eff;
                        const uint32_t c0 = ncommit;
                        if (__any_sync(0xffffffffu, anyeff)) resolve_batch_wide<BEH, RW, N>(L, g, lane, b, ncommit);
#pragma unroll
                        for (int i = 0; i < N; ++i) { nposs += b[i].poss ? 1u : 0u; w[i] = nw[i]; }
                        k += 32 * N;
                        const uint32_t c = ncommit - c0;
                        if (c >= HOT1 || c + prev >= HOT2) form = BEH == AGG ? ONE_PARALLEL : ONE_SERIAL;
                        prev = c;
                    } while (form == WIDE && k + 32 * N <= steps);
                    curw = w[0];
                    form_next = form;
                    ema = 2u * CALM;
                } else if (BEH != AGG && form == ONE_SERIAL) {
                    do {
                        const uint32_t c0 = ncommit;
                        const uint32_t nxtw = draw(k + 32 + lane);
                        unpack(curw, cur);
                        evaluate<BEH, RW>(L, g, cur);
                        resolve_batch<BEH, RW, false, ROUNDS_SERIAL>(L, g, lane, true, cur, ncommit);
                        nposs += cur.poss ? 1u : 0u;
                        curw = nxtw;
                        k += 32;
                        form = form_next;
                        ema = adapt_ema_next(ema, ncommit - c0);
                        form_next = form_of(ema);
                    } while (form == ONE_SERIAL && k + 32 * N <= steps);
                } else {
                    bool more;
                    do {
                        const uint32_t c0 = ncommit;
                        more = narrow_batch(32 * N);
                        // (the decision lags one batch: the loop branch never waits for this batch's commit count)
                        form = form_next;
                        ema = adapt_ema_next(ema, ncommit - c0);
                        form_next = form_of(ema);
                    } while (form == ONE_PARALLEL && more);
                }
            }
        }
        if (k + 32 <= steps) while (narrow_batch()) {}
        while (k < steps) {
            const uint32_t tail = (uint32_t)(steps - k);     // 1..31
            const bool active = lane < tail;
            unpack(curw, cur);
            if (!active) cur.pd = 0xffffffffu;               // never accepted, also when a commit round re-derives
            evaluate<BEH, RW>(L, g, cur);
            if (!active) { cur.eff = false; cur.poss = false; }
            uint32_t R = resolve_batch<BEH, RW, false, ROUNDS>(L, g, lane, active, cur, ncommit);
            if (R > tail) R = tail;
            nposs += (active && cur.poss && lane < R) ? 1u : 0u;
            curw = realign(curw, curw, R);                   // (lanes past the end of the chain get recycled draws: inactive)
            k += R;
        }
    } else {
        // Verification mode: the number of draws a step consumes (2 if blocked, 3 if possible) decides
        // where the next step starts in the WyRand chain, so lanes probe both alignments and a
        // warp scan over the "possible" ballots picks the real ones.  A commit that flips a later
        // lane's possible-ness invalidates that alignment: the batch is cut there.
        //
        // Lane l owns the draws cbase + 2l + 1 (even slot) and cbase + 2l + 2 (odd slot); steps start in lanes 0..29
        // (a step starting in lane 29's odd slot runs through lane 30), so a batch consumes 60..62 draws.  Every raw
        // draw is reduced once to a packed word (as particle | as direction << 16 | as probability << 19; bit 31: one of
        // fastrand's rejection tests fired, p ~ 2e-7) — whichever role the alignment gives it later — and the words of
        // the NEXT batch are computed while this batch's lattice loads are in flight: they start at cbase + 60, and the
        // 0..2 draws the batch turns out to consume beyond that are a shift by at most one lane.
        auto word = [&](uint64_t c) -> uint32_t {
            const uint64_t r = wy_draw(ms, c);
            const uint64_t ln = r * (uint64_t)n, l6 = r * 6ull;
            const uint64_t m1000 = (uint64_t)(uint32_t)r * 1000u;
            const bool rare = ln < (uint64_t)n || l6 < 6ull || (uint32_t)m1000 < 1000u;
            return (uint32_t)__umul64hi(r, (uint64_t)n) | ((uint32_t)__umul64hi(r, 6ull) << 16) | ((uint32_t)(m1000 >> 32) << 19) |
                   (rare ? 0x80000000u : 0u);
        };
        uint64_t k = 0, cbase = 0;                           // steps retired, thread-local draws consumed
        uint32_t wE = word(2 * lane + 1), wO = word(2 * lane + 2);
        while (k < steps) {
            const uint64_t rem = steps - k;
            const uint64_t kE = cbase + 2 * lane + 1, kO = kE + 1;
            const uint32_t nE = __shfl_down_sync(0xffffffffu, wE, 1), nO = __shfl_down_sync(0xffffffffu, wO, 1);
            // a step starting in the even slot uses (E, O, next E); in the odd slot (O, next E, next O)
            uint32_t idxE = wE & 0xffffu, dE = (wO >> 16) & 7u, idxO = wO & 0xffffu, dO = (nE >> 16) & 7u;
            uint32_t pdE = (nE >> 19) & 0x3ffu, pdO = (nO >> 19) & 0x3ffu;
            if (__any_sync(0xffffffffu, ((wE | wO) >> 31) != 0u)) {
                // a rejection test fired somewhere in the batch: the exact reductions with fastrand's redraw loops
                const uint64_t rE = wy_draw(ms, kE), rO = wy_draw(ms, kO);
                const uint64_t rnE = __shfl_down_sync(0xffffffffu, rE, 1);
                const uint32_t rnO = __shfl_down_sync(0xffffffffu, (uint32_t)rO, 1);
                idxE = wy_mod_u64(rE, n, ms, kE); dE = wy_mod_u64(rO, 6, ms, kO);
                idxO = wy_mod_u64(rO, n, ms, kO); dO = wy_mod_u64(rnE, 6, ms, kO + 1);
                pdE = wy_mod_1000((uint32_t)rnE, ms, kE + 2); pdO = wy_mod_1000(rnO, ms, kO + 2);
            }
            const uint32_t PE = __ballot_sync(0xffffffffu, probe_possible<BEH, RW>(L, idxE, dE));
            const uint32_t PO = __ballot_sync(0xffffffffu, probe_possible<BEH, RW>(L, idxO, dO));
            // Which slots start a real step: a blocked step consumes 2 draws, a possible one 3.  Walking lanes in
            // order, the alignment is a 3-state automaton: E (the step starts in this lane's even slot), O (odd
            // slot), S (no step starts here — an odd-slot possible step ran through; the next lane is E again).
            //   E -> possible_E ? O : E      O -> possible_O ? S : O      S -> E
            // Each lane's transition is a map on {E,O,S} (6 bits); the state on entry to lane l is the composition
            // of the maps of lanes 0..l-1 applied to E: a warp scan over function composition (5 shuffle rounds,
            // branch-free) instead of a serial walk.
            // (the next batch's words, if this one is not cut: issued here so that they fill the scan's shuffle latencies)
            const uint32_t pN0 = word(kE + 60), pN1 = word(kO + 60);
            const uint32_t pe = (PE >> lane) & 1u, po = (PO >> lane) & 1u;
            uint32_t map = pe | ((po ? 2u : 1u) << 2);       // image of E | image of O << 2 | image of S (= E = 0) << 4
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t first = __shfl_up_sync(0xffffffffu, map, off);   // composition of the block ending at lane - off
                const uint32_t r0 = (map >> (2u * (first & 3u))) & 3u;
                const uint32_t r1 = (map >> (2u * ((first >> 2) & 3u))) & 3u;
                const uint32_t r2 = (map >> (2u * ((first >> 4) & 3u))) & 3u;
                map = lane >= (uint32_t)off ? (r0 | (r1 << 2) | (r2 << 4)) : map;
            }
            const uint32_t entry = __shfl_up_sync(0xffffffffu, map, 1) & 3u;    // state after lane - 1, started from E
            const uint32_t state = lane == 0 ? 0u : entry;
            const uint32_t realE = __ballot_sync(0xffffffffu, state == 0u && lane < 30u);
            const uint32_t realO = __ballot_sync(0xffffffffu, state == 1u && lane < 30u);
            // draws consumed by lanes 0..29: 60 if the walk leaves lane 29 in E, 61 in O, 62 in S
            const uint32_t shift = __shfl_sync(0xffffffffu, map, 29) & 3u;
            const bool myO = (realO >> lane) & 1u, myE = (realE >> lane) & 1u;
            const uint32_t ord = __popc((realE | realO) & lowmask(lane));
            const bool active = (myE || myO) && (uint64_t)ord < rem;
            Prop p;
            p.idx = myO ? idxO : idxE; p.d = myO ? dO : dE;
            // the probability draw exists only for a possible step; computing it for every lane is harmless
            p.pd = myO ? pdO : pdE;
            if (!active) p.pd = 0xffffffffu;                 // never accepted, also when a commit round re-derives
            evaluate<BEH, RW>(L, g, p);
            if (!active) { p.eff = false; p.poss = false; }
            const uint32_t am = __ballot_sync(0xffffffffu, active);
            const uint32_t R = resolve_batch<BEH, RW, true, ROUNDS>(L, g, lane, active, p, ncommit);
            nposs += (active && p.poss && lane < R) ? 1u : 0u;
            k += __popc(am & lowmask(R));
            if (R < 32) {
                // the cut lane restarts the next batch in the slot it started in: the draw stream moves down by 2R (+1) slots.
                // Its words are already here — this batch's (draws cbase + 1 .. 64) and the prefetched ones (from cbase + 61):
                // even-slot word m of the stream is wE of lane m, or pN0 of lane m - 30 from m = 30 on; same for odd slots
                // (cuts happen in lanes <= 29, so m <= 61 and the prefetched lanes suffice)
                const uint32_t o = __shfl_sync(0xffffffffu, (uint32_t)myO, R);
                cbase += 2 * R + o;
                const uint32_t m = R + lane, m1 = m + 1;
                const uint32_t ea = __shfl_sync(0xffffffffu, wE, m), eb = __shfl_sync(0xffffffffu, pN0, m - 30);
                const uint32_t oa = __shfl_sync(0xffffffffu, wO, m), ob = __shfl_sync(0xffffffffu, pN1, m - 30);
                const uint32_t fa = __shfl_sync(0xffffffffu, wE, m1), fb = __shfl_sync(0xffffffffu, pN0, m1 - 30);
                // (lane 31's words may be stale after an uncut batch's one-lane shift: the prefetched ones take over from m = 30)
                const uint32_t Es = m < 30 ? ea : eb, Os = m < 30 ? oa : ob, Es1 = m1 < 30 ? fa : fb;
                wE = o ? Os : Es;
                wO = o ? Es1 : Os;
            } else {
                // the prefetched words start at cbase + 60; the batch consumed 60 + shift draws
                cbase += 60u + shift;
                const uint32_t a0 = __shfl_down_sync(0xffffffffu, pN0, 1), a1 = __shfl_down_sync(0xffffffffu, pN1, 1);
                wE = shift == 0u ? pN0 : shift == 1u ? pN1 : a0;
                wO = shift == 0u ? pN1 : shift == 1u ? a0 : a1;
            }
        }
    }
    {
        // per-lane counters -> warp totals
        uint32_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)(nposs & 0xffffffu));
        uint32_t mid = __reduce_add_sync(0xffffffffu, (uint32_t)((nposs >> 24) & 0xffffffu));
        uint32_t hi = __reduce_add_sync(0xffffffffu, (uint32_t)(nposs >> 48));
        nposs = (uint64_t)lo + ((uint64_t)mid << 24) + ((uint64_t)hi << 48);
    }

    uint32_t i0, i1;
    fitness<BEH, RW>(P, L, g, in, lane, i0, i1);
    if (lane == 0) {
        DevResult r; r.i0 = i0; r.i1 = i1; r.i2 = (BEH == LOCO) ? g.orient : 0u; r.status = status;
        P.out[in.slot] = r;
        atomicAdd(&P.counters[0], (unsigned long long)steps);
        atomicAdd(&P.counters[1], (unsigned long long)nposs);
        atomicAdd(&P.counters[2], (unsigned long long)ncommit);
    }
    if (P.dump) {
        uint32_t* dst = P.dump + (size_t)(in.slot - P.dump_slot0) * P.dump_stride;
        const uint32_t pw = P.plane_words;
        for (uint32_t w = lane; w < pw; w += 32) {
            dst[w] = L.p0[w]; dst[pw + w] = L.p1[w];
            dst[2 * pw + w] = (BEH == SEP || BEH == COAT) ? L.p2[w] : 0u;
        }
        const uint32_t* pos32 = reinterpret_cast<const uint32_t*>(L.pos);
        for (uint32_t w = lane; w < (g.n + 1) / 2; w += 32) dst[3 * pw + w] = pos32[w];
        if (BEH == LOCO) {
            const uint32_t* l32 = reinterpret_cast<const uint32_t*>(L.light);
            for (uint32_t w = lane; w < SOPS_LIGHT_WORDS; w += 32) dst[3 * pw + P.pos_words + w] = l32[w];
        }
    }
    __syncwarp();
}

// AGG: 3 CTAs of 256 threads per SM = 80 registers per thread (the allocation granule makes 81..88 cost a whole CTA
// of occupancy on the seed sweeps).  The other behaviours measured 4-5 % slower under the same cap on their GA
// batches and keep ptxas' own choice.
#ifndef SOPS_MIN_BLOCKS
#define SOPS_MIN_BLOCKS 1
#endif
template <int BEH, int RW, int MODE, int ROUNDS>
__global__ void __launch_bounds__(256, (BEH == AGG && ROUNDS != ROUNDS_ADAPTIVE) ? 3 : SOPS_MIN_BLOCKS) sops_chain_kernel(const KParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint4 s_dirtab[8];
    __shared__ uint32_t s_plut[SOPS_PLUT_WORDS];
    __shared__ uint4 s_cmp[8][8];                            // per warp: the compacted accepted moves of a batch
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < 6; ++d)
            s_dirtab[d] = dirtab_entry<BEH>(d);
    }
    for (uint32_t k = threadIdx.x; k < SOPS_PLUT_WORDS; k += blockDim.x)
        s_plut[k] = plut_entry((int)(k / 49u), (int)((k % 49u) / 7u), (int)(k % 7u));
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* base = smem + (size_t)warp * P.words_per_warp;
    Lat L;
    L.p0 = base + P.off_p0; L.p1 = base + P.off_p1; L.p2 = base + P.off_p2;
    L.pos = reinterpret_cast<uint16_t*>(base + P.off_pos);
    L.thr = reinterpret_cast<uint16_t*>(base + P.off_thr);
    L.light = reinterpret_cast<uint16_t*>(base + P.off_aux);
    L.tab = reinterpret_cast<uint16_t*>(base + P.off_tab);
    L.inv = P.inv_words != 0u;
    L.dirtab = s_dirtab;
    L.plut = smem_addr(s_plut);
    L.cmp = s_cmp[warp];
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(P.work_counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= P.n_inst) break;
        const Instance in = P.inst[q];
        run_instance<BEH, RW, MODE, ROUNDS>(P, L, in, lane);
    }
}

// =========================================================================================================
// CTA-per-instance variant (fast mode): W warps speculate on 32*W consecutive steps of ONE chain.
// Used when a batch has far fewer instances than the GPU has warp slots (a GA generation: a few hundred
// chains): the per-batch cost (draws, window loads, decision) is paid once per 32*W steps instead of once
// per 32, and the otherwise idle schedulers of the SM work on the same chain.  Commits stay strictly
// ordered: the lowest accepted (warp, lane) publishes its move in shared memory and commits, the block
// synchronises, every later thread patches its register window (same arithmetic as the warp kernel), and a
// `__syncthreads_or` tells everyone whether any decision changed (then the accepted sets are re-voted).
// The retired sequence is exactly the sequential chain, and — being keyed by step index — bit-identical to
// the warp-per-instance kernel.
// =========================================================================================================
template <int BEH>
__device__ __forceinline__ bool patch_window(uint32_t sf, uint32_t ttf, uint32_t wrote, Prop& p, bool& reread) {
    const uint32_t tf = ttf & 0xffffu;
    const uint32_t dU = win25_patch(sf, tf - sf, p.s - 0x0202u);
    reread = (p.s == sf);                                    // the moved particle itself was proposed again
    if (BEH == SEP) {
        const uint32_t x = ttf >> 28;
        if (x & 1) p.U0 ^= dU;
        if (x & 2) p.U1 ^= dU;
        reread |= (p.s == tf);                               // the swap partner was proposed
    } else {
        p.U0 ^= dU;
        if (BEH == LOCO && wrote) {
            const uint32_t b1 = wrote & 0xff, b2 = wrote >> 8;
            const uint32_t bs = (p.aux >> 8) & 0xff, bt = (p.aux >> 16) & 0xff;
            reread |= (b1 && (bs == b1 || bt == b1)) || (b2 && (bs == b2 || bt == b2));
        }
    }
    return dU != 0u || reread;
}

template <int BEH, int RW, int W>
__device__ __forceinline__ void resolve_cta(const Lat& L, const Geo& g, uint32_t tid, bool active, Prop& p, uint32_t& ncommit,
                                            uint32_t* s_ballot /*[W], this super-batch's buffer*/, uint32_t* s_bc /*[8]*/) {
    const uint32_t lane = tid & 31, warp = tid >> 5;
    {
        const uint32_t mine = __ballot_sync(0xffffffffu, p.eff);
        if (lane == 0) s_ballot[warp] = mine;
    }
    __syncthreads();                                         // S1: every thread evaluated; accepted sets published
    uint32_t A[W];
#pragma unroll
    for (int w = 0; w < W; ++w) A[w] = s_ballot[w];
    for (;;) {
        // lowest pending accepted (warp, lane): identical in every thread
        uint32_t wsel = W, Aw = 0;
#pragma unroll
        for (int w = W - 1; w >= 0; --w) if (A[w]) { wsel = (uint32_t)w; Aw = A[w]; }
        if (wsel == W) break;
        const uint32_t f = __ffs(Aw) - 1, gidx = wsel * 32 + f;
#pragma unroll
        for (int w = 0; w < W; ++w) if ((uint32_t)w == wsel) A[w] = Aw & (Aw - 1);
        const bool mine = tid == gidx;
        ++ncommit;
        if (mine) {
            s_bc[0] = p.s; s_bc[1] = p.tt;
            if (BEH == SEP) { s_bc[2] = p.kind; s_bc[3] = p.idx; }
            if (BEH == LOCO) {                               // locomotion.rs:344-519, net effect
                const uint32_t b1 = ((p.aux >> 24) & 1u) ? ((p.aux >> 8) & 0xff) : 0u;
                const uint32_t b2 = ((p.aux >> 25) & 1u) ? ((p.aux >> 16) & 0xff) : 0u;
                const uint32_t sr = p.s & 0xff, sc = p.s >> 8, tr = p.t & 0xff, tc = p.t >> 8;
                if (b1) L.light[b1 - 1] = (uint16_t)((sr - 2) | ((sc - 2) << 8));
                if (b2) L.light[b2 - 1] = (uint16_t)((tr - 2) | ((tc - 2) << 8));
                s_bc[2] = b1 | (b2 << 8);
            }
            if (BEH != SEP) {
                const uint32_t sr = p.s & 0xff, sc = p.s >> 8, tr = p.t & 0xff, tc = p.t >> 8;
                Plane<RW>::atom_xor(L.p0, sr, sc);
                Plane<RW>::atom_xor(L.p0, tr, tc);
                L.pos[p.idx] = (uint16_t)p.t;
            }
            p.eff = false;
        }
        __syncthreads();                                     // S2: the move (and, non-SEP, the lattice update) is visible
        const uint32_t sf = s_bc[0], ttf = s_bc[1];
        const uint32_t wrote = (BEH == LOCO) ? s_bc[2] : 0u;
        if (BEH == SEP) {
            const uint32_t kindf = s_bc[2], tf = ttf & 0xffffu, x = ttf >> 28;
            if (kindf == 2) {                                // partner index: block-wide scan of the position list
                for (uint32_t i = tid; i < g.n; i += 32 * W) if (L.pos[i] == tf) s_bc[4] = i;
                __syncthreads();
            }
            if (mine) {
                const uint32_t sr = sf & 0xff, sc = sf >> 8, tr = tf & 0xff, tc = tf >> 8;
                if (x & 1) { Plane<RW>::atom_xor(L.p0, sr, sc); Plane<RW>::atom_xor(L.p0, tr, tc); }
                if (x & 2) { Plane<RW>::atom_xor(L.p1, sr, sc); Plane<RW>::atom_xor(L.p1, tr, tc); }
                if (kindf == 2) L.pos[s_bc[4]] = (uint16_t)sf;
                L.pos[s_bc[3]] = (uint16_t)tf;
            }
        }
        bool reread;
        const bool touched = patch_window<BEH>(sf, ttf, wrote, p, reread) && active && tid > gidx;
        if (__syncthreads_or(touched)) {                     // S3
            if (touched) {
                if (reread) evaluate<BEH, RW>(L, g, p);
                else derive<BEH>(L, g, p);
            }
            const uint32_t again = __ballot_sync(0xffffffffu, p.eff);
            if (lane == 0) s_ballot[warp] = again;
            __syncthreads();                                 // S4
#pragma unroll
            for (int w = 0; w < W; ++w) A[w] = s_ballot[w];
        }
    }
}

template <int BEH, int RW, int W>
__global__ void __launch_bounds__(32 * W) sops_chain_cta_kernel(const KParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint4 s_dirtab[8];
    __shared__ uint32_t s_ballot[2][8];
    __shared__ uint32_t s_bc[8];
    __shared__ uint32_t s_q;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < 6; ++d)
            s_dirtab[d] = dirtab_entry<BEH>(d);
    }
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Lat L;
    L.p0 = smem + P.off_p0; L.p1 = smem + P.off_p1; L.p2 = smem + P.off_p2;
    L.pos = reinterpret_cast<uint16_t*>(smem + P.off_pos);
    L.thr = reinterpret_cast<uint16_t*>(smem + P.off_thr);
    L.light = reinterpret_cast<uint16_t*>(smem + P.off_aux);
    L.tab = reinterpret_cast<uint16_t*>(smem + P.off_tab);
    L.inv = false;
    L.dirtab = s_dirtab;
    L.plut = 0;                                              // (the parallel commit rounds are a warp-kernel form)
    L.cmp = nullptr;
    for (;;) {
        __syncthreads();                                     // previous instance fully retired (s_q, lattice reusable)
        if (tid == 0) s_q = atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= P.n_inst) break;
        const Instance in = P.inst[q];
        Geo g;
        g.a = in.arena; g.G = 2 * g.a + 1; g.orient = in.orient;
        {
            uint32_t p = in.layers, t = p + in.coat;
            g.n = BEH == AGG ? 3 * p * (p + 1) + 1 : BEH == COAT ? 3 * t * (t + 1) - 3 * p * (p + 1) : 3 * p * (p + 1);
        }
        g.n3 = g.n / 3;
        uint32_t status = 0;
        if (warp == 0) status = init_instance<BEH, RW>(P, L, g, in, lane);
        __syncthreads();

        const uint64_t steps = in.steps, ms = in.move_seed;
        const uint32_t n = g.n;
        const uint32_t key = (uint32_t)ms, khi = (uint32_t)(ms >> 32);
        auto draw = [&](uint64_t st, Prop& qd) {
            uint32_t r0, r1;
            philox2x32_10((uint32_t)st, (uint32_t)(st >> 32) ^ khi, key, r0, r1);
            qd.idx = __umulhi(r0, n);
            const uint64_t m6 = (uint64_t)r1 * 6u;
            qd.d = (uint32_t)(m6 >> 32);
            qd.pd = __umulhi((uint32_t)m6, 1000u);
        };
        uint64_t nposs = 0;
        uint32_t ncommit = 0;
        const uint64_t nsuper = (steps + 32 * W - 1) / (32 * W);
        Prop cur;
        draw((uint64_t)tid, cur);
        for (uint64_t b = 0; b < nsuper; ++b) {
            const uint64_t st = b * (32 * W) + tid;
            const bool active = st < steps;
            Prop nxt;
            draw(st + 32 * W, nxt);
            evaluate<BEH, RW>(L, g, cur);
            if (!active) { cur.eff = false; cur.poss = false; }
            resolve_cta<BEH, RW, W>(L, g, tid, active, cur, ncommit, s_ballot[b & 1], s_bc);
            nposs += (active && cur.poss) ? 1u : 0u;
            cur.idx = nxt.idx; cur.d = nxt.d; cur.pd = nxt.pd;
        }
        {
            uint32_t lo = __reduce_add_sync(0xffffffffu, (uint32_t)(nposs & 0xffffffu));
            uint32_t mid = __reduce_add_sync(0xffffffffu, (uint32_t)((nposs >> 24) & 0xffffffu));
            uint32_t hi = __reduce_add_sync(0xffffffffu, (uint32_t)(nposs >> 48));
            nposs = (uint64_t)lo + ((uint64_t)mid << 24) + ((uint64_t)hi << 48);
            if (lane == 0) atomicAdd(&P.counters[1], (unsigned long long)nposs);
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t i0, i1;
            fitness<BEH, RW>(P, L, g, in, lane, i0, i1);
            if (lane == 0) {
                DevResult r; r.i0 = i0; r.i1 = i1; r.i2 = (BEH == LOCO) ? g.orient : 0u; r.status = status;
                P.out[in.slot] = r;
                atomicAdd(&P.counters[0], (unsigned long long)steps);
                atomicAdd(&P.counters[2], (unsigned long long)ncommit);
            }
            if (P.dump) {
                uint32_t* dst = P.dump + (size_t)(in.slot - P.dump_slot0) * P.dump_stride;
                const uint32_t pw = P.plane_words;
                for (uint32_t w = lane; w < pw; w += 32) {
                    dst[w] = L.p0[w]; dst[pw + w] = L.p1[w];
                    dst[2 * pw + w] = (BEH == SEP || BEH == COAT) ? L.p2[w] : 0u;
                }
                const uint32_t* pos32 = reinterpret_cast<const uint32_t*>(L.pos);
                for (uint32_t w = lane; w < (g.n + 1) / 2; w += 32) dst[3 * pw + w] = pos32[w];
                if (BEH == LOCO) {
                    const uint32_t* l32 = reinterpret_cast<const uint32_t*>(L.light);
                    for (uint32_t w = lane; w < SOPS_LIGHT_WORDS; w += 32) dst[3 * pw + P.pos_words + w] = l32[w];
                }
            }
        }
    }
}

}  // namespace sops
