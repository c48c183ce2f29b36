// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Plain CPU restatement (u8 cell grid + particle vector, exactly the reference's data model) of the
// four SOPS environments of DaymudeLab/EvoSOPS.  Each function cites the reference lines it follows.
// The Rust reference cannot be compiled in this image (no cargo/rustc, crates not vendored), so this
// restatement is the checker.  Pinning status:
//   * init (boundary mask, seeded placement, colour hand-out, particle order): PINNED against the
//     reference's own step-0 snapshot grids (tests/golden/agg_13_9_step0.txt, sep_13_9_step0.txt).
//   * evaluate_fitness (agg, sep, coat): PINNED against hand counts on the reference's snapshot grids.
//   * move stream: PARITY UNPINNED (the reference's fastrand generator is time-seeded); the oracle
//     defines the reproducible stream (see rng.h).
#include "rng.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

namespace oracle {

enum Behavior { AGG = 0, SEP = 1, COAT = 2, LOCO = 3 };

struct Particle { uint8_t x, y, state; };                       // src/SOPSCore/mod.rs:12-16

// src/SOPSCore/mod.rs:80-82 — order matters: the direction draw indexes this table
static const int DIRS[6][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, 1}};
// src/SOPSCore/mod.rs:85-87 (GA table) and 91-93 (theory table, dead code upstream; used by `th`)
static const uint16_t PROB_GA[11] = {1000, 500, 250, 125, 63, 31, 16, 8, 4, 2, 1};
static const uint16_t PROB_THEORY_AGG[6] = {1000, 167, 28, 5, 1, 1};
// src/SOPSCore/separation.rs:82-84
static const uint16_t PROB_THEORY_SEP[16] = {1000, 667, 444, 296, 250, 167, 111, 74, 63, 42, 28, 19, 16, 10, 7, 5};

struct Result { uint32_t i0, i1, i2; float f; double norm; };

struct Env {
    int behavior;
    int a, p, c;                 // arena_layers, particle_layers (object_layers for coat), coat_layers
    int G;                       // grid side 2a+1
    std::vector<uint8_t> grid;   // G*G row-major, grid[i*G+j]
    std::vector<Particle> parts;
    std::vector<uint8_t> genome; // 48 | 600 | 600 | 144 bytes
    const uint16_t* prob; int prob_len;
    float w1, w2;
    uint64_t sim_duration;
    uint64_t steps_done;
    MoveRng rng;
    // counters (not in the reference; used for the roofline's p_poss / p_acc)
    uint64_t n_possible, n_accepted;
    // coat
    std::vector<uint16_t> dist;  // G*G, 0xFFFF = no entry
    // loco
    int orientation;
    std::vector<std::pair<uint8_t, uint8_t>> light;
    float max_c1, max_c2;
    uint64_t max_fitness;        // agg
    uint8_t EMPTY, BOUNDARY;

    Env(int rng_mode, uint64_t move_seed) : rng(rng_mode, move_seed) {}

    uint8_t& at(int i, int j) { return grid[(size_t)i * G + j]; }
    uint8_t at(int i, int j) const { return grid[(size_t)i * G + j]; }
    bool in_array(int i, int j) const { return i >= 0 && i < G && j >= 0 && j < G; }

    // ---- init -----------------------------------------------------------------------------
    // boundary corners: mod.rs:115-122 (identical loops at separation.rs:118-125,
    // coating.rs:249-256, locomotion.rs:178-185)
    void mark_bounds() {
        for (int i = 0; i < a; ++i) {
            int j = 1;
            while (i + a + j < G) {
                at(i, i + a + j) = BOUNDARY;
                at(i + a + j, i) = BOUNDARY;
                ++j;
            }
        }
    }

    int init(int behavior_, const uint8_t* genome_, int a_, int p_, int c_, uint64_t seed,
             float w1_, float w2_, int orientation_, int prob_table) {
        behavior = behavior_; a = a_; p = p_; c = c_;
        G = 2 * a + 1;
        w1 = w1_; w2 = w2_;
        orientation = orientation_;
        steps_done = 0; n_possible = n_accepted = 0;
        grid.assign((size_t)G * G, 0);
        parts.clear();
        EMPTY = 0;
        int glen = 48;
        uint32_t n = 0;
        switch (behavior) {
        case AGG:  BOUNDARY = 2; glen = 48;  n = 6 * p * (1 + p) / 2 + 1; break;   // mod.rs:109
        case SEP:  BOUNDARY = 4; glen = 600; n = 6 * p * (1 + p) / 2; break;       // separation.rs:108
        case COAT: BOUNDARY = 7; glen = 600; {                                      // coating.rs:234-235
            int t = p + c; n = 6 * t * (1 + t) / 2 - 6 * p * (1 + p) / 2; } break;
        case LOCO: BOUNDARY = 2; glen = 144; n = 6 * p * (1 + p) / 2; break;       // locomotion.rs:103
        default: return -1;
        }
        genome.assign(genome_, genome_ + glen);
        if (prob_table == 0) { prob = PROB_GA; prob_len = 11; }
        else if (behavior == SEP || behavior == COAT) { prob = PROB_THEORY_SEP; prob_len = 16; }
        else { prob = PROB_THEORY_AGG; prob_len = 6; }
        for (int g : genome) if (g >= prob_len) return -2;      // the reference would panic (index OOB)

        // free-cell check: the reference's placement loop never terminates otherwise
        int64_t arena_cells = 3LL * a * (a + 1) + 1;
        int64_t obj_cells = (behavior == COAT) ? 3LL * p * (p + 1) + 1 : 0;
        if ((int64_t)n > arena_cells - obj_cells || a > 127 || a < 1) return -3;

        // fitness denominators
        if (behavior == AGG) {                                   // mod.rs:110-111
            uint64_t k = 3 * (uint64_t)p;
            max_fitness = k * (k + 1);
        } else if (behavior == SEP) {                            // separation.rs:112-114 (u16 -> f32)
            max_c1 = (float)(uint16_t)(3 * (3 * p * p + p - 1));
            max_c2 = (float)(uint16_t)(3 * p * p - p - 1);
        } else if (behavior == LOCO) {                           // locomotion.rs:172-173
            max_c1 = (float)(uint16_t)(3 * (3 * p * p + p - 1));
            float np = (float)n;
            float inner = 1.0f + 4.0f * (np - 1.0f) / 3.0f;
            max_c2 = (float)G - ((-1.0f + sqrtf(inner)) / 2.0f);
            if (a > 63) return -4;                               // u8 arithmetic would wrap upstream
            if (orientation < 1 || orientation > 3) return -5;
        }

        // loco: light table before placement — locomotion.rs:107-168
        if (behavior == LOCO) {
            light.assign((size_t)G, std::make_pair((uint8_t)0, (uint8_t)0));
            if (orientation == 3) light[0] = std::make_pair((uint8_t)0, (uint8_t)a);
            if (orientation == 1 || orientation == 2) {          // the `else if orientation == 2` arm at 141 is unreachable
                int i = 1;
                while (i <= a * 2) {
                    light[i].first = (uint8_t)(light[i - 1].first + 1);
                    light[i].second = (uint8_t)(light[i - 1].second + 1);
                    ++i;
                    light[i].first = (uint8_t)((int8_t)light[i - 1].first + 0);
                    light[i].second = (uint8_t)((int8_t)light[i - 1].second - 1);
                    ++i;
                }
            } else {
                int i = 1;
                while (i <= a * 2) {
                    light[i].first = (uint8_t)(light[i - 1].first + 1);
                    light[i].second = (uint8_t)(light[i - 1].second + 0);
                    ++i;
                    light[i].first = (uint8_t)(light[i - 1].first + 0);
                    light[i].second = (uint8_t)(light[i - 1].second + 1);
                    ++i;
                }
            }
        }

        StdRng init_rng(seed);                                   // mod.rs:113
        mark_bounds();

        if (behavior == COAT) {                                  // coating.rs:258-284
            int total_coat = p + c;
            int total_coat_size = total_coat * 2 + 1;
            int object_size = p * 2 + 1;
            int x = a - p, y = a - p;
            // the reference spins forever when this placement test fails
            if (!((y + total_coat_size) < G && (x + total_coat_size) < G)) return -6;
            if (at(x, y) == BOUNDARY || at(x, y + total_coat_size) == BOUNDARY ||
                at(x + total_coat_size, y) == BOUNDARY ||
                at(x + total_coat_size, y + total_coat_size) == BOUNDARY) return -6;
            for (int i = 0; i < object_size; ++i) {
                int j = i;
                while (j < i + p + 1 && j < object_size) {
                    at(x + c + i, y + c + j) = 2;
                    at(x + c + j, y + c + i) = 2;
                    ++j;
                }
            }
        }

        // seeded rejection placement — mod.rs:125-137; colour hand-out separation.rs:126-146;
        // light raise locomotion.rs:199-225
        uint32_t per_colour = n / 3, colour_cnt = 0;
        uint8_t colour = 1;
        while (parts.size() < n) {
            int i = (int)init_rng.uniform_below((uint64_t)G);
            int j = (int)init_rng.uniform_below((uint64_t)G);
            if (at(i, j) != 0) continue;
            if (behavior == SEP) {
                parts.push_back({(uint8_t)i, (uint8_t)j, colour});
                at(i, j) = colour;
                if (++colour_cnt == per_colour) { colour_cnt = 0; ++colour; }
            } else {
                parts.push_back({(uint8_t)i, (uint8_t)j, 0});
                at(i, j) = 1;
            }
            if (behavior == LOCO) {
                if (orientation == 1) {
                    if (2 * i >= j && (uint8_t)(2 * (uint8_t)a) >= (uint8_t)(2 * i - j) &&
                        (uint8_t)j >= light[(uint8_t)(2 * i - j)].second)
                        light[(uint8_t)(2 * i - j)] = std::make_pair((uint8_t)i, (uint8_t)j);
                } else if (orientation == 2) {
                    if (2 * j >= i && (uint8_t)(2 * (uint8_t)a) >= (uint8_t)(2 * j - i) &&
                        (uint8_t)i >= light[(uint8_t)(2 * j - i)].first)
                        light[(uint8_t)(2 * j - i)] = std::make_pair((uint8_t)i, (uint8_t)j);
                } else {
                    if (i + j >= a && i + j < 3 * a && (uint8_t)i >= light[i + j - a].first)
                        light[i + j - a] = std::make_pair((uint8_t)i, (uint8_t)j);
                }
            }
        }
        uint64_t nn = n;
        sim_duration = nn * nn * nn;                             // mod.rs:143
        if (behavior == COAT) save_distance_grid();
        return 0;
    }

    // ---- coat distance grid: coating.rs:101-129 (per-cell BFS) + 341-356 ---------------------------
    // Restated literally: a BFS from every EMPTY/PARTICLE cell that stops at the first OBJECT cell.
    uint16_t get_obj_dist(int si, int sj) const {
        std::vector<uint8_t> visited((size_t)G * G, 0);
        std::vector<int> qi{si}, qj{sj}, qd{0};
        size_t head = 0;
        while (head < qi.size()) {
            int ci = qi[head], cj = qj[head], depth = qd[head];
            ++head;
            for (int k = 0; k < 6; ++k) {
                int ni = ci + DIRS[k][0], nj = cj + DIRS[k][1];
                if (!in_array(ni, nj) || visited[(size_t)ni * G + nj]) continue;
                uint8_t v = at(ni, nj);
                if (v == 2) return (uint16_t)(depth + 1);
                if (v == 0 || v == 1) {
                    qi.push_back(ni); qj.push_back(nj); qd.push_back(depth + 1);
                    visited[(size_t)ni * G + nj] = 1;
                }
            }
        }
        return 0;
    }
    void save_distance_grid() {
        dist.assign((size_t)G * G, 0xFFFF);
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j)
                if (at(i, j) == 0 || at(i, j) == 1) dist[(size_t)i * G + j] = get_obj_dist(i, j);
    }

    // ---- plain 6-neighbour counts ----------------------------------------------------------------
    // mod.rs:166-178 / locomotion.rs:315-327 / coating: n.a.
    int nbr_particles(int i, int j) const {
        int cnt = 0;
        for (int k = 0; k < 6; ++k) {
            int ni = i + DIRS[k][0], nj = j + DIRS[k][1];
            if (in_array(ni, nj) && at(ni, nj) == 1) ++cnt;
        }
        return cnt;
    }
    // separation.rs:255-275 → (any colour, same colour)
    void nbr_sep(int i, int j, int& any, int& same) const {
        any = same = 0;
        for (int k = 0; k < 6; ++k) {
            int ni = i + DIRS[k][0], nj = j + DIRS[k][1];
            if (!in_array(ni, nj)) continue;
            uint8_t v = at(ni, nj);
            if (v != 0 && v != 4) { ++any; if (v == at(i, j)) ++same; }
        }
    }

    // ---- extended neighbourhood (back, mid, front) -------------------------------------------------
    // Generic two-class counter that follows the reference's two loops + seen-cache literally:
    //   class A / class B per cell decided by `classify` (0 none, 1 A only, 2 B only, 3 both).
    // agg/loco  (mod.rs:183-230, locomotion.rs:551-600): A = PARTICLE, B unused
    // sep       (separation.rs:649-715): A = occupied, B = same colour as mover
    // coat      (coating.rs:439-510):    A = PARTICLE, B = OBJECT
    struct Ext { int backA, backB, midA, midB, frontA, frontB; };
    template <class F>
    Ext ext_counts(int x, int y, int di, int dj, F classify) const {
        Ext e{0, 0, 0, 0, 0, 0};
        int mi = x + di, mj = y + dj;
        int seen_i[6], seen_j[6], nseen = 0;
        for (int k = 0; k < 6; ++k) {
            int ni = x + DIRS[k][0], nj = y + DIRS[k][1];
            if (in_array(ni, nj) && !(ni == mi && nj == mj)) {
                seen_i[nseen] = ni; seen_j[nseen] = nj; ++nseen;
                int cl = classify(at(ni, nj));
                if (cl & 1) ++e.backA;
                if (cl & 2) ++e.backB;
            }
        }
        for (int k = 0; k < 6; ++k) {
            int ni = mi + DIRS[k][0], nj = mj + DIRS[k][1];
            if (in_array(ni, nj) && !(ni == x && nj == y)) {
                bool mid = false;
                for (int q = 0; q < nseen; ++q) if (seen_i[q] == ni && seen_j[q] == nj) mid = true;
                int cl = classify(at(ni, nj));
                if (mid) {
                    if (cl & 1) { ++e.midA; --e.backA; }
                    if (cl & 2) { ++e.midB; --e.backB; }
                } else {
                    if (cl & 1) ++e.frontA;
                    if (cl & 2) ++e.frontB;
                }
            }
        }
        return e;
    }

    // separation.rs:205-222 (lookup table) + 323-331 (miss → 0)
    static int sep_dim_idx(int all, int same, int possible) {
        static const int T[4][4] = {{0, -1, -1, -1}, {2, 1, -1, -1}, {5, 4, 3, -1}, {9, 8, 7, 6}};
        if (all < 0 || all > possible || same < 0 || same > all) return 0;
        int v = T[all][same];
        return v < 0 ? 0 : v;
    }
    // coating.rs:302-320 + 426-434
    static int coat_dim_idx(int par, int obj, int possible) {
        if (possible == 2) {
            static const int T2[3][3] = {{0, 1, 2}, {4, 3, -1}, {5, -1, -1}};
            if (par < 0 || par > 2 || obj < 0 || obj > 2) return 0;
            int v = T2[par][obj]; return v < 0 ? 0 : v;
        }
        static const int T3[4][4] = {{0, 1, 2, 3}, {7, 5, 6, -1}, {8, 4, -1, -1}, {9, -1, -1, -1}};
        if (par < 0 || par > 3 || obj < 0 || obj > 3) return 0;
        int v = T3[par][obj]; return v < 0 ? 0 : v;
    }

    static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

    // gene for particle `idx` moving in direction (di,dj)
    int gene_agg(int idx, int di, int dj) const {                 // mod.rs:183-230, 284
        const Particle& pt = parts[idx];
        Ext e = ext_counts(pt.x, pt.y, di, dj, [](uint8_t v) { return v == 1 ? 1 : 0; });
        int b = clampi(e.backA, 0, 3), m = clampi(e.midA, 0, 2), f = clampi(e.frontA, 0, 3);
        return genome[(b * 3 + m) * 4 + f];
    }
    int gene_sep(int idx, int di, int dj) const {                 // separation.rs:649-715
        const Particle& pt = parts[idx];
        uint8_t st = pt.state;
        Ext e = ext_counts(pt.x, pt.y, di, dj, [st](uint8_t v) {
            if (v == 0 || v == 4) return 0;
            return v == st ? 3 : 1;
        });
        int b = clampi(sep_dim_idx(e.backA, e.backB, 3), 0, 9);
        int m = clampi(sep_dim_idx(e.midA, e.midB, 2), 0, 5);
        int f = clampi(sep_dim_idx(e.frontA, e.frontB, 3), 0, 9);
        return genome[(b * 6 + m) * 10 + f];
    }
    int gene_coat(int idx, int di, int dj) const {                // coating.rs:439-510
        const Particle& pt = parts[idx];
        Ext e = ext_counts(pt.x, pt.y, di, dj, [](uint8_t v) { return v == 1 ? 1 : (v == 2 ? 2 : 0); });
        int b = clampi(coat_dim_idx(e.backA, e.backB, 3), 0, 9);
        int m = clampi(coat_dim_idx(e.midA, e.midB, 2), 0, 5);
        int f = clampi(coat_dim_idx(e.frontA, e.frontB, 3), 0, 9);
        return genome[(b * 6 + m) * 10 + f];
    }

    // locomotion.rs:260-290 — all arithmetic in u8 like the reference (a <= 63 keeps it overflow-free)
    int sensing_light(uint8_t x, uint8_t y) const {
        if (orientation == 1) {
            if ((uint8_t)(2 * x) >= y) {
                uint8_t b = (uint8_t)(2 * x - y);
                if ((uint8_t)(2 * (uint8_t)a) >= b && y >= light[b].second) return 1;
            }
        } else if (orientation == 2) {
            if ((uint8_t)(2 * y) >= x) {
                uint8_t b = (uint8_t)(2 * y - x);
                if ((uint8_t)(2 * (uint8_t)a) >= b && x >= light[b].first) return 1;
            }
        } else if (orientation == 3) {
            if ((uint8_t)(x + y) >= (uint8_t)a && (uint8_t)(x + y) <= (uint8_t)(3 * a)) {
                if (x >= light[(uint8_t)(x + y - (uint8_t)a)].first) return 1;
            }
        }
        return 0;
    }
    int gene_loco(int idx, int di, int dj) const {                // locomotion.rs:551-600
        const Particle& pt = parts[idx];
        Ext e = ext_counts(pt.x, pt.y, di, dj, [](uint8_t v) { return v == 1 ? 1 : 0; });
        int cur = sensing_light(pt.x, pt.y);
        int fut = sensing_light((uint8_t)(pt.x + di), (uint8_t)(pt.y + dj));
        int li = (cur == 0 && fut == 1) ? 0 : (cur == 1 && fut == 0) ? 1 : 2;   // locomotion.rs:230-235
        int b = clampi(e.backA, 0, 3), m = clampi(e.midA, 0, 2), f = clampi(e.frontA, 0, 3);
        return genome[((li * 4 + b) * 3 + m) * 4 + f];
    }

    int find_particle_at(int i, int j) const {                    // separation.rs:302,749 (`position`)
        for (size_t q = 0; q < parts.size(); ++q)
            if (parts[q].x == i && parts[q].y == j) return (int)q;
        return -1;
    }

    // ---- commit ------------------------------------------------------------------------------------
    void move_plain(int idx, int ni, int nj) {                     // mod.rs:256-266
        Particle& pt = parts[idx];
        at(pt.x, pt.y) = 0;
        at(ni, nj) = 1;
        pt.x = (uint8_t)ni; pt.y = (uint8_t)nj;
    }
    bool move_sep(int idx, int ni, int nj) {                       // separation.rs:280-318
        Particle& pt = parts[idx];
        if (at(ni, nj) == 4 || at(ni, nj) == at(pt.x, pt.y)) return false;
        if (at(ni, nj) == 0) {
            at(pt.x, pt.y) = 0;
            at(ni, nj) = pt.state;
            pt.x = (uint8_t)ni; pt.y = (uint8_t)nj;
            return true;
        }
        int sw = find_particle_at(ni, nj);
        at(pt.x, pt.y) = at(ni, nj);
        at(ni, nj) = pt.state;
        uint8_t tx = pt.x, ty = pt.y;
        pt.x = (uint8_t)ni; pt.y = (uint8_t)nj;
        parts[sw].x = tx; parts[sw].y = ty;
        return true;
    }
    void move_loco(int idx, int ni, int nj) {                      // locomotion.rs:344-532
        Particle& pt = parts[idx];
        int cur = sensing_light(pt.x, pt.y);
        int fut = sensing_light((uint8_t)ni, (uint8_t)nj);
        // The "same beam" tests (405, 450, 495) cannot hold for a lattice direction, and every
        // "closest particle" scan stops on its first iteration because the mover's own cell is still
        // PARTICLE — so the net effect of 403-519 is the two assignments below.
        auto beam = [&](int i, int j) -> int {
            if (orientation == 1) return 2 * i - j;
            if (orientation == 2) return 2 * j - i;
            return i + j - a;
        };
        if (cur == 1) {
            light[beam(pt.x, pt.y)] = std::make_pair(pt.x, pt.y);
            if (fut == 1) light[beam(ni, nj)] = std::make_pair((uint8_t)ni, (uint8_t)nj);
        } else if (fut == 1) {
            light[beam(ni, nj)] = std::make_pair((uint8_t)ni, (uint8_t)nj);
        }
        at(pt.x, pt.y) = 0;
        at(ni, nj) = 1;
        pt.x = (uint8_t)ni; pt.y = (uint8_t)nj;
    }

    // ---- one chain step: mod.rs:271-290, separation.rs:756-805, coating.rs:535-555,
    //      locomotion.rs:637-654 ---------------------------------------------------------------------
    void step() {
        rng.begin_step(steps_done);
        int idx = (int)rng.particle((uint32_t)parts.size());
        int d = (int)rng.direction();
        int di = DIRS[d][0], dj = DIRS[d][1];
        const Particle& pt = parts[idx];
        int ni = pt.x + di, nj = pt.y + dj;
        ++steps_done;
        if (!in_array(ni, nj)) return;
        uint8_t tv = at(ni, nj);
        if (behavior == SEP) {
            if (tv == 4) return;                                   // particle_move_possible → 0
            ++n_possible;
            int g = gene_sep(idx, di, dj);
            if (tv != 0) {                                         // swap: separation.rs:764-786
                int j2 = find_particle_at(ni, nj);
                int g2 = gene_sep(j2, -di, -dj);
                if (g2 > g) g = g2;
            }
            if (rng.prob_1_1000() <= prob[g]) {
                if (move_sep(idx, ni, nj)) ++n_accepted;
            }
            return;
        }
        if (behavior == COAT) { if (tv != 0) return; }             // coating.rs:515-530
        else { if (tv == 1 || tv == BOUNDARY) return; }            // mod.rs:235-251
        ++n_possible;
        int g = behavior == AGG ? gene_agg(idx, di, dj)
              : behavior == COAT ? gene_coat(idx, di, dj) : gene_loco(idx, di, dj);
        if (rng.prob_1_1000() <= prob[g]) {
            ++n_accepted;
            if (behavior == LOCO) move_loco(idx, ni, nj);
            else move_plain(idx, ni, nj);                          // coat's -=1/+=1 (coating.rs:408-421) is the same on EMPTY targets
        }
    }

    // ---- fitness: mod.rs:296-301, separation.rs:811-825, coating.rs:377-392+561-574,
    //      locomotion.rs:661-688 ---------------------------------------------------------------------
    Result fitness() const {
        Result r{0, 0, 0, 0.0f, 0.0};
        uint32_t n = (uint32_t)parts.size();
        if (behavior == AGG) {
            uint32_t sum = 0;
            for (const Particle& pt : parts) sum += (uint32_t)nbr_particles(pt.x, pt.y);
            uint32_t edges = sum / 2;
            r.i0 = edges; r.i1 = (uint32_t)max_fitness;
            r.f = (float)edges / (float)max_fitness;               // mod.rs:312 (print only)
            r.norm = (double)edges / (double)max_fitness;          // base_ga.rs:409
        } else if (behavior == SEP) {
            uint32_t E = 0, S = 0;
            for (const Particle& pt : parts) { int any, same; nbr_sep(pt.x, pt.y, any, same); E += any; S += same; }
            float e = ((float)E) / 2.0f;
            float s = ((float)S) / (2.0f * 3.0f);
            float c1 = e / max_c1, c2 = s / max_c2;
            float f = w1 * c1 + w2 * c2;
            r.i0 = E; r.i1 = S; r.f = f; r.norm = (double)f;       // sep_ga.rs:441
        } else if (behavior == COAT) {
            std::vector<uint16_t> dv;
            for (uint16_t d : dist) if (d != 0xFFFF) dv.push_back(d);
            std::sort(dv.begin(), dv.end());
            uint32_t min_total = 0;
            for (uint32_t q = 0; q < n; ++q) min_total += dv[q];
            uint32_t par_total = 0;
            for (const Particle& pt : parts) {
                uint16_t d = dist[(size_t)pt.x * G + pt.y];
                if (d != 0xFFFF) par_total += d;
            }
            float f = (float)min_total / (float)par_total;
            r.i0 = min_total; r.i1 = par_total; r.f = f; r.norm = (double)f;
        } else {
            uint32_t sum = 0;
            for (const Particle& pt : parts) sum += (uint32_t)nbr_particles(pt.x, pt.y);
            uint32_t distance = 0;
            if (orientation == 1) for (const Particle& pt : parts) distance += pt.y;
            else if (orientation == 2) for (const Particle& pt : parts) distance += pt.x;
            else for (const Particle& pt : parts)
                distance += (uint32_t)(int32_t)(int8_t)((int8_t)pt.x - (int8_t)pt.y + (int8_t)a);
            float c1 = (float)(sum / 2) / max_c1;
            float c2 = (float)(distance / n) / max_c2;
            float f = w1 * c1 + w2 * c2;
            r.i0 = sum; r.i1 = distance; r.i2 = (uint32_t)orientation; r.f = f; r.norm = (double)f;
        }
        return r;
    }
};

}  // namespace oracle

// ================================================================ C API (ctypes) ================
using oracle::Env;

extern "C" {

struct oracle_result { uint32_t i0, i1, i2; float f; double norm; };

// StdRng probes for the golden vectors
void oracle_stdrng_key(uint64_t seed, uint32_t key_out[8]) {
    oracle::StdRng r(seed);
    for (int k = 0; k < 8; ++k) key_out[k] = r.key[k];
}
void oracle_stdrng_u64(uint64_t seed, int count, uint64_t* out) {
    oracle::StdRng r(seed);
    for (int k = 0; k < count; ++k) out[k] = r.next_u64();
}
void oracle_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int double_rounds, uint32_t out[16]) {
    oracle::chacha_block(key, counter, stream, double_rounds, out);
}
void oracle_philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key, uint32_t out[2]) {
    oracle::philox2x32_10(c0, c1, key, out);
}
void oracle_wyrand_u64(uint64_t seed, int count, uint64_t* out) {
    oracle::WyRand r(seed);
    for (int k = 0; k < count; ++k) out[k] = r.gen_u64();
}
// the move stream as an instance sees it: k-th fastrand::Rng::new() then the typed draw
void oracle_move_draws(uint64_t move_seed, int count, uint32_t n, uint32_t* out /*count*3*/) {
    oracle::WyRand tl(move_seed);
    for (int k = 0; k < count; ++k) {
        { oracle::WyRand r(tl.gen_u64()); out[3 * k + 0] = (uint32_t)r.usize_below(n); }
        { oracle::WyRand r(tl.gen_u64()); out[3 * k + 1] = (uint32_t)r.usize_below(6); }
        { oracle::WyRand r(tl.gen_u64()); out[3 * k + 2] = r.u16_1_to_1000(); }
    }
}

void* oracle_env_create(int behavior, const uint8_t* genome, int arena, int layers, int coat,
                        uint64_t seed, float w1, float w2, int rng_mode, uint64_t move_seed,
                        int orientation, int prob_table, int* err) {
    Env* e = new Env(rng_mode, move_seed);
    int rc = e->init(behavior, genome, arena, layers, coat, seed, w1, w2, orientation, prob_table);
    if (err) *err = rc;
    if (rc != 0) { delete e; return nullptr; }
    return e;
}
void oracle_env_destroy(void* h) { delete (Env*)h; }
int oracle_env_grid_size(void* h) { return ((Env*)h)->G; }
uint32_t oracle_env_particle_count(void* h) { return (uint32_t)((Env*)h)->parts.size(); }
uint64_t oracle_env_sim_duration(void* h) { return ((Env*)h)->sim_duration; }
void oracle_env_step(void* h, uint64_t count) { Env* e = (Env*)h; for (uint64_t k = 0; k < count; ++k) e->step(); }
void oracle_env_fitness(void* h, oracle_result* out) {
    oracle::Result r = ((Env*)h)->fitness();
    out->i0 = r.i0; out->i1 = r.i1; out->i2 = r.i2; out->f = r.f; out->norm = r.norm;
}
void oracle_env_counters(void* h, uint64_t out[3]) {
    Env* e = (Env*)h; out[0] = e->steps_done; out[1] = e->n_possible; out[2] = e->n_accepted;
}
void oracle_env_dump(void* h, uint8_t* grid /*G*G*/, uint8_t* particles /*n*3*/) {
    Env* e = (Env*)h;
    if (grid) memcpy(grid, e->grid.data(), e->grid.size());
    if (particles) for (size_t q = 0; q < e->parts.size(); ++q) {
        particles[3 * q] = e->parts[q].x; particles[3 * q + 1] = e->parts[q].y; particles[3 * q + 2] = e->parts[q].state;
    }
}
// overwrite the lattice with an external grid (for evaluate_fitness known answers on the
// reference's snapshot grids); particles are rebuilt in row-major order.
int oracle_env_load_grid(void* h, const uint8_t* grid) {
    Env* e = (Env*)h;
    e->parts.clear();
    for (int i = 0; i < e->G; ++i) for (int j = 0; j < e->G; ++j) {
        uint8_t v = grid[(size_t)i * e->G + j];
        e->at(i, j) = v;
        bool is_par = (e->behavior == oracle::SEP) ? (v >= 1 && v <= 3) : (v == 1);
        if (is_par) e->parts.push_back({(uint8_t)i, (uint8_t)j, (uint8_t)(e->behavior == oracle::SEP ? v : 0)});
    }
    if (e->behavior == oracle::COAT) e->save_distance_grid();
    return (int)e->parts.size();
}
void oracle_env_dist_grid(void* h, uint16_t* out /*G*G, 0xFFFF = none*/) {
    Env* e = (Env*)h; memcpy(out, e->dist.data(), e->dist.size() * sizeof(uint16_t));
}
void oracle_env_light(void* h, uint8_t* out /*G*2*/) {
    Env* e = (Env*)h;
    for (size_t k = 0; k < e->light.size(); ++k) { out[2 * k] = e->light[k].first; out[2 * k + 1] = e->light[k].second; }
}

// Batched evaluation with a thread pool over independent instances — the stand-in for the
// reference's nested rayon par_iter (base_ga.rs:393-411).  Also bench.py's CPU baseline.
// instance q = (genome g, trial t): g = q / T, t = q % T.
int oracle_eval_batch(int behavior, const uint8_t* genomes, uint32_t P, uint32_t L,
                      const uint16_t* sizes /*T*3: arena,layers,coat*/, const uint64_t* seeds /*T*/,
                      uint32_t T, float w1, float w2, int rng_mode,
                      const uint64_t* move_seeds /*P*T or NULL → seed*/, const uint8_t* orient /*P*T or NULL → 1*/,
                      uint64_t steps_override, int prob_table, int n_threads,
                      oracle_result* out /*P*T*/, uint64_t* counters /*3 totals or NULL*/) {
    std::atomic<uint32_t> next(0);
    std::atomic<int> rc_all(0);
    std::atomic<uint64_t> tot_steps(0), tot_poss(0), tot_acc(0);
    uint32_t N = P * T;
    auto worker = [&]() {
        for (;;) {
            uint32_t q = next.fetch_add(1);
            if (q >= N) break;
            uint32_t g = q / T, t = q % T;
            uint64_t ms = move_seeds ? move_seeds[q] : seeds[t];
            int ori = orient ? orient[q] : 1;
            Env e(rng_mode, ms);
            int rc = e.init(behavior, genomes + (size_t)g * L, sizes[3 * t], sizes[3 * t + 1], sizes[3 * t + 2],
                            seeds[t], w1, w2, ori, prob_table);
            if (rc != 0) { rc_all = rc; continue; }
            uint64_t steps = steps_override ? steps_override : e.sim_duration;
            for (uint64_t k = 0; k < steps; ++k) e.step();
            oracle::Result r = e.fitness();
            out[q].i0 = r.i0; out[q].i1 = r.i1; out[q].i2 = r.i2; out[q].f = r.f; out[q].norm = r.norm;
            tot_steps += e.steps_done; tot_poss += e.n_possible; tot_acc += e.n_accepted;
        }
    };
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    for (int k = 0; k < n_threads; ++k) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
    if (counters) { counters[0] = tot_steps; counters[1] = tot_poss; counters[2] = tot_acc; }
    return rc_all;
}

unsigned long long oracle_lemire_redraws() { return oracle::g_lemire_redraws.load(); }

int oracle_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
