// sops_host.cu — the C ABI of libevosops_b200.so (include/evosops.h) over the chain kernels.
//
// One sops_eval_batch call replaces what one GA generation's nested rayon par_iter does in the
// reference (src/GACore/base_ga.rs:393-411, sep_ga.rs:419-443, coat_ga.rs:410-436, loco_ga.rs:419-433):
// every (genome, size, seed) instance becomes one warp-resident chain.  Host work here is validation,
// cost-sorted instance lists (longest chain first, LPT over devices), H2D/D2H and the reference's
// float fitness formulas applied to the exact integers the kernels return.
//
// No CPU fallback: without a CUDA device sops_ctx_create fails with SOPS_E_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/evosops.h"
#include "sops_kernels.cuh"

using sops::DevResult;
using sops::Instance;
using sops::KParams;

namespace {

const uint16_t PROB_GA[11] = {1000, 500, 250, 125, 63, 31, 16, 8, 4, 2, 1};                 // mod.rs:85-87
const uint16_t PROB_THEORY_AGG[6] = {1000, 167, 28, 5, 1, 1};                               // mod.rs:91-93
const uint16_t PROB_THEORY_SEP[16] = {1000, 667, 444, 296, 250, 167, 111, 74,               // separation.rs:82-84
                                      63, 42, 28, 19, 16, 10, 7, 5};

struct Shape { uint32_t G, n; uint64_t dur; };

int genome_len(int b) { return b == SOPS_AGG ? 48 : b == SOPS_LOCO ? 144 : 600; }

// particle counts: mod.rs:109, separation.rs:108, coating.rs:234-235, locomotion.rs:103
int shape_of(int b, sops_size s, Shape* out) {
    uint64_t a = s.arena, p = s.layers, c = s.coat;
    if (a < 1 || a > 127) return SOPS_E_SIZE;
    uint64_t n;
    switch (b) {
    case SOPS_AGG: n = 3 * p * (p + 1) + 1; break;
    case SOPS_SEP: case SOPS_LOCO: n = 3 * p * (p + 1); break;
    case SOPS_COAT: { uint64_t t = p + c; n = 3 * t * (t + 1) - 3 * p * (p + 1); } break;
    default: return SOPS_E_ARG;
    }
    uint64_t arena_cells = 3 * a * (a + 1) + 1;
    uint64_t obj_cells = (b == SOPS_COAT) ? 3 * p * (p + 1) + 1 : 0;
    if (n == 0 || n + obj_cells > arena_cells) return SOPS_E_SIZE;   // the reference's placement loop would spin
    if (b == SOPS_COAT) {                                              // coating.rs:258-260 placement test
        uint64_t G = 2 * a + 1, side = 2 * (p + c) + 1, x = a - p;
        if (p > a) return SOPS_E_SIZE;
        if (!(x + side < G)) return SOPS_E_SIZE;
        if (side > a) return SOPS_E_SIZE;                              // corner (x, x+side) must not be BOUNDARY
    }
    out->G = (uint32_t)(2 * a + 1);
    out->n = (uint32_t)n;
    out->dur = n * n * n;
    return 0;
}

// coating.rs:101-129,341-356 as one multi-source BFS from the object cells through free cells
// (equivalent to the reference's per-cell BFS: moves are symmetric).  0xFFFF = no map entry.
void coat_dist_grid(sops_size s, std::vector<uint16_t>& dist) {
    const int a = s.arena, o = s.layers, c = s.coat, G = 2 * a + 1;
    static const int D[6][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, 1}};
    std::vector<uint8_t> kind((size_t)G * G, 0);                       // 0 free, 1 object, 2 boundary
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j)
            if (i - j > a || j - i > a) kind[(size_t)i * G + j] = 2;
    const int base = a - o + c;
    std::vector<int> q;
    for (int u = 0; u <= 2 * o; ++u)
        for (int v = 0; v <= 2 * o; ++v)
            if (u - v <= o && v - u <= o) { kind[(size_t)(base + u) * G + base + v] = 1; q.push_back((base + u) * G + base + v); }
    dist.assign((size_t)G * G, 0xFFFF);
    std::vector<uint16_t> d((size_t)G * G, 0);
    std::vector<uint8_t> seen((size_t)G * G, 0);
    for (int cell : q) seen[cell] = 1;
    size_t head = 0;
    while (head < q.size()) {
        int cell = q[head++], ci = cell / G, cj = cell % G;
        for (int k = 0; k < 6; ++k) {
            int ni = ci + D[k][0], nj = cj + D[k][1];
            if (ni < 0 || nj < 0 || ni >= G || nj >= G) continue;
            size_t nc = (size_t)ni * G + nj;
            if (seen[nc] || kind[nc] != 0) continue;
            seen[nc] = 1;
            d[nc] = (uint16_t)(d[cell] + 1);
            q.push_back((int)nc);
        }
    }
    for (size_t k = 0; k < dist.size(); ++k)
        if (kind[k] == 0) dist[k] = seen[k] ? d[k] : 0;                // unreachable free cell -> 0 (coating.rs:128)
}

struct Group {                      // one kernel launch: instances of one row-width class on one device
    int rw;
    uint32_t first, count;          // range in the device's instance array
    uint32_t plane_words, pos_words, words_per_warp;
    uint32_t off_p0, off_p1, off_p2, off_pos, off_thr, off_aux, off_tab, inv_words;
    uint32_t dump_stride;
    size_t dump_off;                // word offset of this group's dump area
    int wpb, blocks;
    int cta_w;                      // 0: warp-per-instance kernel; else warps per instance of the CTA kernel
    int rounds;                     // commit-round form of the warp kernel (sops::ROUNDS_*)
    size_t smem;
};

struct Dev {
    int id = 0, n_sm = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: the second row-width class of a mixed batch runs concurrently
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    // device buffers (grown on demand)
    Instance* d_inst = nullptr; size_t cap_inst = 0;
    uint8_t* d_genomes = nullptr; size_t cap_genomes = 0;
    DevResult* d_out = nullptr; size_t cap_out = 0;
    uint16_t* d_coat = nullptr; size_t cap_coat = 0;
    uint32_t* d_dump = nullptr; size_t cap_dump = 0;
    unsigned long long* d_counters = nullptr;   // 3 counters + 2 work counters (as u32 pairs)
    // pinned host staging
    Instance* h_inst = nullptr; size_t hcap_inst = 0;
    DevResult* h_out = nullptr; size_t hcap_out = 0;
    unsigned long long* h_counters = nullptr;
    uint32_t* h_dump = nullptr; size_t hcap_dump = 0;
    // batch
    uint32_t n_inst = 0;
    std::vector<uint32_t> global_of;            // local slot -> global instance id
    std::vector<Group> groups;
    bool launched = false;
    float last_ms = 0.f;
};

template <class T>
cudaError_t grow_dev(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 4 + 64;
    cudaError_t e = cudaMalloc((void**)p, want * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}
template <class T>
cudaError_t grow_host(T** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    size_t want = need + need / 4 + 64;
    cudaError_t e = cudaMallocHost((void**)p, want * sizeof(T));
    if (e == cudaSuccess) *cap = want;
    return e;
}

}  // namespace

struct sops_ctx {
    std::vector<Dev> devs;
    std::string err;
    // batch description
    bool prepared = false, collected = false;
    int behavior = 0, rng_mode = 0;
    uint32_t P = 0, T = 0, flags = 0;
    float w1 = 0, w2 = 0;
    uint8_t* h_genomes = nullptr; size_t hcap_genomes = 0;      // pinned
    uint16_t* h_coat = nullptr; size_t hcap_coat = 0;           // pinned
    size_t coat_elems = 0;
    std::vector<sops_size> sizes;
    std::vector<uint8_t> orient;                                // per instance (loco)
    std::vector<uint32_t> coat_min_total;                       // per trial
    std::vector<std::pair<int, uint32_t>> where;                // global instance -> (device index, local slot)
    uint16_t prob[16];
    uint64_t totals[6] = {0, 0, 0, 0, 0, 0};
    double last_p_acc[4] = {-1.0, -1.0, -1.0, -1.0};            // committed / attempts of the last collected batch, per behaviour
    std::vector<uint64_t> trial_steps;                          // sops_ctx_set_trial_steps: consumed by the next prepare
    uint64_t h2d_bytes = 0, d2h_bytes = 0, launches = 0;
};

namespace {

int fail(sops_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, SOPS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

typedef void (*kernel_fn)(const KParams);

template <int BEH, int ROUNDS>
kernel_fn pick_kernel(int rw, int mode) {
    if (rw == 1) return mode == 0 ? sops::sops_chain_kernel<BEH, 1, 0, ROUNDS> : sops::sops_chain_kernel<BEH, 1, 1, ROUNDS>;
    if (rw == 2) return mode == 0 ? sops::sops_chain_kernel<BEH, 2, 0, ROUNDS> : sops::sops_chain_kernel<BEH, 2, 1, ROUNDS>;
    if (rw == 4) return mode == 0 ? sops::sops_chain_kernel<BEH, 4, 0, ROUNDS> : sops::sops_chain_kernel<BEH, 4, 1, ROUNDS>;
    return mode == 0 ? sops::sops_chain_kernel<BEH, 8, 0, ROUNDS> : sops::sops_chain_kernel<BEH, 8, 1, ROUNDS>;
}
// row width class of a padded grid side: 32-bit words per lattice row
int rw_of(uint32_t pg) { return pg <= 32 ? 1 : pg <= 64 ? 2 : pg <= 128 ? 4 : 8; }
// CTA-per-chain variant: 4 warps per chain (opt-in, SOPS_F_FORCE_CTA; see DESIGN.md for when it pays)
kernel_fn kernel_for_cta(int beh, int rw) {
    switch (beh) {
    case SOPS_AGG: return rw == 1 ? sops::sops_chain_cta_kernel<sops::AGG, 1, 4> : sops::sops_chain_cta_kernel<sops::AGG, 2, 4>;
    case SOPS_SEP: return rw == 1 ? sops::sops_chain_cta_kernel<sops::SEP, 1, 4> : sops::sops_chain_cta_kernel<sops::SEP, 2, 4>;
    case SOPS_COAT: return rw == 1 ? sops::sops_chain_cta_kernel<sops::COAT, 1, 4> : sops::sops_chain_cta_kernel<sops::COAT, 2, 4>;
    default: return rw == 1 ? sops::sops_chain_cta_kernel<sops::LOCO, 1, 4> : sops::sops_chain_cta_kernel<sops::LOCO, 2, 4>;
    }
}
// rounds: sops::ROUNDS_PARALLEL or sops::ROUNDS_SERIAL (every behaviour), sops::ROUNDS_ADAPTIVE (fast mode)
// (separation keeps a cell -> particle map for it: one- and two-word rows only, 2 or 6 KB per chain)
bool has_parallel_rounds(int beh, int rw) { return beh != SOPS_SEP || rw <= 2; }
template <int ROUNDS>
kernel_fn pick_sep_narrow(int rw, int mode) {
    if (rw == 1) return mode == 0 ? sops::sops_chain_kernel<sops::SEP, 1, 0, ROUNDS> : sops::sops_chain_kernel<sops::SEP, 1, 1, ROUNDS>;
    return mode == 0 ? sops::sops_chain_kernel<sops::SEP, 2, 0, ROUNDS> : sops::sops_chain_kernel<sops::SEP, 2, 1, ROUNDS>;
}
// adaptive rounds: aggregation in fast mode only (the one behaviour whose calm chains gain from them)
// adaptive rounds: fast mode; every behaviour (separation: one- and two-word rows, like its parallel rounds)
bool has_adaptive_rounds(int beh, int mode, int rw) { return mode == SOPS_RNG_FAST && (beh != SOPS_SEP || rw <= 2); }
template <int BEH>
kernel_fn pick_adaptive(int rw) {
    if (rw == 1) return sops::sops_chain_kernel<BEH, 1, 1, sops::ROUNDS_ADAPTIVE>;
    if (rw == 2) return sops::sops_chain_kernel<BEH, 2, 1, sops::ROUNDS_ADAPTIVE>;
    if (rw == 4) return sops::sops_chain_kernel<BEH, 4, 1, sops::ROUNDS_ADAPTIVE>;
    return sops::sops_chain_kernel<BEH, 8, 1, sops::ROUNDS_ADAPTIVE>;
}
kernel_fn kernel_for(int beh, int rw, int mode, int rounds) {
    if (rounds == sops::ROUNDS_ADAPTIVE && has_adaptive_rounds(beh, mode, rw)) {
        if (beh == SOPS_SEP) return rw == 1 ? sops::sops_chain_kernel<sops::SEP, 1, 1, sops::ROUNDS_ADAPTIVE> : sops::sops_chain_kernel<sops::SEP, 2, 1, sops::ROUNDS_ADAPTIVE>;
        return beh == SOPS_LOCO ? pick_adaptive<sops::LOCO>(rw) : beh == SOPS_COAT ? pick_adaptive<sops::COAT>(rw) : pick_adaptive<sops::AGG>(rw);
    }
    const bool par = rounds != sops::ROUNDS_SERIAL && has_parallel_rounds(beh, rw);
    switch (beh) {
    case SOPS_AGG: return par ? pick_kernel<sops::AGG, sops::ROUNDS_PARALLEL>(rw, mode) : pick_kernel<sops::AGG, sops::ROUNDS_SERIAL>(rw, mode);
    case SOPS_SEP: return par ? pick_sep_narrow<sops::ROUNDS_PARALLEL>(rw, mode) : pick_kernel<sops::SEP, sops::ROUNDS_SERIAL>(rw, mode);
    case SOPS_COAT: return par ? pick_kernel<sops::COAT, sops::ROUNDS_PARALLEL>(rw, mode) : pick_kernel<sops::COAT, sops::ROUNDS_SERIAL>(rw, mode);
    default: return par ? pick_kernel<sops::LOCO, sops::ROUNDS_PARALLEL>(rw, mode) : pick_kernel<sops::LOCO, sops::ROUNDS_SERIAL>(rw, mode);
    }
}

// shared-memory layout of one warp-resident instance, sized for the largest instance of the group
void layout_group(Group& g, int beh, uint32_t max_pg, uint32_t max_n) {
    const bool three = (beh == SOPS_SEP || beh == SOPS_COAT);
    g.plane_words = ((max_pg * (uint32_t)g.rw) + 1u) & ~1u;     // keep 64-bit rows 8-byte aligned
    g.pos_words = ((max_n + 1) / 2 + 1u) & ~1u;
    uint32_t thr_words = (uint32_t)genome_len(beh) / 2;
    uint32_t off = 0;
    g.off_p0 = off; off += g.plane_words;
    g.off_p1 = off; off += g.plane_words;
    g.off_p2 = off; if (three) off += g.plane_words;
    g.off_pos = off; off += g.pos_words;
    g.off_thr = off; off += thr_words;
    g.off_aux = off; if (beh == SOPS_LOCO) off += SOPS_LOCO_AUX_WORDS;
    g.off_tab = off; if (beh == SOPS_AGG) off += SOPS_AGG_TAB_WORDS;
    g.inv_words = 0;
    if (beh == SOPS_SEP && g.rounds != sops::ROUNDS_SERIAL) {     // cell -> particle map, u16 per cell, 32 * rw cells per row
        g.inv_words = (max_pg * 32u * (uint32_t)g.rw + 1u) / 2u;
        off += g.inv_words;
    }
    off = (off + 1u) & ~1u;
    off |= 2u;                                                  // stride = 2 mod 4 words: spreads warps over banks, keeps 8-byte alignment
    g.words_per_warp = off;
    g.dump_stride = 3 * g.plane_words + g.pos_words + (beh == SOPS_LOCO ? SOPS_LIGHT_WORDS : 0u);
}

}  // namespace

extern "C" {

int sops_ctx_create(sops_ctx** out, const int* device_ids, int n_dev) {
    if (!out || n_dev < 1) return SOPS_E_ARG;
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return SOPS_E_CUDA;
    sops_ctx* ctx = new sops_ctx();
    ctx->devs.resize((size_t)n_dev);
    for (int k = 0; k < n_dev; ++k) {
        Dev& d = ctx->devs[(size_t)k];
        d.id = device_ids ? device_ids[k] : k;
        if (d.id < 0 || d.id >= have) { delete ctx; return SOPS_E_ARG; }
        cudaError_t e = cudaSetDevice(d.id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&d.ev1);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, d.id);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d.d_counters, 8 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&d.h_counters, 8 * sizeof(unsigned long long));
        if (e != cudaSuccess) { sops_ctx_destroy(ctx); return SOPS_E_CUDA; }
    }
    *out = ctx;
    return 0;
}

void sops_ctx_destroy(sops_ctx* ctx) {
    if (!ctx) return;
    for (Dev& d : ctx->devs) {
        cudaSetDevice(d.id);
        if (d.stream) cudaStreamSynchronize(d.stream);
        cudaFree(d.d_inst); cudaFree(d.d_genomes); cudaFree(d.d_out); cudaFree(d.d_coat); cudaFree(d.d_dump);
        cudaFree(d.d_counters);
        cudaFreeHost(d.h_inst); cudaFreeHost(d.h_out); cudaFreeHost(d.h_counters); cudaFreeHost(d.h_dump);
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.ev_join) cudaEventDestroy(d.ev_join);
        if (d.stream2) cudaStreamDestroy(d.stream2);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    cudaFreeHost(ctx->h_genomes);
    cudaFreeHost(ctx->h_coat);
    delete ctx;
}

const char* sops_last_error(const sops_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sops_ctx_set_trial_steps(sops_ctx* ctx, const uint64_t* steps, uint32_t T) {
    if (!ctx) return SOPS_E_ARG;
    ctx->trial_steps.clear();
    if (steps && T) ctx->trial_steps.assign(steps, steps + T);
    return 0;
}

int sops_instance_shape(int behavior, sops_size size, uint32_t* grid_side, uint32_t* n_particles, uint64_t* sim_duration) {
    Shape s;
    int rc = shape_of(behavior, size, &s);
    if (rc) return rc;
    if (grid_side) *grid_side = s.G;
    if (n_particles) *n_particles = s.n;
    if (sim_duration) *sim_duration = s.dur;
    return 0;
}

int sops_coat_distance_grid(sops_size size, uint16_t* dist) {
    if (!dist) return SOPS_E_ARG;
    Shape s;
    int rc = shape_of(SOPS_COAT, size, &s);
    if (rc) return rc;
    std::vector<uint16_t> d;
    coat_dist_grid(size, d);
    memcpy(dist, d.data(), d.size() * sizeof(uint16_t));
    return 0;
}

int sops_batch_prepare(sops_ctx* ctx, int behavior, const uint8_t* genomes, uint32_t P,
                       const sops_size* trial_sizes, const uint64_t* trial_seeds, uint32_t T,
                       uint8_t gran, float w1, float w2, int rng_mode,
                       const uint64_t* move_seeds, const uint8_t* loco_orient,
                       uint64_t steps_override, uint32_t flags) {
    if (!ctx) return SOPS_E_ARG;
    (void)gran;   // stored but never read by SOPSCore (mod.rs:34); gene range is checked against the table
    ctx->prepared = false; ctx->collected = false;
    if (!genomes || !trial_sizes || !trial_seeds || P == 0 || T == 0) return fail(ctx, SOPS_E_ARG, "null pointer or empty batch");
    if (behavior < SOPS_AGG || behavior > SOPS_LOCO) return fail(ctx, SOPS_E_ARG, "bad behavior");
    if (rng_mode != SOPS_RNG_VERIFY && rng_mode != SOPS_RNG_FAST) return fail(ctx, SOPS_E_ARG, "bad rng_mode");
    if ((uint64_t)P * T > 0x7fffffffull) return fail(ctx, SOPS_E_ARG, "batch too large");
    const int L = genome_len(behavior);
    const uint32_t N = P * T;

    // probability table (gene_probability / theory_gene_probability)
    int plen;
    memset(ctx->prob, 0, sizeof(ctx->prob));
    if (!(flags & SOPS_F_THEORY_TABLE)) { plen = 11; memcpy(ctx->prob, PROB_GA, sizeof(PROB_GA)); }
    else if (behavior == SOPS_SEP || behavior == SOPS_COAT) { plen = 16; memcpy(ctx->prob, PROB_THEORY_SEP, sizeof(PROB_THEORY_SEP)); }
    else { plen = 6; memcpy(ctx->prob, PROB_THEORY_AGG, sizeof(PROB_THEORY_AGG)); }
    for (size_t k = 0; k < (size_t)P * L; ++k)
        if (genomes[k] >= plen) return fail(ctx, SOPS_E_GENE, "gene value indexes past the probability table");

    // shapes
    std::vector<Shape> shp(T);
    for (uint32_t t = 0; t < T; ++t) {
        int rc = shape_of(behavior, trial_sizes[t], &shp[t]);
        if (rc) return fail(ctx, rc, "invalid size tuple at trial " + std::to_string(t));
        if (shp[t].G + 4 > 255) return fail(ctx, SOPS_E_UNSUPPORTED, "arena_layers > 125 not implemented by the kernels (u8 padded coordinates)");
        if (behavior == SOPS_LOCO && trial_sizes[t].arena > 63)
            return fail(ctx, SOPS_E_UNSUPPORTED, "loco is defined for arena_layers <= 63 (u8 arithmetic in locomotion.rs wraps beyond)");
            }
    if (loco_orient && behavior == SOPS_LOCO)
        for (uint32_t q = 0; q < N; ++q)
            if (loco_orient[q] < 1 || loco_orient[q] > 3) return fail(ctx, SOPS_E_ARG, "loco orientation must be 1..3");

    ctx->behavior = behavior; ctx->rng_mode = rng_mode; ctx->P = P; ctx->T = T; ctx->flags = flags;
    ctx->w1 = w1; ctx->w2 = w2;
    ctx->sizes.assign(trial_sizes, trial_sizes + T);
    ctx->orient.assign(N, 1);
    if (loco_orient) ctx->orient.assign(loco_orient, loco_orient + N);

    // coating: one distance table per distinct size (coat_ga.rs:508-537 precomputes the same thing)
    std::vector<uint32_t> coat_off(T, 0);
    ctx->coat_min_total.assign(T, 0);
    ctx->coat_elems = 0;
    if (behavior == SOPS_COAT) {
        std::vector<std::vector<uint16_t>> tabs;
        std::vector<uint32_t> tab_of(T);
        for (uint32_t t = 0; t < T; ++t) {
            uint32_t found = UINT32_MAX;
            for (uint32_t u = 0; u < t; ++u)
                if (trial_sizes[u].arena == trial_sizes[t].arena && trial_sizes[u].layers == trial_sizes[t].layers &&
                    trial_sizes[u].coat == trial_sizes[t].coat) { found = u; break; }
            if (found != UINT32_MAX) { coat_off[t] = coat_off[found]; ctx->coat_min_total[t] = ctx->coat_min_total[found]; continue; }
            std::vector<uint16_t> d;
            coat_dist_grid(trial_sizes[t], d);
            coat_off[t] = (uint32_t)ctx->coat_elems;
            // get_min_total_dist (coating.rs:377-392): sum of the n smallest map entries
            std::vector<uint16_t> v;
            for (uint16_t x : d) if (x != 0xFFFF) v.push_back(x);
            std::sort(v.begin(), v.end());
            uint32_t mt = 0;
            for (uint32_t k = 0; k < shp[t].n && k < v.size(); ++k) mt += v[k];
            ctx->coat_min_total[t] = mt;
            ctx->coat_elems += d.size();
            tabs.push_back(std::move(d));
        }
        CK(grow_host(&ctx->h_coat, &ctx->hcap_coat, ctx->coat_elems));
        size_t off = 0;
        for (auto& d : tabs) { memcpy(ctx->h_coat + off, d.data(), d.size() * sizeof(uint16_t)); off += d.size(); }
    }

    CK(grow_host(&ctx->h_genomes, &ctx->hcap_genomes, (size_t)P * L));
    memcpy(ctx->h_genomes, genomes, (size_t)P * L);

    // cost-sorted instance order (longest chain first), then LPT over devices
    std::vector<uint64_t> per_trial;
    per_trial.swap(ctx->trial_steps);                          // consumed by this call
    if (!per_trial.empty() && per_trial.size() != T) return fail(ctx, SOPS_E_ARG, "sops_ctx_set_trial_steps: length differs from T");
    std::vector<uint64_t> steps_of(T);
    for (uint32_t t = 0; t < T; ++t) {
        steps_of[t] = (flags & SOPS_F_INIT_ONLY) ? 0 : !per_trial.empty() ? per_trial[t] : (steps_override ? steps_override : shp[t].dur);
    }
    std::vector<uint32_t> trial_order(T);
    for (uint32_t t = 0; t < T; ++t) trial_order[t] = t;
    std::stable_sort(trial_order.begin(), trial_order.end(), [&](uint32_t x, uint32_t y) { return steps_of[x] > steps_of[y]; });

    const size_t nd = ctx->devs.size();
    std::vector<uint64_t> load(nd, 0);
    std::vector<std::vector<uint32_t>> assigned(nd);       // global ids per device, already cost-descending
    ctx->where.assign(N, std::make_pair(0, 0u));
    size_t rr = 0;
    for (uint32_t ti = 0; ti < T; ++ti) {
        uint32_t t = trial_order[ti];
        for (uint32_t g = 0; g < P; ++g) {
            size_t best = 0;
            if (nd > 1) {
                best = rr % nd;
                for (size_t k = 0; k < nd; ++k) { size_t cand = (rr + k) % nd; if (load[cand] < load[best]) best = cand; }
                ++rr;
            }
            load[best] += steps_of[t] + 1;
            assigned[best].push_back(g * T + t);
        }
    }

    ctx->h2d_bytes = 0;
    for (size_t di = 0; di < nd; ++di) {
        Dev& d = ctx->devs[di];
        CK(cudaSetDevice(d.id));
        d.launched = false;
        d.n_inst = (uint32_t)assigned[di].size();
        d.global_of = assigned[di];
        d.groups.clear();
        if (d.n_inst == 0) continue;
        // split by row-width class (stable: keeps the cost order inside each class)
        std::vector<uint32_t> cls[4];
        for (uint32_t q : d.global_of) { const int rw = rw_of(shp[q % T].G + 4); cls[rw == 1 ? 0 : rw == 2 ? 1 : rw == 4 ? 2 : 3].push_back(q); }
        d.global_of.clear();
        CK(grow_host(&d.h_inst, &d.hcap_inst, (size_t)d.n_inst));
        size_t dump_words = 0;
        for (int c = 0; c < 4; ++c) {
            if (cls[c].empty()) continue;
            Group grp;
            memset(&grp, 0, sizeof(grp));
            grp.rw = 1 << c;
            grp.first = (uint32_t)d.global_of.size();
            grp.count = (uint32_t)cls[c].size();
            uint32_t max_pg = 0, max_n = 0;
            for (uint32_t q : cls[c]) {
                uint32_t t = q % T;
                max_pg = std::max(max_pg, shp[t].G + 4);
                max_n = std::max(max_n, shp[t].n);
                Instance in;
                memset(&in, 0, sizeof(in));
                in.seed = trial_seeds[t];
                in.move_seed = move_seeds ? move_seeds[q] : trial_seeds[t];
                in.steps = steps_of[t];
                in.genome = q / T;
                in.slot = (uint32_t)d.global_of.size();
                in.aux = coat_off[t];
                in.arena = (uint8_t)trial_sizes[t].arena;
                in.layers = (uint8_t)trial_sizes[t].layers;
                in.coat = (uint8_t)trial_sizes[t].coat;
                in.orient = ctx->orient[q];
                d.h_inst[in.slot] = in;
                ctx->where[q] = std::make_pair((int)di, in.slot);
                d.global_of.push_back(q);
            }
            // kernel choice: one warp per chain, unless the caller asks for the 4-warp CTA-per-chain variant
            // (fast mode only; it pays when few chains with very low acceptance run, i.e. late GA generations)
            grp.cta_w = (rng_mode == SOPS_RNG_FAST && (flags & SOPS_F_FORCE_CTA) && !(flags & SOPS_F_FORCE_WARP) && grp.rw <= 2) ? 4 : 0;
            // How many warps per SM?  The list is cost-sorted and a GA batch is skewed (chain length goes with n^3), and a
            // chain is a dependent instruction stream that slows down with every other chain on its warp scheduler.  So
            // no more warps are resident than it takes to start every LONG chain (at least half the longest) at once,
            // rounded up to whole schedulers; the shorter chains are pulled as warps free up.  C2 (250 long of 750): one
            // warp per scheduler; coating pop 200 (1000 long of 2000): 8 per SM, 13 % faster than 16; aggregation pop 400
            // (2000 long of 6000): 16 per SM, 25 % faster than 24.  Equal-cost batches (seed sweeps) fill the machine.
            uint32_t W;                                         // resident warps per SM this batch asks for
            {
                double longest = 0;
                for (uint32_t q : cls[c]) longest = std::max(longest, (double)steps_of[q % T]);
                uint32_t n_long = 0;
                for (uint32_t q : cls[c]) if ((double)steps_of[q % T] * 2.0 >= longest) ++n_long;
                W = (n_long + (uint32_t)d.n_sm - 1) / (uint32_t)d.n_sm;               // warps per SM to hold the long chains
                W = std::max(4u, (W + 3u) & ~3u);
                if (grp.count < 4u * (uint32_t)d.n_sm) W = std::max(1u, (grp.count + (uint32_t)d.n_sm - 1) / (uint32_t)d.n_sm);
            }
            // commit rounds of the warp kernel: all accepted lanes at once where the behaviour has that form (the caller
            // may force a form; the bits are the same)
            grp.rounds = has_parallel_rounds(behavior, grp.rw) ? sops::ROUNDS_PARALLEL : sops::ROUNDS_SERIAL;
            // aggregation chains that accept little (a converged GA population, seed sweeps of an evolved genome) run
            // 7-10 % faster through the serial form; the acceptance rate of this context's previous aggregation batch is
            // the predictor (unknown: parallel)
            if (behavior == SOPS_AGG && ctx->last_p_acc[SOPS_AGG] >= 0.0 && ctx->last_p_acc[SOPS_AGG] < 0.02) grp.rounds = sops::ROUNDS_SERIAL;
            // ... and with at most two chains per warp scheduler (a GA generation) each chain picks its form itself
            // (fast mode): a late-generation aggregation population of 50 / 100 / 200
            // runs 40 / 42 / 25 % faster than through the serial form, a generation-0 one as fast as through the parallel
            // form (population 400, four per scheduler: 3-7 % slower)
            if (has_adaptive_rounds(behavior, rng_mode, grp.rw) && W <= 8u) grp.rounds = sops::ROUNDS_ADAPTIVE;
            // (locomotion, coating, separation: the single-batch loop of their adaptive kernels is 1-2 % slower than the
            // parallel-rounds kernel's, so a population whose previous batch was hot on average stays with that one;
            // aggregation's adaptive kernel is the faster one on hot chains too)
            if (grp.rounds == sops::ROUNDS_ADAPTIVE && behavior != SOPS_AGG && ctx->last_p_acc[behavior] >= 0.06 &&
                has_parallel_rounds(behavior, grp.rw))
                grp.rounds = sops::ROUNDS_PARALLEL;
            if (flags & SOPS_F_ROUNDS_SERIAL) grp.rounds = sops::ROUNDS_SERIAL;
            if ((flags & SOPS_F_ROUNDS_PARALLEL) && has_parallel_rounds(behavior, grp.rw)) grp.rounds = sops::ROUNDS_PARALLEL;
            if (flags & SOPS_F_ROUNDS_ADAPTIVE)
                grp.rounds = has_adaptive_rounds(behavior, rng_mode, grp.rw) ? sops::ROUNDS_ADAPTIVE
                           : has_parallel_rounds(behavior, grp.rw) ? sops::ROUNDS_PARALLEL : sops::ROUNDS_SERIAL;
            if (grp.cta_w) grp.rounds = sops::ROUNDS_SERIAL;
            layout_group(grp, behavior, max_pg, max_n);
            grp.dump_off = dump_words;
            if (flags & SOPS_F_DUMP) dump_words += (size_t)grp.count * grp.dump_stride;
            const size_t smem_per_warp = (size_t)grp.words_per_warp * 4;
            int per_sm = 0;
            if (grp.cta_w) {
                kernel_fn fn = kernel_for_cta(behavior, grp.rw);
                grp.wpb = grp.cta_w;
                grp.smem = smem_per_warp;
                if (grp.smem + 1024 > 48 * 1024)                   // + the kernels' static shared memory
                    CK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grp.smem));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, grp.cta_w * 32, grp.smem));
                if (per_sm < 1) return fail(ctx, SOPS_E_UNSUPPORTED, "instance does not fit in shared memory");
                grp.blocks = (int)std::min(grp.count, (uint32_t)per_sm * (uint32_t)d.n_sm);
            } else {
                // launch geometry: spread few instances over all SMs, otherwise fill the machine with
                // 8-warp CTAs that pull instances from a work counter
                kernel_fn fn = kernel_for(behavior, grp.rw, rng_mode, grp.rounds);
                int wpb = 8;
                uint32_t resident_cap = 0;                      // CTAs; 0 = whatever fits
                if (W <= 16u) {
                    wpb = (int)std::min(W, 8u);
                    resident_cap = (uint32_t)d.n_sm * ((W + (uint32_t)wpb - 1) / (uint32_t)wpb);
                }
                if (wpb < 1) wpb = 1;
                while (wpb > 1 && smem_per_warp * (size_t)wpb + 1024 > 227u * 1024u) wpb /= 2;   // largest arenas: fewer warps per CTA
                grp.wpb = wpb;
                grp.smem = smem_per_warp * (size_t)wpb;
                if (grp.smem + 1024 > 48 * 1024)                   // + the kernels' static shared memory
                    CK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grp.smem));
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, wpb * 32, grp.smem));
                if (per_sm < 1) return fail(ctx, SOPS_E_UNSUPPORTED, "instance does not fit in shared memory");
                uint32_t want = (grp.count + (uint32_t)wpb - 1) / (uint32_t)wpb;
                // a batch that fits the machine once is spread over ALL SMs (warps pull chains from the work counter, the
                // few surplus warps exit at once): 1000 chains run as 148 x 7 warps, not 143 x 7 with five SMs idle
                if (grp.count >= (uint32_t)d.n_sm && want < (uint32_t)d.n_sm) want = (uint32_t)d.n_sm;
                uint32_t cap = (uint32_t)per_sm * (uint32_t)d.n_sm;
                if (resident_cap) cap = std::min(cap, resident_cap);
                grp.blocks = (int)std::min(want, cap);
            }
            d.groups.push_back(grp);
        }
        CK(grow_dev(&d.d_inst, &d.cap_inst, (size_t)d.n_inst));
        CK(grow_dev(&d.d_out, &d.cap_out, (size_t)d.n_inst));
        CK(grow_host(&d.h_out, &d.hcap_out, (size_t)d.n_inst));
        CK(grow_dev(&d.d_genomes, &d.cap_genomes, (size_t)P * L));
        CK(cudaMemcpyAsync(d.d_inst, d.h_inst, (size_t)d.n_inst * sizeof(Instance), cudaMemcpyHostToDevice, d.stream));
        CK(cudaMemcpyAsync(d.d_genomes, ctx->h_genomes, (size_t)P * L, cudaMemcpyHostToDevice, d.stream));
        ctx->h2d_bytes += (uint64_t)d.n_inst * sizeof(Instance) + (uint64_t)P * L;
        if (behavior == SOPS_COAT) {
            CK(grow_dev(&d.d_coat, &d.cap_coat, ctx->coat_elems));
            CK(cudaMemcpyAsync(d.d_coat, ctx->h_coat, ctx->coat_elems * sizeof(uint16_t), cudaMemcpyHostToDevice, d.stream));
            ctx->h2d_bytes += ctx->coat_elems * sizeof(uint16_t);
        }
        if (flags & SOPS_F_DUMP) {
            CK(grow_dev(&d.d_dump, &d.cap_dump, dump_words));
            CK(grow_host(&d.h_dump, &d.hcap_dump, dump_words));
        }
    }
    ctx->prepared = true;
    return 0;
}

int sops_batch_launch(sops_ctx* ctx) {
    if (!ctx) return SOPS_E_ARG;
    if (!ctx->prepared) return fail(ctx, SOPS_E_STATE, "launch before prepare");
    ctx->collected = false;
    ctx->launches = 0;
    for (Dev& d : ctx->devs) {
        if (d.n_inst == 0) continue;
        CK(cudaSetDevice(d.id));
        CK(cudaMemsetAsync(d.d_counters, 0, 8 * sizeof(unsigned long long), d.stream));
        CK(cudaEventRecord(d.ev0, d.stream));
        if (d.groups.size() > 1) {                           // fork point: after the counter reset, before any chain kernel
            CK(cudaEventRecord(d.ev_fork, d.stream));
            CK(cudaStreamWaitEvent(d.stream2, d.ev_fork, 0));
        }
        for (size_t gi = 0; gi < d.groups.size(); ++gi) {
            const Group& g = d.groups[gi];
            KParams kp;
            memset(&kp, 0, sizeof(kp));
            kp.inst = d.d_inst + g.first;
            kp.n_inst = g.count;
            kp.genomes = d.d_genomes;
            kp.genome_len = (uint32_t)genome_len(ctx->behavior);
            memcpy(kp.prob, ctx->prob, sizeof(kp.prob));
            kp.work_counter = reinterpret_cast<uint32_t*>(d.d_counters + 4) + gi;
            kp.out = d.d_out;
            kp.counters = d.d_counters;
            kp.dump = (ctx->flags & SOPS_F_DUMP) ? d.d_dump + g.dump_off : nullptr;
            kp.dump_stride = g.dump_stride;
            kp.dump_slot0 = g.first;
            kp.coat_dist = d.d_coat;
            kp.words_per_warp = g.words_per_warp;
            kp.off_p0 = g.off_p0; kp.off_p1 = g.off_p1; kp.off_p2 = g.off_p2;
            kp.off_pos = g.off_pos; kp.off_thr = g.off_thr; kp.off_aux = g.off_aux; kp.off_tab = g.off_tab;
            kp.plane_words = g.plane_words; kp.pos_words = g.pos_words;
            kp.inv_words = g.inv_words;
            kernel_fn fn = g.cta_w ? kernel_for_cta(ctx->behavior, g.rw) : kernel_for(ctx->behavior, g.rw, ctx->rng_mode, g.rounds);
            cudaStream_t st = d.stream;
            if (gi > 0) st = d.stream2;                      // independent chains: the two classes run concurrently
            fn<<<g.blocks, g.wpb * 32, g.smem, st>>>(kp);
            CK(cudaGetLastError());
            if (gi > 0) {
                CK(cudaEventRecord(d.ev_join, st));
                CK(cudaStreamWaitEvent(d.stream, d.ev_join, 0));
            }
            ++ctx->launches;
        }
        CK(cudaEventRecord(d.ev1, d.stream));
        d.launched = true;
    }
    return 0;
}

int sops_batch_sync(sops_ctx* ctx, double* device_ms) {
    if (!ctx) return SOPS_E_ARG;
    if (!ctx->prepared) return fail(ctx, SOPS_E_STATE, "sync before prepare");
    double worst = 0.0;
    for (Dev& d : ctx->devs) {
        if (d.n_inst == 0) continue;
        if (!d.launched) return fail(ctx, SOPS_E_STATE, "sync before launch");
        CK(cudaSetDevice(d.id));
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaEventElapsedTime(&d.last_ms, d.ev0, d.ev1));
        worst = std::max(worst, (double)d.last_ms);
    }
    if (device_ms) *device_ms = worst;
    return 0;
}

int sops_batch_collect(sops_ctx* ctx, sops_result* out) {
    if (!ctx || !out) return SOPS_E_ARG;
    if (!ctx->prepared) return fail(ctx, SOPS_E_STATE, "collect before prepare");
    ctx->d2h_bytes = 0;
    for (Dev& d : ctx->devs) {
        if (d.n_inst == 0) continue;
        if (!d.launched) return fail(ctx, SOPS_E_STATE, "collect before launch");
        CK(cudaSetDevice(d.id));
        CK(cudaMemcpyAsync(d.h_out, d.d_out, (size_t)d.n_inst * sizeof(DevResult), cudaMemcpyDeviceToHost, d.stream));
        CK(cudaMemcpyAsync(d.h_counters, d.d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
        ctx->d2h_bytes += (uint64_t)d.n_inst * sizeof(DevResult) + 8 * sizeof(unsigned long long);
        if (ctx->flags & SOPS_F_DUMP) {
            size_t words = 0;
            for (const Group& g : d.groups) words += (size_t)g.count * g.dump_stride;
            CK(cudaMemcpyAsync(d.h_dump, d.d_dump, words * 4, cudaMemcpyDeviceToHost, d.stream));
            ctx->d2h_bytes += words * 4;
        }
    }
    uint64_t tot[3] = {0, 0, 0};
    bool init_reject = false;
    for (Dev& d : ctx->devs) {
        if (d.n_inst == 0) continue;
        CK(cudaSetDevice(d.id));
        CK(cudaStreamSynchronize(d.stream));
        CK(cudaEventElapsedTime(&d.last_ms, d.ev0, d.ev1));
        for (int k = 0; k < 3; ++k) tot[k] += d.h_counters[k];
        for (uint32_t j = 0; j < d.n_inst; ++j) {
            const DevResult r = d.h_out[j];
            const uint32_t q = d.global_of[j], t = q % ctx->T;
            const sops_size sz = ctx->sizes[t];
            const uint32_t p = sz.layers;
            sops_result o;
            memset(&o, 0, sizeof(o));
            if (r.status) init_reject = true;
            switch (ctx->behavior) {
            case SOPS_AGG: {                                   // mod.rs:296-301, base_ga.rs:409
                uint32_t edges = r.i0 / 2;
                uint64_t mx = (uint64_t)(3 * p) * (3 * p + 1);
                o.i0 = edges; o.i1 = (uint32_t)mx;
                o.f = (float)edges / (float)mx;
                o.norm = (double)edges / (double)mx;
            } break;
            case SOPS_SEP: {                                   // separation.rs:811-825
                float max_c1 = (float)(uint16_t)(3 * (3 * p * p + p - 1));
                float max_c2 = (float)(uint16_t)(3 * p * p - p - 1);
                float e = ((float)r.i0) / 2.0f;
                float s = ((float)r.i1) / (2.0f * 3.0f);
                float c1 = e / max_c1, c2 = s / max_c2;
                float f = ctx->w1 * c1 + ctx->w2 * c2;
                o.i0 = r.i0; o.i1 = r.i1; o.f = f; o.norm = (double)f;
            } break;
            case SOPS_COAT: {                                  // coating.rs:561-574
                uint32_t mt = ctx->coat_min_total[t];
                float f = (float)mt / (float)r.i1;
                o.i0 = mt; o.i1 = r.i1; o.f = f; o.norm = (double)f;
            } break;
            default: {                                         // locomotion.rs:661-688, 172-173
                Shape sh;
                shape_of(SOPS_LOCO, sz, &sh);
                float max_c1 = (float)(uint16_t)(3 * (3 * p * p + p - 1));
                float np = (float)sh.n;
                float inner = 1.0f + 4.0f * (np - 1.0f) / 3.0f;
                float max_c2 = (float)sh.G - ((-1.0f + sqrtf(inner)) / 2.0f);
                float c1 = (float)(r.i0 / 2) / max_c1;
                float c2 = (float)(r.i1 / sh.n) / max_c2;
                float f = ctx->w1 * c1 + ctx->w2 * c2;
                o.i0 = r.i0; o.i1 = r.i1; o.i2 = r.i2; o.f = f; o.norm = (double)f;
            } break;
            }
            out[q] = o;
        }
    }
    ctx->totals[0] = tot[0]; ctx->totals[1] = tot[1]; ctx->totals[2] = tot[2];
    if (tot[0] >= 1000000) ctx->last_p_acc[ctx->behavior] = (double)tot[2] / (double)tot[0];
    ctx->totals[3] = ctx->launches; ctx->totals[4] = ctx->h2d_bytes; ctx->totals[5] = ctx->d2h_bytes;
    ctx->collected = true;
    if (init_reject) return fail(ctx, SOPS_E_INIT_REJECT, "a placement draw hit the Uniform zone rejection");
    return 0;
}

int sops_eval_batch(sops_ctx* ctx, int behavior, const uint8_t* genomes, uint32_t P,
                    const sops_size* trial_sizes, const uint64_t* trial_seeds, uint32_t T,
                    uint8_t gran, float w1, float w2, int rng_mode,
                    const uint64_t* move_seeds, const uint8_t* loco_orient,
                    uint64_t steps_override, uint32_t flags, sops_result* out) {
    if (!out) return fail(ctx, SOPS_E_ARG, "null out");
    int rc = sops_batch_prepare(ctx, behavior, genomes, P, trial_sizes, trial_seeds, T, gran, w1, w2, rng_mode,
                                move_seeds, loco_orient, steps_override, flags);
    if (rc) return rc;
    rc = sops_batch_launch(ctx);
    if (rc) return rc;
    return sops_batch_collect(ctx, out);
}

int sops_batch_counters(sops_ctx* ctx, uint64_t out[6]) {
    if (!ctx || !out) return SOPS_E_ARG;
    if (!ctx->collected) return fail(ctx, SOPS_E_STATE, "counters before collect");
    for (int k = 0; k < 6; ++k) out[k] = ctx->totals[k];
    return 0;
}

int sops_dump_state(sops_ctx* ctx, uint32_t instance, uint8_t* grid, uint8_t* particles) {
    if (!ctx) return SOPS_E_ARG;
    if (!ctx->collected || !(ctx->flags & SOPS_F_DUMP)) return fail(ctx, SOPS_E_STATE, "dump needs a collected batch run with SOPS_F_DUMP");
    if (instance >= ctx->P * ctx->T) return fail(ctx, SOPS_E_ARG, "instance out of range");
    const std::pair<int, uint32_t> w = ctx->where[instance];
    const Dev& d = ctx->devs[(size_t)w.first];
    const Group* grp = nullptr;
    for (const Group& g : d.groups) if (w.second >= g.first && w.second < g.first + g.count) grp = &g;
    if (!grp) return fail(ctx, SOPS_E_STATE, "instance not found");
    const uint32_t* src = d.h_dump + grp->dump_off + (size_t)(w.second - grp->first) * grp->dump_stride;
    const uint32_t pw = grp->plane_words, rw = (uint32_t)grp->rw;
    const sops_size sz = ctx->sizes[instance % ctx->T];
    Shape sh;
    shape_of(ctx->behavior, sz, &sh);
    const int a = sz.arena, G = (int)sh.G;
    auto bit = [&](uint32_t plane, int r, int c) -> uint32_t {
        return (src[plane * pw + (uint32_t)r * rw + ((uint32_t)c >> 5)] >> (c & 31)) & 1u;
    };
    if (grid) {
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j) {
                uint8_t v;
                const bool bnd = (i - j > a) || (j - i > a);
                switch (ctx->behavior) {
                case SOPS_SEP: v = bnd ? 4 : (uint8_t)(bit(0, i + 2, j + 2) | (bit(1, i + 2, j + 2) << 1)); break;
                case SOPS_COAT: v = bnd ? 7 : bit(2, i + 2, j + 2) ? 2 : (uint8_t)bit(0, i + 2, j + 2); break;
                default: v = bnd ? 2 : (uint8_t)bit(0, i + 2, j + 2); break;
                }
                grid[(size_t)i * G + j] = v;
            }
    }
    if (particles) {
        const uint16_t* pos = reinterpret_cast<const uint16_t*>(src + 3 * pw);
        const uint32_t n3 = sh.n / 3;
        for (uint32_t k = 0; k < sh.n; ++k) {
            particles[3 * k] = (uint8_t)((pos[k] & 0xff) - 2);
            particles[3 * k + 1] = (uint8_t)((pos[k] >> 8) - 2);
            particles[3 * k + 2] = ctx->behavior == SOPS_SEP ? (uint8_t)(1 + (k >= n3) + (k >= 2 * n3)) : 0;
        }
    }
    return 0;
}

}  // extern "C"
