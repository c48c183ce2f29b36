// sops_core.hpp — C++ host mirror of the reference's SOPSCore module API over the C ABI.
//
// The reference (Rust) creates one environment per (genome, size, seed) inside nested rayon par_iters:
//   SOPSEnvironment::init_sops_env(&genome, arena, particle, seed, gran)           src/SOPSCore/mod.rs:106
//   SOPSepEnvironment::init_sops_env(.., w1, w2)                                   separation.rs:100
//   SOPSCoatEnvironment::init_sops_env(.., object, coat, .., dist)                 coating.rs:229
//   SOPSLocoEnvironment::init_sops_env(.., w1, w2)                                 locomotion.rs:98
//   .simulate(take_snaps) .evaluate_fitness() .get_max_fitness() .get_participant_cnt() .print_grid()
// Here the same names exist for single instances (each call is one tiny batch), and `Evaluator` is the
// batched form the GA uses: one eval() per generation instead of the nested par_iter.
// Rust is not available in this image, so this mirror is C++; the Rust binding is in INTEGRATION.md.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/evosops.h"

namespace evosops {

enum Behavior { Agg = SOPS_AGG, Sep = SOPS_SEP, Coat = SOPS_COAT, Loco = SOPS_LOCO };

inline int genome_len(int b) { return b == Agg ? 48 : b == Loco ? 144 : 600; }
inline const char* behavior_debug_name(int b) { return b == Agg ? "Agg" : b == Sep ? "Sep" : b == Coat ? "Coat" : "Loco"; }

struct Trial { sops_size size; uint64_t seed; };

// What evaluates a population: the CUDA library in the product, a stub in the host-logic tests.
class Evaluator {
public:
    virtual ~Evaluator() {}
    // out[g * trials.size() + t]; move_seeds / orient may be empty (defaults of include/evosops.h)
    virtual void eval(int behavior, const std::vector<uint8_t>& genomes, uint32_t P, const std::vector<Trial>& trials,
                      uint8_t gran, float w1, float w2, const std::vector<uint64_t>& move_seeds,
                      const std::vector<uint8_t>& orient, uint64_t steps_override, uint32_t flags,
                      std::vector<sops_result>& out) = 0;
    // final lattice (reference cell codes) + particles of instance q of the last eval() run with SOPS_F_DUMP
    virtual void dump(uint32_t instance, std::vector<uint8_t>& grid, std::vector<uint8_t>& particles, uint32_t G, uint32_t n) = 0;
    // attempts / possible / committed / launches / h2d bytes / d2h bytes of the last eval() (zeros if unknown)
    virtual void counters(uint64_t out[6]) { for (int k = 0; k < 6; ++k) out[k] = 0; }
    // per-trial chain lengths for the next eval() (sops_ctx_set_trial_steps); false if the evaluator cannot do it
    virtual bool set_trial_steps(const std::vector<uint64_t>&) { return false; }
};

class CudaEvaluator : public Evaluator {
public:
    explicit CudaEvaluator(int n_dev = 1, int rng_mode = SOPS_RNG_FAST) : rng_mode_(rng_mode) {
        int rc = sops_ctx_create(&ctx_, nullptr, n_dev);
        if (rc) throw std::runtime_error("sops_ctx_create failed (no CUDA device? there is no CPU fallback), rc=" + std::to_string(rc));
    }
    ~CudaEvaluator() override { sops_ctx_destroy(ctx_); }
    void eval(int behavior, const std::vector<uint8_t>& genomes, uint32_t P, const std::vector<Trial>& trials,
              uint8_t gran, float w1, float w2, const std::vector<uint64_t>& move_seeds,
              const std::vector<uint8_t>& orient, uint64_t steps_override, uint32_t flags,
              std::vector<sops_result>& out) override {
        std::vector<sops_size> sz(trials.size());
        std::vector<uint64_t> sd(trials.size());
        for (size_t t = 0; t < trials.size(); ++t) { sz[t] = trials[t].size; sd[t] = trials[t].seed; }
        out.assign((size_t)P * trials.size(), sops_result());
        int rc = sops_eval_batch(ctx_, behavior, genomes.data(), P, sz.data(), sd.data(), (uint32_t)trials.size(), gran, w1, w2,
                                 rng_mode_, move_seeds.empty() ? nullptr : move_seeds.data(),
                                 orient.empty() ? nullptr : orient.data(), steps_override, flags, out.data());
        if (rc) throw std::runtime_error(std::string("sops_eval_batch: ") + sops_last_error(ctx_));
    }
    void counters(uint64_t out[6]) override { if (sops_batch_counters(ctx_, out)) for (int k = 0; k < 6; ++k) out[k] = 0; }
    bool set_trial_steps(const std::vector<uint64_t>& steps) override {
        return sops_ctx_set_trial_steps(ctx_, steps.empty() ? nullptr : steps.data(), (uint32_t)steps.size()) == 0;
    }
    void dump(uint32_t instance, std::vector<uint8_t>& grid, std::vector<uint8_t>& particles, uint32_t G, uint32_t n) override {
        grid.assign((size_t)G * G, 0);
        particles.assign((size_t)n * 3, 0);
        int rc = sops_dump_state(ctx_, instance, grid.data(), particles.data());
        if (rc) throw std::runtime_error(std::string("sops_dump_state: ") + sops_last_error(ctx_));
    }
private:
    sops_ctx* ctx_ = nullptr;
    int rng_mode_;
};

// print_grid(): mod.rs:153-161 (identical in the other three environments)
inline void print_grid(FILE* f, const std::vector<uint8_t>& grid, uint32_t G) {
    std::fprintf(f, "SOPS grid\n");
    for (uint32_t i = 0; i < G; ++i) {
        for (uint32_t j = 0; j < G; ++j) std::fprintf(f, " %u ", (unsigned)grid[(size_t)i * G + j]);
        std::fprintf(f, "\n");
    }
}

// One environment = one (genome, size, seed) instance, the reference's per-instance API.
class SOPSEnvironmentBase {
public:
    SOPSEnvironmentBase(Evaluator& ev, int behavior, const uint8_t* genome, sops_size size, uint64_t seed, uint8_t gran,
                        float w1, float w2, uint64_t move_seed, uint8_t orient, uint32_t flags)
        : ev_(ev), behavior_(behavior), genome_(genome, genome + genome_len(behavior)), size_(size), seed_(seed), gran_(gran),
          w1_(w1), w2_(w2), move_seed_(move_seed), orient_(orient), flags_(flags) {
        if (sops_instance_shape(behavior, size, &G_, &n_, &sim_duration_)) throw std::invalid_argument("invalid size tuple");
        run(0, SOPS_F_INIT_ONLY);                    // init_sops_env: the initial configuration
    }
    // the chain restarted from the seeded initial configuration and run for `steps` moves; streams are
    // counter based, so a run of k steps is the k-step prefix of the full chain (used for snapshots)
    sops_result run_prefix(uint64_t steps) { run(steps, 0); return last_; }
    sops_result evaluate() const { return last_; }
    uint32_t grid_size() const { return G_; }
    uint64_t sim_duration() const { return sim_duration_; }
    void print_grid(FILE* f = stdout) const { evosops::print_grid(f, grid_, G_); }
    const std::vector<uint8_t>& grid() const { return grid_; }
    const std::vector<uint8_t>& particles() const { return parts_; }
protected:
    void run(uint64_t steps, uint32_t extra) {
        std::vector<Trial> tr(1);
        tr[0].size = size_; tr[0].seed = seed_;
        std::vector<sops_result> out;
        std::vector<uint64_t> ms(1, move_seed_);
        std::vector<uint8_t> ori(1, orient_);
        ev_.eval(behavior_, genome_, 1, tr, gran_, w1_, w2_, ms, ori, steps, flags_ | extra | SOPS_F_DUMP, out);
        last_ = out[0];
        ev_.dump(0, grid_, parts_, G_, n_);
    }
    Evaluator& ev_;
    int behavior_;
    std::vector<uint8_t> genome_;
    sops_size size_;
    uint64_t seed_;
    uint8_t gran_;
    float w1_, w2_;
    uint64_t move_seed_;
    uint8_t orient_;
    uint32_t flags_;
    uint32_t G_ = 0, n_ = 0;
    uint64_t sim_duration_ = 0;
    sops_result last_{};
    std::vector<uint8_t> grid_, parts_;
};

// src/SOPSCore/mod.rs:25-327
class SOPSEnvironment : public SOPSEnvironmentBase {
public:
    static SOPSEnvironment init_sops_env(Evaluator& ev, const uint8_t genome[48], uint16_t arena_layers, uint16_t particle_layers,
                                         uint64_t seed, uint8_t granularity, uint64_t move_seed, uint32_t flags = 0) {
        return SOPSEnvironment(ev, genome, sops_size{arena_layers, particle_layers, 0}, seed, granularity, move_seed, flags);
    }
    uint32_t evaluate_fitness() const { return last_.i0; }                      // edges
    uint32_t simulate(bool take_snaps, FILE* log = stdout) {
        if (take_snaps) {                                                         // mod.rs:309: after step == n and step == n^2
            const uint64_t n = n_;
            for (uint64_t at : {n, n * n}) {
                if (at >= sim_duration_) continue;
                sops_result r = run_prefix(at + 1);
                print_grid(log);
                std::fprintf(log, "Edge Count: %u\n", r.i0);
                std::fprintf(log, "Fitness: %s\n", fmt_f32((float)r.i0 / (float)r.i1).c_str());
            }
        }
        return run_prefix(sim_duration_).i0;
    }
    uint64_t get_max_fitness() const { return last_.i1; }
    size_t get_participant_cnt() const { return n_; }
    static std::string fmt_f32(float v);
private:
    SOPSEnvironment(Evaluator& ev, const uint8_t* g, sops_size s, uint64_t seed, uint8_t gran, uint64_t ms, uint32_t flags)
        : SOPSEnvironmentBase(ev, Agg, g, s, seed, gran, 0.f, 0.f, ms, 1, flags) {}
};

// separation.rs / coating.rs / locomotion.rs: f32 fitness, snapshots with "Step {}" headers
class SOPSFloatEnvironment : public SOPSEnvironmentBase {
public:
    SOPSFloatEnvironment(Evaluator& ev, int behavior, const uint8_t* g, sops_size s, uint64_t seed, uint8_t gran, float w1, float w2,
                         uint64_t ms, uint8_t orient, uint32_t flags)
        : SOPSEnvironmentBase(ev, behavior, g, s, seed, gran, w1, w2, ms, orient, flags) {}
    float evaluate_fitness() const { return last_.f; }
    float simulate(bool take_snaps, FILE* log = stdout) {
        if (take_snaps) {
            const uint64_t n = n_;
            std::vector<uint64_t> at;
            if (behavior_ == Coat) at = {n, n * n, n * n * 20, n * n * 40, n * n * 60, n * n * 70};   // coating.rs:582
            else at = {n, n * n};                                                                     // separation.rs:833 (n^3.. never reached)
            for (uint64_t a : at) {
                if (a >= sim_duration_) continue;
                sops_result r = run_prefix(a + 1);
                std::fprintf(log, "Step %llu\n", (unsigned long long)a);
                print_grid(log);
                std::fprintf(log, "Fitness: %s\n", SOPSEnvironment::fmt_f32(r.f).c_str());
            }
        }
        return run_prefix(sim_duration_).f;
    }
    size_t get_participant_cnt() const { return n_; }
};

struct SOPSepEnvironment : SOPSFloatEnvironment {
    static SOPSepEnvironment init_sops_env(Evaluator& ev, const uint8_t genome[600], uint16_t arena_layers, uint16_t particle_layers,
                                           uint64_t seed, uint8_t granularity, float w1, float w2, uint64_t move_seed, uint32_t flags = 0) {
        return SOPSepEnvironment(ev, genome, sops_size{arena_layers, particle_layers, 0}, seed, granularity, w1, w2, move_seed, flags);
    }
private:
    SOPSepEnvironment(Evaluator& ev, const uint8_t* g, sops_size s, uint64_t seed, uint8_t gran, float w1, float w2, uint64_t ms, uint32_t fl)
        : SOPSFloatEnvironment(ev, Sep, g, s, seed, gran, w1, w2, ms, 1, fl) {}
};
struct SOPSCoatEnvironment : SOPSFloatEnvironment {
    static SOPSCoatEnvironment init_sops_env(Evaluator& ev, const uint8_t genome[600], uint16_t arena_layers, uint16_t object_layers,
                                             uint16_t coat_layers, uint64_t seed, uint8_t granularity, float w1, float w2,
                                             uint64_t move_seed, uint32_t flags = 0) {
        return SOPSCoatEnvironment(ev, genome, sops_size{arena_layers, object_layers, coat_layers}, seed, granularity, w1, w2, move_seed, flags);
    }
    // coating.rs:341-356: the (size-only) BFS distance grid; 0xFFFF where the reference has no map entry
    std::vector<uint16_t> save_distance_grid() const {
        std::vector<uint16_t> d((size_t)G_ * G_);
        sops_coat_distance_grid(size_, d.data());
        return d;
    }
private:
    SOPSCoatEnvironment(Evaluator& ev, const uint8_t* g, sops_size s, uint64_t seed, uint8_t gran, float w1, float w2, uint64_t ms, uint32_t fl)
        : SOPSFloatEnvironment(ev, Coat, g, s, seed, gran, w1, w2, ms, 1, fl) {}
};
struct SOPSLocoEnvironment : SOPSFloatEnvironment {
    // `orientation` replaces locomotion.rs:109's unseeded thread_rng().gen_range(1..4)
    static SOPSLocoEnvironment init_sops_env(Evaluator& ev, const uint8_t genome[144], uint16_t arena_layers, uint16_t particle_layers,
                                             uint64_t seed, uint8_t granularity, float w1, float w2, uint64_t move_seed,
                                             uint8_t orientation, uint32_t flags = 0) {
        return SOPSLocoEnvironment(ev, genome, sops_size{arena_layers, particle_layers, 0}, seed, granularity, w1, w2, move_seed, orientation, flags);
    }
private:
    SOPSLocoEnvironment(Evaluator& ev, const uint8_t* g, sops_size s, uint64_t seed, uint8_t gran, float w1, float w2, uint64_t ms,
                        uint8_t orient, uint32_t fl)
        : SOPSFloatEnvironment(ev, Loco, g, s, seed, gran, w1, w2, ms, orient, fl) {}
};

// The standalone genome run (`-e gm`, main.rs:252-274 and its three siblings) over ALL (size, seed) trials at once: the
// reference runs them under into_par_iter, each trial printing its lattice at step 0, after its snapshot steps and at
// the end.  Here every snapshot stage is ONE batched evaluation of all trials (per-trial chain lengths: the snapshot
// steps n, n^2, ... differ per size) — a stage re-runs the counter-based chains from the seeded initial configuration,
// so a k-step run is the k-step prefix of the full chain.  `stage_steps[s][t]` = moves of trial t in stage s
// (0 = the initial configuration), `shown[s][t]` = whether the reference prints that snapshot for that trial.
struct StandaloneRun {
    std::vector<std::vector<uint64_t>> stage_steps;
    std::vector<std::vector<uint8_t>> shown;
    std::vector<std::vector<sops_result>> results;           // [stage][trial]
    std::vector<std::vector<std::vector<uint8_t>>> grids;    // [stage][trial] lattice in the reference's cell codes
    std::vector<uint32_t> G, n;
    std::vector<uint64_t> duration;
    uint64_t launches = 0;
};
inline std::vector<uint64_t> snapshot_steps(int behavior, uint64_t n) {
    if (behavior == Coat) return {n, n * n, n * n * 20, n * n * 40, n * n * 60, n * n * 70};        // coating.rs:582
    return {n, n * n};                               // mod.rs:309; separation.rs:833 / locomotion.rs (n^3.. is never reached)
}
inline StandaloneRun run_standalone(Evaluator& ev, int behavior, const std::vector<uint8_t>& genome, const std::vector<Trial>& trials,
                                    uint8_t gran, float w1, float w2, const std::vector<uint64_t>& move_seeds,
                                    const std::vector<uint8_t>& orient, uint64_t final_steps_override, uint32_t flags) {
    StandaloneRun R;
    const size_t T = trials.size();
    R.G.resize(T); R.n.resize(T); R.duration.resize(T);
    size_t n_snap = 0;
    for (size_t t = 0; t < T; ++t) {
        if (sops_instance_shape(behavior, trials[t].size, &R.G[t], &R.n[t], &R.duration[t])) throw std::invalid_argument("invalid size tuple");
        n_snap = std::max(n_snap, snapshot_steps(behavior, R.n[t]).size());
    }
    const size_t n_stage = n_snap + 2;                       // initial, snapshots, final
    R.stage_steps.assign(n_stage, std::vector<uint64_t>(T, 0));
    R.shown.assign(n_stage, std::vector<uint8_t>(T, 1));
    for (size_t t = 0; t < T; ++t) {
        const std::vector<uint64_t> at = snapshot_steps(behavior, R.n[t]);
        const uint64_t total = final_steps_override ? final_steps_override : R.duration[t];
        for (size_t k = 0; k < n_snap; ++k) {
            // the snapshot is taken after step index `at` has run, i.e. after at + 1 moves, and only if the chain gets there
            const bool reached = k < at.size() && at[k] < R.duration[t] && !final_steps_override;
            R.shown[1 + k][t] = reached ? 1 : 0;
            R.stage_steps[1 + k][t] = reached ? at[k] + 1 : 0;
        }
        R.stage_steps[n_stage - 1][t] = total;
    }
    R.results.resize(n_stage);
    R.grids.assign(n_stage, std::vector<std::vector<uint8_t>>(T));
    for (size_t s = 0; s < n_stage; ++s) {
        bool any = s == 0 || s + 1 == n_stage;
        for (size_t t = 0; t < T; ++t) any = any || R.shown[s][t];
        if (!any) continue;                                  // no trial prints this snapshot
        uint32_t fl = flags | SOPS_F_DUMP;
        if (s == 0) fl |= SOPS_F_INIT_ONLY;
        else if (!ev.set_trial_steps(R.stage_steps[s])) throw std::runtime_error("evaluator cannot take per-trial chain lengths");
        ev.eval(behavior, genome, 1, trials, gran, w1, w2, move_seeds, orient, 0, fl, R.results[s]);
        uint64_t c[6];
        ev.counters(c);
        R.launches += c[3];
        for (size_t t = 0; t < T; ++t) {
            std::vector<uint8_t> parts;
            ev.dump((uint32_t)t, R.grids[s][t], parts, R.G[t], R.n[t]);
        }
    }
    return R;
}

}  // namespace evosops
