// ga.hpp — C++ mirror of the reference's GACore (host side of the drop-in; stays on the CPU).
//
//   GeneticAlgo  src/GACore/base_ga.rs   aggregation: genome cache, fixed mutation rate
//   SepGA        src/GACore/sep_ga.rs    separation: fastrand mutation draw, adaptive mutation rate, no cache
//   CoatGA       src/GACore/coat_ga.rs   coating: as sep (+ distance grids, here inside the library)
//   LocoGA       src/GACore/loco_ga.rs   locomotion: zip+repeat trial list, cache, thresholds 0.2 / 0.02
// One class parameterised by GaSpec instead of four copies.  The nested rayon par_iter of step_through
// (base_ga.rs:393-424) is ONE Evaluator::eval call per generation.  The reference's host randomness
// (thread_rng / fastrand) is unseeded; here it is an explicit xoshiro256** stream (--ga-seed) drawing the
// same distributions, so runs are repeatable.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

#include "fmt.hpp"
#include "sops_core.hpp"

namespace evosops {

struct HostRng {                               // xoshiro256** seeded by SplitMix64
    uint64_t s[4];
    explicit HostRng(uint64_t seed) {
        for (int k = 0; k < 4; ++k) {
            seed += 0x9E3779B97F4A7C15ULL;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            s[k] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    uint64_t below(uint64_t n) {               // uniform in [0, n), unbiased (Lemire)
        uint64_t x = next();
        unsigned __int128 m = (unsigned __int128)x * n;
        uint64_t lo = (uint64_t)m;
        if (lo < n) { uint64_t t = (0 - n) % n; while (lo < t) { x = next(); m = (unsigned __int128)x * n; lo = (uint64_t)m; } }
        return (uint64_t)(m >> 64);
    }
    uint64_t inclusive(uint64_t lo, uint64_t hi) { return lo + below(hi - lo + 1); }
    bool bernoulli(double p) { return (double)(next() >> 11) * (1.0 / 9007199254740992.0) < p; }
};

struct GaSpec {
    int behavior;
    int dims[4]; int nd;                       // genome array dims (GACore/mod.rs:13-45)
    bool cache;                                // genome -> fitness cache consulted before evaluation (agg, loco)
    bool frand_mut;                            // mutation draw u64(1..=10000) <= rate*10000 (sep/coat/loco) vs Uniform(1..=100) <= rate*100 (agg)
    bool adaptive;                             // diversity-driven mutation rate x10 / /10 (sep, coat, loco)
    float upper_t, lower_t;                    // sep/coat 0.3 / 0.08, loco 0.2 / 0.02
    bool zip_trials;                           // loco: zip(sizes, seeds) each repeated len(seeds) times (loco_ga.rs:384-389)
    const char* struct_name;                   // Debug name of the genome struct
    const char* log_tag;                       // genomic_data_<tag>_<rand>.log
    const char* weights_fmt;                   // init banner (sep_ga.rs:94, loco_ga.rs:94)
    bool print_raw_avg_div;                    // coat prints the un-normalised window average too (coat_ga.rs:553)
};

inline GaSpec spec_for(int behavior) {
    switch (behavior) {
    case Agg:  return GaSpec{Agg, {4, 3, 4, 0}, 3, true, false, false, 0.f, 0.f, false, "Genome", "Agg", nullptr, false};
    case Sep:  return GaSpec{Sep, {10, 6, 10, 0}, 3, false, true, true, 0.3f, 0.08f, false, "SepGenome", "Sep",
                             "Weights: %s x Total Edges + %s x Avg. Same Clr Edges\n", false};
    case Coat: return GaSpec{Coat, {10, 6, 10, 0}, 3, false, true, true, 0.3f, 0.08f, false, "CoatGenome", "Coat",
                             "Weights: %s x Total Edges + %s x Avg. Same Clr Edges\n", true};
    default:   return GaSpec{Loco, {3, 4, 3, 4}, 4, true, true, true, 0.2f, 0.02f, true, "LocoGenome", "Loco",
                             "Weights: %s x Total Edges + %s x Avg. Distance\n", false};
    }
}

struct GenomeRec { std::vector<uint8_t> string; double fitness = 0.0; };

class GeneticAlgo {
public:
    // init_ga: base_ga.rs:72-116 (random genes in 0..=granularity)
    GeneticAlgo(Evaluator& ev, int behavior, uint16_t population_size, uint16_t max_gen, uint16_t elitist_cnt, double mut_rate,
                uint8_t granularity, bool perform_cross, std::vector<sops_size> sizes, std::vector<uint64_t> trial_seeds,
                float w1, float w2, uint32_t random_seed, uint64_t ga_seed, const std::string& out_dir, FILE* log, bool verbose = false)
        : ev_(ev), spec_(spec_for(behavior)), max_gen_(max_gen), elitist_cnt_(elitist_cnt), mut_rate_(mut_rate), granularity_(granularity),
          perform_cross_(perform_cross), sizes_(std::move(sizes)), trial_seeds_(std::move(trial_seeds)), w1_(w1), w2_(w2),
          random_seed_(random_seed), rng_(ga_seed), out_dir_(out_dir), log_(log), verbose_(verbose) {
        L_ = genome_len(behavior);
        if (spec_.weights_fmt) {
            std::fprintf(log_, spec_.weights_fmt, fmt_display(w1).c_str(), fmt_display(w2).c_str());
            std::fprintf(log_, "Thresholds: UPPER: %s, LOWER %s\n", fmt_display(spec_.upper_t).c_str(), fmt_display(spec_.lower_t).c_str());
        }
        population_.resize(population_size);
        for (auto& g : population_) {
            g.string.resize(L_);
            for (auto& gene : g.string) gene = (uint8_t)rng_.inclusive(0, granularity);
        }
        max_div_ = (uint32_t)(granularity - 1) * (uint32_t)L_;                   // base_ga.rs:114
    }

    std::vector<GenomeRec>& population() { return population_; }
    double mut_rate() const { return mut_rate_; }
    bool lower_hit() const { return lower_hit_; }
    uint64_t evaluated_instances() const { return evaluated_; }

    // mutate_genome: base_ga.rs:119-141 / sep_ga.rs:135-157
    std::vector<uint8_t> mutate_genome(const std::vector<uint8_t>& genome) {
        std::vector<uint8_t> out = genome;
        for (int k = 0; k < L_; ++k) {
            bool hit;
            if (spec_.frand_mut) hit = (double)rng_.inclusive(1, 10000) <= mut_rate_ * 10000.0;
            else hit = (double)rng_.inclusive(1, 100) <= mut_rate_ * 100.0;
            if (hit) {
                const bool up = rng_.bernoulli(0.3);                                // mut_sign(): Bernoulli(0.3)
                unsigned v = up ? (unsigned)genome[k] + 1u : (genome[k] == 0 ? 0u : (unsigned)genome[k] - 1u);
                out[k] = (uint8_t)std::min<unsigned>(v, granularity_);             // .clamp(0, granularity)
            }
        }
        return out;
    }
    // generate_offspring: 2-point crossover, base_ga.rs:165-188 (the gene AT the lower point comes from parent1)
    std::vector<uint8_t> generate_offspring(const std::vector<uint8_t>& p1, const std::vector<uint8_t>& p2) {
        const uint64_t c1 = rng_.inclusive(0, (uint64_t)L_ - 1), c2 = rng_.inclusive(0, (uint64_t)L_ - 1);
        const uint64_t lo = c1 <= c2 ? c1 : c2, hi = c1 > c2 ? c1 : c2;
        std::vector<uint8_t> out(L_);
        for (uint64_t cnt = 0; cnt < (uint64_t)L_; ++cnt) out[cnt] = (cnt > lo && cnt < hi) ? p2[cnt] : p1[cnt];
        return out;
    }

    // generate_new_pop: base_ga.rs:257-338
    void generate_new_pop() {
        const uint16_t P = (uint16_t)population_.size();
        const GenomeRec* best = &population_[0];                                    // Iterator::max_by keeps the LAST maximum
        for (const auto& g : population_) if (!(g.fitness < best->fitness)) best = &g;
        std::fprintf(log_, "Best Genome -> %s { string: %s, fitness: %s }\n", spec_.struct_name,
                     fmt_genome_debug(best->string.data(), spec_.dims, spec_.nd).c_str(), fmt_fixed(best->fitness, 5).c_str());
        append_genomic_data();
        std::vector<std::vector<uint8_t>> selected, crossed;
        for (uint16_t k = 0; k < P; ++k) {                                          // tournament selection (294-310)
            const uint64_t i1 = rng_.below(P);
            uint64_t i2;
            do { i2 = rng_.below(P); } while (i1 == i2);
            selected.push_back(population_[i1].fitness > population_[i2].fitness ? population_[i1].string : population_[i2].string);
        }
        for (uint16_t k = 0; k < P; ++k) {                                          // 2-point crossover (313-325)
            const uint64_t i1 = rng_.below(P);
            uint64_t i2;
            do { i2 = rng_.below(P); } while (i1 == i2);
            crossed.push_back(generate_offspring(selected[i1], selected[i2]));
        }
        std::vector<GenomeRec> next(P);
        for (uint16_t k = 0; k < P; ++k) { next[k].string = mutate_genome(crossed[k]); next[k].fitness = 0.0; }
        population_ = std::move(next);
    }

    // the flattened trial list: base_ga.rs:356-362 (sizes-major, seeds-minor) / loco_ga.rs:384-389 (zip + repeat)
    std::vector<Trial> trials_vec() const {
        std::vector<Trial> t;
        if (spec_.zip_trials) {
            const size_t m = std::min(sizes_.size(), trial_seeds_.size());
            for (size_t k = 0; k < m; ++k)
                for (size_t r = 0; r < trial_seeds_.size(); ++r) t.push_back(Trial{sizes_[k], trial_seeds_[k]});
        } else {
            for (const auto& sz : sizes_) for (uint64_t sd : trial_seeds_) t.push_back(Trial{sz, sd});
        }
        return t;
    }

    // step_through: base_ga.rs:344-477; returns avg_pop_diversity
    float step_through(uint16_t gen) {
        if (spec_.adaptive) {
            std::fprintf(log_, "Mutation Rate:%s\n", fmt_display(mut_rate_).c_str());
            std::fprintf(log_, "Diversity State:%s\n", lower_hit_ ? "LOWER_HIT" : "INIT");
        }
        const std::vector<Trial> trials = trials_vec();
        const size_t T = trials.size();
        if (spec_.cache)                                                            // 366-391
            for (auto& g : population_) {
                auto it = cache_.find(key_of(g));
                if (it != cache_.end() && it->second > -1.0) g.fitness = it->second;
            }
        std::vector<size_t> todo;                                                   // 396-398: skip if gen > 0 && fitness > 0.0
        for (size_t i = 0; i < population_.size(); ++i)
            if (!(gen > 0 && population_[i].fitness > 0.0)) todo.push_back(i);
        if (!todo.empty() && T > 0) {
            std::vector<uint8_t> flat(todo.size() * (size_t)L_);
            for (size_t k = 0; k < todo.size(); ++k) std::memcpy(&flat[k * L_], population_[todo[k]].string.data(), (size_t)L_);
            // every evaluation draws fresh move randomness (the reference's fastrand stream is never re-seeded)
            std::vector<uint64_t> ms(todo.size() * T);
            for (auto& v : ms) v = rng_.next();
            std::vector<uint8_t> orient;
            if (spec_.behavior == Loco) { orient.resize(todo.size() * T); for (auto& o : orient) o = (uint8_t)rng_.inclusive(1, 3); }   // locomotion.rs:109
            std::vector<sops_result> out;
            // kernel hint (ours): when a generation accepts < 1 % of its attempts the 4-warp-per-chain kernel can be
            // faster.  Not for aggregation: its warp kernel (hashed decision, one chain per warp scheduler) measured
            // faster than the CTA form on late-generation populations too.  Results do not depend on the choice.
            const bool use_cta = low_acceptance_ && spec_.behavior != Agg;
            ev_.eval(spec_.behavior, flat, (uint32_t)todo.size(), trials, granularity_, w1_, w2_, ms, orient, 0,
                     use_cta ? (uint32_t)SOPS_F_FORCE_CTA : 0u, out);
            uint64_t cnt[6];
            ev_.counters(cnt);
            if (cnt[0] > 0) low_acceptance_ = (double)cnt[2] / (double)cnt[0] < 0.01;
            evaluated_ += todo.size() * T;
            for (size_t k = 0; k < todo.size(); ++k) {
                double tot = 0.0;
                for (size_t t = 0; t < T; ++t) tot += out[k * T + t].norm;          // 401-411 (.sum())
                population_[todo[k]].fitness = tot / (double)T;                     // 422-423
            }
        }
        if (spec_.cache) for (const auto& g : population_) cache_[key_of(g)] = g.fitness;   // 426-430
        double fit_sum = 0.0;
        for (const auto& g : population_) fit_sum += g.fitness;
        std::fprintf(log_, "Avg. Fitness -> %s\n", fmt_display(fit_sum / (double)population_.size()).c_str());
        // population diversity: mean pairwise L1 distance (443-473), u16 per pair, f32 accumulation
        float pop_diversity = 0.f;
        size_t pairs = 0;
        for (size_t i = 0; i < population_.size(); ++i)
            for (size_t j = i + 1; j < population_.size(); ++j) {
                uint16_t dis = 0;
                const auto& a = population_[i].string; const auto& b = population_[j].string;
                for (int k = 0; k < L_; ++k) dis = (uint16_t)(dis + (a[k] > b[k] ? a[k] - b[k] : b[k] - a[k]));
                pop_diversity += (float)dis;
                ++pairs;
            }
        const float avg_pop_diversity = pop_diversity / (float)pairs;
        std::fprintf(log_, "Population diversity -> %s\n", fmt_display(avg_pop_diversity / (float)max_div_).c_str());
        generate_new_pop();
        return avg_pop_diversity;
    }

    // run_through: base_ga.rs:483-504 / sep_ga.rs:525-565
    void run_through() {
        std::deque<float> diversity_q;
        for (uint16_t gen = 0; gen < max_gen_; ++gen) {
            std::fprintf(log_, "Starting Gen:%u\n", (unsigned)gen);
            const auto t0 = std::chrono::steady_clock::now();
            if (!spec_.adaptive) {
                step_through(gen);
            } else {
                if (diversity_q.size() == 10) diversity_q.pop_front();              // BUFFER_LEN
                diversity_q.push_back(step_through(gen));
                float sum = 0.f;
                for (float v : diversity_q) sum += v;
                const float avg_div = sum / (float)diversity_q.size();
                const float norm_avg_div = avg_div / (float)max_div_;
                if (spec_.print_raw_avg_div)
                    std::fprintf(log_, "Avg. Population diversity for last %d gen -> %s\n", 10, fmt_display(avg_div).c_str());
                std::fprintf(log_, "Avg. Population diversity for last %zu gen -> %s\n", diversity_q.size(), fmt_display(norm_avg_div).c_str());
                if (!lower_hit_) { if (norm_avg_div <= spec_.lower_t) { mut_rate_ *= 10.0; lower_hit_ = true; } }
                else if (norm_avg_div >= spec_.upper_t) { mut_rate_ /= 10.0; lower_hit_ = false; }
            }
            const auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
            std::fprintf(log_, "Generation Elapsed Time: %llds\n", (long long)secs);
            if (verbose_)                                                          // extension (-v): the reference only has whole seconds
                std::fprintf(log_, "Generation Wall Time: %.3fs  Evaluated Instances: %llu  Kernel: %s\n",
                             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(),
                             (unsigned long long)evaluated_, (low_acceptance_ && spec_.behavior != Agg) ? "cta" : "warp");
            std::fflush(log_);
        }
    }

private:
    std::string key_of(const GenomeRec& g) const { return std::string(g.string.begin(), g.string.end()); }
    // per genome: L gene bytes + f64 fitness big-endian, appended every generation (base_ga.rs:275-291)
    void append_genomic_data() {
        std::string path = out_dir_ + "/genomic_data_" + spec_.log_tag + "_" + std::to_string(random_seed_) + ".log";
        FILE* f = std::fopen(path.c_str(), "ab");
        if (!f) throw std::runtime_error("Failed to create genomic data file!");
        for (const auto& g : population_) {
            std::fwrite(g.string.data(), 1, g.string.size(), f);
            uint64_t bits;
            std::memcpy(&bits, &g.fitness, 8);
            unsigned char be[8];
            for (int k = 0; k < 8; ++k) be[k] = (unsigned char)(bits >> (56 - 8 * k));
            std::fwrite(be, 1, 8, f);
        }
        std::fclose(f);
    }

    Evaluator& ev_;
    GaSpec spec_;
    int L_ = 0;
    uint16_t max_gen_, elitist_cnt_;               // elitist_cnt is stored and never read upstream either (base_ga.rs:19)
    double mut_rate_;
    uint8_t granularity_;
    bool perform_cross_;
    std::vector<sops_size> sizes_;
    std::vector<uint64_t> trial_seeds_;
    float w1_, w2_;
    uint32_t random_seed_;
    uint32_t max_div_ = 1;
    bool lower_hit_ = false;
    bool low_acceptance_ = false;                  // last generation's accepted fraction < 1 %
    HostRng rng_;
    std::string out_dir_;
    FILE* log_;
    bool verbose_ = false;
    std::vector<GenomeRec> population_;
    std::unordered_map<std::string, double> cache_;
    uint64_t evaluated_ = 0;
};

}  // namespace evosops
