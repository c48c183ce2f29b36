// evosops — C++ mirror of the reference's CLI (src/main.rs) on top of libevosops_b200.so.
//
// Same flags as the clap struct at main.rs:28-78:
//   -b/--behavior {agg,sep,coat,loco}  -e/--exp {ga,gm,th}  -g/--gen  -p/--pop  --gran (default 20)
//   -m/--mut (0.08)  --eli  -k/--ks "(a,p[,c])" (repeatable)  -s/--seed (repeatable)  --path FILE  -v  --snaps
// Same side effects: stdout is redirected into ./output/<Behavior:?>_<Exp:?>_<#sizes>_sizes_<#seeds>_trials_gran_<g>_<rand>.log
// (main.rs:114-127) and the GA appends ./output/genomic_data_<B>_<rand>.log.
// Differences, all deliberate: --path is only required for gm/th (upstream asserts it for ga too,
// main.rs:108-110); `-e th` runs the theory probability tables (upstream: todo!()); invalid sizes / genes are
// errors instead of hangs / panics.  Extensions: --rng {fast,verify}, --devices N, --ga-seed S, --out DIR,
// --no-redirect, --steps N (cap chains, for smoke runs).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "ga.hpp"

using namespace evosops;

namespace {

struct Args {
    int behavior = -1, exp = -1;                 // exp: 0 GA, 1 GM, 2 TH
    uint16_t max_generations = 0, population = 0, elitist_count = 0;
    uint8_t granularity = 20;
    double mutation_rate = 0.08;
    std::vector<std::string> particle_sizes;
    std::vector<uint64_t> seeds;
    std::string path;
    bool verbose = false, snaps = false;
    // extensions
    int rng_mode = SOPS_RNG_FAST, devices = 1;
    uint64_t ga_seed = 0, steps = 0;
    bool have_ga_seed = false, redirect = true;
    std::string out_dir = "./output";
};

[[noreturn]] void usage(const char* why) {
    std::fprintf(stderr, "error: %s\n\nUsage: evosops --behavior <agg|sep|coat|loco> --exp <ga|gm|th> --ks <(a,p[,c])>... --seed <SEED>... "
                         "[--gen N] [--pop N] [--gran G] [--mut R] [--eli N] [--path FILE] [-v] [--snaps]\n", why);
    std::exit(2);
}

Args parse(int argc, char** argv) {
    Args a;
    auto value = [&](int& i, const std::string& arg, const std::string& attached) -> std::string {
        if (!attached.empty()) return attached;
        if (i + 1 >= argc) usage(("missing value for " + arg).c_str());
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        std::string arg = argv[i], name, attached;
        if (arg.rfind("--", 0) == 0) {
            size_t eq = arg.find('=');
            name = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) attached = arg.substr(eq + 1);
        } else if (arg.size() >= 2 && arg[0] == '-') {
            name = std::string(1, arg[1]);
            attached = arg.substr(2);
            if (!attached.empty() && attached[0] == '=') attached = attached.substr(1);
            static const char* longs[][2] = {{"b", "behavior"}, {"e", "exp"}, {"g", "gen"}, {"p", "pop"}, {"m", "mut"}, {"k", "ks"},
                                             {"s", "seed"}, {"v", "verbose"}};
            bool found = false;
            for (auto& l : longs) if (name == l[0]) { name = l[1]; found = true; }
            if (!found) usage(("unexpected argument " + arg).c_str());
        } else usage(("unexpected argument " + arg).c_str());
        if (name == "behavior") {
            std::string v = value(i, arg, attached);
            a.behavior = v == "agg" ? Agg : v == "sep" ? Sep : v == "coat" ? Coat : v == "loco" ? Loco : -1;
            if (a.behavior < 0) usage("invalid value for --behavior (agg, sep, coat, loco)");
        } else if (name == "exp") {
            std::string v = value(i, arg, attached);
            a.exp = v == "ga" ? 0 : v == "gm" ? 1 : v == "th" ? 2 : -1;
            if (a.exp < 0) usage("invalid value for --exp (ga, gm, th)");
        } else if (name == "gen") a.max_generations = (uint16_t)std::stoul(value(i, arg, attached));
        else if (name == "pop") a.population = (uint16_t)std::stoul(value(i, arg, attached));
        else if (name == "gran") a.granularity = (uint8_t)std::stoul(value(i, arg, attached));
        else if (name == "mut") a.mutation_rate = std::stod(value(i, arg, attached));
        else if (name == "eli") a.elitist_count = (uint16_t)std::stoul(value(i, arg, attached));
        else if (name == "ks") a.particle_sizes.push_back(value(i, arg, attached));
        else if (name == "seed") a.seeds.push_back(std::stoull(value(i, arg, attached)));
        else if (name == "path") a.path = value(i, arg, attached);
        else if (name == "verbose") a.verbose = true;
        else if (name == "snaps") a.snaps = true;
        else if (name == "rng") { std::string v = value(i, arg, attached); a.rng_mode = v == "verify" ? SOPS_RNG_VERIFY : SOPS_RNG_FAST; }
        else if (name == "devices") a.devices = std::stoi(value(i, arg, attached));
        else if (name == "ga-seed") { a.ga_seed = std::stoull(value(i, arg, attached)); a.have_ga_seed = true; }
        else if (name == "steps") a.steps = std::stoull(value(i, arg, attached));
        else if (name == "out") a.out_dir = value(i, arg, attached);
        else if (name == "no-redirect") a.redirect = false;
        else usage(("unexpected argument " + arg).c_str());
    }
    if (a.behavior < 0) usage("--behavior is required");
    if (a.exp < 0) usage("--exp is required");
    if (a.particle_sizes.empty()) usage("--ks is required");
    if (a.seeds.empty()) usage("--seed is required");
    return a;
}

// main.rs:141-147: split on '(' ')' ',' and keep what parses as u16
std::vector<uint16_t> size_numbers(const std::string& s) {
    std::vector<uint16_t> out;
    std::string tok;
    auto flush = [&]() {
        if (tok.empty()) return;
        bool ok = true;
        unsigned long v = 0;
        for (char c : tok) { if (c < '0' || c > '9') { ok = false; break; } v = v * 10 + (unsigned long)(c - '0'); if (v > 65535) { ok = false; break; } }
        if (ok) out.push_back((uint16_t)v);
        tok.clear();
    };
    for (char c : s) { if (c == '(' || c == ')' || c == ',') flush(); else tok += c; }
    flush();
    return out;
}

std::string debug_strings(const std::vector<std::string>& v) {
    std::string s = "[";
    for (size_t k = 0; k < v.size(); ++k) { if (k) s += ", "; s += "\"" + v[k] + "\""; }
    return s + "]";
}
std::string debug_u64s(const std::vector<uint64_t>& v) {
    std::string s = "[";
    for (size_t k = 0; k < v.size(); ++k) { if (k) s += ", "; s += std::to_string(v[k]); }
    return s + "]";
}

}  // namespace

std::string SOPSEnvironment::fmt_f32(float v) { return fmt_display(v); }

int main(int argc, char** argv) {
    Args args = parse(argc, argv);
    static const char* exp_names[] = {"GA", "GM", "TH"};
    if (args.exp != 0 && args.path.empty()) usage("--path FILE is required for gm / th");

    const uint64_t now_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                                std::chrono::system_clock::now().time_since_epoch()).count();
    HostRng name_rng(args.have_ga_seed ? args.ga_seed : now_ns);
    const uint32_t random_trial_seed = (uint32_t)name_rng.next();                  // main.rs:113 fastrand::u32(..)
    const std::string file_name = std::string(behavior_debug_name(args.behavior)) + "_" + exp_names[args.exp] + "_" +
                                  std::to_string(args.particle_sizes.size()) + "_sizes_" + std::to_string(args.seeds.size()) +
                                  "_trials_gran_" + std::to_string((unsigned)args.granularity) + "_" + std::to_string(random_trial_seed);
    std::printf("Running the experiment... \nPlease check: \"%s\" file in ./output folder\n", file_name.c_str());
    std::fflush(stdout);
    mkdir(args.out_dir.c_str(), 0755);
    FILE* log = stdout;
    if (args.redirect) {
        const std::string log_path = args.out_dir + "/" + file_name + ".log";      // get_temp_filepath, main.rs:23-26
        log = std::fopen(log_path.c_str(), "w");
        if (!log) { std::fprintf(stderr, "cannot open %s\n", log_path.c_str()); return 1; }
    }

    std::fprintf(log, "Target Behavior: %s\n", behavior_debug_name(args.behavior));
    std::fprintf(log, "Experiment Type: %s\n", exp_names[args.exp]);
    std::fprintf(log, "Particle Sizes: %s\n", debug_strings(args.particle_sizes).c_str());
    std::fprintf(log, "Initialization Seeds: %s\n", debug_u64s(args.seeds).c_str());
    std::fprintf(log, "Representation Granularity: %u\n", (unsigned)args.granularity);

    std::vector<sops_size> sizes;
    for (const auto& s : args.particle_sizes) {
        std::vector<uint16_t> c = size_numbers(s);
        const size_t need = args.behavior == Coat ? 3 : 2;
        if (c.size() < need) usage(("size tuple needs " + std::to_string(need) + " numbers: " + s).c_str());
        sizes.push_back(sops_size{c[0], c[1], (uint16_t)(args.behavior == Coat ? c[2] : 0)});
    }

    int rc = 0;
    try {
        CudaEvaluator ev(args.devices, args.rng_mode);
        const float w1 = args.behavior == Loco ? 0.75f : 0.65f, w2 = args.behavior == Loco ? 0.25f : 0.35f;   // main.rs:171,177,182
        if (args.exp == 0) {
            std::fprintf(log, "Population Size: %u\n", (unsigned)args.population);
            std::fprintf(log, "Max Generations: %u\n", (unsigned)args.max_generations);
            std::fprintf(log, "Mutation Rate: %s\n", fmt_debug(args.mutation_rate).c_str());
            std::fprintf(log, "Elitist Count: %u\n", (unsigned)args.elitist_count);
            std::fprintf(log, "Cross-over: true\n");
            static const char* titles[] = {"Aggregation", "Separation", "Coating", "Locomotion"};
            std::fprintf(log, "\nStarting %s GA Experiment...\n\n", titles[args.behavior]);
            GeneticAlgo ga(ev, args.behavior, args.population, args.max_generations, args.elitist_count, args.mutation_rate,
                           args.granularity, true, sizes, args.seeds, w1, w2, random_trial_seed,
                           args.have_ga_seed ? args.ga_seed : now_ns, args.out_dir, log, args.verbose);
            ga.run_through();
        } else {
            std::fprintf(log, "Snapshots: %s\n", args.snaps ? "true" : "false");
            std::fprintf(log, "Genome file path: %s\n", args.path.c_str());
            std::vector<uint8_t> all_entries = read_genome_file(args.path);
            const int L = genome_len(args.behavior);
            if ((int)all_entries.size() < L) throw std::runtime_error("genome file holds " + std::to_string(all_entries.size()) +
                                                                      " genes, need " + std::to_string(L));
            all_entries.resize((size_t)L);
            const GaSpec spec = spec_for(args.behavior);
            static const char* titles[] = {"Aggregation", "Separation", "Coating", "Locomotion"};
            std::fprintf(log, "\nStarting %s Single Genome Trial...\n\n", titles[args.behavior]);
            std::fprintf(log, "Read Genome:\n%s\n", fmt_genome_debug(all_entries.data(), spec.dims, spec.nd).c_str());
            const uint32_t flags = args.exp == 2 ? SOPS_F_THEORY_TABLE : 0;          // `th`: upstream todo!(); here the theory tables
            const float gw1 = 0.65f, gw2 = 0.35f;                                    // GM uses 0.65 / 0.35 for every behaviour (main.rs:311,366,421)
            HostRng mrng(args.have_ga_seed ? args.ga_seed : now_ns);
            double fitness_tot = 0.0;
            // every (size, seed) trial of the run (main.rs:244-250), evaluated together: one launch per snapshot stage
            // (the reference fans the trials out with into_par_iter, main.rs:252-274)
            std::vector<Trial> trials;
            std::vector<uint64_t> move_seeds;
            std::vector<uint8_t> orient;
            for (const auto& sz : sizes)
                for (uint64_t seed : args.seeds) {
                    trials.push_back(Trial{sz, seed});
                    move_seeds.push_back(mrng.next());
                    if (args.behavior != Agg) orient.push_back((uint8_t)mrng.inclusive(1, 3));
                }
            const auto t0 = std::chrono::steady_clock::now();
            const StandaloneRun R = run_standalone(ev, args.behavior, all_entries, trials, args.granularity, gw1, gw2, move_seeds,
                                                   orient, args.steps, flags);
            const auto secs = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
            const size_t last = R.results.size() - 1;
            for (size_t t = 0; t < trials.size(); ++t) {
                const std::vector<uint64_t> at = snapshot_steps(args.behavior, R.n[t]);
                if (args.behavior == Agg) {
                    const sops_result& r0 = R.results[0][t];
                    print_grid(log, R.grids[0][t], R.G[t]);
                    std::fprintf(log, "Edge Count: %u\n", r0.i0);
                    std::fprintf(log, "Max Fitness: %llu\n", (unsigned long long)r0.i1);
                    std::fprintf(log, "Starting Fitness: %s\n", fmt_display((float)r0.i0 / (float)r0.i1).c_str());
                    for (size_t k = 1; k < last; ++k) {
                        if (!R.shown[k][t]) continue;
                        const sops_result& r = R.results[k][t];
                        print_grid(log, R.grids[k][t], R.G[t]);
                        std::fprintf(log, "Edge Count: %u\n", r.i0);
                        std::fprintf(log, "Fitness: %s\n", SOPSEnvironment::fmt_f32((float)r.i0 / (float)r.i1).c_str());
                    }
                    const sops_result& rf = R.results[last][t];
                    print_grid(log, R.grids[last][t], R.G[t]);
                    std::fprintf(log, "Edge Count: %u\n", rf.i0);
                    const double tf = (double)rf.i0 / (double)rf.i1;
                    std::fprintf(log, "Fitness: %s\n", fmt_display(tf).c_str());
                    std::fprintf(log, "Trial Elapsed Time: %llds\n", (long long)secs);
                    fitness_tot += tf;
                } else {
                    print_grid(log, R.grids[0][t], R.G[t]);
                    std::fprintf(log, "Starting Fitness: %s\n", fmt_display(R.results[0][t].f).c_str());
                    for (size_t k = 1; k < last; ++k) {
                        if (!R.shown[k][t]) continue;
                        std::fprintf(log, "Step %llu\n", (unsigned long long)at[k - 1]);
                        print_grid(log, R.grids[k][t], R.G[t]);
                        std::fprintf(log, "Fitness: %s\n", SOPSEnvironment::fmt_f32(R.results[k][t].f).c_str());
                    }
                    const float f = R.results[last][t].f;
                    print_grid(log, R.grids[last][t], R.G[t]);
                    std::fprintf(log, "Fitness: %s\n", fmt_display(f).c_str());
                    std::fprintf(log, "Trial Elapsed Time: %llds\n", (long long)secs);
                    fitness_tot = (double)((float)fitness_tot + f);                  // f32 sum upstream
                }
            }
            if (args.verbose) std::fprintf(log, "Kernel launches for %zu trials: %llu\n", trials.size(), (unsigned long long)R.launches);
            if (args.behavior == Agg) std::fprintf(log, "Total Fitness: %s\n", fmt_display(fitness_tot).c_str());
            else std::fprintf(log, "Total Fitness: %s\n", fmt_display((float)fitness_tot).c_str());
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "evosops: %s\n", e.what());
        std::fprintf(log, "Error: %s\n", e.what());
        rc = 1;
    }
    if (log != stdout) std::fclose(log);
    return rc;
}
