// evosops_hosttest — exercises the host-side mirror (GA operators, trial lists, log and file formats) with a
// STUB evaluator, so the CPU test-suite can check it without a GPU.  Not part of the product path.
//   hosttest ga <behavior> <pop> <gen> <gran> <mut> <ga_seed> <out_dir>       runs a GA against the stub
//   hosttest parse <file>                                                    prints the parsed genome bytes
//   hosttest trials <behavior>                                               prints the flattened trial list
#include <cstdlib>
#include <cstring>
#include <string>

#include "ga.hpp"

using namespace evosops;

// fitness = 1 - mean(gene)/granularity, independent of the trial: a GA must drive genes towards 0
class StubEvaluator : public Evaluator {
public:
    explicit StubEvaluator(int gran) : gran_(gran) {}
    void eval(int behavior, const std::vector<uint8_t>& genomes, uint32_t P, const std::vector<Trial>& trials, uint8_t, float, float,
              const std::vector<uint64_t>& ms, const std::vector<uint8_t>& orient, uint64_t, uint32_t, std::vector<sops_result>& out) override {
        const int L = genome_len(behavior);
        out.assign((size_t)P * trials.size(), sops_result());
        if (!ms.empty() && ms.size() != (size_t)P * trials.size()) std::abort();
        if (behavior == Loco && orient.size() != (size_t)P * trials.size()) std::abort();
        for (uint32_t g = 0; g < P; ++g) {
            double sum = 0;
            for (int k = 0; k < L; ++k) sum += genomes[(size_t)g * L + k];
            const double f = 1.0 - sum / ((double)L * gran_);
            for (size_t t = 0; t < trials.size(); ++t) { out[g * trials.size() + t].norm = f; out[g * trials.size() + t].f = (float)f; }
        }
        ++calls;
        instances += (size_t)P * trials.size();
    }
    void dump(uint32_t, std::vector<uint8_t>&, std::vector<uint8_t>&, uint32_t, uint32_t) override {}
    size_t calls = 0, instances = 0;
private:
    int gran_;
};

std::string evosops::SOPSEnvironment::fmt_f32(float v) { return fmt_display(v); }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    auto beh = [](const char* s) { return !strcmp(s, "agg") ? Agg : !strcmp(s, "sep") ? Sep : !strcmp(s, "coat") ? Coat : Loco; };
    if (cmd == "parse" && argc == 3) {
        for (uint8_t v : read_genome_file(argv[2])) std::printf("%u ", (unsigned)v);
        std::printf("\n");
        return 0;
    }
    if (cmd == "fmt") {
        std::printf("%s %s %s %s %s\n", fmt_display(0.65f).c_str(), fmt_display(0.08).c_str(), fmt_debug(1.0).c_str(),
                    fmt_display(1.0).c_str(), fmt_display(0.5194300637633971).c_str());
        return 0;
    }
    if (cmd == "trials" && argc == 3) {
        StubEvaluator ev(10);
        std::vector<sops_size> sizes = {{6, 3, 1}, {10, 7, 2}, {13, 9, 3}};
        GeneticAlgo ga(ev, beh(argv[2]), 2, 1, 0, 0.08, 10, true, sizes, {11, 22, 33}, 0.65f, 0.35f, 7, 1, "/tmp", stderr);
        for (const Trial& t : ga.trials_vec()) std::printf("(%u,%u,%u):%llu ", t.size.arena, t.size.layers, t.size.coat, (unsigned long long)t.seed);
        std::printf("\n");
        return 0;
    }
    if (cmd == "ga" && argc == 9) {
        const int b = beh(argv[2]);
        const int gran = std::atoi(argv[5]);
        StubEvaluator ev(gran);
        std::vector<sops_size> sizes = {{6, 3, 1}, {10, 7, 2}};
        GeneticAlgo ga(ev, b, (uint16_t)std::atoi(argv[3]), (uint16_t)std::atoi(argv[4]), 0, std::atof(argv[6]), (uint8_t)gran, true, sizes,
                       {1, 2, 3}, 0.65f, 0.35f, 424242u, std::strtoull(argv[7], nullptr, 10), argv[8], stdout);
        ga.run_through();
        double best = 0;
        for (auto& g : ga.population()) (void)g;
        std::printf("EVAL_CALLS %zu INSTANCES %zu MUT %s LOWER_HIT %d\n", ev.calls, ev.instances, fmt_display(ga.mut_rate()).c_str(), (int)ga.lower_hit());
        (void)best;
        return 0;
    }
    return 2;
}
