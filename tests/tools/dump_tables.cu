// Test tool (CPU): prints the kernel header's compile-time direction tables so tests/test_kernel_tables.py can
// check them against a from-scratch restatement of the lattice geometry.  Host code only — no CUDA calls.
#include <cstdio>
#include "../../evosops_b200/csrc/sops_kernels.cuh"

int main() {
    using namespace sops;
    std::printf("{\"ws\": %d, \"dirs\": [", WS);
    for (int d = 0; d < 6; ++d) {
        std::printf("%s{\"dr\": %d, \"dc\": %d, \"delta\": %d, \"tbit\": %u, \"back\": %u, \"mid\": %u, \"front\": %u, \"all\": %u, \"magic\": %u}",
                    d ? ", " : "", dir_dr(d), dir_dc(d), dir_delta(d), dir_tbit(d), dir_back5(d), dir_mid5(d), dir_front5(d),
                    dir_all5(d), dir_magic(d));
    }
    std::printf("], \"patch\": [");
    // win25_patch over every (source offset, direction) around a window whose centre is the padded cell (40, 50)
    const uint32_t centre = 40u | (50u << 8), org = centre - 0x0202u;
    bool first = true;
    for (int orow = -38; orow <= 38; orow += (orow >= -9 && orow < 9) ? 1 : 29)
        for (int ocol = -48; ocol <= 48; ocol += (ocol >= -9 && ocol < 9) ? 1 : 39)
            for (int d = 0; d < 6; ++d) {
                const uint32_t sf = (uint32_t)(40 + orow) | ((uint32_t)(50 + ocol) << 8);
                const uint32_t tf = (sf + (uint32_t)dir_delta(d)) & 0xffffu;
                std::printf("%s[%d, %d, %d, %u]", first ? "" : ", ", orow, ocol, d, win25_patch(sf, tf - sf, org));
                first = false;
            }
    std::printf("], \"plut\": [");
    // The parallel commit rounds' patch look-up (plut_fetch in the header), restated on the host with the header's own
    // table (plut_entry), byte weights (SOPS_PLUT_DP4A) and validity test: v = m + (0x0303 - s_l) with m = s_f | d << 16 |
    // lane_f << 24; valid iff ((v | v + 0x0101) & 0xF8F8) == 0; entry address = dot(bytes(v), bytes(weights)); counted only
    // for an earlier lane (m < lane_l << 24).  Dumped: [row offset, col offset, d, lane_f, lane_l, patch word or 0].
    first = true;
    const uint32_t wts = SOPS_PLUT_DP4A;
    for (int orow = -38; orow <= 38; orow += (orow >= -9 && orow < 9) ? 1 : 29)
        for (int ocol = -48; ocol <= 48; ocol += (ocol >= -9 && ocol < 9) ? 1 : 39)
            for (int d = 0; d < 6; ++d)
                for (int pair = 0; pair < 3; ++pair) {
                    const uint32_t lane_f = pair == 0 ? 3u : pair == 1 ? 17u : 31u, lane_l = pair == 0 ? 9u : pair == 1 ? 17u : 0u;
                    const uint32_t sl = centre;
                    const uint32_t sf = (uint32_t)(40 + orow) | ((uint32_t)(50 + ocol) << 8);
                    const uint32_t m = sf | ((uint32_t)d << 16) | (lane_f << 24), bias = 0x0303u - sl, mykey = lane_l << 24;
                    const uint32_t v = m + bias;
                    const uint32_t z = (v | (v + 0x0101u)) & 0xF8F8u;
                    uint32_t word = 0;
                    if (z == 0u && m < mykey) {
                        uint32_t addr = 0;
                        for (int b = 0; b < 4; ++b) addr += ((v >> (8 * b)) & 0xffu) * ((wts >> (8 * b)) & 0xffu);
                        const uint32_t k = addr / 4u;
                        word = k < SOPS_PLUT_WORDS ? plut_entry((int)(k / 49u), (int)((k % 49u) / 7u), (int)(k % 7u)) : 0xdeadbeefu;
                    }
                    std::printf("%s[%d, %d, %d, %u, %u, %u]", first ? "" : ", ", orow, ocol, d, lane_f, lane_l, word);
                    first = false;
                }
    std::printf("], \"adapt\": [");
    // The adaptive rounds' controller (adapt_* in the header): thresholds per behaviour, the running mean's fixed points
    // under a constant commit count per batch (from a hot start, 400 batches), and the form picked there.
    for (int beh = 0; beh < 4; ++beh) {
        std::printf("%s{\"beh\": %d, \"hot1\": %u, \"hot2\": %u, \"calm\": %u, \"mid\": %u, \"wide_n\": %d, \"steady\": [",
                    beh ? ", " : "", beh, adapt_hot1(beh), adapt_hot2(beh), adapt_calm(beh), adapt_mid(beh), SOPS_WIDE_N);
        for (uint32_t c = 0; c <= 8; ++c) {
            uint32_t ema = 4096;
            for (int b = 0; b < 400; ++b) ema = adapt_ema_next(ema, c);
            const int form = beh == 0 ? adapt_form<0>(ema) : beh == 1 ? adapt_form<1>(ema) : beh == 2 ? adapt_form<2>(ema) : adapt_form<3>(ema);
            std::printf("%s[%u, %u, %d]", c ? ", " : "", c, ema, form);
        }
        std::printf("]}");
    }
    std::printf("]}\n");
    return 0;
}
