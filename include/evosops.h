/* evosops.h — C ABI of libevosops_b200.so: batched SOPS genome-fitness evaluation on B200 (sm_100a).
 *
 * The reference (DaymudeLab/EvoSOPS) has no FFI; the seam this library replaces is the Rust module
 * API that GACore and main.rs call once per (genome, size, seed) inside a nested rayon par_iter:
 *
 *   SOPSEnvironment::init_sops_env(&genome, arena, particle, seed, gran)      src/SOPSCore/mod.rs:106
 *   SOPSepEnvironment::init_sops_env(.., w1, w2)                              src/SOPSCore/separation.rs:100
 *   SOPSCoatEnvironment::init_sops_env(.., object, coat, .., dist_grid)       src/SOPSCore/coating.rs:229
 *   SOPSLocoEnvironment::init_sops_env(.., w1, w2)                            src/SOPSCore/locomotion.rs:98
 *   .simulate(take_snaps) / .evaluate_fitness() / .get_max_fitness()          mod.rs:306,296,320
 *   callers: src/GACore/base_ga.rs:393-411, sep_ga.rs:419-443, coat_ga.rs:410-436,
 *            loco_ga.rs:419-433, src/main.rs:252-274,305-323,360-380,415-433
 *
 * One sops_eval_batch call evaluates a whole population x trial list (what one GA generation's
 * nested par_iter does).  Conventions: every function returns 0 on success or a negative SOPS_E_*
 * code (message via sops_last_error); nothing aborts; all buffers are caller-owned HOST memory and no
 * pointer is retained after return; a ctx is used from one thread at a time.  There is no CPU
 * fallback: without a CUDA device sops_ctx_create fails.
 */
#ifndef EVOSOPS_H
#define EVOSOPS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sops_ctx sops_ctx;

enum { SOPS_AGG = 0, SOPS_SEP = 1, SOPS_COAT = 2, SOPS_LOCO = 3 };   /* main.rs:80-86 Behavior */
enum { SOPS_RNG_VERIFY = 0, SOPS_RNG_FAST = 1 };

/* genome bytes per behaviour, row-major over the Rust array dims:
 * agg [[[u8;4];3];4]=48 (GACore/mod.rs:7), sep/coat [[[u8;10];6];10]=600, loco [[[[u8;4];3];4];3]=144 */
#define SOPS_GENOME_LEN_AGG 48
#define SOPS_GENOME_LEN_SEP 600
#define SOPS_GENOME_LEN_COAT 600
#define SOPS_GENOME_LEN_LOCO 144

enum {
    SOPS_E_ARG = -1,        /* null pointer / bad enum / empty batch                              */
    SOPS_E_GENE = -2,       /* gene value indexes past the probability table (reference panics)   */
    SOPS_E_SIZE = -3,       /* more particles than free cells (reference never terminates), a>127 */
    SOPS_E_UNSUPPORTED = -4,/* arena_layers > 125 (loco: > 63): outside the kernels' u8 coordinates   */
    SOPS_E_CUDA = -5,       /* CUDA runtime error                                                 */
    SOPS_E_STATE = -6,      /* call order (e.g. collect before launch, dump without SOPS_F_DUMP)  */
    SOPS_E_INIT_REJECT = -7 /* a placement draw hit the Uniform zone rejection (p < 2^-58)        */
};

/* (arena_layers, particle_layers[, coat_layers]) — the `-k (a,p[,c])` tuple, main.rs:141-147,176.
 * For SOPS_COAT `layers` is object_layers. */
typedef struct { uint16_t arena, layers, coat; } sops_size;

/* Per-instance outcome.  Integers are exact; f is the reference's f32 fitness; norm is the f64 value
 * the GA sums over trials (base_ga.rs:409, sep_ga.rs:441).
 *   agg : i0 = edges (mod.rs:296-301), i1 = max_fitness 3p(3p+1)
 *   sep : i0 = E (sum of occupied nbrs), i1 = S (sum of same-colour nbrs)   separation.rs:811-825
 *   coat: i0 = min_total_dist, i1 = par_total_dist                          coating.rs:561-574
 *   loco: i0 = sum of nbr counts, i1 = distance sum, i2 = orientation       locomotion.rs:661-688 */
typedef struct { uint32_t i0, i1, i2; float f; double norm; } sops_result;

enum {
    SOPS_F_DUMP = 1,        /* keep final lattices + particle lists for sops_dump_state            */
    SOPS_F_INIT_ONLY = 2,   /* run 0 chain steps (init_sops_env + evaluate_fitness only)           */
    SOPS_F_THEORY_TABLE = 4,/* use theory_gene_probability() (mod.rs:91-93, separation.rs:82-84)   */
    /* Kernel choice (results are bit-identical either way; default: one warp per chain) */
    SOPS_F_FORCE_WARP = 8,  /* one warp per chain (the default, spelled out)                        */
    SOPS_F_FORCE_CTA = 16,  /* one 4-warp CTA per chain (fast mode): pays for few, low-acceptance chains */
    /* Commit-round form of the warp kernel (bit-identical; default: chosen from the batch size) */
    SOPS_F_ROUNDS_SERIAL = 32,   /* one commit per round, only lanes whose window changed re-derive          */
    SOPS_F_ROUNDS_PARALLEL = 64, /* all accepted lanes of a batch commit at once (fixed-point patch rounds;
                                    separation: arenas up to 29 layers, else serial)                         */
    SOPS_F_ROUNDS_ADAPTIVE = 128 /* fast mode: every chain switches on its own acceptance rate between the parallel
                                    rounds, the serial rounds and (not separation) serial rounds over three batches
                                    evaluated side by side (the default for batches that leave at most two chains per
                                    warp scheduler, i.e. GA generations; where the form does not exist — verification
                                    mode, separation arenas above 29 layers — the flag falls back to the parallel or
                                    serial one)                                                                  */
};

/* Create a context over n_dev CUDA devices (device_ids NULL -> devices 0..n_dev-1). */
int sops_ctx_create(sops_ctx** out, const int* device_ids, int n_dev);
void sops_ctx_destroy(sops_ctx* ctx);
const char* sops_last_error(const sops_ctx* ctx);

/* Evaluate P genomes x T trials.  Instance q = g*T + t runs
 *   env = init_sops_env(genomes[g], sizes[t], seeds[t], ...); env.simulate(false)
 * trial_sizes/trial_seeds are the flattened trial list the GA builds (base_ga.rs:356-362 sizes-major,
 * seeds-minor; loco_ga.rs:384-389 zip+repeat).  move_seeds[P*T] (NULL -> trial seed) is the state the
 * instance's fastrand thread-local generator starts from (verify) / the Philox key (fast).
 * loco_orient[P*T] in 1..3 (NULL -> 1) replaces locomotion.rs:109's thread_rng draw.
 * steps_override 0 -> n^3 (mod.rs:143).  out[P*T] in instance order. */
int sops_eval_batch(sops_ctx* ctx, int behavior, const uint8_t* genomes, uint32_t P,
                    const sops_size* trial_sizes, const uint64_t* trial_seeds, uint32_t T,
                    uint8_t gran, float w1, float w2, int rng_mode,
                    const uint64_t* move_seeds, const uint8_t* loco_orient,
                    uint64_t steps_override, uint32_t flags, sops_result* out);

/* Per-trial chain lengths for the NEXT sops_eval_batch / sops_batch_prepare on this context (copied; consumed by that
 * call, which must pass the same T): trial t runs steps[t] moves instead of n^3 / steps_override.  The standalone
 * `-e gm` path uses it to take the reference's snapshots (after n, n^2, ... steps, mod.rs:309, separation.rs:833,
 * coating.rs:582 — counts that differ per size) of ALL trials with one launch per snapshot.  NULL or T = 0 clears. */
int sops_ctx_set_trial_steps(sops_ctx* ctx, const uint64_t* steps, uint32_t T);

/* The same call split in three so a caller (bench.py) can time the device part with the inputs
 * already resident in HBM: prepare = validate + H2D, launch = enqueue the chain kernels (async),
 * collect = wait + D2H + float fitness.  launch may be repeated on one prepared batch. */
int sops_batch_prepare(sops_ctx* ctx, int behavior, const uint8_t* genomes, uint32_t P,
                       const sops_size* trial_sizes, const uint64_t* trial_seeds, uint32_t T,
                       uint8_t gran, float w1, float w2, int rng_mode,
                       const uint64_t* move_seeds, const uint8_t* loco_orient,
                       uint64_t steps_override, uint32_t flags);
int sops_batch_launch(sops_ctx* ctx);
int sops_batch_collect(sops_ctx* ctx, sops_result* out);

/* Wait for the launched kernels without copying results; device_ms = max over devices of the
 * CUDA-event time of the last launch (first kernel start to last kernel end on that device). */
int sops_batch_sync(sops_ctx* ctx, double* device_ms);

/* Totals over the last collected batch: out[0] = move attempts, out[1] = attempts whose target was
 * enterable (p_poss numerator), out[2] = committed moves/swaps (p_acc numerator),
 * out[3] = kernel launches, out[4] = h2d bytes, out[5] = d2h bytes. */
int sops_batch_counters(sops_ctx* ctx, uint64_t out[6]);

/* Final state of instance q of the last collected batch (needs SOPS_F_DUMP): grid[G*G] with the
 * reference's cell codes (print_grid, mod.rs:153-161) and particles[n*3] = (x, y, state) in
 * participants order.  Either pointer may be NULL. */
int sops_dump_state(sops_ctx* ctx, uint32_t instance, uint8_t* grid, uint8_t* particles);

/* Shape helpers (pure host arithmetic): grid side 2a+1, particle count, chain length n^3. */
int sops_instance_shape(int behavior, sops_size size, uint32_t* grid_side, uint32_t* n_particles,
                        uint64_t* sim_duration);

/* Coating distance grid (coating.rs:101-129,341-356): dist[G*G] u16, 0xFFFF where the reference has
 * no map entry (object/boundary cells).  Host arithmetic; the GA precomputes it per size
 * (coat_ga.rs:508-537). */
int sops_coat_distance_grid(sops_size size, uint16_t* dist);

#ifdef __cplusplus
}
#endif
#endif /* EVOSOPS_H */
