// Replacement for the nested par_iter of GeneticAlgo::step_through (EvoSOPS src/GACore/base_ga.rs:393-424).
// `self.ctx: evosops_ffi::Context` is created once in init_ga with `Context::new(&[0])`.
// The cache prefill above (366-391), the cache insert, diversity and generate_new_pop below stay as they are.
{
    use evosops_ffi::{Behavior, RngMode, SopsSize};
    // same skip rule as base_ga.rs:396-398
    let todo: Vec<usize> = (0..self.population.len())
        .filter(|&i| !(gen > 0 && self.population[i].fitness > 0.0))
        .collect();
    if !todo.is_empty() {
        let flat: Vec<u8> = todo.iter()
            .flat_map(|&i| self.population[i].string.iter().flatten().flatten().copied())
            .collect();
        let trials: Vec<(SopsSize, u64)> = trials_vec.iter()
            .map(|t| (SopsSize { arena: t.0.0, layers: t.0.1, coat: 0 }, t.1))
            .collect();
        // fresh move randomness per evaluation, like the reference's never re-seeded fastrand stream
        let move_seeds: Vec<u64> = (0..todo.len() * trials.len()).map(|_| fastrand::u64(..)).collect();
        let out = self.ctx
            .eval_batch(Behavior::Agg, &flat, &trials, granularity, 0.0, 0.0, RngMode::Fast, Some(&move_seeds), None, 0, 0)
            .expect("sops_eval_batch");
        let t = trials.len();
        for (k, &i) in todo.iter().enumerate() {
            let tot: f64 = out[k * t..(k + 1) * t].iter().map(|r| r.norm).sum();
            self.population[i].fitness = tot / (t as f64);          // base_ga.rs:422-423
        }
    }
}
