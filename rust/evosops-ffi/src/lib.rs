//! Rust binding of `include/evosops.h` (libevosops_b200.so).
//!
//! One `Context::eval_batch` call replaces the nested rayon `par_iter` of `GACore::*::step_through`
//! (EvoSOPS `src/GACore/base_ga.rs:393-424`, `sep_ga.rs:419-457`, `coat_ga.rs:410-436`, `loco_ga.rs:419-433`):
//! every (genome, size, seed) instance runs as one warp-resident chain on the GPU.
//!
//! NOT compiled in the repository's build image (no cargo/rustc there); kept in sync with the header by hand.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct SopsCtx { _private: [u8; 0] }

/// `(arena_layers, particle_layers | object_layers, coat_layers)` — the `-k (a,p[,c])` tuple (main.rs:141-147).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct SopsSize { pub arena: u16, pub layers: u16, pub coat: u16 }

/// Per-instance outcome; see the header for the meaning of `i0..i2` per behaviour.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct SopsResult { pub i0: u32, pub i1: u32, pub i2: u32, pub f: f32, pub norm: f64 }

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Behavior { Agg = 0, Sep = 1, Coat = 2, Loco = 3 }

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum RngMode { Verify = 0, Fast = 1 }

pub const F_DUMP: u32 = 1;
pub const F_INIT_ONLY: u32 = 2;
pub const F_THEORY_TABLE: u32 = 4;
pub const F_FORCE_WARP: u32 = 8;
pub const F_FORCE_CTA: u32 = 16;
pub const F_ROUNDS_SERIAL: u32 = 32;
pub const F_ROUNDS_PARALLEL: u32 = 64;
pub const F_ROUNDS_ADAPTIVE: u32 = 128;

extern "C" {
    fn sops_ctx_create(out: *mut *mut SopsCtx, device_ids: *const c_int, n_dev: c_int) -> c_int;
    fn sops_ctx_destroy(ctx: *mut SopsCtx);
    fn sops_last_error(ctx: *const SopsCtx) -> *const c_char;
    fn sops_eval_batch(ctx: *mut SopsCtx, behavior: c_int, genomes: *const u8, p: u32,
                       trial_sizes: *const SopsSize, trial_seeds: *const u64, t: u32,
                       gran: u8, w1: f32, w2: f32, rng_mode: c_int,
                       move_seeds: *const u64, loco_orient: *const u8,
                       steps_override: u64, flags: u32, out: *mut SopsResult) -> c_int;
    fn sops_batch_counters(ctx: *mut SopsCtx, out: *mut u64) -> c_int;
    fn sops_ctx_set_trial_steps(ctx: *mut SopsCtx, steps: *const u64, t: u32) -> c_int;
    fn sops_dump_state(ctx: *mut SopsCtx, instance: u32, grid: *mut u8, particles: *mut u8) -> c_int;
    fn sops_instance_shape(behavior: c_int, size: SopsSize, grid_side: *mut u32, n_particles: *mut u32,
                           sim_duration: *mut u64) -> c_int;
}

#[derive(Debug)]
pub struct SopsError { pub code: i32, pub message: String }

/// Owns the streams and device buffers of one or more GPUs.  Not `Sync`: use from one thread at a time.
pub struct Context { raw: *mut SopsCtx }

unsafe impl Send for Context {}

impl Context {
    pub fn new(device_ids: &[i32]) -> Result<Self, SopsError> {
        let mut raw: *mut SopsCtx = std::ptr::null_mut();
        let rc = unsafe { sops_ctx_create(&mut raw, device_ids.as_ptr(), device_ids.len() as c_int) };
        if rc != 0 { return Err(SopsError { code: rc, message: "sops_ctx_create failed (no CUDA device?)".into() }); }
        Ok(Context { raw })
    }

    fn err(&self, code: i32) -> SopsError {
        let message = unsafe { CStr::from_ptr(sops_last_error(self.raw)) }.to_string_lossy().into_owned();
        SopsError { code, message }
    }

    /// `genomes`: `P * L` bytes, row-major over the Rust genome array (`[[[u8;4];3];4]` flattens to exactly this).
    /// `trials`: the flattened `(size, seed)` list the GA builds (sizes-major / seeds-minor; loco: zip + repeat).
    /// Returns `P * trials.len()` results, instance `g * T + t`.
    #[allow(clippy::too_many_arguments)]
    pub fn eval_batch(&mut self, behavior: Behavior, genomes: &[u8], trials: &[(SopsSize, u64)], gran: u8,
                      w1: f32, w2: f32, rng: RngMode, move_seeds: Option<&[u64]>, loco_orient: Option<&[u8]>,
                      steps_override: u64, flags: u32) -> Result<Vec<SopsResult>, SopsError> {
        let l = match behavior { Behavior::Agg => 48, Behavior::Loco => 144, _ => 600 };
        assert!(genomes.len() % l == 0, "genome bytes must be a multiple of the behaviour's genome length");
        let p = genomes.len() / l;
        let sizes: Vec<SopsSize> = trials.iter().map(|t| t.0).collect();
        let seeds: Vec<u64> = trials.iter().map(|t| t.1).collect();
        if let Some(ms) = move_seeds { assert_eq!(ms.len(), p * trials.len()); }
        if let Some(o) = loco_orient { assert_eq!(o.len(), p * trials.len()); }
        let mut out = vec![SopsResult::default(); p * trials.len()];
        let rc = unsafe {
            sops_eval_batch(self.raw, behavior as c_int, genomes.as_ptr(), p as u32, sizes.as_ptr(), seeds.as_ptr(),
                            trials.len() as u32, gran, w1, w2, rng as c_int,
                            move_seeds.map_or(std::ptr::null(), |m| m.as_ptr()),
                            loco_orient.map_or(std::ptr::null(), |o| o.as_ptr()),
                            steps_override, flags, out.as_mut_ptr())
        };
        if rc != 0 { return Err(self.err(rc)); }
        Ok(out)
    }

    /// Per-trial chain lengths for the next `eval_batch` (the standalone run's snapshots); an empty slice clears.
    pub fn set_trial_steps(&mut self, steps: &[u64]) -> Result<(), SopsError> {
        let rc = unsafe { sops_ctx_set_trial_steps(self.raw, if steps.is_empty() { std::ptr::null() } else { steps.as_ptr() }, steps.len() as u32) };
        if rc != 0 { return Err(self.err(rc)); }
        Ok(())
    }

    /// (attempts, possible, committed, kernel launches, h2d bytes, d2h bytes) of the last batch.
    pub fn counters(&mut self) -> Result<[u64; 6], SopsError> {
        let mut c = [0u64; 6];
        let rc = unsafe { sops_batch_counters(self.raw, c.as_mut_ptr()) };
        if rc != 0 { return Err(self.err(rc)); }
        Ok(c)
    }

    /// Final lattice (`print_grid` cell codes, row-major `G*G`) and particles (`x, y, state` triples) of one instance
    /// of the last batch; the batch must have been run with `F_DUMP`.
    pub fn dump_state(&mut self, instance: u32, behavior: Behavior, size: SopsSize) -> Result<(Vec<u8>, Vec<u8>), SopsError> {
        let (mut g, mut n, mut d) = (0u32, 0u32, 0u64);
        let rc = unsafe { sops_instance_shape(behavior as c_int, size, &mut g, &mut n, &mut d) };
        if rc != 0 { return Err(SopsError { code: rc, message: "invalid size".into() }); }
        let mut grid = vec![0u8; (g * g) as usize];
        let mut parts = vec![0u8; (n * 3) as usize];
        let rc = unsafe { sops_dump_state(self.raw, instance, grid.as_mut_ptr(), parts.as_mut_ptr()) };
        if rc != 0 { return Err(self.err(rc)); }
        Ok((grid, parts))
    }
}

impl Drop for Context {
    fn drop(&mut self) { unsafe { sops_ctx_destroy(self.raw) } }
}
