//! qcfock-sys -- raw bindings of `include/qcfock.h` and a safe wrapper.
//!
//! Mirrors the header one to one; `tests/test_abi.py` checks the struct layouts against gcc.
//! NOT compiled in the engine's repository (no Rust toolchain there).
#![allow(non_camel_case_types)]

use nalgebra::DMatrix;
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_double, c_int, c_longlong, c_void};
use std::path::Path;

#[repr(C)] pub struct qcf_ctx { _private: [u8; 0] }
#[repr(C)] pub struct qcf_system { _private: [u8; 0] }

#[repr(C)]
pub struct qcf_basis {
    pub n_atoms: c_int, pub z: *const c_int, pub xyz: *const c_double,
    pub n_shells: c_int, pub shell_atom: *const c_int, pub shell_l: *const c_int,
    pub shell_nprim: *const c_int, pub shell_prim_off: *const c_int,
    pub exps: *const c_double, pub coefs: *const c_double, pub cartesian: c_int,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct qcf_opts {
    pub screen_tau: c_double, pub device: c_int, pub rank: c_int, pub world_size: c_int,
    pub block_threads: c_int, pub n_gpus: c_int, pub deterministic: c_int,
}
pub const QCF_TAU_NONE: f64 = -1.0;

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct qcf_stats_t {
    pub n_basis: c_int, pub n_shells: c_int, pub n_pairs: c_int, pub n_groups: c_int,
    pub quartets: c_longlong, pub quartets_total: c_longlong, pub model_flops: c_double,
    pub kernel_ms: c_double, pub total_ms: c_double, pub launches: c_int,
    pub prim_pairs: c_longlong, pub prim_pairs_kept: c_longlong,
    pub create_ms: c_double, pub host_ms: c_double, pub n_devices: c_int, pub graph_launches: c_int,
    pub rank_imbalance: c_double,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct qcf_scf_info {
    pub iteration: c_int, pub converged: c_int, pub electronic_energy: c_double, pub density_rms: c_double,
    pub build_ms: c_double, pub linalg_ms: c_double, pub wall_ms: c_double,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct qcf_launch_rec {
    pub la: c_int, pub lb: c_int, pub kab: c_int, pub lc: c_int, pub ld: c_int, pub kcd: c_int,
    pub nbra: c_int, pub nket: c_int, pub quartets: c_longlong, pub flops_per_prim_quartet: c_double, pub ms: f32,
}

#[link(name = "qcfock")]
extern "C" {
    // loaders: BasisSet::load / MolecularSystem::load (qchem-cli/src/main.rs:76-77)
    pub fn qcf_system_load(basis_json: *const c_char, molecule_json: *const c_char, out: *mut *mut qcf_system) -> c_int;
    pub fn qcf_system_basis(sys: *const qcf_system) -> *const qcf_basis;
    pub fn qcf_system_n_electrons(sys: *const qcf_system) -> c_int;
    pub fn qcf_system_n_basis(sys: *const qcf_system) -> c_int;
    pub fn qcf_system_nuclear_repulsion(sys: *const qcf_system) -> c_double;
    pub fn qcf_system_error(sys: *const qcf_system) -> *const c_char;
    pub fn qcf_system_free(sys: *mut qcf_system);
    // engine
    pub fn qcf_create(b: *const qcf_basis, o: *const qcf_opts, out: *mut *mut qcf_ctx) -> c_int;
    pub fn qcf_nbasis(ctx: *const qcf_ctx) -> c_int;
    pub fn qcf_build_rhf(ctx: *mut qcf_ctx, p: *const c_double, g: *mut c_double) -> c_int;
    pub fn qcf_build_uhf(ctx: *mut qcf_ctx, pa: *const c_double, pb: *const c_double, ga: *mut c_double, gb: *mut c_double) -> c_int;
    pub fn qcf_build_jk(ctx: *mut qcf_ctx, nd: c_int, p: *const *const c_double, j: *const *mut c_double, k: *const *mut c_double) -> c_int;
    pub fn qcf_build_rhf_incremental(ctx: *mut qcf_ctx, p: *const c_double, g: *mut c_double, reset: c_int) -> c_int;
    pub fn qcf_build_uhf_incremental(ctx: *mut qcf_ctx, pa: *const c_double, pb: *const c_double, ga: *mut c_double,
                                     gb: *mut c_double, reset: c_int) -> c_int;
    pub fn qcf_build_rhf_dev(ctx: *mut qcf_ctx, dp: *const c_double, dg: *mut c_double, stream: *mut c_void) -> c_int;
    pub fn qcf_build_uhf_dev(ctx: *mut qcf_ctx, dpa: *const c_double, dpb: *const c_double, dga: *mut c_double,
                             dgb: *mut c_double, stream: *mut c_void) -> c_int;
    pub fn qcf_one_electron(ctx: *mut qcf_ctx, s: *mut c_double, t: *mut c_double, v: *mut c_double) -> c_int;
    // device-resident SCF iteration (rhf.rs:66-104, uhf.rs:79-189 on the GPU)
    pub fn qcf_scf_init(ctx: *mut qcf_ctx, s: *const c_double, h: *const c_double, unrestricted: c_int, n_alpha: c_int,
                        n_beta: c_int, full_rebuild_every: c_int) -> c_int;
    pub fn qcf_scf_step(ctx: *mut qcf_ctx, epsilon: c_double, out: *mut qcf_scf_info) -> c_int;
    pub fn qcf_scf_get(ctx: *mut qcf_ctx, what: c_int, spin: c_int, out: *mut c_double) -> c_int;
    // diagnostics / parity hooks (used by the test-suite, not by the SCF drivers)
    pub fn qcf_eri_quartet(ctx: *mut qcf_ctx, sa: c_int, sb: c_int, sc: c_int, sd: c_int, out: *mut c_double) -> c_int;
    pub fn qcf_schwarz(ctx: *mut qcf_ctx, q: *mut c_double) -> c_int;
    pub fn qcf_boys(ctx: *mut qcf_ctx, mmax: c_int, n: c_int, t: *const c_double, f: *mut c_double) -> c_int;
    pub fn qcf_fp64_peak(ctx: *mut qcf_ctx, tflops: *mut c_double) -> c_int;
    pub fn qcf_launch_profile(ctx: *mut qcf_ctx, max_rec: c_int, out: *mut qcf_launch_rec) -> c_int;
    pub fn qcf_device_times(ctx: *mut qcf_ctx, max_dev: c_int, ms: *mut c_double) -> c_int;
    pub fn qcf_stats(ctx: *const qcf_ctx, out: *mut qcf_stats_t) -> c_int;
    pub fn qcf_last_error(ctx: *const qcf_ctx) -> *const c_char;
    pub fn qcf_destroy(ctx: *mut qcf_ctx);
}

/// Error of the engine: status code + the text of `qcf_last_error`.
#[derive(Debug)]
pub struct FockError { pub code: i32, pub message: String }
impl std::fmt::Display for FockError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result { write!(f, "qcfock error {}: {}", self.code, self.message) }
}
impl std::error::Error for FockError {}

/// Basis + molecule loaded by the library's own JSON loaders (replaces `BasisSet::load` + `MolecularSystem::load`).
pub struct System { raw: *mut qcf_system }
impl System {
    pub fn load(basis_json: &Path, molecule_json: &Path) -> Result<Self, FockError> {
        let b = CString::new(basis_json.to_string_lossy().as_bytes()).unwrap();
        let m = CString::new(molecule_json.to_string_lossy().as_bytes()).unwrap();
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { qcf_system_load(b.as_ptr(), m.as_ptr(), &mut raw) };
        if rc != 0 {
            let message = if raw.is_null() { "qcf_system_load failed".into() }
                          else { unsafe { CStr::from_ptr(qcf_system_error(raw)) }.to_string_lossy().into_owned() };
            if !raw.is_null() { unsafe { qcf_system_free(raw) } }
            return Err(FockError { code: rc, message });
        }
        Ok(System { raw })
    }
    pub fn n_basis(&self) -> usize { unsafe { qcf_system_n_basis(self.raw) as usize } }
    pub fn n_electrons(&self) -> usize { unsafe { qcf_system_n_electrons(self.raw) as usize } }
    pub fn nuclear_repulsion(&self) -> f64 { unsafe { qcf_system_nuclear_repulsion(self.raw) } }
}
impl Drop for System { fn drop(&mut self) { unsafe { qcf_system_free(self.raw) } } }

/// Owns a `qcf_ctx`.  `build_rhf` / `build_uhf` are the per-iteration calls that stand in for
/// `rhf.rs:67-68` and `uhf.rs:90-91`.
pub struct FockEngine { ctx: *mut qcf_ctx, n: usize }

impl FockEngine {
    /// `n_gpus` > 1: one context drives that many GPUs from the calling thread.
    pub fn new(system: &System, tau: f64, n_gpus: i32, deterministic: bool) -> Result<Self, FockError> {
        let opts = qcf_opts { screen_tau: tau, device: 0, rank: 0, world_size: 1, block_threads: 0, n_gpus,
                              deterministic: deterministic as c_int };
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { qcf_create(qcf_system_basis(system.raw), &opts, &mut ctx) };
        if rc != 0 {
            let message = if ctx.is_null() { "qcf_create failed".into() }
                          else { unsafe { CStr::from_ptr(qcf_last_error(ctx)) }.to_string_lossy().into_owned() };
            if !ctx.is_null() { unsafe { qcf_destroy(ctx) } }
            return Err(FockError { code: rc, message });
        }
        let n = unsafe { qcf_nbasis(ctx) } as usize;
        Ok(FockEngine { ctx, n })
    }
    fn check(&self, rc: c_int) -> Result<(), FockError> {
        if rc == 0 { return Ok(()); }
        let message = unsafe { CStr::from_ptr(qcf_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(FockError { code: rc, message })
    }
    /// (S, T, V): stands in for molint::overlap / kinetic / nuclear (rhf.rs:41-43).
    pub fn one_electron(&mut self) -> Result<(DMatrix<f64>, DMatrix<f64>, DMatrix<f64>), FockError> {
        let (mut s, mut t, mut v) = (DMatrix::zeros(self.n, self.n), DMatrix::zeros(self.n, self.n), DMatrix::zeros(self.n, self.n));
        self.check(unsafe { qcf_one_electron(self.ctx, s.as_mut_ptr(), t.as_mut_ptr(), v.as_mut_ptr()) })?;
        Ok((s, t, v))
    }
    /// G = J[P] - K[P]/2 (replaces rhf.rs:58-62 + 152-167).  DMatrix is column-major; P and G are symmetric.
    pub fn build_rhf(&mut self, density: &DMatrix<f64>) -> Result<DMatrix<f64>, FockError> {
        let mut g = DMatrix::zeros(self.n, self.n);
        self.check(unsafe { qcf_build_rhf(self.ctx, density.as_ptr(), g.as_mut_ptr()) })?;
        Ok(g)
    }
    /// (G_alpha, G_beta) = (J[Pa+Pb] - K[Pa], J[Pa+Pb] - K[Pb]) (replaces the two calls at uhf.rs:90-91).
    pub fn build_uhf(&mut self, pa: &DMatrix<f64>, pb: &DMatrix<f64>) -> Result<(DMatrix<f64>, DMatrix<f64>), FockError> {
        let (mut ga, mut gb) = (DMatrix::zeros(self.n, self.n), DMatrix::zeros(self.n, self.n));
        self.check(unsafe { qcf_build_uhf(self.ctx, pa.as_ptr(), pb.as_ptr(), ga.as_mut_ptr(), gb.as_mut_ptr()) })?;
        Ok((ga, gb))
    }
    pub fn build_rhf_incremental(&mut self, density: &DMatrix<f64>, reset: bool) -> Result<DMatrix<f64>, FockError> {
        let mut g = DMatrix::zeros(self.n, self.n);
        self.check(unsafe { qcf_build_rhf_incremental(self.ctx, density.as_ptr(), g.as_mut_ptr(), reset as c_int) })?;
        Ok(g)
    }
    pub fn stats(&self) -> Result<qcf_stats_t, FockError> {
        let mut st = qcf_stats_t::default();
        self.check(unsafe { qcf_stats(self.ctx, &mut st) })?;
        Ok(st)
    }
}
impl Drop for FockEngine { fn drop(&mut self) { unsafe { qcf_destroy(self.ctx) } } }
