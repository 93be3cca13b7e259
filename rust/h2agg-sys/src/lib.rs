//! SOURCE ONLY: written against include/h2agg.h; this image has no cargo/rustc, so this crate has
//! never been compiled here.  It is the binding a maintainer adds to the patched halo2_proofs
//!: zero-copy because Fr/Fq/G1Affine/G1 in halo2curves 0.2.1 are plain
//! `[u64; 4]` Montgomery limbs (checked at start-up below).
#![allow(non_camel_case_types)]
use halo2curves::bn256::{Fr, G1Affine, G1};
use std::os::raw::{c_char, c_int, c_void};
use std::sync::OnceLock;

#[repr(C)]
pub struct h2agg_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn h2agg_init(device_id: c_int, out: *mut *mut h2agg_ctx) -> c_int;
    pub fn h2agg_destroy(ctx: *mut h2agg_ctx);
    pub fn h2agg_last_error(ctx: *mut h2agg_ctx) -> *const c_char;
    pub fn h2agg_host_register(ctx: *mut h2agg_ctx, p: *const c_void, bytes: usize) -> c_int;
    pub fn h2agg_srs_register(ctx: *mut h2agg_ctx, bases: *const u64, n: usize, out_id: *mut u64) -> c_int;
    pub fn h2agg_srs_release(ctx: *mut h2agg_ctx, id: u64) -> c_int;
    pub fn h2agg_msm_g1(ctx: *mut h2agg_ctx, srs_id: u64, bases: *const u64, scalars: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn h2agg_msm_g1_batch(ctx: *mut h2agg_ctx, srs_id: u64, cols: *const *const u64, n_cols: usize, n: usize, out_affine: *mut u64) -> c_int;
    pub fn h2agg_ntt_fr(ctx: *mut h2agg_ctx, a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn h2agg_intt_fr(ctx: *mut h2agg_ctx, a: *mut u64, omega_inv: *const u64, n_inv: *const u64, log_n: u32) -> c_int;
    pub fn h2agg_coeff_to_extended(ctx: *mut h2agg_ctx, coeffs: *const u64, k: u32, ext_k: u32, zeta: *const u64, omega_ext: *const u64, out: *mut u64) -> c_int;
    pub fn h2agg_extended_to_coeff(ctx: *mut h2agg_ctx, a: *mut u64, ext_k: u32, omega_ext_inv: *const u64, ext_n_inv: *const u64, zeta: *const u64, out_len: usize) -> c_int;
    pub fn h2agg_eval_polynomial(ctx: *mut h2agg_ctx, poly: *const u64, n: usize, point: *const u64, out: *mut u64) -> c_int;
    pub fn h2agg_kate_division(ctx: *mut h2agg_ctx, a: *const u64, n: usize, b: *const u64, q: *mut u64) -> c_int;
    pub fn h2agg_permute_expression_pair(ctx: *mut h2agg_ctx, input: *const u64, table: *const u64, usable_rows: usize, permuted_input: *mut u64, permuted_table: *mut u64) -> c_int;
    pub fn h2agg_commit_round_resident(ctx: *mut h2agg_ctx, srs_id: u64, lagrange_cols: *const *const u64, n_cols: usize, k: u32, omega_inv: *const u64, n_inv: *const u64, out_affine: *mut u64, d_lagrange_out: *const *mut c_void, d_coeff_out: *const *mut c_void, ext_k: u32, zeta: *const u64, omega_ext: *const u64, d_ext_out: *const *mut c_void) -> c_int;
    // deferred transforms: the NTT passes of a commit round on the background stream, joined before the quotient /
    // evaluation round reads a coefficient or extended form (include/h2agg.h)
    pub fn h2agg_set_defer_transforms(ctx: *mut h2agg_ctx, enable: c_int) -> c_int;
    pub fn h2agg_transforms_join(ctx: *mut h2agg_ctx) -> c_int;
    pub fn h2agg_transforms_dev(ctx: *mut h2agg_ctx, d_lagrange_cols: *const *const c_void, n_cols: usize, k: u32, omega_inv: *const u64, n_inv: *const u64, d_coeff_out: *const *mut c_void, ext_k: u32, zeta: *const u64, omega_ext: *const u64, d_ext_out: *const *mut c_void) -> c_int;
    // multi-GPU: a window range per column in one batched call (window-sharded MSMs beside whole ones)
    pub fn h2agg_msm_g1_batch_ranges_dev(ctx: *mut h2agg_ctx, srs_id: u64, d_bases: *const c_void, d_cols: *const *const c_void, n_cols: usize, n: usize, win_begins: *const c_int, win_ends: *const c_int, d_out160s: *mut c_void) -> c_int;
}

struct Ctx(*mut h2agg_ctx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // the library serialises entry points internally

static CTX: OnceLock<Ctx> = OnceLock::new();

fn ctx() -> *mut h2agg_ctx {
    CTX.get_or_init(|| {
        // layout guards: the shim is only zero-copy if these hold
        assert_eq!(std::mem::size_of::<Fr>(), 32);
        assert_eq!(std::mem::size_of::<G1Affine>(), 64);
        assert_eq!(std::mem::size_of::<G1>(), 96);
        let one: [u64; 4] = unsafe { std::mem::transmute(Fr::one()) };
        assert_eq!(one, [0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], "Fr is not 4x64 Montgomery");
        let dev: c_int = std::env::var("H2AGG_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p = std::ptr::null_mut();
        let rc = unsafe { h2agg_init(dev, &mut p) };
        assert_eq!(rc, 0, "h2agg_init failed: {}", last_error(std::ptr::null_mut()));
        Ctx(p)
    })
    .0
}

fn last_error(c: *mut h2agg_ctx) -> String {
    unsafe { std::ffi::CStr::from_ptr(h2agg_last_error(c)).to_string_lossy().into_owned() }
}

fn check(rc: c_int) {
    // halo2's functions have no error channel: panic like the asserts they already contain
    if rc != 0 {
        panic!("h2agg: {}", last_error(ctx()));
    }
}

/// Drop-in body for `halo2_proofs::arithmetic::best_multiexp::<G1Affine>`.
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    let mut out = [0u64; 12];
    check(unsafe { h2agg_msm_g1(ctx(), 0, bases.as_ptr() as *const u64, coeffs.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr()) });
    unsafe { std::mem::transmute(out) }
}

/// SRS-resident variant used by the patched `ParamsKZG::{commit, commit_lagrange}`.
pub struct ResidentSrs(u64);
impl ResidentSrs {
    pub fn new(bases: &[G1Affine]) -> Self {
        let mut id = 0u64;
        check(unsafe { h2agg_srs_register(ctx(), bases.as_ptr() as *const u64, bases.len(), &mut id) });
        ResidentSrs(id)
    }
    pub fn msm(&self, coeffs: &[Fr]) -> G1 {
        let mut out = [0u64; 12];
        check(unsafe { h2agg_msm_g1(ctx(), self.0, std::ptr::null(), coeffs.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr()) });
        unsafe { std::mem::transmute(out) }
    }
    /// one commit round -> affine commitments
    pub fn msm_many(&self, cols: &[&[Fr]]) -> Vec<G1Affine> {
        let n = cols[0].len();
        let ptrs: Vec<*const u64> = cols.iter().map(|c| { assert_eq!(c.len(), n); c.as_ptr() as *const u64 }).collect();
        let mut out = vec![G1Affine::default(); cols.len()];
        check(unsafe { h2agg_msm_g1_batch(ctx(), self.0, ptrs.as_ptr(), ptrs.len(), n, out.as_mut_ptr() as *mut u64) });
        out
    }
}
impl Drop for ResidentSrs {
    fn drop(&mut self) {
        unsafe { h2agg_srs_release(ctx(), self.0) };
    }
}

/// Drop-in body for `halo2_proofs::arithmetic::best_fft::<Fr>`.
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    check(unsafe { h2agg_ntt_fr(ctx(), a.as_mut_ptr() as *mut u64, &omega as *const Fr as *const u64, log_n) });
}

pub fn ifft(a: &mut [Fr], omega_inv: Fr, divisor: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    check(unsafe { h2agg_intt_fr(ctx(), a.as_mut_ptr() as *mut u64, &omega_inv as *const Fr as *const u64, &divisor as *const Fr as *const u64, log_n) });
}

pub fn coeff_to_extended(coeffs: &[Fr], k: u32, ext_k: u32, zeta: Fr, omega_ext: Fr) -> Vec<Fr> {
    assert_eq!(coeffs.len(), 1 << k);
    let mut out = vec![Fr::zero(); 1 << ext_k];
    check(unsafe { h2agg_coeff_to_extended(ctx(), coeffs.as_ptr() as *const u64, k, ext_k, &zeta as *const Fr as *const u64, &omega_ext as *const Fr as *const u64, out.as_mut_ptr() as *mut u64) });
    out
}

pub fn extended_to_coeff(a: &mut Vec<Fr>, ext_k: u32, omega_ext_inv: Fr, ext_n_inv: Fr, zeta: Fr, out_len: usize) {
    assert_eq!(a.len(), 1 << ext_k);
    check(unsafe { h2agg_extended_to_coeff(ctx(), a.as_mut_ptr() as *mut u64, ext_k, &omega_ext_inv as *const Fr as *const u64, &ext_n_inv as *const Fr as *const u64, &zeta as *const Fr as *const u64, out_len) });
    a.truncate(out_len);
}

/// Drop-in body for the sort + BTreeMap loop of `plonk::lookup::prover::permute_expression_pair` (usable rows only;
/// the caller appends its blinding rows).  `Err(())` = an input value is missing from the table, which halo2 reports
/// as `Error::ConstraintSystemFailure`.
pub fn permute_expression_pair(input: &[Fr], table: &[Fr]) -> Result<(Vec<Fr>, Vec<Fr>), ()> {
    assert_eq!(input.len(), table.len());
    let (mut a, mut s) = (vec![Fr::zero(); input.len()], vec![Fr::zero(); input.len()]);
    let rc = unsafe { h2agg_permute_expression_pair(ctx(), input.as_ptr() as *const u64, table.as_ptr() as *const u64, input.len(), a.as_mut_ptr() as *mut u64, s.as_mut_ptr() as *mut u64) };
    if rc == 4 {
        return Err(());
    }
    check(rc);
    Ok((a, s))
}

/// Drop-in bodies for `halo2_proofs::arithmetic::{eval_polynomial, kate_division}`.
pub fn eval_polynomial(poly: &[Fr], point: Fr) -> Fr {
    let mut out = Fr::zero();
    check(unsafe { h2agg_eval_polynomial(ctx(), poly.as_ptr() as *const u64, poly.len(), &point as *const Fr as *const u64, &mut out as *mut Fr as *mut u64) });
    out
}
pub fn kate_division(a: &[Fr], b: Fr) -> Vec<Fr> {
    let mut q = vec![Fr::zero(); a.len() - 1];
    check(unsafe { h2agg_kate_division(ctx(), a.as_ptr() as *const u64, a.len(), &b as *const Fr as *const u64, q.as_mut_ptr() as *mut u64) });
    q
}
