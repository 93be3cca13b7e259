//! SOURCE ONLY -- written against include/h2agg.h and the reference's traits; never compiled here (no cargo/rustc).
//!
//! The B200 chip set: a recording implementation of the reference's plugin surface
//!   ArithCommonChip   halo2-snark-aggregator-api/src/arith/common.rs:3-42
//!   ArithFieldChip    halo2-snark-aggregator-api/src/arith/field.rs:6-105      (B200ScalarChip; also the NativeChip)
//!   ArithEccChip      halo2-snark-aggregator-api/src/arith/ecc.rs:5-61         (B200EccChip)
//!   Encode            halo2-snark-aggregator-api/src/transcript/encode.rs:3-23 (B200PoseidonEncode)
//! with exactly the operand semantics of the circuit adapters it replaces
//! (halo2-snark-aggregator-circuit/src/chips/{scalar_chip,ecc_chip,encode_chip}.rs).  `verify_aggregation_proofs_in_chip`
//! (api/src/systems/halo2/verify.rs:835-942) is generic over these traits, so the reference's own verifier code drives
//! this chip set unchanged; the calls only RECORD (row layout + the values the chain needs), and
//! `B200Context::expand_advice` runs the sm_100a kernel that writes the five advice columns in one pass
//! (h2agg_witness_expand / _dev).  Python twin with the same method names: halo2_snark_aggregator_b200/witness.py,
//! exercised against the circuit-chip oracle in tests/test_aggregation_cpu.py and tests/test_gpu_witness.py.
#![allow(non_camel_case_types)]
use halo2_proofs::arithmetic::{CurveAffine, Field, FieldExt};
use halo2_proofs::plonk::Error;
use halo2_snark_aggregator_api::arith::{common::ArithCommonChip, ecc::ArithEccChip, field::ArithFieldChip};
use halo2_snark_aggregator_api::transcript::encode::Encode;
use halo2curves::bn256::{Fq, Fr, G1Affine};
use halo2curves::group::ff::PrimeField;
use std::fmt;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct h2agg_witness {
    _private: [u8; 0],
}

extern "C" {
    fn h2agg_wit_new() -> *mut h2agg_witness;
    fn h2agg_wit_set_threads(n: c_int) -> c_int; // host threads a multi_exp records its independent sections on
    fn h2agg_wit_free(w: *mut h2agg_witness);
    fn h2agg_wit_error(w: *mut h2agg_witness) -> *const c_char;
    fn h2agg_wit_rows(w: *mut h2agg_witness) -> u64;
    fn h2agg_wit_assign_point(w: *mut h2agg_witness, xy: *const u64) -> i64;
    fn h2agg_wit_assign_constant_point(w: *mut h2agg_witness, xy: *const u64) -> i64;
    fn h2agg_wit_assign_scalar(w: *mut h2agg_witness, s: *const u64) -> i64;
    fn h2agg_wit_ecc_assign_identity(w: *mut h2agg_witness) -> i64;
    fn h2agg_wit_ecc_add(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_ecc_sub(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_ecc_reduce(w: *mut h2agg_witness, a: i64) -> i64;
    fn h2agg_wit_ecc_mul(w: *mut h2agg_witness, a: i64, s: i64) -> i64;
    fn h2agg_wit_ecc_shamir(w: *mut h2agg_witness, pts: *const i64, scalars: *const i64, n: usize) -> i64;
    fn h2agg_wit_ecc_constant_mul(w: *mut h2agg_witness, base_xy: *const u64, s: i64) -> i64;
    fn h2agg_wit_point_value(w: *mut h2agg_witness, h: i64, out_xy: *mut u64, is_identity: *mut c_int) -> c_int;
    fn h2agg_wit_field_assign_const(w: *mut h2agg_witness, c: *const u64) -> i64;
    fn h2agg_wit_field_add(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_field_sub(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_field_mul(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_field_square(w: *mut h2agg_witness, a: i64) -> i64;
    fn h2agg_wit_field_div(w: *mut h2agg_witness, a: i64, b: i64) -> i64;
    fn h2agg_wit_field_sum_with_coeff_and_constant(w: *mut h2agg_witness, elems: *const i64, coeffs: *const u64, n: usize, constant: *const u64) -> i64;
    fn h2agg_wit_field_mul_add_constant(w: *mut h2agg_witness, a: i64, b: i64, c: *const u64) -> i64;
    fn h2agg_wit_scalar_value(w: *mut h2agg_witness, h: i64, out: *mut u64) -> c_int;
    fn h2agg_wit_scalar_cell(w: *mut h2agg_witness, h: i64, column: *mut u32, row: *mut u32) -> c_int;
    fn h2agg_wit_encode_point(w: *mut h2agg_witness, point: i64, out: *mut i64) -> c_int;
    fn h2agg_wit_ecc_assert_equal(w: *mut h2agg_witness, a: i64, b: i64) -> c_int;
    fn h2agg_wit_assert_not_identity(w: *mut h2agg_witness, p: i64) -> c_int;
    fn h2agg_wit_expose_final_pair(w: *mut h2agg_witness, w_x: i64, w_g: i64, out: *mut i64) -> c_int;
}

/// `Context` of the chips: the recorder.  Display prints the row offset like the circuit Context
/// (halo2-ecc-circuit-lib/src/gates/base_gate.rs:142-146).
pub struct B200Context {
    w: *mut h2agg_witness,
}
impl B200Context {
    pub fn new() -> Self {
        B200Context { w: unsafe { h2agg_wit_new() } }
    }
    pub fn raw(&self) -> *mut h2agg_witness {
        self.w
    }
    fn check(&self, h: i64) -> Result<i64, Error> {
        if h < 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(h2agg_wit_error(self.w)) }.to_string_lossy().into_owned();
            log::error!("h2agg witness recorder: {}", msg);
            return Err(Error::Synthesis);
        }
        Ok(h)
    }
}
impl B200Context {
    fn scalar(&self, h: i64) -> Result<AssignedScalar, Error> {
        let h = self.check(h)?;
        let mut out = [0u64; 4];
        assert_eq!(unsafe { h2agg_wit_scalar_value(self.w, h, out.as_mut_ptr()) }, 0);
        Ok(AssignedScalar { h, value: from_repr(out) })
    }
    fn point(&self, h: i64) -> Result<AssignedPoint, Error> {
        let h = self.check(h)?;
        let (mut out, mut ident) = ([0u64; 8], 0 as c_int);
        assert_eq!(unsafe { h2agg_wit_point_value(self.w, h, out.as_mut_ptr(), &mut ident) }, 0);
        let value = if ident != 0 {
            G1Affine::identity()
        } else {
            let (x, y): (Fq, Fq) = unsafe { std::mem::transmute::<[u64; 8], (Fq, Fq)>(out) };
            G1Affine::from_xy(x, y).unwrap()
        };
        Ok(AssignedPoint { h, value })
    }
}
impl Drop for B200Context {
    fn drop(&mut self) {
        unsafe { h2agg_wit_free(self.w) }
    }
}
impl fmt::Display for B200Context {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "(offset: {})", unsafe { h2agg_wit_rows(self.w) })
    }
}

/// AssignedValue<Fr> / AssignedPoint: a handle into the recorder plus the value, read back once when the handle is
/// made -- the traits' `to_value(&self, v)` has no context argument (common.rs:36), and the reference calls it
/// (verify.rs:727-728 for the pairing check, params.rs:65 for omega).
#[derive(Clone, Copy, Debug)]
pub struct AssignedScalar {
    pub h: i64,
    pub value: Fr,
}
#[derive(Clone, Copy, Debug)]
pub struct AssignedPoint {
    pub h: i64,
    pub value: G1Affine,
}

fn repr(s: &Fr) -> [u64; 4] {
    // canonical little-endian limbs (PrimeField::to_repr); the ABI takes canonical scalars
    let b = s.to_repr();
    let mut out = [0u64; 4];
    for i in 0..4 {
        out[i] = u64::from_le_bytes(b.as_ref()[8 * i..8 * i + 8].try_into().unwrap());
    }
    out
}
fn from_repr(l: [u64; 4]) -> Fr {
    let mut b = [0u8; 32];
    for i in 0..4 {
        b[8 * i..8 * i + 8].copy_from_slice(&l[i].to_le_bytes());
    }
    Fr::from_repr(b).unwrap()
}
fn xy(p: &G1Affine) -> [u64; 8] {
    // Montgomery limbs as they sit in memory (identity = zeros), the layout of include/h2agg.h
    if bool::from(p.is_identity()) {
        return [0; 8];
    }
    unsafe { std::mem::transmute::<G1Affine, [u64; 8]>(*p) }
}

// ------------------------------------------------------------------------------------------------ ScalarChip
pub struct B200ScalarChip;

impl ArithCommonChip for B200ScalarChip {
    type Context = B200Context;
    type Value = Fr;
    type AssignedValue = AssignedScalar;
    type Error = Error;

    fn add(&self, ctx: &mut B200Context, a: &AssignedScalar, b: &AssignedScalar) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_add(ctx.w, a.h, b.h) })
    }
    fn sub(&self, ctx: &mut B200Context, a: &AssignedScalar, b: &AssignedScalar) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_sub(ctx.w, a.h, b.h) })
    }
    fn assign_zero(&self, ctx: &mut B200Context) -> Result<AssignedScalar, Error> {
        self.assign_const(ctx, Fr::zero())
    }
    fn assign_one(&self, ctx: &mut B200Context) -> Result<AssignedScalar, Error> {
        self.assign_const(ctx, Fr::one())
    }
    fn assign_const(&self, ctx: &mut B200Context, c: Fr) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_assign_const(ctx.w, repr(&c).as_ptr()) })
    }
    fn assign_var(&self, ctx: &mut B200Context, v: Fr) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_assign_scalar(ctx.w, repr(&v).as_ptr()) })
    }
    fn to_value(&self, v: &AssignedScalar) -> Result<Fr, Error> {
        Ok(v.value)
    }
    fn normalize(&self, _ctx: &mut B200Context, v: &AssignedScalar) -> Result<AssignedScalar, Error> {
        Ok(*v)
    }
}

impl B200ScalarChip {
    /// (advice column, row): the halo2 `Cell` constrain_instance binds (verify_circuit.rs:357-367)
    pub fn cell(&self, ctx: &B200Context, v: &AssignedScalar) -> (u32, u32) {
        let (mut c, mut r) = (0u32, 0u32);
        assert_eq!(unsafe { h2agg_wit_scalar_cell(ctx.w, v.h, &mut c, &mut r) }, 0);
        (c, r)
    }
}

impl ArithFieldChip for B200ScalarChip {
    type Field = Fr;
    type AssignedField = AssignedScalar;

    fn mul(&self, ctx: &mut B200Context, a: &AssignedScalar, b: &AssignedScalar) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_mul(ctx.w, a.h, b.h) })
    }
    fn div(&self, ctx: &mut B200Context, a: &AssignedScalar, b: &AssignedScalar) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_div(ctx.w, a.h, b.h) })
    }
    fn square(&self, ctx: &mut B200Context, a: &AssignedScalar) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_square(ctx.w, a.h) })
    }
    fn sum_with_coeff_and_constant(&self, ctx: &mut B200Context, a_with_coeff: Vec<(&AssignedScalar, Fr)>, b: Fr) -> Result<AssignedScalar, Error> {
        let hs: Vec<i64> = a_with_coeff.iter().map(|(h, _)| h.h).collect();
        let cs: Vec<u64> = a_with_coeff.iter().flat_map(|(_, c)| repr(c)).collect();
        ctx.scalar(unsafe { h2agg_wit_field_sum_with_coeff_and_constant(ctx.w, hs.as_ptr(), cs.as_ptr(), hs.len(), repr(&b).as_ptr()) })
    }
    fn mul_add_constant(&self, ctx: &mut B200Context, a: &AssignedScalar, b: &AssignedScalar, c: Fr) -> Result<AssignedScalar, Error> {
        ctx.scalar(unsafe { h2agg_wit_field_mul_add_constant(ctx.w, a.h, b.h, repr(&c).as_ptr()) })
    }
    // sum_with_constant, mul_add, mul_add_accumulate, pow_constant: the trait's provided methods (field.rs:37-104)
}

// ------------------------------------------------------------------------------------------------ EccChip
pub struct B200EccChip;

impl ArithCommonChip for B200EccChip {
    type Context = B200Context;
    type Value = G1Affine;
    type AssignedValue = AssignedPoint;
    type Error = Error;

    fn add(&self, ctx: &mut B200Context, a: &AssignedPoint, b: &AssignedPoint) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_add(ctx.w, a.h, b.h) })
    }
    fn sub(&self, ctx: &mut B200Context, a: &AssignedPoint, b: &AssignedPoint) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_sub(ctx.w, a.h, b.h) })
    }
    fn assign_zero(&self, ctx: &mut B200Context) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_assign_identity(ctx.w) })
    }
    fn assign_one(&self, ctx: &mut B200Context) -> Result<AssignedPoint, Error> {
        self.assign_const(ctx, G1Affine::generator()) // assign_constant_point_from_scalar(1), chips/ecc_chip.rs:58-61
    }
    fn assign_const(&self, ctx: &mut B200Context, c: G1Affine) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_assign_constant_point(ctx.w, xy(&c).as_ptr()) })
    }
    fn assign_var(&self, ctx: &mut B200Context, v: G1Affine) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_assign_point(ctx.w, xy(&v).as_ptr()) })
    }
    fn to_value(&self, v: &AssignedPoint) -> Result<G1Affine, Error> {
        Ok(v.value)
    }
    fn normalize(&self, ctx: &mut B200Context, v: &AssignedPoint) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_reduce(ctx.w, v.h) })
    }
}

impl B200EccChip {
    /// what Halo2VerifierCircuits::synthesize does around the chips (verify_circuit.rs:264-368, 487-496)
    pub fn assert_equal(&self, ctx: &mut B200Context, a: &AssignedPoint, b: &AssignedPoint) -> Result<(), Error> {
        ctx.check(unsafe { h2agg_wit_ecc_assert_equal(ctx.w, a.h, b.h) } as i64).map(|_| ())
    }
    pub fn assert_not_identity(&self, ctx: &mut B200Context, p: &AssignedPoint) -> Result<(), Error> {
        ctx.check(unsafe { h2agg_wit_assert_not_identity(ctx.w, p.h) } as i64).map(|_| ())
    }
    pub fn expose_final_pair(&self, ctx: &mut B200Context, w_x: &AssignedPoint, w_g: &AssignedPoint) -> Result<[AssignedScalar; 4], Error> {
        let mut out = [0i64; 4];
        ctx.check(unsafe { h2agg_wit_expose_final_pair(ctx.w, w_x.h, w_g.h, out.as_mut_ptr()) } as i64)?;
        Ok([ctx.scalar(out[0])?, ctx.scalar(out[1])?, ctx.scalar(out[2])?, ctx.scalar(out[3])?])
    }
}

impl ArithEccChip for B200EccChip {
    type Point = G1Affine;
    type AssignedPoint = AssignedPoint;
    type Scalar = Fr;
    type AssignedScalar = AssignedScalar;
    type Native = Fr;
    type AssignedNative = AssignedScalar;
    type ScalarChip = B200ScalarChip;
    type NativeChip = B200ScalarChip;

    fn scalar_mul(&self, ctx: &mut B200Context, lhs: &AssignedScalar, rhs: &AssignedPoint) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_mul(ctx.w, rhs.h, lhs.h) })
    }
    fn scalar_mul_constant(&self, ctx: &mut B200Context, lhs: &AssignedScalar, rhs: G1Affine) -> Result<AssignedPoint, Error> {
        ctx.point(unsafe { h2agg_wit_ecc_constant_mul(ctx.w, xy(&rhs).as_ptr(), lhs.h) })
    }
    // multi_exp -> EccChipOps::shamir, as the circuit adapter overrides it (chips/ecc_chip.rs:125-132)
    fn multi_exp(&self, ctx: &mut B200Context, points: Vec<AssignedPoint>, scalars: Vec<AssignedScalar>) -> Result<AssignedPoint, Error> {
        let p: Vec<i64> = points.iter().map(|x| x.h).collect();
        let s: Vec<i64> = scalars.iter().map(|x| x.h).collect();
        ctx.point(unsafe { h2agg_wit_ecc_shamir(ctx.w, p.as_ptr(), s.as_ptr(), p.len()) })
    }
}

// ------------------------------------------------------------------------------------------------ Encode
pub struct B200PoseidonEncode;

impl Encode<B200EccChip> for B200PoseidonEncode {
    fn encode_point(ctx: &mut B200Context, _: &B200ScalarChip, _: &B200ScalarChip, _pchip: &B200EccChip, v: &AssignedPoint) -> Result<Vec<AssignedScalar>, Error> {
        let mut out = [0i64; 2];
        ctx.check(unsafe { h2agg_wit_encode_point(ctx.w, v.h, out.as_mut_ptr()) } as i64)?;
        Ok(vec![ctx.scalar(out[0])?, ctx.scalar(out[1])?])
    }
    fn encode_scalar(_: &mut B200Context, _: &B200ScalarChip, _: &B200ScalarChip, v: &AssignedScalar) -> Result<Vec<AssignedScalar>, Error> {
        Ok(vec![*v])
    }
    fn decode_scalar(_: &mut B200Context, _: &B200ScalarChip, _: &B200ScalarChip, v: &[AssignedScalar]) -> Result<AssignedScalar, Error> {
        Ok(v[0])
    }
}
