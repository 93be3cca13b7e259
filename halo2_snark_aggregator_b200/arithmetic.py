"""Mirror of halo2_proofs::arithmetic::{best_multiexp, best_fft} (boundary B1).

Same names, argument meaning and error behaviour as the Rust functions the reference's
create_proof reaches (halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994):
  best_multiexp(coeffs: &[Scalar], bases: &[G1Affine]) -> G1      asserts coeffs.len() == bases.len()
  best_fft(a: &mut [Fr], omega: Fr, log_n: u32)                   asserts a.len() == 1 << log_n, in place
Arrays are numpy uint64 in the Rust memory layout (4 limbs per Fr/Fq, Montgomery form).
"""
import numpy as np

from .context import default_context


def best_multiexp(coeffs, bases, ctx=None):
    """Returns the (normalised) Jacobian point as 12 uint64 limbs."""
    assert coeffs.dtype == np.uint64 and bases.dtype == np.uint64
    assert coeffs.size // 4 == bases.size // 8, "best_multiexp: coeffs.len() != bases.len()"
    ctx = ctx or default_context()
    return ctx.msm_g1(coeffs, bases)


def best_fft(a, omega, log_n, ctx=None):
    """In-place NTT, natural order in and out."""
    assert a.dtype == np.uint64 and a.size == 4 << log_n, "best_fft: a.len() != 1 << log_n"
    ctx = ctx or default_context()
    ctx.ntt_fr(a, np.ascontiguousarray(omega, dtype=np.uint64), log_n)


def eval_polynomial(poly, point, ctx=None):
    """halo2_proofs::arithmetic::eval_polynomial(poly: &[F], point: F) -> F"""
    ctx = ctx or default_context()
    return ctx.eval_polynomial(poly, np.ascontiguousarray(point, dtype=np.uint64))


def kate_division(a, b, ctx=None):
    """halo2_proofs::arithmetic::kate_division(a, b) -> Vec<F> of len a.len() - 1"""
    ctx = ctx or default_context()
    return ctx.kate_division(a, np.ascontiguousarray(b, dtype=np.uint64))
