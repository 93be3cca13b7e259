"""ctypes binding of the CPU oracle (oracle/cpu_halo2.cpp).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle_h2.so")
sys.path.insert(0, os.path.join(ORACLE_DIR, "py"))

c_vp = ctypes.c_void_p


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "cpu_halo2.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liboracle_h2.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_o = None


def oracle():
    global _o
    if _o is None:
        lib = ctypes.CDLL(build_oracle())
        try:  # a .so built on another CPU generation may not run here: rebuild once
            lib.oracle_hw_threads()
        except Exception:
            lib = ctypes.CDLL(build_oracle(force=True))
        u64, sz, u32, ui, ci = ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint, ctypes.c_int
        lib.oracle_gen_scalars.argtypes = [u64, ci, sz, sz, c_vp, ui]
        lib.oracle_gen_bases.argtypes = [u64, sz, sz, c_vp, ui]
        lib.oracle_field_op.argtypes = [ci, ci, c_vp, c_vp, c_vp, sz]
        lib.oracle_to_mont.argtypes = [ci, c_vp, c_vp, sz]
        lib.oracle_from_mont.argtypes = [ci, c_vp, c_vp, sz]
        lib.oracle_best_multiexp.argtypes = [c_vp, c_vp, sz, ui, c_vp]
        lib.oracle_msm_naive.argtypes = [c_vp, c_vp, sz, c_vp]
        lib.oracle_g1_sum.argtypes = [c_vp, sz, c_vp]
        lib.oracle_g1_on_curve.argtypes = [c_vp]
        lib.oracle_g1_on_curve.restype = ci
        lib.oracle_best_fft.argtypes = [c_vp, c_vp, u32, ui]
        lib.oracle_dft_naive.argtypes = [c_vp, c_vp, u32, c_vp]
        lib.oracle_ifft.argtypes = [c_vp, c_vp, c_vp, u32, ui]
        lib.oracle_coeff_to_extended.argtypes = [c_vp, u32, u32, c_vp, c_vp, c_vp, ui]
        lib.oracle_extended_to_coeff.argtypes = [c_vp, u32, c_vp, c_vp, c_vp, ui]
        lib.oracle_eval_polynomial.argtypes = [c_vp, sz, c_vp, c_vp]
        lib.oracle_kate_division.argtypes = [c_vp, sz, c_vp, c_vp]
        lib.oracle_batch_invert.argtypes = [c_vp, sz]
        lib.oracle_grand_product.argtypes = [c_vp, c_vp, sz, c_vp]
        lib.oracle_evaluate_h.argtypes = [c_vp, sz, ctypes.POINTER(c_vp), c_vp, u32, u32] + [c_vp] * 8 + [sz, c_vp, ui]
        lib.oracle_permute_expression_pair.argtypes = [c_vp, c_vp, sz, c_vp, c_vp]
        lib.oracle_permute_expression_pair.restype = ci
        lib.oracle_compress_expressions.argtypes = [c_vp, ctypes.POINTER(c_vp), c_vp, u32, c_vp, c_vp, ui]
        lib.oracle_lookup_product.argtypes = [c_vp] * 4 + [sz, c_vp, c_vp, c_vp]
        lib.oracle_permutation_product.argtypes = [ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), sz, u32] + [c_vp] * 7
        lib.oracle_hw_threads.restype = ui
        _o = lib
    return _o


def P(a):
    return c_vp(a.ctypes.data) if a is not None else None


def threads():
    return int(oracle().oracle_hw_threads())


def gen_scalars(seed, kind, n, first=0):
    out = np.empty(4 * n, dtype=np.uint64)
    oracle().oracle_gen_scalars(seed, kind, first, n, P(out), threads())
    return out


def gen_bases(seed, n, first=0):
    out = np.empty(8 * n, dtype=np.uint64)
    oracle().oracle_gen_bases(seed, first, n, P(out), threads())
    return out


def field_op(field, op, a, b=None):
    out = np.empty_like(a)
    oracle().oracle_field_op(field, op, P(a), P(b), P(out), a.size // 4)
    return out


def to_mont(field, canon):
    out = np.empty_like(canon)
    oracle().oracle_to_mont(field, P(canon), P(out), canon.size // 4)
    return out


def from_mont(field, mont):
    out = np.empty_like(mont)
    oracle().oracle_from_mont(field, P(mont), P(out), mont.size // 4)
    return out


def best_multiexp(scalars, bases, nthreads=None):
    out = np.zeros(12, dtype=np.uint64)
    oracle().oracle_best_multiexp(P(scalars), P(bases), scalars.size // 4, nthreads or threads(), P(out))
    return out


def msm_naive(scalars, bases):
    out = np.zeros(12, dtype=np.uint64)
    oracle().oracle_msm_naive(P(scalars), P(bases), scalars.size // 4, P(out))
    return out


def g1_sum(points12):
    out = np.zeros(12, dtype=np.uint64)
    oracle().oracle_g1_sum(P(points12), points12.size // 12, P(out))
    return out


def on_curve(aff8):
    return bool(oracle().oracle_g1_on_curve(P(np.ascontiguousarray(aff8))))


def best_fft(a, omega, log_n, nthreads=None):
    oracle().oracle_best_fft(P(a), P(omega), log_n, nthreads or threads())
    return a


def dft_naive(a, omega, log_n):
    out = np.empty_like(a)
    oracle().oracle_dft_naive(P(a), P(omega), log_n, P(out))
    return out


def ifft(a, omega_inv, divisor, log_n, nthreads=None):
    oracle().oracle_ifft(P(a), P(omega_inv), P(divisor), log_n, nthreads or threads())
    return a


def coeff_to_extended(coeffs, k, ext_k, zeta, omega_ext, nthreads=None):
    out = np.empty(4 << ext_k, dtype=np.uint64)
    oracle().oracle_coeff_to_extended(P(coeffs), k, ext_k, P(zeta), P(omega_ext), P(out), nthreads or threads())
    return out


def extended_to_coeff(a, ext_k, omega_ext_inv, ext_n_inv, zeta, out_len, nthreads=None):
    oracle().oracle_extended_to_coeff(P(a), ext_k, P(omega_ext_inv), P(ext_n_inv), P(zeta), nthreads or threads())
    return a[: 4 * out_len]


def eval_polynomial(poly, point):
    out = np.zeros(4, dtype=np.uint64)
    oracle().oracle_eval_polynomial(P(poly), poly.size // 4, P(point), P(out))
    return out


def kate_division(a, b):
    n = a.size // 4
    q = np.zeros(4 * max(n - 1, 0), dtype=np.uint64)
    if n >= 2:
        oracle().oracle_kate_division(P(a), n, P(b), P(q))
    return q


def batch_invert(a):
    out = a.copy()
    oracle().oracle_batch_invert(P(out), out.size // 4)
    return out


def grand_product(num, den):
    z = np.zeros_like(num)
    oracle().oracle_grand_product(P(num), P(den), num.size // 4, P(z))
    return z


def permute_expression_pair(inp, tab):
    """-> (status, permuted_input, permuted_table); status 1 = an input value is missing from the table"""
    a, b = np.zeros_like(inp), np.zeros_like(tab)
    rc = oracle().oracle_permute_expression_pair(P(inp), P(tab), inp.size // 4, P(a), P(b))
    return rc, a, b


def compress_expressions(words, consts, cols, k, theta, nthreads=None):
    out = np.empty(4 << k, dtype=np.uint64)
    arr = (c_vp * len(cols))(*[c.ctypes.data for c in cols])
    oracle().oracle_compress_expressions(P(words), arr, P(consts) if consts.size else None, k, P(theta), P(out), nthreads or threads())
    return out


def lookup_product(A, S, Ap, Sp, beta, gamma):
    z = np.empty_like(A)
    oracle().oracle_lookup_product(P(A), P(S), P(Ap), P(Sp), A.size // 4, P(beta), P(gamma), P(z))
    return z


def permutation_product(values, sigmas, k, omega, beta_delta_start, delta, beta, gamma, last_z=None):
    z = np.empty(4 << k, dtype=np.uint64)
    v = (c_vp * len(values))(*[c.ctypes.data for c in values])
    s = (c_vp * len(sigmas))(*[c.ctypes.data for c in sigmas])
    oracle().oracle_permutation_product(v, s, len(values), k, P(omega), P(beta_delta_start), P(delta), P(beta), P(gamma),
                                        P(last_z) if last_z is not None else None, P(z))
    return z


def evaluate_h(plan_words, consts, ext_cols, k, ext_k, y, beta, gamma, theta, omega_ext, zeta, delta, t_evals=None,
               nthreads=None):
    """plan-driven C++ restatement of evaluate_h (+ division by the vanishing polynomial when t_evals is given)"""
    out = np.empty(4 << ext_k, dtype=np.uint64)
    cols = (c_vp * len(ext_cols))(*[c.ctypes.data for c in ext_cols])
    oracle().oracle_evaluate_h(P(plan_words), plan_words.size, cols, P(consts) if consts.size else None, k, ext_k, P(y),
                               P(beta), P(gamma), P(theta), P(omega_ext), P(zeta), P(delta), P(t_evals),
                               (t_evals.size // 4) if t_evals is not None else 0, P(out), nthreads or threads())
    return out
