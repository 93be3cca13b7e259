"""Context = one GPU, one stream, resident SRS and twiddle tables (h2agg_ctx)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import H2aggError, c_vp


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return c_vp(a)
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need contiguous uint64 ndarray"
    return c_vp(a.ctypes.data)


class Context:
    """Owns an h2agg_ctx. Raises H2aggError if the library or a B200 is missing (no CPU fallback)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = c_vp()
        rc = self.lib.h2agg_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise H2aggError("h2agg_init failed (%d): %s" % (rc, self.lib.h2agg_last_error(None).decode()))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.h2agg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise H2aggError("h2agg error %d: %s" % (rc, self.lib.h2agg_last_error(self.h).decode()))

    # -- plumbing
    def set_stream(self, cuda_stream_ptr):
        self.check(self.lib.h2agg_set_stream(self.h, c_vp(cuda_stream_ptr)))

    def synchronize(self):
        self.check(self.lib.h2agg_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.h2agg_launch_count(self.h))

    KERNEL_CLASSES = ("msm_accumulate", "msm_digits_sort", "msm_reduce", "ntt_pass", "msm_total", "witness_expand", "evaluate_h")

    def kernel_timing(self, enable):
        self.check(self.lib.h2agg_kernel_timing(self.h, 1 if enable else 0))

    def kernel_times(self):
        """{class: (total_ms, launches)} since the last call (synchronises)."""
        nc = len(self.KERNEL_CLASSES)
        ms = (ctypes.c_double * nc)()
        cnt = (ctypes.c_uint64 * nc)()
        self.check(self.lib.h2agg_kernel_times(self.h, ms, cnt, nc))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def host_register(self, arr):
        self.check(self.lib.h2agg_host_register(self.h, _ptr(arr), arr.nbytes))

    def host_unregister(self, arr):
        self.check(self.lib.h2agg_host_unregister(self.h, _ptr(arr)))

    def set_msm_window(self, c):
        self.check(self.lib.h2agg_set_msm_window(self.h, int(c)))

    def set_ntt_radix_cap(self, log2_radix):
        self.check(self.lib.h2agg_set_ntt_radix_cap(self.h, int(log2_radix)))

    def set_srs_precompute(self, enable):
        self.check(self.lib.h2agg_set_srs_precompute(self.h, 1 if enable else 0))

    def srs_config(self, sid):
        """(table_mode, c, n_windows) for MSMs against a registered SRS."""
        t, c, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.check(self.lib.h2agg_srs_config(self.h, sid, ctypes.byref(t), ctypes.byref(c), ctypes.byref(w)))
        return bool(t.value), c.value, w.value

    def msm_config(self, n):
        c, w = ctypes.c_int(), ctypes.c_int()
        self.check(self.lib.h2agg_msm_config(self.h, n, ctypes.byref(c), ctypes.byref(w)))
        return c.value, w.value

    def dev_alloc(self, nbytes):
        p = c_vp()
        self.check(self.lib.h2agg_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, p):
        self.check(self.lib.h2agg_dev_free(self.h, c_vp(p)))

    def h2d(self, d_dst, arr):
        self.check(self.lib.h2agg_memcpy_h2d(self.h, c_vp(d_dst), _ptr(arr), arr.nbytes))

    def d2h(self, d_src, n_u64):
        out = np.empty(n_u64, dtype=np.uint64)
        self.check(self.lib.h2agg_memcpy_d2h(self.h, _ptr(out), c_vp(d_src), out.nbytes))
        return out

    def synth_scalars_dev(self, seed, kind, first, n, d_out):
        self.check(self.lib.h2agg_synth_scalars_dev(self.h, seed, kind, first, n, c_vp(d_out)))

    def synth_bases_dev(self, seed, first, n, d_out):
        self.check(self.lib.h2agg_synth_bases_dev(self.h, seed, first, n, c_vp(d_out)))

    # -- SRS
    def srs_register(self, bases):
        sid = ctypes.c_uint64()
        self.check(self.lib.h2agg_srs_register(self.h, _ptr(bases), bases.size // 8, ctypes.byref(sid)))
        return sid.value

    def srs_register_dev(self, d_bases, n):
        sid = ctypes.c_uint64()
        self.check(self.lib.h2agg_srs_register_dev(self.h, c_vp(d_bases), n, ctypes.byref(sid)))
        return sid.value

    def srs_release(self, sid):
        self.check(self.lib.h2agg_srs_release(self.h, sid))

    # -- K1
    def msm_g1(self, scalars, bases=None, srs_id=0, n=None, windows=None):
        n = scalars.size // 4 if n is None else n
        out = np.zeros(12, dtype=np.uint64)
        if windows is None:
            self.check(self.lib.h2agg_msm_g1(self.h, srs_id, _ptr(bases), _ptr(scalars), n, _ptr(out)))
        else:
            self.check(self.lib.h2agg_msm_g1_windows(self.h, srs_id, _ptr(bases), _ptr(scalars), n, windows[0], windows[1], _ptr(out)))
        return out

    def msm_g1_dev(self, d_scalars, n, d_out160, d_bases=0, srs_id=0, windows=None):
        if windows is None:
            self.check(self.lib.h2agg_msm_g1_dev(self.h, srs_id, c_vp(d_bases), c_vp(d_scalars), n, c_vp(d_out160)))
        else:
            self.check(self.lib.h2agg_msm_g1_windows_dev(self.h, srs_id, c_vp(d_bases), c_vp(d_scalars), n, windows[0], windows[1], c_vp(d_out160)))

    def msm_g1_batch(self, srs_id, cols, n=None):
        n = cols[0].size // 4 if n is None else n
        arr = (c_vp * len(cols))(*[c.ctypes.data if isinstance(c, np.ndarray) else c for c in cols])
        out = np.zeros(8 * len(cols), dtype=np.uint64)
        self.check(self.lib.h2agg_msm_g1_batch(self.h, srs_id, arr, len(cols), n, _ptr(out)))
        return out.reshape(len(cols), 8)

    def msm_g1_batch_dev(self, d_cols, n, d_out160s, d_bases=0, srs_id=0, windows=None):
        """windows: None (all), one (begin, end) for the batch, or a list with one (begin, end) / None per column."""
        arr = (c_vp * len(d_cols))(*d_cols)
        if windows is None:
            self.check(self.lib.h2agg_msm_g1_batch_dev(self.h, srs_id, c_vp(d_bases), arr, len(d_cols), n, c_vp(d_out160s)))
        elif isinstance(windows, list):
            assert len(windows) == len(d_cols)
            lo = (ctypes.c_int * len(d_cols))(*[0 if w is None else w[0] for w in windows])
            hi = (ctypes.c_int * len(d_cols))(*[-1 if w is None else w[1] for w in windows])
            self.check(self.lib.h2agg_msm_g1_batch_ranges_dev(self.h, srs_id, c_vp(d_bases), arr, len(d_cols), n, lo, hi, c_vp(d_out160s)))
        else:
            self.check(self.lib.h2agg_msm_g1_batch_windows_dev(self.h, srs_id, c_vp(d_bases), arr, len(d_cols), n, windows[0], windows[1], c_vp(d_out160s)))

    def g1_sum_dev(self, d_points, m, stride_bytes, n_out, d_out160s):
        """out[j] = sum_i Jacobian point at d_points + i*stride + j*160 (device; combines gathered window shards)."""
        self.check(self.lib.h2agg_g1_sum_dev(self.h, c_vp(d_points), m, stride_bytes, n_out, c_vp(d_out160s)))

    def g1_sum(self, points_jac):
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.h2agg_g1_sum(self.h, _ptr(points_jac), points_jac.size // 12, _ptr(out)))
        return out

    # -- K2 / K3 (host pointers, in place like the reference)
    def ntt_fr(self, a, omega, log_n):
        self.check(self.lib.h2agg_ntt_fr(self.h, _ptr(a), _ptr(omega), log_n))

    def intt_fr(self, a, omega_inv, n_inv, log_n):
        self.check(self.lib.h2agg_intt_fr(self.h, _ptr(a), _ptr(omega_inv), _ptr(n_inv), log_n))

    def intt_fr_batch(self, cols, omega_inv, n_inv, log_n):
        arr = (c_vp * len(cols))(*[c.ctypes.data for c in cols])
        self.check(self.lib.h2agg_intt_fr_batch(self.h, arr, len(cols), _ptr(omega_inv), _ptr(n_inv), log_n))

    def coeff_to_extended_batch(self, coeff_cols, out_cols, k, ext_k, zeta, omega_ext):
        a = (c_vp * len(coeff_cols))(*[c.ctypes.data for c in coeff_cols])
        b = (c_vp * len(out_cols))(*[c.ctypes.data for c in out_cols])
        self.check(self.lib.h2agg_coeff_to_extended_batch(self.h, a, b, len(coeff_cols), k, ext_k, _ptr(zeta), _ptr(omega_ext)))

    def commit_round(self, srs_id, cols, k, omega_inv, n_inv, coeff_out=None, ext_k=0, zeta=None, omega_ext=None, ext_out=None):
        """Fused commit round -> affine commitments (len(cols), 8); coefficients / extended evaluations land in the given arrays."""
        nc = len(cols)
        a = (c_vp * nc)(*[c.ctypes.data for c in cols])
        co = (c_vp * nc)(*[c.ctypes.data for c in coeff_out]) if coeff_out is not None else None
        eo = (c_vp * nc)(*[(c.ctypes.data if c is not None else None) for c in ext_out]) if ext_out is not None else None
        out = np.zeros(8 * nc, dtype=np.uint64)
        self.check(self.lib.h2agg_commit_round(self.h, srs_id, a, nc, k, _ptr(omega_inv), _ptr(n_inv), _ptr(out), co, ext_k,
                                               _ptr(zeta), _ptr(omega_ext), eo))
        return out.reshape(nc, 8)

    def set_defer_transforms(self, enable):
        """commit_round_resident / _dev: NTT passes on the background stream; see include/h2agg.h"""
        self.check(self.lib.h2agg_set_defer_transforms(self.h, 1 if enable else 0))

    def transforms_join(self):
        self.check(self.lib.h2agg_transforms_join(self.h))

    def transforms_dev(self, d_cols, k, omega_inv, n_inv, d_coeff_out, ext_k=0, zeta=None, omega_ext=None, d_ext_out=None):
        nc = len(d_cols)
        a = (c_vp * nc)(*d_cols)
        co = (c_vp * nc)(*d_coeff_out)
        eo = (c_vp * nc)(*d_ext_out) if d_ext_out is not None else None
        self.check(self.lib.h2agg_transforms_dev(self.h, a, nc, k, _ptr(omega_inv), _ptr(n_inv), co, ext_k, _ptr(zeta), _ptr(omega_ext), eo))

    def commit_round_resident(self, srs_id, cols, k, omega_inv, n_inv, d_coeff_out, ext_k=0, zeta=None, omega_ext=None, d_ext_out=None,
                              d_lagrange_out=None):
        """Commit round whose coefficient / extended (/ Lagrange) forms stay in HBM (device pointers) -> affine commitments (len(cols), 8)."""
        nc = len(cols)
        a = (c_vp * nc)(*[c.ctypes.data for c in cols])
        co = (c_vp * nc)(*d_coeff_out)
        eo = (c_vp * nc)(*d_ext_out) if d_ext_out is not None else None
        lo = (c_vp * nc)(*d_lagrange_out) if d_lagrange_out is not None else None
        out = np.zeros(8 * nc, dtype=np.uint64)
        self.check(self.lib.h2agg_commit_round_resident(self.h, srs_id, a, nc, k, _ptr(omega_inv), _ptr(n_inv), _ptr(out), lo, co, ext_k,
                                                        _ptr(zeta), _ptr(omega_ext), eo))
        return out.reshape(nc, 8)

    def commit_round_dev(self, srs_id, d_cols, k, omega_inv, n_inv, d_coeff_out, ext_k=0, zeta=None, omega_ext=None, d_ext_out=None):
        """The same round for Lagrange columns already in HBM."""
        nc = len(d_cols)
        a = (c_vp * nc)(*d_cols)
        co = (c_vp * nc)(*d_coeff_out)
        eo = (c_vp * nc)(*d_ext_out) if d_ext_out is not None else None
        out = np.zeros(8 * nc, dtype=np.uint64)
        self.check(self.lib.h2agg_commit_round_dev(self.h, srs_id, a, nc, k, _ptr(omega_inv), _ptr(n_inv), _ptr(out), co, ext_k,
                                                   _ptr(zeta), _ptr(omega_ext), eo))
        return out.reshape(nc, 8)

    # -- N3: lookup / permutation arguments on resident Lagrange columns
    def compress_expressions_dev(self, words, consts, d_columns, k, theta, d_out):
        """words: uint32 array { n_exprs, POLY x n_exprs }; consts: Montgomery limbs (n_consts*4) or empty."""
        cols = (c_vp * len(d_columns))(*d_columns)
        self.check(self.lib.h2agg_compress_expressions_dev(self.h, words.ctypes.data, words.size, cols, len(d_columns),
                                                           consts.ctypes.data if consts.size else None, consts.size // 4, k,
                                                           _ptr(theta), c_vp(d_out)))

    def lookup_product_dev(self, d_a, d_s, d_ap, d_sp, n, beta, gamma, d_z):
        self.check(self.lib.h2agg_lookup_product_dev(self.h, c_vp(d_a), c_vp(d_s), c_vp(d_ap), c_vp(d_sp), n, _ptr(beta), _ptr(gamma), c_vp(d_z)))

    def lookup_products_dev(self, d_a, d_s, d_ap, d_sp, n, beta, gamma, d_z):
        """all lookups at once (lists of device pointers): lookup i runs on lane i mod 8"""
        m = len(d_z)
        arr = lambda v: (c_vp * m)(*v)
        self.check(self.lib.h2agg_lookup_products_dev(self.h, m, arr(d_a), arr(d_s), arr(d_ap), arr(d_sp), n, _ptr(beta), _ptr(gamma), arr(d_z)))

    def permutation_product_dev(self, d_values, d_sigmas, k, omega, beta_delta_start, delta, beta, gamma, d_last_z, d_z):
        v = (c_vp * len(d_values))(*d_values)
        s = (c_vp * len(d_sigmas))(*d_sigmas)
        self.check(self.lib.h2agg_permutation_product_dev(self.h, v, s, len(d_values), k, _ptr(omega), _ptr(beta_delta_start), _ptr(delta),
                                                          _ptr(beta), _ptr(gamma), c_vp(d_last_z) if d_last_z else None, c_vp(d_z)))

    def ntt_fr_dev(self, d_a, omega, log_n, scale=None):
        self.check(self.lib.h2agg_ntt_fr_dev(self.h, c_vp(d_a), _ptr(omega), _ptr(scale), log_n))

    def coeff_to_extended(self, coeffs, k, ext_k, zeta, omega_ext):
        out = np.empty(4 << ext_k, dtype=np.uint64)
        self.check(self.lib.h2agg_coeff_to_extended(self.h, _ptr(coeffs), k, ext_k, _ptr(zeta), _ptr(omega_ext), _ptr(out)))
        return out

    def extended_to_coeff(self, a, ext_k, omega_ext_inv, ext_n_inv, zeta, out_len):
        self.check(self.lib.h2agg_extended_to_coeff(self.h, _ptr(a), ext_k, _ptr(omega_ext_inv), _ptr(ext_n_inv), _ptr(zeta), out_len))
        return a[: 4 * out_len]

    def coeff_to_extended_dev(self, d_coeffs, k, ext_k, zeta, omega_ext, d_out):
        self.check(self.lib.h2agg_coeff_to_extended_dev(self.h, c_vp(d_coeffs), k, ext_k, _ptr(zeta), _ptr(omega_ext), c_vp(d_out)))

    def extended_to_coeff_dev(self, d_a, ext_k, omega_ext_inv, ext_n_inv, zeta, out_len):
        self.check(self.lib.h2agg_extended_to_coeff_dev(self.h, c_vp(d_a), ext_k, _ptr(omega_ext_inv), _ptr(ext_n_inv), _ptr(zeta), out_len))

    # -- N2: evaluation / Kate division
    def eval_polynomial(self, poly, point):
        out = np.zeros(4, dtype=np.uint64)
        self.check(self.lib.h2agg_eval_polynomial(self.h, _ptr(poly), poly.size // 4, _ptr(point), _ptr(out)))
        return out

    def eval_polynomial_dev(self, d_poly, n, point, d_out32):
        self.check(self.lib.h2agg_eval_polynomial_dev(self.h, c_vp(d_poly), n, _ptr(point), c_vp(d_out32)))

    def eval_polynomials_dev(self, d_polys, n, point, d_out):
        """d_out[i] = polys[i](point): many polynomials of n coefficients at one point"""
        arr = (c_vp * len(d_polys))(*d_polys)
        self.check(self.lib.h2agg_eval_polynomials_dev(self.h, arr, len(d_polys), n, _ptr(point), c_vp(d_out)))

    def kate_division(self, a, b):
        n = a.size // 4
        q = np.zeros(4 * max(n - 1, 0), dtype=np.uint64)
        self.check(self.lib.h2agg_kate_division(self.h, _ptr(a), n, _ptr(b), _ptr(q) if q.size else c_vp(a.ctypes.data)))
        return q

    def kate_division_dev(self, d_a, n, b, d_q):
        self.check(self.lib.h2agg_kate_division_dev(self.h, c_vp(d_a), n, _ptr(b), c_vp(d_q)))

    # -- N1: quotient numerator on the extended coset (evaluate_h [+ divide_by_vanishing_poly])
    def evaluate_h_dev(self, plan, d_columns, k, ext_k, y, beta, gamma, theta, d_out, divide=True, rows=None, col_row0=None):
        """plan: plonk.QuotientPlan; d_columns: device pointers in plan.columns order; challenges: 4-limb arrays.
        rows = (row_begin, row_count) evaluates a window only; col_row0[c] = global row of element 0 of column c's buffer."""
        from . import plonk

        assert len(d_columns) == len(plan.columns), "one device column per plan.columns entry"
        cols = (c_vp * len(d_columns))(*d_columns)
        w_ext = plonk.fr_mont(pow(plonk.ROOT_OF_UNITY, 1 << (28 - ext_k), plonk.R_MOD))
        zeta, delta = plonk.fr_mont(plonk.ZETA), plonk.fr_mont(plonk.DELTA)
        keep = [cols, w_ext, zeta, delta, plan.words, plan.consts]
        a = _lib.QuotientArgs()
        a.k, a.ext_k = k, ext_k
        a.plan, a.n_plan_words = plan.words.ctypes.data, plan.words.size
        a.d_columns, a.n_columns = ctypes.cast(cols, c_vp).value, len(d_columns)
        a.consts, a.n_consts = (plan.consts.ctypes.data if plan.consts.size else None), plan.consts.size // 4
        a.y, a.beta, a.gamma, a.theta = y.ctypes.data, beta.ctypes.data, gamma.ctypes.data, theta.ctypes.data
        a.omega_ext, a.zeta, a.delta = w_ext.ctypes.data, zeta.ctypes.data, delta.ctypes.data
        if divide:
            t = np.concatenate([plonk.fr_mont(v) for v in plonk.t_evaluations(k, ext_k)])
            keep.append(t)
            a.t_evaluations, a.t_len = t.ctypes.data, t.size // 4
        else:
            a.t_evaluations, a.t_len = None, 0
        if rows is None:
            self.check(self.lib.h2agg_evaluate_h_dev(self.h, ctypes.byref(a), c_vp(d_out)))
        else:
            r0 = np.ascontiguousarray(col_row0, dtype=np.uint64) if col_row0 is not None else None
            self.check(self.lib.h2agg_evaluate_h_rows_dev(self.h, ctypes.byref(a), int(rows[0]), int(rows[1]),
                                                          _ptr(r0), c_vp(d_out)))
        del keep

    def poly_fold_dev(self, d_polys, n, v, d_out):
        """out[j] = sum_i polys[i][j] * v^(m-1-i) (GWC: poly_batch = poly_batch * v + poly)"""
        arr = (c_vp * len(d_polys))(*d_polys)
        self.check(self.lib.h2agg_poly_fold_dev(self.h, arr, len(d_polys), n, _ptr(v), c_vp(d_out)))

    def poly_lincomb_dev(self, d_polys, weights, n, d_out):
        """out[j] = sum_i weights[i] * polys[i][j]; weights: (m, 4) or flat Montgomery limbs"""
        m = len(d_polys)
        arr = (c_vp * max(m, 1))(*d_polys)
        w = np.ascontiguousarray(weights, dtype=np.uint64).reshape(-1) if m else np.zeros(4, dtype=np.uint64)
        self.check(self.lib.h2agg_poly_lincomb_dev(self.h, arr, _ptr(w), m, n, c_vp(d_out)))

    # -- N3: grand-product scans
    def batch_invert(self, a):
        out = a.copy()
        self.check(self.lib.h2agg_batch_invert(self.h, _ptr(out), out.size // 4))
        return out

    def grand_product(self, num, den):
        z = np.zeros_like(num)
        self.check(self.lib.h2agg_grand_product(self.h, _ptr(num), _ptr(den), num.size // 4, _ptr(z)))
        return z

    def grand_product_dev(self, d_num, d_den, n, d_z):
        self.check(self.lib.h2agg_grand_product_dev(self.h, c_vp(d_num), c_vp(d_den), n, c_vp(d_z)))

    def sort_fr(self, a):
        out = a.copy()
        self.check(self.lib.h2agg_sort_fr(self.h, _ptr(out), out.size // 4))
        return out

    def sort_fr_dev(self, d_a, n):
        self.check(self.lib.h2agg_sort_fr_dev(self.h, c_vp(d_a), n))

    def permute_expression_pair(self, inp, tab):
        """lookup::prover::permute_expression_pair over the usable rows -> (permuted_input, permuted_table);
        raises H2aggError (status 4) when an input value is not in the table."""
        a, s = np.zeros_like(inp), np.zeros_like(tab)
        self.check(self.lib.h2agg_permute_expression_pair(self.h, _ptr(inp), _ptr(tab), inp.size // 4, _ptr(a), _ptr(s)))
        return a, s

    def permute_expression_pair_dev(self, d_inp, d_tab, u, d_pin, d_ptab):
        self.check(self.lib.h2agg_permute_expression_pair_dev(self.h, c_vp(d_inp), c_vp(d_tab), u, c_vp(d_pin), c_vp(d_ptab)))

    # -- N4: PrimeField::to_repr / from_repr in bulk
    def fr_to_repr(self, limbs):
        """Montgomery limbs (n*4 uint64) -> n*32 bytes, little-endian canonical"""
        out = np.empty_like(limbs)
        self.check(self.lib.h2agg_fr_repr(self.h, 0, _ptr(limbs), _ptr(out), limbs.size // 4))
        return out.tobytes()

    def fr_from_repr(self, data):
        """n*32 bytes -> Montgomery limbs; raises H2aggError (status 4) on a non-canonical value"""
        a = np.frombuffer(data, dtype=np.uint64).copy()
        out = np.empty_like(a)
        self.check(self.lib.h2agg_fr_repr(self.h, 1, _ptr(a), _ptr(out), a.size // 4))
        return out

    # -- N4: G1Affine::to_bytes / from_bytes in bulk (the point codec of ParamsKZG files and Poseidon transcripts)
    def g1_decompress(self, data):
        """n*32 bytes -> affine Montgomery limbs (n*8 uint64); raises H2aggError (status 4) on an invalid encoding"""
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        n = buf.size // 32
        out = np.empty(8 * n, dtype=np.uint64)
        self.check(self.lib.h2agg_g1_decompress(self.h, c_vp(buf.ctypes.data), _ptr(out), n))
        return out

    def g1_decompress_dev(self, d_in, d_out, n):
        self.check(self.lib.h2agg_g1_decompress_dev(self.h, c_vp(d_in), c_vp(d_out), n))

    def g1_compress(self, affine):
        n = affine.size // 8
        out = np.empty(32 * n, dtype=np.uint8)
        self.check(self.lib.h2agg_g1_compress(self.h, _ptr(affine), c_vp(out.ctypes.data), n))
        return out.tobytes()

    # -- field helpers (device)
    def field_op(self, field, op, a, b=None):
        out = np.empty_like(a)
        self.check(self.lib.h2agg_field_op(self.h, field, op, _ptr(a), _ptr(b), _ptr(out), a.size // 4))
        return out


_default = None


def default_context():
    global _default
    if _default is None:
        _default = Context(0)
    return _default
