"""ctypes loader for libh2agg.so (the C ABI declared in include/h2agg.h).

The product path has no CPU fallback: if the shared library is missing, or no B200 is present,
loading / context creation raises.  PyTorch is not needed to use the library; bench.py and the
multi-GPU host layer use it only for pinned buffers, events and torch.distributed.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libh2agg.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "h2agg.h")

c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_vp = ctypes.c_void_p


class H2aggError(RuntimeError):
    pass


class QuotientArgs(ctypes.Structure):
    """h2agg_quotient_args (include/h2agg.h)."""
    _fields_ = [("k", ctypes.c_uint32), ("ext_k", ctypes.c_uint32),
                ("plan", c_vp), ("n_plan_words", ctypes.c_size_t),
                ("d_columns", c_vp), ("n_columns", ctypes.c_size_t),
                ("consts", c_vp), ("n_consts", ctypes.c_size_t),
                ("y", c_vp), ("beta", c_vp), ("gamma", c_vp), ("theta", c_vp),
                ("omega_ext", c_vp), ("zeta", c_vp), ("delta", c_vp),
                ("t_evaluations", c_vp), ("t_len", ctypes.c_size_t)]


_lib = None


def declared_symbols():
    """Every function name include/h2agg.h declares (used by the CPU-side export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2agg_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise H2aggError(
            "libh2agg.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the product path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    sz, u32, u64, ci, i64 = ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64
    sig = {
        "h2agg_init": (ci, [ci, ctypes.POINTER(c_vp)]),
        "h2agg_destroy": (None, [c_vp]),
        "h2agg_last_error": (ctypes.c_char_p, [c_vp]),
        "h2agg_version": (ctypes.c_char_p, []),
        "h2agg_set_stream": (ci, [c_vp, c_vp]),
        "h2agg_synchronize": (ci, [c_vp]),
        "h2agg_launch_count": (u64, [c_vp]),
        "h2agg_kernel_timing": (ci, [c_vp, ci]),
        "h2agg_kernel_times": (ci, [c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64), ci]),
        "h2agg_host_register": (ci, [c_vp, c_vp, sz]),
        "h2agg_host_unregister": (ci, [c_vp, c_vp]),
        "h2agg_set_msm_window": (ci, [c_vp, ci]),
        "h2agg_set_ntt_radix_cap": (ci, [c_vp, ci]),
        "h2agg_set_srs_precompute": (ci, [c_vp, ci]),
        "h2agg_srs_config": (ci, [c_vp, u64, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]),
        "h2agg_msm_config": (ci, [c_vp, sz, ctypes.POINTER(ci), ctypes.POINTER(ci)]),
        "h2agg_srs_register": (ci, [c_vp, c_vp, sz, ctypes.POINTER(u64)]),
        "h2agg_srs_register_dev": (ci, [c_vp, c_vp, sz, ctypes.POINTER(u64)]),
        "h2agg_srs_release": (ci, [c_vp, u64]),
        "h2agg_msm_g1": (ci, [c_vp, u64, c_vp, c_vp, sz, c_vp]),
        "h2agg_msm_g1_batch": (ci, [c_vp, u64, ctypes.POINTER(c_vp), sz, sz, c_vp]),
        "h2agg_msm_g1_batch_dev": (ci, [c_vp, u64, c_vp, ctypes.POINTER(c_vp), sz, sz, c_vp]),
        "h2agg_msm_g1_batch_windows_dev": (ci, [c_vp, u64, c_vp, ctypes.POINTER(c_vp), sz, sz, ci, ci, c_vp]),
        "h2agg_msm_g1_batch_ranges_dev": (ci, [c_vp, u64, c_vp, ctypes.POINTER(c_vp), sz, sz, ctypes.POINTER(ci), ctypes.POINTER(ci), c_vp]),
        "h2agg_g1_sum_dev": (ci, [c_vp, c_vp, sz, sz, sz, c_vp]),
        "h2agg_msm_g1_dev": (ci, [c_vp, u64, c_vp, c_vp, sz, c_vp]),
        "h2agg_msm_g1_windows": (ci, [c_vp, u64, c_vp, c_vp, sz, ci, ci, c_vp]),
        "h2agg_msm_g1_windows_dev": (ci, [c_vp, u64, c_vp, c_vp, sz, ci, ci, c_vp]),
        "h2agg_g1_sum": (ci, [c_vp, c_vp, sz, c_vp]),
        "h2agg_ntt_fr": (ci, [c_vp, c_vp, c_vp, u32]),
        "h2agg_intt_fr": (ci, [c_vp, c_vp, c_vp, c_vp, u32]),
        "h2agg_intt_fr_batch": (ci, [c_vp, ctypes.POINTER(c_vp), sz, c_vp, c_vp, u32]),
        "h2agg_coeff_to_extended_batch": (ci, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), sz, u32, u32, c_vp, c_vp]),
        "h2agg_commit_round": (ci, [c_vp, u64, ctypes.POINTER(c_vp), sz, u32, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), u32, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "h2agg_commit_round_resident": (ci, [c_vp, u64, ctypes.POINTER(c_vp), sz, u32, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), u32, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "h2agg_commit_round_dev": (ci, [c_vp, u64, ctypes.POINTER(c_vp), sz, u32, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp), u32, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "h2agg_compress_expressions_dev": (ci, [c_vp, c_vp, sz, ctypes.POINTER(c_vp), sz, c_vp, sz, u32, c_vp, c_vp]),
        "h2agg_lookup_product_dev": (ci, [c_vp, c_vp, c_vp, c_vp, c_vp, sz, c_vp, c_vp, c_vp]),
        "h2agg_lookup_products_dev": (ci, [c_vp, sz, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), sz, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "h2agg_permutation_product_dev": (ci, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), sz, u32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "h2agg_ntt_fr_dev": (ci, [c_vp, c_vp, c_vp, c_vp, u32]),
        "h2agg_coeff_to_extended": (ci, [c_vp, c_vp, u32, u32, c_vp, c_vp, c_vp]),
        "h2agg_extended_to_coeff": (ci, [c_vp, c_vp, u32, c_vp, c_vp, c_vp, sz]),
        "h2agg_coeff_to_extended_dev": (ci, [c_vp, c_vp, u32, u32, c_vp, c_vp, c_vp]),
        "h2agg_extended_to_coeff_dev": (ci, [c_vp, c_vp, u32, c_vp, c_vp, c_vp, sz]),
        "h2agg_eval_polynomial": (ci, [c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_eval_polynomial_dev": (ci, [c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_eval_polynomials_dev": (ci, [c_vp, ctypes.POINTER(c_vp), sz, sz, c_vp, c_vp]),
        "h2agg_kate_division": (ci, [c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_kate_division_dev": (ci, [c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_batch_invert": (ci, [c_vp, c_vp, sz]),
        "h2agg_batch_invert_dev": (ci, [c_vp, c_vp, sz]),
        "h2agg_grand_product": (ci, [c_vp, c_vp, c_vp, sz, c_vp]),
        "h2agg_grand_product_dev": (ci, [c_vp, c_vp, c_vp, sz, c_vp]),
        "h2agg_permute_expression_pair": (ci, [c_vp, c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_permute_expression_pair_dev": (ci, [c_vp, c_vp, c_vp, sz, c_vp, c_vp]),
        "h2agg_sort_fr": (ci, [c_vp, c_vp, sz]),
        "h2agg_sort_fr_dev": (ci, [c_vp, c_vp, sz]),
        "h2agg_evaluate_h_dev": (ci, [c_vp, ctypes.POINTER(QuotientArgs), c_vp]),
        "h2agg_poly_fold_dev": (ci, [c_vp, ctypes.POINTER(c_vp), sz, sz, c_vp, c_vp]),
        "h2agg_poly_lincomb_dev": (ci, [c_vp, ctypes.POINTER(c_vp), c_vp, sz, sz, c_vp]),
        "h2agg_evaluate_h_rows_dev": (ci, [c_vp, ctypes.POINTER(QuotientArgs), u64, u64, c_vp, c_vp]),
        "h2agg_set_defer_transforms": (ci, [c_vp, ci]),
        "h2agg_transforms_join": (ci, [c_vp]),
        "h2agg_transforms_dev": (ci, [c_vp, ctypes.POINTER(c_vp), sz, u32, c_vp, c_vp, ctypes.POINTER(c_vp), u32, c_vp, c_vp, ctypes.POINTER(c_vp)]),
        "h2agg_wit_new": (c_vp, []),
        "h2agg_wit_set_threads": (ci, [ci]),
        "h2agg_wit_free": (None, [c_vp]),
        "h2agg_wit_error": (ctypes.c_char_p, [c_vp]),
        "h2agg_wit_rows": (u64, [c_vp]),
        "h2agg_wit_ops": (u64, [c_vp]),
        "h2agg_wit_assign_point": (ctypes.c_int64, [c_vp, c_vp]),
        "h2agg_wit_assign_constant_point": (ctypes.c_int64, [c_vp, c_vp]),
        "h2agg_wit_assign_scalar": (ctypes.c_int64, [c_vp, c_vp]),
        "h2agg_wit_ecc_add": (ctypes.c_int64, [c_vp, ctypes.c_int64, ctypes.c_int64]),
        "h2agg_wit_ecc_sub": (ctypes.c_int64, [c_vp, ctypes.c_int64, ctypes.c_int64]),
        "h2agg_wit_ecc_double": (ctypes.c_int64, [c_vp, ctypes.c_int64]),
        "h2agg_wit_ecc_reduce": (ctypes.c_int64, [c_vp, ctypes.c_int64]),
        "h2agg_wit_ecc_mul": (ctypes.c_int64, [c_vp, ctypes.c_int64, ctypes.c_int64]),
        "h2agg_wit_ecc_shamir": (ctypes.c_int64, [c_vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), sz]),
        "h2agg_wit_ecc_constant_mul": (ctypes.c_int64, [c_vp, c_vp, ctypes.c_int64]),
        "h2agg_wit_point_value": (ci, [c_vp, ctypes.c_int64, c_vp, ctypes.POINTER(ci)]),
        "h2agg_wit_field_assign_const": (i64, [c_vp, c_vp]),
        "h2agg_wit_field_add": (i64, [c_vp, i64, i64]),
        "h2agg_wit_field_sub": (i64, [c_vp, i64, i64]),
        "h2agg_wit_field_mul": (i64, [c_vp, i64, i64]),
        "h2agg_wit_field_square": (i64, [c_vp, i64]),
        "h2agg_wit_field_div": (i64, [c_vp, i64, i64]),
        "h2agg_wit_field_sum_with_coeff_and_constant": (i64, [c_vp, ctypes.POINTER(i64), c_vp, sz, c_vp]),
        "h2agg_wit_field_mul_add_constant": (i64, [c_vp, i64, i64, c_vp]),
        "h2agg_wit_scalar_value": (ci, [c_vp, i64, c_vp]),
        "h2agg_wit_scalar_cell": (ci, [c_vp, i64, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
        "h2agg_wit_encode_point": (ci, [c_vp, i64, ctypes.POINTER(i64)]),
        "h2agg_wit_ecc_assign_identity": (i64, [c_vp]),
        "h2agg_wit_ecc_assert_equal": (ci, [c_vp, i64, i64]),
        "h2agg_wit_assert_not_identity": (ci, [c_vp, i64]),
        "h2agg_wit_expose_final_pair": (ci, [c_vp, i64, i64, ctypes.POINTER(i64)]),
        "h2agg_witness_expand": (ci, [c_vp, c_vp, ctypes.POINTER(c_vp), sz]),
        "h2agg_witness_expand_dev": (ci, [c_vp, c_vp, ctypes.POINTER(c_vp), sz]),
        "h2agg_fr_repr": (ci, [c_vp, ci, c_vp, c_vp, sz]),
        "h2agg_g1_decompress": (ci, [c_vp, c_vp, c_vp, sz]),
        "h2agg_g1_decompress_dev": (ci, [c_vp, c_vp, c_vp, sz]),
        "h2agg_g1_compress": (ci, [c_vp, c_vp, c_vp, sz]),
        "h2agg_field_mul": (ci, [c_vp, ci, c_vp, c_vp, c_vp, sz]),
        "h2agg_field_op": (ci, [c_vp, ci, ci, c_vp, c_vp, c_vp, sz]),
        "h2agg_dev_alloc": (ci, [c_vp, sz, ctypes.POINTER(c_vp)]),
        "h2agg_dev_free": (ci, [c_vp, c_vp]),
        "h2agg_memcpy_h2d": (ci, [c_vp, c_vp, c_vp, sz]),
        "h2agg_memcpy_d2h": (ci, [c_vp, c_vp, c_vp, sz]),
        "h2agg_synth_scalars_dev": (ci, [c_vp, u64, ci, u64, u64, c_vp]),
        "h2agg_synth_bases_dev": (ci, [c_vp, u64, u64, u64, c_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    lib._h2agg_signatures = sig
    _lib = lib
    return lib
