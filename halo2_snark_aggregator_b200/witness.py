"""Host mirror of the reference's in-circuit chip surface for the witness path (W1-W6).

The reference drives its aggregation circuit through three chips plus an encoder, all bound to one `Context`
(halo2-snark-aggregator-circuit/src/chips/{ecc_chip,scalar_chip,encode_chip}.rs):
    EccChip            ArithCommonChip + ArithEccChip   (api/src/arith/{common,ecc}.rs)
    ScalarChip         ArithCommonChip + ArithFieldChip (api/src/arith/field.rs)   -- also the NativeChip
    PoseidonEncodeChip Encode                           (api/src/transcript/encode.rs)
`B200Context` is the shared recording context (an h2agg_witness); `B200EccChip`, `B200ScalarChip`, `B200EncodeChip`
keep the reference's method names and argument meaning -- add / sub / scalar_mul / scalar_mul_constant / multi_exp /
assign_var / assign_const / normalize / to_value, mul / div / square / sum_with_coeff_and_constant / mul_add_constant
(+ the trait's provided compositions), encode_point / encode_scalar / decode_scalar -- but every call only records the
row layout and the values the chain needs; `expand()` runs the B200 kernel and returns the 5 advice columns.
Point values are affine Montgomery limb arrays (8 x u64, zeros = identity); scalar values are canonical ints < r."""
import ctypes

import numpy as np

from . import _lib
from ._lib import H2aggError, c_vp

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_M64 = (1 << 64) - 1


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return c_vp(a.ctypes.data)


def _limbs(v):
    v %= R_MOD
    return np.array([(v >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


_U64x4 = ctypes.c_uint64 * 4


def _c4(v):
    """canonical scalar -> 4 x u64 for a C-ABI argument (plain ctypes: the ScalarChip calls are tiny, numpy
    array construction would dominate them)"""
    v %= R_MOD
    return _U64x4(v & _M64, (v >> 64) & _M64, (v >> 128) & _M64, v >> 192)


class B200Context:
    """`Context` of the circuit chips (halo2-ecc-circuit-lib/src/gates/base_gate.rs:113-140): the row offset plus, here,
    the op records the expansion kernel consumes."""

    def __init__(self):
        self.lib = _lib.load()
        self.h = c_vp(self.lib.h2agg_wit_new())

    def close(self):
        if self.h:
            self.lib.h2agg_wit_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _h(self, v):
        if v < 0:
            raise H2aggError("witness recorder: " + self.lib.h2agg_wit_error(self.h).decode())
        return v

    @property
    def offset(self):
        return int(self.lib.h2agg_wit_rows(self.h))

    def rows(self):
        return self.offset

    def ops(self):
        return int(self.lib.h2agg_wit_ops(self.h))

    def __str__(self):      # Context: Display prints the current row offset (base_gate.rs:142-146)
        return "(offset: %d)" % self.offset

    # ---- expansion: the five advice columns
    def expand(self, ctx, n_rows=None):
        """-> (5, n_rows, 4) uint64: the advice columns in Montgomery form."""
        n_rows = self.rows() if n_rows is None else n_rows
        cols = np.zeros((5, n_rows, 4), dtype=np.uint64)
        ptrs = (c_vp * 5)(*[cols[c].ctypes.data for c in range(5)])
        ctx.check(self.lib.h2agg_witness_expand(ctx.h, self.h, ptrs, n_rows))
        return cols

    def expand_dev(self, ctx, d_cols, n_rows):
        ptrs = (c_vp * 5)(*d_cols)
        ctx.check(self.lib.h2agg_witness_expand_dev(ctx.h, self.h, ptrs, n_rows))


class B200ScalarChip:
    """ArithFieldChip over the base gate = ScalarChip (chips/scalar_chip.rs:17-127).  Handles are ints."""

    def __init__(self, wctx):
        self.w = wctx
        self.lib = wctx.lib

    # -- ArithCommonChip
    def add(self, a, b):
        return self.w._h(self.lib.h2agg_wit_field_add(self.w.h, a, b))

    def sub(self, a, b):
        return self.w._h(self.lib.h2agg_wit_field_sub(self.w.h, a, b))

    def assign_zero(self):
        return self.assign_const(0)

    def assign_one(self):
        return self.assign_const(1)

    def assign_const(self, c):
        return self.w._h(self.lib.h2agg_wit_field_assign_const(self.w.h, _c4(c)))

    def assign_var(self, v):
        return self.w._h(self.lib.h2agg_wit_assign_scalar(self.w.h, _c4(v)))

    def to_value(self, h):
        out = np.zeros(4, dtype=np.uint64)
        if self.lib.h2agg_wit_scalar_value(self.w.h, h, _p(out)) != 0:
            raise H2aggError("bad scalar handle")
        return sum(int(x) << (64 * i) for i, x in enumerate(out))

    def normalize(self, v):
        return v

    def cell(self, h):
        """(advice column, row) of an AssignedValue: the Cell constrain_instance binds (verify_circuit.rs:357-367)"""
        c, r = ctypes.c_uint32(), ctypes.c_uint32()
        if self.lib.h2agg_wit_scalar_cell(self.w.h, h, ctypes.byref(c), ctypes.byref(r)) != 0:
            raise H2aggError("bad scalar handle")
        return c.value, r.value

    # -- ArithFieldChip: required methods
    def mul(self, a, b):
        return self.w._h(self.lib.h2agg_wit_field_mul(self.w.h, a, b))

    def div(self, a, b):
        return self.w._h(self.lib.h2agg_wit_field_div(self.w.h, a, b))

    def square(self, a):
        return self.w._h(self.lib.h2agg_wit_field_square(self.w.h, a))

    def sum_with_coeff_and_constant(self, a_with_coeff, b):
        n = len(a_with_coeff)
        hs = (ctypes.c_int64 * max(n, 1))(*[h for h, _ in a_with_coeff])
        words = []
        for _, c in a_with_coeff:
            c %= R_MOD
            words += [c & _M64, (c >> 64) & _M64, (c >> 128) & _M64, c >> 192]
        co = (ctypes.c_uint64 * max(4 * n, 4))(*words)
        return self.w._h(self.lib.h2agg_wit_field_sum_with_coeff_and_constant(self.w.h, hs, co, n, _c4(b)))

    def mul_add_constant(self, a, b, c):
        return self.w._h(self.lib.h2agg_wit_field_mul_add_constant(self.w.h, a, b, _c4(c)))

    # -- ArithFieldChip: provided methods (api/src/arith/field.rs:37-104)
    def sum_with_constant(self, a, b):
        return self.sum_with_coeff_and_constant([(x, 1) for x in a], b)

    def mul_add(self, a, b, c):
        return self.add(self.mul(a, b), c)

    def mul_add_accumulate(self, a, b):
        acc = self.assign_zero()
        for v in a:
            acc = self.mul_add(acc, b, v)
        return acc

    def pow_constant(self, base, exponent):
        assert exponent >= 1
        acc, second_bit = base, 1
        while second_bit <= exponent:
            second_bit <<= 1
        second_bit >>= 2
        while second_bit > 0:
            acc = self.square(acc)
            if exponent & second_bit:
                acc = self.mul(acc, base)
            second_bit >>= 1
        return acc


class B200EccChip:
    """ArithEccChip = EccChip over NativeEccChip (chips/ecc_chip.rs:28-133).  `B200EccChip()` alone creates its own
    context (the multi_exp-only use of round 1); pass a B200Context to share it with the other chips."""

    def __init__(self, wctx=None):
        self.w = wctx if wctx is not None else B200Context()
        self.lib = self.w.lib
        self.h = self.w.h
        self.scalar_chip = B200ScalarChip(self.w)

    def close(self):
        self.w.close()
        self.h = None

    def _h(self, v):
        return self.w._h(v)

    # ---- ArithCommonChip / ArithEccChip
    def assign_var(self, xy):  # transcript point: assign_point with the on-curve check
        return self._h(self.lib.h2agg_wit_assign_point(self.h, _p(np.ascontiguousarray(xy, dtype=np.uint64))))

    def assign_const(self, xy):
        return self._h(self.lib.h2agg_wit_assign_constant_point(self.h, _p(np.ascontiguousarray(xy, dtype=np.uint64))))

    def assign_zero(self):
        return self._h(self.lib.h2agg_wit_ecc_assign_identity(self.h))

    def assign_scalar(self, canonical_int):
        return self.scalar_chip.assign_var(canonical_int)

    def add(self, a, b):
        return self._h(self.lib.h2agg_wit_ecc_add(self.h, a, b))

    def sub(self, a, b):
        return self._h(self.lib.h2agg_wit_ecc_sub(self.h, a, b))

    def double(self, a):
        return self._h(self.lib.h2agg_wit_ecc_double(self.h, a))

    def normalize(self, a):
        return self._h(self.lib.h2agg_wit_ecc_reduce(self.h, a))

    def scalar_mul(self, s, a):
        return self._h(self.lib.h2agg_wit_ecc_mul(self.h, a, s))

    def scalar_mul_constant(self, s, base_xy):
        return self._h(self.lib.h2agg_wit_ecc_constant_mul(self.h, _p(np.ascontiguousarray(base_xy, dtype=np.uint64)), s))

    def multi_exp(self, points, scalars):
        n = len(points)
        pa = (ctypes.c_int64 * n)(*points)
        sa = (ctypes.c_int64 * n)(*scalars)
        return self._h(self.lib.h2agg_wit_ecc_shamir(self.h, pa, sa, n))

    def to_value(self, h):
        out = np.zeros(8, dtype=np.uint64)
        ident = ctypes.c_int()
        if self.lib.h2agg_wit_point_value(self.h, h, _p(out), ctypes.byref(ident)) != 0:
            raise H2aggError("bad point handle")
        return out, bool(ident.value)

    # ---- what Halo2VerifierCircuits::synthesize does around the chips (verify_circuit.rs:264-368, 487-496)
    def assert_equal(self, a, b):
        if self.lib.h2agg_wit_ecc_assert_equal(self.h, a, b) != 0:
            self._h(-1)

    def assert_not_identity(self, p):
        if self.lib.h2agg_wit_assert_not_identity(self.h, p) != 0:
            self._h(-1)

    def expose_final_pair(self, w_x, w_g):
        """-> 4 AssignedValue handles: the cells constrain_instance binds to instance rows 0..3"""
        out = (ctypes.c_int64 * 4)()
        if self.lib.h2agg_wit_expose_final_pair(self.h, w_x, w_g, out) != 0:
            self._h(-1)
        return list(out)

    # ---- layout / expansion (kept on the chip for the single-chip use)
    def rows(self):
        return self.w.rows()

    def ops(self):
        return self.w.ops()

    def expand(self, ctx, n_rows=None):
        return self.w.expand(ctx, n_rows)

    def expand_dev(self, ctx, d_cols, n_rows):
        return self.w.expand_dev(ctx, d_cols, n_rows)


class B200EncodeChip:
    """Encode = PoseidonEncodeChip (chips/encode_chip.rs:14-51)."""

    def __init__(self, wctx):
        self.w = wctx
        self.lib = wctx.lib

    def encode_point(self, p):
        out = (ctypes.c_int64 * 2)()
        if self.lib.h2agg_wit_encode_point(self.w.h, p, out) != 0:
            self.w._h(-1)
        return [out[0], out[1]]

    def encode_scalar(self, s):
        return [s]

    def decode_scalar(self, v):
        return v[0]
