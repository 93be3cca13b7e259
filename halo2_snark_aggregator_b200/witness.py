"""Host mirror of the reference's in-circuit ECC chip surface for the witness path (W1-W5).

`B200EccChip` plays the role of halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs (`EccChip`,
the Circuit implementation of `ArithEccChip`): same method names and meaning --
add / sub / scalar_mul / scalar_mul_constant / multi_exp / assign_var / assign_const / normalize /
to_value (halo2-snark-aggregator-api/src/arith/{common,ecc}.rs) -- but every call only records;
`expand()` runs the B200 kernel and returns the 5 advice columns."""
import ctypes

import numpy as np

from . import _lib
from ._lib import H2aggError, c_vp


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return c_vp(a.ctypes.data)


class B200EccChip:
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

    # ---- ArithCommonChip / ArithEccChip
    def assign_var(self, xy):  # transcript point: assign_point with the on-curve check
        return self._h(self.lib.h2agg_wit_assign_point(self.h, _p(np.ascontiguousarray(xy, dtype=np.uint64))))

    def assign_const(self, xy):
        return self._h(self.lib.h2agg_wit_assign_constant_point(self.h, _p(np.ascontiguousarray(xy, dtype=np.uint64))))

    def assign_scalar(self, canonical_int):
        limbs = np.array([(canonical_int >> (64 * i)) & ((1 << 64) - 1) for i in range(4)], dtype=np.uint64)
        return self._h(self.lib.h2agg_wit_assign_scalar(self.h, _p(limbs)))

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

    # ---- layout / expansion
    def rows(self):
        return int(self.lib.h2agg_wit_rows(self.h))

    def ops(self):
        return int(self.lib.h2agg_wit_ops(self.h))

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
