"""Mirror of halo2_proofs ParamsKZG::{commit_lagrange, commit}:
    commit_lagrange(poly, _blind) = best_multiexp(poly.values, g_lagrange)
    commit(poly, _blind)          = best_multiexp(poly.values, g[..poly.len()])
with the SRS registered once and kept resident in HBM (get_params_cached,
halo2-snark-aggregator-circuit/src/verify_circuit.rs:701-731 keeps it for the life of the run)."""
import numpy as np

from .context import default_context


class ParamsKZG:
    def __init__(self, k, g, g_lagrange, ctx=None):
        self.k, self.n = k, 1 << k
        self.ctx = ctx or default_context()
        assert g.size == 8 * self.n and g_lagrange.size == 8 * self.n
        self._g = self.ctx.srs_register(np.ascontiguousarray(g))
        self._gl = self.ctx.srs_register(np.ascontiguousarray(g_lagrange))

    def commit_lagrange(self, values, blind=None):
        assert values.size == 4 * self.n, "commit_lagrange: poly.len() != n"
        return self.ctx.msm_g1(values, srs_id=self._gl)

    def commit(self, coeffs, blind=None):
        assert coeffs.size <= 4 * self.n, "commit: poly.len() > n"
        return self.ctx.msm_g1(coeffs, srs_id=self._g)

    def commit_lagrange_many(self, columns):
        """One commit round (e.g. the 5 advice columns): returns affine points, shape (len, 8)."""
        return self.ctx.msm_g1_batch(self._gl, list(columns), self.n)

    def release(self):
        self.ctx.srs_release(self._g)
        self.ctx.srs_release(self._gl)
