"""Mirror of halo2_proofs ParamsKZG:
    commit_lagrange(poly, _blind) = best_multiexp(poly.values, g_lagrange)
    commit(poly, _blind)          = best_multiexp(poly.values, g[..poly.len()])
    read / write                  the `verify_circuit.params` / `sample_circuit_*.params` / HALO2_PARAMS_k files
with the SRS registered once and kept resident in HBM (get_params_cached,
halo2-snark-aggregator-circuit/src/verify_circuit.rs:701-731 keeps it for the life of the run).

File layout (halo2_proofs poly/kzg/commitment.rs `ParamsKZG::write`, PSE v2022_09_10 -- an external crate, restated from its
published source, not pinned against a file the real crate wrote):
    k            u32 little-endian
    g            2^k compressed G1 points (32 B each, halo2curves G1Affine::to_bytes)
    g_lagrange   2^k compressed G1 points
    g2, s_g2     two compressed G2 points (64 B each) -- carried through as opaque bytes: the prover never touches G2
Reading decodes the 2 x 2^k points on the GPU (h2agg_g1_decompress: one Fq square root each) straight into the two
resident SRS tables."""
import numpy as np

from .context import default_context


class ParamsKZG:
    def __init__(self, k, g, g_lagrange, ctx=None, g2_bytes=bytes(64), s_g2_bytes=bytes(64)):
        self.k, self.n = k, 1 << k
        self.ctx = ctx or default_context()
        assert g.size == 8 * self.n and g_lagrange.size == 8 * self.n
        self._g = self.ctx.srs_register(np.ascontiguousarray(g))
        self._gl = self.ctx.srs_register(np.ascontiguousarray(g_lagrange))
        self._host = (g, g_lagrange)
        self.g2_bytes, self.s_g2_bytes = bytes(g2_bytes), bytes(s_g2_bytes)

    @property
    def srs_g(self):
        return self._g

    @property
    def srs_g_lagrange(self):
        return self._gl

    def commit_lagrange(self, values, blind=None):
        assert values.size == 4 * self.n, "commit_lagrange: poly.len() != n"
        return self.ctx.msm_g1(values, srs_id=self._gl)

    def commit(self, coeffs, blind=None):
        assert coeffs.size <= 4 * self.n, "commit: poly.len() > n"
        return self.ctx.msm_g1(coeffs, srs_id=self._g)

    def commit_lagrange_many(self, columns):
        """One commit round (e.g. the 5 advice columns): returns affine points, shape (len, 8)."""
        return self.ctx.msm_g1_batch(self._gl, list(columns), self.n)

    # ---- ParamsKZG::read / write ---------------------------------------------------------------------------------
    @classmethod
    def read(cls, data, ctx=None):
        """bytes of a params file -> ParamsKZG with both tables resident (raises H2aggError, status 4, on a bad point)"""
        ctx = ctx or default_context()
        data = bytes(data)
        k = int.from_bytes(data[:4], "little")
        n = 1 << k
        need = 4 + 2 * 32 * n + 128
        if k > 28 or len(data) < need:
            raise ValueError("params file too short for k = %d: %d < %d bytes" % (k, len(data), need))
        g = ctx.g1_decompress(data[4:4 + 32 * n])
        gl = ctx.g1_decompress(data[4 + 32 * n:4 + 64 * n])
        off = 4 + 64 * n
        return cls(k, g, gl, ctx, data[off:off + 64], data[off + 64:off + 128])

    def write(self):
        g, gl = self._host
        return self.k.to_bytes(4, "little") + self.ctx.g1_compress(np.ascontiguousarray(g)) + \
            self.ctx.g1_compress(np.ascontiguousarray(gl)) + self.g2_bytes + self.s_g2_bytes

    def release(self):
        self.ctx.srs_release(self._g)
        self.ctx.srs_release(self._gl)
