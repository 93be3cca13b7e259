"""Mirror of halo2_proofs::poly::EvaluationDomain for the transforms on the hot path
(SURVEY.md App. B3).  Constructor arguments as in Rust: EvaluationDomain::new(j, k) with
j = cs.degree(); the aggregation circuit has j = 5 -> extended_k = k + 2.
Field constants are computed with Python ints (host-side setup, not the hot path)."""
import numpy as np

from .context import default_context

_R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_S = 28
_ROOT_OF_UNITY = pow(7, (_R - 1) >> _S, _R)
_ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
_M64 = (1 << 64) - 1


def fr_to_limbs(x):
    v = (x % _R) * (1 << 256) % _R
    return np.array([(v >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


class EvaluationDomain:
    def __init__(self, j, k, ctx=None):
        self.ctx = ctx or default_context()
        self.k, self.n = k, 1 << k
        self.quotient_poly_degree = j - 1
        self.extended_k = k
        while (1 << self.extended_k) < self.n * self.quotient_poly_degree:
            self.extended_k += 1
        ext_omega = pow(_ROOT_OF_UNITY, 1 << (_S - self.extended_k), _R)
        omega = pow(ext_omega, 1 << (self.extended_k - k), _R)
        self._omega, self._omega_inv = omega, pow(omega, -1, _R)
        self._ext_omega, self._ext_omega_inv = ext_omega, pow(ext_omega, -1, _R)
        self.omega = fr_to_limbs(omega)
        self.omega_inv = fr_to_limbs(self._omega_inv)
        self.extended_omega = fr_to_limbs(ext_omega)
        self.extended_omega_inv = fr_to_limbs(self._ext_omega_inv)
        self.g_coset = fr_to_limbs(_ZETA)
        self.ifft_divisor = fr_to_limbs(pow(self.n, -1, _R))
        self.extended_ifft_divisor = fr_to_limbs(pow(1 << self.extended_k, -1, _R))

    def extended_len(self):
        return 1 << self.extended_k

    # host-pointer forms (in place like the Rust methods that take Polynomial by value)
    def lagrange_to_coeff(self, a):
        assert a.size == 4 * self.n
        self.ctx.intt_fr(a, self.omega_inv, self.ifft_divisor, self.k)
        return a

    def coeff_to_extended(self, a):
        assert a.size == 4 * self.n
        return self.ctx.coeff_to_extended(a, self.k, self.extended_k, self.g_coset, self.extended_omega)

    def extended_to_coeff(self, a):
        assert a.size == 4 * self.extended_len()
        return self.ctx.extended_to_coeff(a, self.extended_k, self.extended_omega_inv, self.extended_ifft_divisor,
                                          self.g_coset, self.n * self.quotient_poly_degree)

    # a whole round at once (pipelined over the context's lanes)
    def lagrange_to_coeff_many(self, cols):
        self.ctx.intt_fr_batch(cols, self.omega_inv, self.ifft_divisor, self.k)

    def coeff_to_extended_many(self, cols, outs):
        self.ctx.coeff_to_extended_batch(cols, outs, self.k, self.extended_k, self.g_coset, self.extended_omega)

    # device-resident forms (asynchronous on the context's stream)
    def lagrange_to_coeff_dev(self, d_a):
        self.ctx.ntt_fr_dev(d_a, self.omega_inv, self.k, scale=self.ifft_divisor)

    def coeff_to_extended_dev(self, d_coeffs, d_out):
        self.ctx.coeff_to_extended_dev(d_coeffs, self.k, self.extended_k, self.g_coset, self.extended_omega, d_out)

    def extended_to_coeff_dev(self, d_a):
        self.ctx.extended_to_coeff_dev(d_a, self.extended_k, self.extended_omega_inv, self.extended_ifft_divisor,
                                       self.g_coset, self.n * self.quotient_poly_degree)
