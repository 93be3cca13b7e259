"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/cpu_halo2.cpp header).

Independent big-integer reference for the BN254 arithmetic of the hot path: plain Python ints,
`pow` and `%`, affine-coordinate textbook group law, O(n^2) DFT.  It shares nothing with either the
C++ oracle or the CUDA kernels and is what the golden vectors in tests/golden/ are minted from
(the reference holds no vectors for this path, so they are minted here from the
mathematical definition of the reference's call-site semantics:
  best_multiexp(coeffs, bases) = sum_i coeffs[i] * bases[i]            (create_proof, verify_circuit.rs:986)
  best_fft(a, omega, log_n)[k]  = sum_j a[j] * omega^(j k)             natural order in and out
  EvaluationDomain::{ifft, coeff_to_extended, extended_to_coeff}       SURVEY.md App. B3).
"""
import struct

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr
MONT = 1 << 256
B = 3
S = 28
GENERATOR = 7
T = (R - 1) >> S
ROOT_OF_UNITY = pow(GENERATOR, T, R)  # order 2^28
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23  # cube root of unity in Fr (halo2curves Fr::ZETA)
assert pow(ZETA, 3, R) == 1 and ZETA != 1
assert pow(ROOT_OF_UNITY, 1 << 28, R) == 1 and pow(ROOT_OF_UNITY, 1 << 27, R) != 1
M64 = (1 << 64) - 1


def omega(k):
    return pow(ROOT_OF_UNITY, 1 << (S - k), R)


# ---- Montgomery <-> limbs -------------------------------------------------------------------------
def to_mont_limbs(x, mod):
    v = (x % mod) * MONT % mod
    return [(v >> (64 * i)) & M64 for i in range(4)]


def from_mont_limbs(limbs, mod):
    v = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    return v * pow(MONT, -1, mod) % mod


def pack_fr(vals):
    out = []
    for x in vals:
        out += to_mont_limbs(x, R)
    return out


def unpack_fr(limbs):
    return [from_mont_limbs(limbs[4 * i:4 * i + 4], R) for i in range(len(limbs) // 4)]


def pack_points(pts):
    """pts: list of None (identity) or (x, y) canonical ints -> flat u64 list, 8 per point."""
    out = []
    for p in pts:
        if p is None:
            out += [0] * 8
        else:
            out += to_mont_limbs(p[0], P) + to_mont_limbs(p[1], P)
    return out


def unpack_point(limbs8):
    x, y = from_mont_limbs(limbs8[0:4], P), from_mont_limbs(limbs8[4:8], P)
    return None if (x == 0 and y == 0) else (x, y)


def unpack_jacobian(limbs12):
    x, y, z = (from_mont_limbs(limbs12[4 * i:4 * i + 4], P) for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, P)
    return (x * zi * zi % P, y * zi * zi * zi % P)


# ---- G1, affine textbook law (None = identity) --------------------------------------------------------
def on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - B) % P == 0


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, P) % P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x = (lam * lam - p[0] - q[0]) % P
    return (x, (lam * (p[0] - x) - p[1]) % P)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % P)


def g1_mul(k, p):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def msm(scalars, points):
    acc = None
    for s, p in zip(scalars, points):
        acc = g1_add(acc, g1_mul(s, p))
    return acc


G1_GEN = (1, 2)

# ---- transforms -------------------------------------------------------------------------------------


def dft(a, w):
    n = len(a)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * w % R
    return [sum(a[j] * pw[(j * k) % n] for j in range(n)) % R for k in range(n)]


def ifft(a, k):
    n = len(a)
    w_inv = pow(omega(k), -1, R)
    n_inv = pow(n, -1, R)
    return [x * n_inv % R for x in dft(a, w_inv)]


def coeff_to_extended(coeffs, k, ext_k):
    n, en = 1 << k, 1 << ext_k
    a = [c * pow(ZETA, i % 3, R) % R for i, c in enumerate(coeffs)] + [0] * (en - n)
    return dft(a, omega(ext_k))


def extended_to_coeff(a, ext_k, out_len):
    en = 1 << ext_k
    w_inv = pow(omega(ext_k), -1, R)
    d = pow(en, -1, R)
    zi = pow(ZETA, -1, R)
    c = [x * d % R for x in dft(a, w_inv)]
    c = [x * pow(zi, i % 3, R) % R for i, x in enumerate(c)]
    return c[:out_len]


# ---- deterministic synthetic inputs: must match oracle_gen_* in cpu_halo2.cpp and synth.cu --------------
def _splitmix(state):
    state = (state + 0x9e3779b97f4a7c15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M64
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M64
    return state, z ^ (z >> 31)


def _stream_seed(seed, index):
    s = (seed ^ ((index * 0xd1342543de82ef95 + 0x2545f4914f6cdd1d) & M64)) & M64
    s, _ = _splitmix(s)
    return s


def _draw_below(s, mod):
    while True:
        w = []
        for _ in range(4):
            s, z = _splitmix(s)
            w.append(z)
        w[3] &= 0x3fffffffffffffff
        v = sum(x << (64 * i) for i, x in enumerate(w))
        if v < mod:
            return s, v


def gen_scalar(seed, kind, index):
    s = _stream_seed(seed, index)
    s, z = _splitmix(s)
    sel = z % 100
    if kind == 0:
        return _draw_below(s, R)[1]
    if kind == 1:
        if sel < 70:
            return _splitmix(s)[1] & 0x1ffff
        if sel < 80:
            return _splitmix(s)[1] & 1
        return 0
    if kind == 2:
        if sel < 50:
            s, lo = _splitmix(s)
            s, hi = _splitmix(s)
            return lo | ((hi & 0xf) << 64)
        if sel < 70:
            return _draw_below(s, R)[1]
        return 0
    return _splitmix(s)[1] & 0x1ffff


def gen_base(seed, index):
    s = _stream_seed(seed, index)
    s, x = _draw_below(s, P)
    while True:
        rhs = (x * x * x + B) % P
        y = pow(rhs, (P + 1) // 4, P)
        if y * y % P == rhs and y != 0:
            if y & 1:
                y = P - y
            return (x, y)
        x = (x + 1) % P


def limbs_to_bytes(limbs):
    return struct.pack("<%dQ" % len(limbs), *limbs)
