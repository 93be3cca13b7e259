"""TEST INFRASTRUCTURE (oracle) -- Poseidon over BN254 Fr as the reference's aggregation circuit uses it.

The reference hashes its in-circuit Fiat-Shamir transcript with `PoseidonChip<A, T = 9, RATE = 8>`
(halo2-snark-aggregator-api/src/hash/poseidon.rs:150-231; T, RATE, R_F = 8, R_P = 63 fixed at
halo2-snark-aggregator-circuit/src/verify_circuit.rs:128-133).  The constants come from the external crate
`poseidon 0.2.0` (privacy-scaling-explorations/poseidon @ 0b9965fb, Cargo.lock:2517-2519), which is NOT under
/root/reference.  What is restated here is that crate's published algorithm:

  Grain LFSR      the Poseidon paper's parameter generator: 80-bit state = field tag (1, 2 bits) | s-box tag (0, 4 bits)
                  | field size 254 (12 bits) | t (12) | R_F (10) | R_P (10) | thirty 1-bits; 160 bits discarded; output
                  bits in pairs (the second is kept when the first is 1); field elements from 254 bits MSB first,
                  round constants by rejection sampling, MDS seeds reduced mod r.
  MDS             Cauchy matrix 1 / (x_i + y_j) from 2 t seeds drawn after the round constants.
  Spec::new       optimised constants (start / partial / end) and the factorisation of the partial rounds into sparse
                  matrices (Poseidon paper, appendix B), as the chip consumes them: constants().start()/partial()/end(),
                  mds_matrices().{mds, pre_sparse_mds, sparse_matrices}.
  State::default  [2^64, 0, ..., 0].

Pins (tests/test_poseidon_cpu.py):
  * the generator reproduces a known answer that does not come from this repository: circomlib's
    poseidon([1, 2]) = 0x115cc0f5...189a, whose constants are this LFSR at (t = 3, R_F = 8, R_P = 57);
  * the optimised permutation (what the chip computes) equals the textbook permutation on the raw constants for
    (T = 9, R_F = 8, R_P = 63) and other shapes -- the factorisation is checked, not trusted.
Parity against the crate's own numbers stays unpinned (no Rust toolchain, no vendored source).
"""
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
NUM_BITS = 254


class Grain:
    def __init__(self, t, r_f, r_p, field_bits=NUM_BITS):
        bits = []

        def push(value, length):
            for i in reversed(range(length)):
                bits.append((value >> i) & 1)

        push(1, 2)            # prime field
        push(0, 4)            # x^alpha s-box
        push(field_bits, 12)
        push(t, 12)
        push(r_f, 10)
        push(r_p, 10)
        bits += [1] * 30
        assert len(bits) == 80
        self.state = bits
        self.field_bits = field_bits
        for _ in range(160):
            self._step()

    def _step(self):
        s = self.state
        new = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(new)
        return new

    def next_bit(self):
        while True:
            first = self._step()
            second = self._step()
            if first:
                return second

    def _next_int(self):
        v = 0
        for _ in range(self.field_bits):
            v = (v << 1) | self.next_bit()
        return v

    def next_field_element(self, modulus=R):
        while True:
            v = self._next_int()
            if v < modulus:
                return v

    def next_field_element_without_rejection(self, modulus=R):
        return self._next_int() % modulus


def generate(t, r_f, r_p, modulus=R):
    """-> (round constants [r_f + r_p][t], mds [t][t])   (poseidon crate: Grain::generate)"""
    g = Grain(t, r_f, r_p)
    constants = [[g.next_field_element(modulus) for _ in range(t)] for _ in range(r_f + r_p)]
    xs = [g.next_field_element_without_rejection(modulus) for _ in range(t)]
    ys = [g.next_field_element_without_rejection(modulus) for _ in range(t)]
    mds = [[pow((x + y) % modulus, -1, modulus) for y in ys] for x in xs]
    return constants, mds


# ---- small dense linear algebra mod r ------------------------------------------------------------------------
def mat_identity(n):
    return [[1 if i == j else 0 for j in range(n)] for i in range(n)]


def mat_transpose(m):
    return [list(r) for r in zip(*m)]


def mat_mul(a, b):
    n, k, p = len(a), len(b), len(b[0])
    return [[sum(a[i][x] * b[x][j] for x in range(k)) % R for j in range(p)] for i in range(n)]


def mat_vec(m, v):
    return [sum(a * b for a, b in zip(row, v)) % R for row in m]


def mat_invert(m):
    n = len(m)
    a = [list(row) + ident for row, ident in zip(m, mat_identity(n))]
    for c in range(n):
        p = next(r for r in range(c, n) if a[r][c] % R)
        a[c], a[p] = a[p], a[c]
        inv = pow(a[c][c], -1, R)
        a[c] = [x * inv % R for x in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % R for x, y in zip(a[r], a[c])]
    return [row[n:] for row in a]


class Spec:
    """poseidon::Spec::new(r_f, r_p) for width t: what PoseidonChip reads (API/hash/poseidon.rs:203-230)."""

    def __init__(self, t, r_f, r_p):
        self.t, self.r_f, self.r_p = t, r_f, r_p
        self.raw_constants, self.mds = generate(t, r_f, r_p)
        self._optimise_constants()
        self._sparse_matrices()

    def _optimise_constants(self):
        t, r_f, r_p, c = self.t, self.r_f, self.r_p, self.raw_constants
        half = r_f // 2
        inv = mat_invert(self.mds)
        start = [list(c[0])] + [mat_vec(inv, c[i]) for i in range(1, half)]
        acc = list(c[half + r_p])
        partial = [0] * r_p
        for i in reversed(range(r_p)):          # rounds half + r_p - 1 ... half
            tmp = mat_vec(inv, acc)
            partial[i] = tmp[0]
            tmp[0] = 0
            acc = [(a + b) % R for a, b in zip(tmp, c[half + i])]
        start.append(mat_vec(inv, acc))
        end = [mat_vec(inv, c[i]) for i in range(half + r_p + 1, r_f + r_p)]
        assert len(start) == half + 1 and len(end) == half - 1
        self.start, self.partial, self.end = start, partial, end

    @staticmethod
    def _factorise(m):
        """m = m' * m'' with m' = diag(1, m_hat) and m'' sparse (first row + first column + identity)."""
        t = len(m)
        w = [m[i][0] for i in range(1, t)]
        m_hat = [row[1:] for row in m[1:]]
        w_hat = mat_vec(mat_invert(m_hat), w)
        prime = mat_identity(t)
        for i in range(1, t):
            for j in range(1, t):
                prime[i][j] = m_hat[i - 1][j - 1]
        pp = mat_identity(t)
        pp[0] = list(m[0])
        for i in range(1, t):
            pp[i][0] = w_hat[i - 1]
        ppt = mat_transpose(pp)
        return prime, {"row": list(ppt[0]), "col_hat": [ppt[i][0] for i in range(1, t)]}

    def _sparse_matrices(self):
        mds_t = mat_transpose(self.mds)
        acc = [list(r) for r in mds_t]
        sparse = []
        for _ in range(self.r_p):
            prime, pp = self._factorise(acc)
            acc = mat_mul(mds_t, prime)
            sparse.append(pp)
        sparse.reverse()
        self.sparse_matrices = sparse
        self.pre_sparse_mds = mat_transpose(acc)


def state_default(t):
    """poseidon::State::default(): the capacity word is 2^64"""
    return [1 << 64] + [0] * (t - 1)


def permute_textbook(state, constants, mds, r_f, r_p):
    """add round constants -> s-box -> MDS, r_f/2 full + r_p partial + r_f/2 full rounds (Poseidon paper, fig. 2)"""
    t = len(state)
    s = list(state)
    half = r_f // 2
    for rnd in range(r_f + r_p):
        s = [(a + b) % R for a, b in zip(s, constants[rnd])]
        if rnd < half or rnd >= half + r_p:
            s = [pow(x, 5, R) for x in s]
        else:
            s[0] = pow(s[0], 5, R)
        s = mat_vec(mds, s)
    assert len(s) == t
    return s


def permute_optimised(spec, state, inputs):
    """The value-level walk of PoseidonChip::permutation (API/hash/poseidon.rs:196-230), absorbing `inputs`
    (len < T) first (absorb_with_pre_constants, :47-88: s[0] += c0; s[i] += input + c_i; s[len + 1] += c + 1)."""
    t = spec.t
    assert len(inputs) < t
    s = list(state)
    pre = spec.start[0]
    off = len(inputs) + 1
    s[0] = (s[0] + pre[0]) % R
    for i, v in enumerate(inputs):
        s[i + 1] = (s[i + 1] + v + pre[i + 1]) % R
    for i in range(off, t):
        s[i] = (s[i] + pre[i] + (1 if i == off else 0)) % R
    half = spec.r_f // 2

    def sbox_full(consts):
        return [(pow(x, 5, R) + c) % R for x, c in zip(s, consts)]

    for consts in spec.start[1:half]:
        s = sbox_full(consts)
        s = mat_vec(spec.mds, s)
    s = sbox_full(spec.start[-1])
    s = mat_vec(spec.pre_sparse_mds, s)
    for c, sp in zip(spec.partial, spec.sparse_matrices):
        s[0] = (pow(s[0], 5, R) + c) % R
        first = sum(a * b for a, b in zip(sp["row"], s)) % R
        rest = [(e * s[0] + x) % R for e, x in zip(sp["col_hat"], s[1:])]
        s = [first] + rest
    for consts in spec.end:
        s = sbox_full(consts)
        s = mat_vec(spec.mds, s)
    s = sbox_full([0] * t)
    s = mat_vec(spec.mds, s)
    return s


class PoseidonSponge:
    """Value-level PoseidonChip (update / squeeze, API/hash/poseidon.rs:172-194) = poseidon::Poseidon of the crate."""

    def __init__(self, spec):
        self.spec = spec
        self.state = state_default(spec.t)
        self.absorbing = []

    def update(self, elements):
        self.absorbing += [e % R for e in elements]

    def squeeze(self):
        rate = self.spec.t - 1
        inputs, self.absorbing = self.absorbing, []
        padding_offset = 0
        for i in range(0, len(inputs), rate):
            chunk = inputs[i:i + rate]
            padding_offset = rate - len(chunk)
            self.state = permute_optimised(self.spec, self.state, chunk)
        if padding_offset == 0:
            self.state = permute_optimised(self.spec, self.state, [])
        return self.state[1]


_SPECS = {}


def spec(t=9, r_f=8, r_p=63):
    key = (t, r_f, r_p)
    if key not in _SPECS:
        _SPECS[key] = Spec(t, r_f, r_p)
    return _SPECS[key]
