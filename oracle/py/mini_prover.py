"""TEST INFRASTRUCTURE (oracle) -- a plain-Python halo2 prover (KZG + GWC) for TINY circuits.

Role in the tests: the reference's `sample_circuit_random_run`
(halo2-snark-aggregator-circuit/src/sample_circuit.rs:56-124): it produces the INNER proof, written with the Poseidon
transcript, that `verify_aggregation_proofs_in_chip` (halo2-snark-aggregator-api/src/systems/halo2/verify.rs:835-942)
consumes -- here through oracle/py/verifier_ref.py.  halo2_proofs is an external crate (not under /root/reference,
Cargo.lock:1549-1551); `create_proof`'s order is restated from the order in which the reference's verifier READS a
proof back (verify.rs:342-478) and the equations it checks (params.rs, permutation.rs, lookup.rs, vanish.rs,
multiopen.rs), so a proof made here is accepted by the restated verifier iff both sides agree with those files.

The SRS is a toy one with a KNOWN trapdoor s (like ParamsKZG::unsafe_setup, which also draws s locally):
commit(f) = [f(s)] G needs one scalar multiplication, and the final pairing check e(w_x, [s]_2) = e(w_g, [1]_2) of
evaluate_multiopen_proof (verify.rs:733-740) becomes s * w_x == w_g in G1.

A constraint system is a dict:
  num_advice, num_fixed, num_instance, gates [expr], lookups [([input expr], [table expr])], perm_columns [(kind, idx)],
  advice_queries / fixed_queries / instance_queries [(column, rotation)] in halo2's REGISTRATION order, degree,
  blinding_factors                     (expressions are nested tuples as in quotient_ref.py)
"""
import bn254_ref as ref
import lookup_ref as lr
import poseidon_ref as pos
import quotient_ref as qr

R = qr.R
P = ref.P
G = ref.G1_GEN


# ---------------------------------------------------------------------------------------------------- formats
def point_to_bytes(pt):
    """halo2curves 0.2.1 compressed G1Affine: 32-byte LE x, parity of y in the top bit of byte 31;
    the identity is all zero"""
    if pt is None:
        return bytes(32)
    x, y = pt
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << 7
    return bytes(b)


def point_from_bytes(b):
    if b == bytes(32):
        return None
    sign = b[31] >> 7
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x7F]), "little")
    assert x < P, "invalid point encoding in proof"
    y2 = (x * x * x + 3) % P
    y = pow(y2, (P + 1) // 4, P)          # p = 3 mod 4
    assert y * y % P == y2, "invalid point encoding in proof"
    if (y & 1) != sign:
        y = P - y
    return (x, y)


class PoseidonWrite:
    """The inner proofs' transcript (sample_circuit.rs:74-83): value-level dual of PoseidonTranscriptRead
    (api/src/systems/halo2/transcript.rs:10-179) with PoseidonEncode (api/src/mock/transcript_encode.rs:28-74):
    a point is absorbed as (x mod r, y mod r), a scalar as itself; T = 9, RATE = 8, R_F = 8, R_P = 63."""

    def __init__(self):
        self.sponge = pos.PoseidonSponge(pos.spec(9, 8, 63))
        self.out = bytearray()

    def common_point(self, pt):
        x, y = (0, 0) if pt is None else pt
        self.sponge.update([x % R, y % R])

    def common_scalar(self, s):
        self.sponge.update([s % R])

    def write_point(self, pt):
        self.common_point(pt)
        self.out += point_to_bytes(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.out += (s % R).to_bytes(32, "little")

    def squeeze_challenge(self):
        return self.sponge.squeeze()


# ---------------------------------------------------------------------------------------------------- setup
class ToySrs:
    def __init__(self, k, s):
        self.k, self.n, self.s = k, 1 << k, s % R

    def commit(self, coeffs):
        return ref.g1_mul(qr.horner(coeffs, self.s), G)

    def g_lagrange(self, i):
        """L_i(s) G with L_i(X) = (X^n - 1) w^i / (n (X - w^i))"""
        w_i = pow(qr.omega(self.k), i, R)
        li = (pow(self.s, self.n, R) - 1) * w_i % R * pow(self.n * (self.s - w_i) % R, -1, R) % R
        return ref.g1_mul(li, G)


def extended_k(cs, k):
    q, e = cs["degree"] - 1, 0
    while (1 << e) < q:
        e += 1
    return k + e


def chunk_len(cs):
    return cs["degree"] - 2


def num_sets(cs):
    c = chunk_len(cs)
    return (len(cs["perm_columns"]) + c - 1) // c if cs["perm_columns"] else 0


def desc_of(cs):
    return dict(gates=cs["gates"], lookups=cs["lookups"], perm_columns=cs["perm_columns"], chunk_len=chunk_len(cs),
                last_rotation=-(cs["blinding_factors"] + 1))


def opening_queries(cs):
    """create_proof's opening queries in the order the reference's verifier rebuilds them (params.rs:156-224)."""
    out = [(("instance", c), r) for c, r in cs["instance_queries"]]
    out += [(("advice", c), r) for c, r in cs["advice_queries"]]
    last = -(cs["blinding_factors"] + 1)
    sets = num_sets(cs)
    for s in range(sets):
        out += [(("perm_z", s), 0), (("perm_z", s), 1)]
    for s in reversed(range(sets - 1)):
        out.append((("perm_z", s), last))
    for i in range(len(cs["lookups"])):
        out += [(("lookup_z", i), 0), (("lookup_input", i), 0), (("lookup_table", i), 0), (("lookup_input", i), -1), (("lookup_z", i), 1)]
    out += [(("fixed", c), r) for c, r in cs["fixed_queries"]]
    out += [(("sigma", j), 0) for j in range(len(cs["perm_columns"]))]
    out += [(("h", 0), 0), (("random", 0), 0)]
    return out


def eval_write_order(cs):
    """the order the evaluations are written in = read back (verify.rs:446-462, :198-230, :294-312)"""
    out = [(("instance", c), r) for c, r in cs["instance_queries"]]
    out += [(("advice", c), r) for c, r in cs["advice_queries"]]
    out += [(("fixed", c), r) for c, r in cs["fixed_queries"]]
    out.append((("random", 0), 0))
    out += [(("sigma", j), 0) for j in range(len(cs["perm_columns"]))]
    last = -(cs["blinding_factors"] + 1)
    sets = num_sets(cs)
    for s in range(sets):
        out += [(("perm_z", s), 0), (("perm_z", s), 1)]
        if s + 1 < sets:
            out.append((("perm_z", s), last))
    for i in range(len(cs["lookups"])):
        out += [(("lookup_z", i), 0), (("lookup_z", i), 1), (("lookup_input", i), 0), (("lookup_input", i), -1), (("lookup_table", i), 0)]
    return out


def keygen(cs, k, fixed_cols, copy_cycles, srs, transcript_repr=0x1234567):
    """-> (vk, pk).  copy_cycles: lists of (perm column index, row) cells that must be equal.
    transcript_repr stands for the hash of vk.pinned() (verify.rs:56-72): an external-crate format, not restated --
    prover and verifier only have to agree on the scalar."""
    n = 1 << k
    w = qr.omega(k)
    ident = [[pow(qr.DELTA, j, R) * pow(w, i, R) % R for i in range(n)] for j in range(len(cs["perm_columns"]))]
    sigma = [list(col) for col in ident]
    for cyc in copy_cycles:
        for (j, i), (j2, i2) in zip(cyc, cyc[1:] + cyc[:1]):
            sigma[j][i] = ident[j2][i2]
    bf = cs["blinding_factors"]
    last = n - bf - 1
    pk = dict(cs=cs, k=k, n=n, srs=srs, transcript_repr=transcript_repr % R,
              fixed=[list(c) for c in fixed_cols], sigma=sigma,
              l0=[1 if i == 0 else 0 for i in range(n)], l_last=[1 if i == last else 0 for i in range(n)],
              l_active=[1 if i < last else 0 for i in range(n)])
    vk = dict(cs=cs, k=k, n=n, omega=w, transcript_repr=transcript_repr % R,
              fixed_commitments=[srs.commit(ref.ifft(c, k)) for c in pk["fixed"]],
              permutation_commitments=[srs.commit(ref.ifft(c, k)) for c in sigma],
              g_lagrange=srs.g_lagrange)
    return vk, pk


def kate_division(a, z):
    """quotient of a(X) by (X - z), remainder dropped"""
    q = [0] * (len(a) - 1)
    carry = 0
    for i in reversed(range(1, len(a))):
        carry = (a[i] + carry * z) % R
        q[i - 1] = carry
    return q


def create_proof(pk, advice_cols, instance_cols, rng):
    """-> proof bytes (the instance values themselves are not part of the proof)"""
    cs, k, n, srs = pk["cs"], pk["k"], pk["n"], pk["srs"]
    bf = cs["blinding_factors"]
    u = n - bf - 1
    ext_k = extended_k(cs, k)
    w = qr.omega(k)
    t = PoseidonWrite()
    t.common_scalar(pk["transcript_repr"])
    lag, coeff = {}, {}

    def add(name, values):
        lag[name] = list(values)
        coeff[name] = ref.ifft(lag[name], k)
        return srs.commit(coeff[name])

    for i, c in enumerate(pk["fixed"]):
        add(("fixed", i), c)
    for j, c in enumerate(pk["sigma"]):
        add(("sigma", j), c)
    add(("l0", 0), pk["l0"]); add(("l_last", 0), pk["l_last"]); add(("l_active_row", 0), pk["l_active"])
    for i, c in enumerate(instance_cols):                       # absorbed, not written (verify.rs:74-92)
        t.common_point(add(("instance", i), list(c) + [0] * (n - len(c))))
    for i, c in enumerate(advice_cols):
        assert len(c) == u
        t.write_point(add(("advice", i), list(c) + [rng.randrange(R) for _ in range(n - u)]))
    theta = t.squeeze_challenge()
    comp = {}
    for i, (ins, tabs) in enumerate(cs["lookups"]):
        A = lr.compress_expressions(ins, lag, n, theta)
        S = lr.compress_expressions(tabs, lag, n, theta)
        comp[i] = (A, S)
        pa, ps = lr.permute_expression_pair(A[:u], S[:u])
        t.write_point(add(("lookup_input", i), pa + [rng.randrange(R) for _ in range(n - u)]))
        t.write_point(add(("lookup_table", i), ps + [rng.randrange(R) for _ in range(n - u)]))
    beta = t.squeeze_challenge()
    gamma = t.squeeze_challenge()
    ch = chunk_len(cs)
    last_z = 1
    for s in range(num_sets(cs)):
        cols = cs["perm_columns"][s * ch:(s + 1) * ch]
        z = lr.permutation_product([lag[c] for c in cols], [lag[("sigma", s * ch + j)] for j in range(len(cols))], k, w, beta, gamma,
                                   s * ch, last_z)
        last_z = z[u]
        t.write_point(add(("perm_z", s), z[:n - bf] + [rng.randrange(R) for _ in range(bf)]))
    for i in range(len(cs["lookups"])):
        z = lr.lookup_product(comp[i][0], comp[i][1], lag[("lookup_input", i)], lag[("lookup_table", i)], beta, gamma)
        assert z[u] == 1, "lookup argument does not close: witness outside the table?"
        t.write_point(add(("lookup_z", i), z[:n - bf] + [rng.randrange(R) for _ in range(bf)]))
    coeff[("random", 0)] = [rng.randrange(R) for _ in range(n)]
    t.write_point(srs.commit(coeff[("random", 0)]))
    y = t.squeeze_challenge()
    ext = {nm: qr.coeff_to_extended(c, k, ext_k) for nm, c in coeff.items() if nm != ("random", 0)}
    h = qr.divide_by_vanishing_poly(qr.evaluate_h(desc_of(cs), ext, k, ext_k, y, beta, gamma, theta), k, ext_k)
    q = cs["degree"] - 1
    h_coeff = ref.extended_to_coeff(h, ext_k, n * q)
    pieces = [h_coeff[i * n:(i + 1) * n] for i in range(q)]
    for p_ in pieces:
        t.write_point(srs.commit(p_))
    x = t.squeeze_challenge()
    xn = pow(x, n, R)
    coeff[("h", 0)] = [sum(pieces[i][j] * pow(xn, i, R) for i in range(q)) % R for j in range(n)]

    def ev(q_):
        nm, rot = q_
        return qr.horner(coeff[nm], x * pow(w, rot, R) % R)

    for q_ in eval_write_order(cs):
        t.write_scalar(ev(q_))
    v = t.squeeze_challenge()
    queries = opening_queries(cs)
    for rot in dict.fromkeys(r for _, r in queries):          # point sets in order of first appearance
        z = x * pow(w, rot, R) % R
        acc = [0] * n
        for nm, r2 in reversed(queries):                        # query i of a point is weighted v^i (multiopen.rs:55-61)
            if r2 == rot:
                acc = [(a * v + b) % R for a, b in zip(acc, coeff[nm])]
        t.write_point(srs.commit(kate_division(acc, z)))
    return bytes(t.out)
