"""TEST INFRASTRUCTURE (oracle) -- the reference's "verifier as a schema" restated over a chip interface.

The reference's aggregation circuit is the halo2 verifier written against three plugin traits
(halo2-snark-aggregator-api/src/arith/{common,ecc,field}.rs + transcript/encode.rs); the same code runs with plain
values (api/src/mock), with the circuit chips (circuit/src/chips) and with the Solidity generator.  This file restates
that CALLER, call for call, so that the op stream it emits -- and therefore the row layout of the witness -- is the one
the reference produces:

  PoseidonChip                      api/src/hash/poseidon.rs:6-231
  PoseidonTranscriptRead            api/src/systems/halo2/transcript.rs:10-179
  VerifierParamsBuilder.build_params   api/src/systems/halo2/verify.rs:56-571
  VerifierParams.queries            api/src/systems/halo2/params.rs:59-224 (+ lagrange.rs:17-39, expression.rs:18-114,
                                    permutation.rs:54-182, lookup.rs:34-165, vanish.rs:18-75)
  batch_multi_open_proofs           api/src/systems/halo2/multiopen.rs:23-102
  EvaluationQuerySchema.eval        api/src/systems/halo2/evaluation.rs:172-305
  assign_instance_commitment, verify_single_proof_no_eval, evaluate_multiopen_proof,
  verify_aggregation_proofs_in_chip    api/src/systems/halo2/verify.rs:574-942
  synthesize                        circuit/src/verify_circuit.rs:242-504 (what Halo2VerifierCircuits does around the chips)

Chips are duck-typed: the methods are the trait methods (same names, `ctx` dropped: a chip owns its context).
Three families are used by the tests:
  Mock*   plain values (api/src/mock/arith/{ecc,field}.rs, mock/transcript_encode.rs)   -- accepts / rejects a proof
  Ref*    circuit chips over oracle/py/ecc_chip_ref.py (circuit/src/chips/*.rs)        -- the witness oracle
  B200*   the product's recording chips (halo2_snark_aggregator_b200/witness.py)         -- the thing under test
Points cross the chip boundary as canonical affine tuples (x, y) or None for the identity; scalars as ints < r.
"""
import bn254_ref as ref
import ecc_chip_ref as E
import mini_prover as mp
import poseidon_ref as pos
import quotient_ref as qr

R = qr.R
DELTA = qr.DELTA


# ================================================================================================ Poseidon chip
class PoseidonChip:
    """api/src/hash/poseidon.rs: the sponge written against ArithFieldChip."""

    def __init__(self, chip, t=9, r_f=8, r_p=63):
        self.chip, self.spec, self.t = chip, pos.spec(t, r_f, r_p), t
        self.s = [chip.assign_const(x) for x in pos.state_default(t)]     # :151-157
        self.absorbing = []

    def update(self, elements):
        self.absorbing += list(elements)

    def squeeze(self):                                                      # :172-194
        rate = self.t - 1
        inputs, self.absorbing = self.absorbing, []
        padding_offset = 0
        for i in range(0, len(inputs), rate):
            chunk = inputs[i:i + rate]
            padding_offset = rate - len(chunk)
            self._permutation(chunk)
        if padding_offset == 0:
            self._permutation([])
        return self.s[1]

    def _x5c(self, x, c):                                                   # :10-19
        c2 = self.chip.mul(x, x)
        c4 = self.chip.mul(c2, c2)
        return self.chip.mul_add_constant(x, c4, c)

    def _sbox_full(self, consts):
        self.s = [self._x5c(x, c) for x, c in zip(self.s, consts)]

    def _apply_mds(self, mds):                                              # :90-112
        self.s = [self.chip.sum_with_coeff_and_constant(list(zip(self.s, row)), 0) for row in mds]

    def _apply_sparse(self, sp):                                            # :114-143
        res = [self.chip.sum_with_coeff_and_constant(list(zip(self.s, sp["row"])), 0)]
        for e, x in zip(sp["col_hat"], self.s[1:]):
            res.append(self.chip.sum_with_coeff_and_constant([(self.s[0], e), (x, 1)], 0))
        self.s = res

    def _permutation(self, inputs):                                         # :196-230
        spec, chip, t = self.spec, self.chip, self.t
        half = spec.r_f // 2
        pre = spec.start[0]
        assert len(inputs) < t
        off = len(inputs) + 1
        s = self.s                                                          # absorb_with_pre_constants :47-88
        s[0] = chip.sum_with_constant([s[0]], pre[0])
        for i, v in enumerate(inputs):
            s[i + 1] = chip.sum_with_constant([s[i + 1], v], pre[i + 1])
        for i in range(off, t):
            s[i] = chip.sum_with_constant([s[i]], (pre[i] + (1 if i == off else 0)) % R)
        for consts in spec.start[1:half]:
            self._sbox_full(consts)
            self._apply_mds(spec.mds)
        self._sbox_full(spec.start[-1])
        self._apply_mds(spec.pre_sparse_mds)
        for c, sp in zip(spec.partial, spec.sparse_matrices):
            self.s[0] = self._x5c(self.s[0], c)
            self._apply_sparse(sp)
        for consts in spec.end:
            self._sbox_full(consts)
            self._apply_mds(spec.mds)
        self._sbox_full([0] * t)
        self._apply_mds(spec.mds)


class PoseidonTranscriptRead:
    """api/src/systems/halo2/transcript.rs"""

    def __init__(self, data, nchip, r_f=8, r_p=63):
        self.hash = PoseidonChip(nchip, 9, r_f, r_p)
        self.data, self.pos = bytes(data), 0

    def _take(self, n):
        if self.pos + n > len(self.data):
            raise EOFError("transcript exhausted")
        b = self.data[self.pos:self.pos + n]
        self.pos += n
        return b

    def read_point(self, chips):
        pt = mp.point_from_bytes(self._take(32))
        p = chips.pchip.assign_var(pt)
        self.common_point(chips, p)
        return p

    def read_scalar(self, chips):
        v = int.from_bytes(self._take(32), "little")
        assert v < R, "invalid field element encoding in proof"
        s = chips.schip.assign_var(v)
        self.common_scalar(chips, s)
        return s

    def squeeze_challenge_scalar(self, chips):
        return chips.encode.decode_scalar([self.hash.squeeze()])

    def common_point(self, chips, p):
        self.hash.update(chips.encode.encode_point(p))

    def common_scalar(self, chips, s):
        self.hash.update(chips.encode.encode_scalar(s))


class Chips:
    def __init__(self, nchip, schip, pchip, encode):
        self.nchip, self.schip, self.pchip, self.encode = nchip, schip, pchip, encode


# ================================================================================================ chip families
class _FieldProvided:
    """provided methods of ArithFieldChip (api/src/arith/field.rs:37-104)"""

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

    def assign_zero(self):
        return self.assign_const(0)

    def assign_one(self):
        return self.assign_const(1)

    def normalize(self, v):
        return v


class MockFieldChip(_FieldProvided):
    """api/src/mock/arith/field.rs: plain Fr values"""

    def add(self, a, b): return (a + b) % R
    def sub(self, a, b): return (a - b) % R
    def mul(self, a, b): return a * b % R
    def square(self, a): return a * a % R
    def div(self, a, b): return a * pow(b, -1, R) % R
    def assign_const(self, c): return c % R
    def assign_var(self, v): return v % R
    def to_value(self, v): return v
    def sum_with_coeff_and_constant(self, a, b): return (sum(x * c for x, c in a) + b) % R
    def mul_add_constant(self, a, b, c): return (a * b + c) % R


class MockEccChip:
    """api/src/mock/arith/ecc.rs: plain G1 values (None = identity)"""

    def add(self, a, b): return ref.g1_add(a, b)
    def sub(self, a, b): return ref.g1_add(a, ref.g1_neg(b))
    def assign_zero(self): return None
    def assign_one(self): return ref.G1_GEN
    def assign_const(self, c): return c
    def assign_var(self, v): return v
    def to_value(self, v): return v
    def normalize(self, v): return v
    def scalar_mul(self, s, p): return ref.g1_mul(s, p)
    def scalar_mul_constant(self, s, p): return ref.g1_mul(s, p)

    def multi_exp(self, points, scalars):
        acc = None
        for p, s in zip(points, scalars):
            acc = ref.g1_add(acc, ref.g1_mul(s, p))
        return acc


class MockEncode:
    """api/src/mock/transcript_encode.rs:28-74"""

    def encode_point(self, p):
        x, y = (0, 0) if p is None else p
        return [x % R, y % R]

    def encode_scalar(self, s): return [s]
    def decode_scalar(self, v): return v[0]


def mock_chips():
    f = MockFieldChip()
    return Chips(f, f, MockEccChip(), MockEncode())


class RefScalarChip(_FieldProvided):
    """circuit/src/chips/scalar_chip.rs over the base-gate recipes of ecc_chip_ref.py"""

    def __init__(self, ctx):
        self.ctx = ctx

    def add(self, a, b): return E.bg_add(self.ctx, a, b)
    def sub(self, a, b): return E.sum_with_constant(self.ctx, [(a, 1), (b, -1)], 0)
    def assign_const(self, c): return E.bg_assign_constant(self.ctx, c % R)
    def assign_var(self, v): return E.bg_assign(self.ctx, v % R)
    def to_value(self, v): return v.value
    def mul(self, a, b): return E.bg_mul(self.ctx, a, b)
    def square(self, a): return E.bg_mul(self.ctx, a, a)
    def div(self, a, b): return E.bg_div_unsafe(self.ctx, a, b)
    def sum_with_coeff_and_constant(self, a, b): return E.sum_with_constant(self.ctx, [(x, c % R) for x, c in a], b % R)
    def mul_add_constant(self, a, b, c): return E.bg_mul_add_constant(self.ctx, a, b, c % R)
    def cell(self, v): return (v.col, v.row)


class RefEccChip:
    """circuit/src/chips/ecc_chip.rs:28-133 -- note which operands the adapter clones"""

    def __init__(self, ctx):
        self.ctx = ctx

    def add(self, a, b): return E.ecc_add(self.ctx, a.clone(), b.clone())
    def sub(self, a, b): return E.ecc_sub(self.ctx, a.clone(), b)
    def assign_zero(self): return E.assign_identity(self.ctx)
    def assign_one(self): return E.assign_constant_point(self.ctx, ref.G1_GEN)
    def assign_const(self, c): return E.assign_constant_point(self.ctx, c)
    def assign_var(self, v): return E.assign_point(self.ctx, v)
    def to_value(self, v): return None if v.z.value == 1 else (v.x.w(), v.y.w())
    def normalize(self, v): return E.ecc_reduce(self.ctx, v.clone())
    def scalar_mul(self, s, p): return E.ecc_mul(self.ctx, p.clone(), s)
    def scalar_mul_constant(self, s, p): return E.ecc_constant_mul(self.ctx, p, s, ref.g1_add)
    def multi_exp(self, points, scalars): return E.ecc_shamir(self.ctx, [p.clone() for p in points], scalars)
    # around the chips (verify_circuit.rs:264-368, 487-496)
    def assert_equal(self, a, b): E.ecc_assert_equal(self.ctx, a.clone(), b)
    def assert_not_identity(self, p): E.bg_assert_constant(self.ctx, p.z, 0)
    def expose_final_pair(self, w_x, w_g): return E.expose_final_pair(self.ctx, w_x, w_g)


class RefEncodeChip:
    """circuit/src/chips/encode_chip.rs:14-51"""

    def __init__(self, ctx):
        self.ctx = ctx

    def encode_point(self, p):
        px, py = E._clone_int(p.x), E._clone_int(p.y)
        return [E.native(self.ctx, px), E.native(self.ctx, py)]

    def encode_scalar(self, s): return [s]
    def decode_scalar(self, v): return v[0]


def ref_chips(ctx=None):
    ctx = ctx if ctx is not None else E.Context()
    f = RefScalarChip(ctx)
    return Chips(f, f, RefEccChip(ctx), RefEncodeChip(ctx)), ctx


# ================================================================================================ EvaluationQuerySchema
def _has_commitment(s):
    k = s[0]
    if k == "commit":
        return True
    if k in ("eval", "scalar"):
        return False
    return s[1][1] or s[2][1]


def q_commit(key, commitment, eval_=None): return ("commit", key, commitment, eval_)
def q_eval(key, commitment, eval_): return ("eval", key, commitment, eval_)
def q_scalar(s): return ("scalar", s)
def q_add(a, b): return ("add", (a, _has_commitment(a)), (b, _has_commitment(b)))
def q_mul(a, b): return ("mul", (a, _has_commitment(a)), (b, _has_commitment(b)))


def eval_prepare(s, chips, one, scalar):
    """evaluation.rs:211-305 -> [(key, point | None, scalar | None)]"""
    schip = chips.schip
    k = s[0]
    if k == "commit":
        return [[s[1], s[2], scalar]]
    if k == "eval":
        e = schip.mul(scalar, s[3]) if scalar is not None else s[3]
        return [["", None, e]]
    if k == "scalar":
        v = schip.mul(s[1], scalar) if scalar is not None else s[1]
        return [["", None, v]]
    l, r = s[1], s[2]
    if k == "add":
        if not l[1] and not r[1]:
            lv = eval_prepare(l[0], chips, one, None)
            rv = eval_prepare(r[0], chips, one, None)
            assert len(lv) == 1 and len(rv) == 1
            total = schip.add(lv[0][2], rv[0][2])
            if scalar is not None:
                total = schip.mul(scalar, total)
            return [["", None, total]]
        res = []
        for side in (l, r):
            for ev in eval_prepare(side[0], chips, one, scalar):
                found = next((p for p in res if p[0] == ev[0]), None)
                if found is not None:
                    found[2] = schip.add(found[2] if found[2] is not None else one, ev[2] if ev[2] is not None else one)
                else:
                    res.append(ev)
        return res
    # mul
    if not l[1]:
        sv, rem = eval_prepare(l[0], chips, one, None), r[0]
    else:
        sv, rem = eval_prepare(r[0], chips, one, None), l[0]
    assert len(sv) == 1
    v = sv[0][2]
    if scalar is not None:
        v = schip.mul(scalar, v)
    return eval_prepare(rem, chips, one, v)


def schema_eval(s, chips, one):
    """evaluation.rs:172-209 -> (point, scalar | None, point names)"""
    points = eval_prepare(s, chips, one, None)
    names = [p[0] for p in points]
    sc = next((p[2] for p in points if p[0] == ""), None)
    p_wo_scalar = [p[1] for p in points if p[2] is None and p[1] is not None]
    pl = [(p[1], p[2]) for p in points if p[1] is not None and p[2] is not None]
    acc = chips.pchip.multi_exp([a for a, _ in pl], [b for _, b in pl])
    for p in p_wo_scalar:
        acc = chips.pchip.add(acc, p)
    return acc, sc, names


# ================================================================================================ expressions
def convert_expression(e, schip):
    """verify.rs:168-196: constants become assigned cells (in traversal order)"""
    k = e[0]
    if k == "const":
        return ("const", schip.assign_const(e[1]))
    if k in ("fixed", "advice", "instance"):
        return e
    if k == "neg":
        return ("neg", convert_expression(e[1], schip))
    if k == "scaled":
        inner = convert_expression(e[1], schip)
        return ("scaled", inner, schip.assign_const(e[2]))
    return (k, convert_expression(e[1], schip), convert_expression(e[2], schip))


def chip_evaluate(e, schip, query, zero):
    """expression.rs:18-114; query(kind, col, rot) -> assigned eval"""
    k = e[0]
    if k == "const":
        return e[1]
    if k in ("fixed", "advice", "instance"):
        return query(k, e[1], e[2])
    if k == "neg":
        a = chip_evaluate(e[1], schip, query, zero)
        return schip.sub(zero, a)
    if k == "scaled":
        a = chip_evaluate(e[1], schip, query, zero)
        return schip.mul(e[2], a)
    a = chip_evaluate(e[1], schip, query, zero)
    b = chip_evaluate(e[2], schip, query, zero)
    return schip.add(a, b) if k == "sum" else schip.mul(a, b)


# ================================================================================================ build_params + queries
class VerifierParams:
    pass


def _rotate_omega(schip, x, omega, at):
    base, exp = (pow(omega, -1, R), -at) if at < 0 else (omega, at)
    return schip.sum_with_coeff_and_constant([(x, pow(base, exp, R))], 0)


def build_params(chips, vk, assigned_instances, transcript, key):
    """verify.rs:342-571 for ONE proof of `vk` (the reference's multi-proof handling is marked FIXME there)"""
    schip, pchip = chips.schip, chips.pchip
    cs = vk["cs"]
    t = transcript
    sq = lambda: t.squeeze_challenge_scalar(chips)
    pt = lambda: t.read_point(chips)
    sc = lambda: t.read_scalar(chips)

    t.common_scalar(chips, schip.assign_const(vk["transcript_repr"]))         # init_transcript :56-72
    for p in assigned_instances:                                              # squeeze_instance_commitment :74-92
        t.common_point(chips, p)
    schip.assign_const(0)                                                     # assigned_scalar_zero :355-356
    advice_commitments = [pt() for _ in range(cs["num_advice"])]
    theta = sq()
    lookups_permuted = [(pt(), pt()) for _ in cs["lookups"]]
    beta, gamma = sq(), sq()
    n_sets = mp.num_sets(cs)
    permutations_committed = [pt() for _ in range(n_sets)]
    lookups_committed = [pt() for _ in cs["lookups"]]
    random_commitment = pt()
    y = sq()
    h_commitments = [pt() for _ in range(cs["degree"] - 1)]
    l = cs["blinding_factors"] + 1
    n = vk["n"]
    omega = vk["omega"]
    x = sq()
    instance_evals = [sc() for _ in cs["instance_queries"]]
    advice_evals = [sc() for _ in cs["advice_queries"]]
    fixed_evals = [sc() for _ in cs["fixed_queries"]]
    random_eval = sc()
    permutation_evals = [sc() for _ in vk["permutation_commitments"]]
    # build_permutation_evaluated :198-289
    sets = []
    for i, c in enumerate(permutations_committed):
        ev, nx = sc(), sc()
        last = sc() if i + 1 < len(permutations_committed) else None
        sets.append(dict(commitment=c, eval=ev, next_eval=nx, last_eval=last))
    qlists = {"advice": (cs["advice_queries"], advice_evals), "fixed": (cs["fixed_queries"], fixed_evals),
              "instance": (cs["instance_queries"], instance_evals)}
    perm_column_evals = [qlists[kind][1][qlists[kind][0].index((idx, 0))] for kind, idx in cs["perm_columns"]]
    # build_lookup_evaluated :291-340
    lookups = []
    for j, ((pin, ptab), zc, (ins, tabs)) in enumerate(zip(lookups_permuted, lookups_committed, cs["lookups"])):
        product_eval, product_next_eval, a_eval, a_inv_eval, s_eval = sc(), sc(), sc(), sc(), sc()
        lookups.append(dict(inputs=[convert_expression(e, schip) for e in ins], tables=[convert_expression(e, schip) for e in tabs],
                            permuted_input=pin, permuted_table=ptab, product=zc, product_eval=product_eval,
                            product_next_eval=product_next_eval, permuted_input_eval=a_eval, permuted_input_inv_eval=a_inv_eval,
                            permuted_table_eval=s_eval, key="%s_0_%d" % (key, j)))
    fixed_commitments = [pchip.assign_const(c) for c in vk["fixed_commitments"]]
    v = sq()
    w = []
    while True:                                                               # :471-474 read points until the proof ends
        try:
            w.append(pt())
        except EOFError:
            break
    u = sq()
    p = VerifierParams()
    p.x_next = _rotate_omega(schip, x, omega, 1)
    p.x_last = _rotate_omega(schip, x, omega, -l)
    p.x_inv = _rotate_omega(schip, x, omega, -1)
    p.xn = schip.pow_constant(x, n)
    p.key = key
    p.cs, p.l, p.n_rows = cs, l, n
    p.gates = [convert_expression(g, schip) for g in cs["gates"]]             # struct fields in source order :486-570
    p.lookups, p.sets, p.perm_column_evals = lookups, sets, perm_column_evals
    p.instance_commitments, p.instance_evals = assigned_instances, instance_evals
    p.advice_commitments, p.advice_evals = advice_commitments, advice_evals
    p.fixed_commitments, p.fixed_evals = fixed_commitments, fixed_evals
    p.permutation_commitments = [pchip.assign_const(c) for c in vk["permutation_commitments"]]
    p.permutation_evals = permutation_evals
    p.vanish_commitments, p.random_commitment, p.random_eval = h_commitments, random_commitment, random_eval
    p.beta, p.gamma, p.theta = beta, gamma, theta
    p.delta = schip.assign_const(DELTA)
    p.x, p.y, p.u, p.v = x, y, u, v
    p.omega_value = omega
    p.omega = schip.assign_const(omega)
    p.w = w
    p.zero = schip.assign_const(0)
    p.one = schip.assign_const(1)
    p.n = schip.assign_const(n)
    return p


def _lagrange_commits(p, schip):
    """lagrange.rs:17-39"""
    ws = [p.one]
    for i in range(1, p.l + 1):
        ws.append(schip.div(ws[i - 1], p.omega))
    out = []
    for wi in ws:
        a = schip.div(wi, p.n)
        b = schip.sub(p.xn, p.one)
        ab = schip.mul(a, b)
        c = schip.sub(p.x, wi)
        out.append(schip.div(ab, c))
    return out


def _permutation_expressions(p, schip, l_0, l_last, l_blind):
    """permutation.rs:54-136"""
    res = []
    sets, one, beta, gamma, delta, x = p.sets, p.one, p.beta, p.gamma, p.delta, p.x
    chunk = mp.chunk_len(p.cs)
    if sets:
        z_x = sets[0]["eval"]
        res.append(schip.mul(l_0, schip.sub(one, z_x)))
        z_x = sets[-1]["eval"]
        res.append(schip.mul(l_last, schip.sub(schip.mul(z_x, z_x), z_x)))
    for s_, prev in zip(sets[1:], sets):
        res.append(schip.mul(schip.sub(s_["eval"], prev["last_eval"]), l_0))
    t0 = schip.mul(beta, x)
    t1 = schip.sub(one, schip.add(l_last, l_blind))
    for ci, st in enumerate(sets):
        evals = p.perm_column_evals[ci * chunk:(ci + 1) * chunk]
        pevals = p.permutation_evals[ci * chunk:(ci + 1) * chunk]
        left, right = st["next_eval"], st["eval"]
        delta_pow = one if ci == 0 else schip.pow_constant(delta, ci * chunk)
        d = schip.mul(t0, delta_pow)
        for ev, pe in zip(evals, pevals):
            t2 = schip.add(ev, gamma)
            left = schip.mul(schip.add(t2, schip.mul(beta, pe)), left)
            right = schip.mul(schip.add(t2, d), right)
            d = schip.mul(delta, d)
        res.append(schip.mul(schip.sub(left, right), t1))
    return res


def _lookup_expressions(p, lk, schip, query, l_0, l_last, l_blind):
    """lookup.rs:34-118"""
    one, zero, beta, gamma, theta = p.one, p.zero, p.beta, p.gamma, p.theta
    z_wx, z_x = lk["product_next_eval"], lk["product_eval"]
    a_x, s_x, a_invwx = lk["permuted_input_eval"], lk["permuted_table_eval"], lk["permuted_input_inv_eval"]
    left = schip.mul(schip.mul(z_wx, schip.add(a_x, beta)), schip.add(s_x, gamma))
    input_evals = [chip_evaluate(e, schip, query, zero) for e in lk["inputs"]]
    input_eval = schip.mul_add_accumulate(input_evals, theta)
    table_evals = [chip_evaluate(e, schip, query, zero) for e in lk["tables"]]
    table_eval = schip.mul_add_accumulate(table_evals, theta)
    t0 = schip.sub(one, schip.add(l_last, l_blind))
    t1 = schip.sub(a_x, s_x)
    e1 = schip.mul(l_0, schip.sub(one, z_x))
    e2 = schip.mul(l_last, schip.sub(schip.mul(z_x, z_x), z_x))
    # ((left - ((product_eval * (input_eval + beta)) * (table_eval + gamma))) * t0): operands evaluated left to right
    inner = schip.mul(schip.mul(z_x, schip.add(input_eval, beta)), schip.add(table_eval, gamma))
    e3 = schip.mul(schip.sub(left, inner), t0)
    e4 = schip.mul(l_0, t1)
    e5 = schip.mul(schip.mul(t1, schip.sub(a_x, a_invwx)), t0)
    return [e1, e2, e3, e4, e5]


def _eq(rotation, key, point, commitment, eval_):
    """EvaluationQuery::new (evaluation.rs:108-128): Commitment(s) + Eval(s)"""
    return dict(rotation=rotation, point=point, s=q_add(q_commit(key, commitment, eval_), q_eval(key, commitment, eval_)))


def queries(p, chips):
    """params.rs:74-224"""
    schip = chips.schip
    ls = _lagrange_commits(p, schip)
    l_0, l_last = ls[0], ls[p.l]
    l_blind = schip.sum_with_constant(ls[1:p.l], 0)
    cs = p.cs
    qidx = {"advice": cs["advice_queries"], "fixed": cs["fixed_queries"], "instance": cs["instance_queries"]}
    evs = {"advice": p.advice_evals, "fixed": p.fixed_evals, "instance": p.instance_evals}

    def query(kind, col, rot):
        return evs[kind][qidx[kind].index((col, rot))]

    expression = [chip_evaluate(g, schip, query, p.zero) for g in p.gates]
    expression += _permutation_expressions(p, schip, l_0, l_last, l_blind)
    for lk in p.lookups:
        expression += _lookup_expressions(p, lk, schip, query, l_0, l_last, l_blind)
    out = []
    rot = lambda at: _rotate_omega(schip, p.x, p.omega_value, at)              # x_rotate_omega :59-72: a row per query
    for qi, (col, at) in enumerate(cs["instance_queries"]):
        out.append(_eq(at, "%s_instance_commitments%d" % (p.key, col), rot(at), p.instance_commitments[col], p.instance_evals[qi]))
    for qi, (col, at) in enumerate(cs["advice_queries"]):
        out.append(_eq(at, "%s_advice_commitments%d" % (p.key, col), rot(at), p.advice_commitments[col], p.advice_evals[qi]))
    pkey = "%s_0" % p.key                                                      # permutation.rs:138-181
    for i, st in enumerate(p.sets):
        k_ = "%s_permutation_product_commitment_%d" % (pkey, i)
        out.append(_eq(0, k_, p.x, st["commitment"], st["eval"]))
        out.append(_eq(1, k_, p.x_next, st["commitment"], st["next_eval"]))
    for i in reversed(range(len(p.sets) - 1)):
        st = p.sets[i]
        out.append(_eq(-p.l, "%s_permutation_product_commitment_%d" % (pkey, i), p.x_last, st["commitment"], st["last_eval"]))
    for lk in p.lookups:                                                       # lookup.rs:121-165
        k_ = lk["key"]
        out.append(_eq(0, k_ + "_product_commitment", p.x, lk["product"], lk["product_eval"]))
        out.append(_eq(0, k_ + "_permuted_input_commitment", p.x, lk["permuted_input"], lk["permuted_input_eval"]))
        out.append(_eq(0, k_ + "_permuted_table_commitment", p.x, lk["permuted_table"], lk["permuted_table_eval"]))
        out.append(_eq(-1, k_ + "_permuted_input_commitment", p.x_inv, lk["permuted_input"], lk["permuted_input_inv_eval"]))
        out.append(_eq(1, k_ + "_product_commitment", p.x_next, lk["product"], lk["product_next_eval"]))
    for qi, (col, at) in enumerate(cs["fixed_queries"]):
        out.append(_eq(at, "%s_fixed_commitments%d" % (p.key, col), rot(at), p.fixed_commitments[col], p.fixed_evals[qi]))
    for i, (c, e) in enumerate(zip(p.permutation_commitments, p.permutation_evals)):   # permutation.rs:34-51
        out.append(_eq(0, "%s_permutation_commitments%d" % (p.key, i), p.x, c, e))
    # vanish.rs:18-75
    expected_h = schip.mul_add_accumulate(expression, p.y)
    expected_h = schip.div(expected_h, schip.sub(p.xn, p.one))
    h_commitment = None
    for i, c in enumerate(reversed(p.vanish_commitments)):
        term = q_commit("%s_h_commitment%d" % (p.key, i), c, None)
        h_commitment = term if h_commitment is None else q_add(q_mul(q_scalar(p.xn), h_commitment), term)
    out.append(dict(rotation=0, point=p.x, s=q_add(h_commitment, q_scalar(expected_h))))
    out.append(_eq(0, "%s_random_commitment" % p.key, p.x, p.random_commitment, p.random_eval))
    return out


def batch_multi_open_proofs(p, chips):
    """multiopen.rs:23-102 -> (w_x, w_g) schemas"""
    qs = queries(p, chips)
    points = []          # [(rotation, point, [schemas])] in order of first appearance
    for q in qs:
        hit = next((e for e in points if e[0] == q["rotation"]), None)
        if hit is not None:
            hit[2].append(q["s"])
        else:
            points.append((q["rotation"], q["point"], [q["s"]]))
    assert len(p.w) == len(points), "W commitments vs opening points"
    proofs = []
    for i, (_, point, schemas) in enumerate(points):
        acc = None
        for s in reversed(schemas):
            acc = s if acc is None else q_add(q_mul(q_scalar(p.v), acc), s)
        proofs.append((acc, point, p.w[i]))
    w_x = w_g = None
    for i in reversed(range(len(proofs))):
        s, point, w = proofs[i]
        wq = q_commit("%s_w%d" % (p.key, i), w, None)
        w_x = wq if w_x is None else q_add(q_mul(q_scalar(p.u), w_x), wq)
        if w_g is None:
            w_g = q_add(q_mul(q_scalar(point), wq), s)
        else:
            w_g = q_add(q_add(q_mul(q_scalar(p.u), w_g), q_mul(q_scalar(point), wq)), s)
    return w_x, w_g


# ================================================================================================ verify.rs drivers
def assign_instance_commitment(chips, instances, vk):
    """verify.rs:574-649.  instances: one list of values per instance column"""
    schip, pchip = chips.schip, chips.pchip
    plain, assigned = [], []
    for col in instances:
        a = [schip.assign_var(v) for v in col]
        plain += a
        assigned.append(a)
    commitments = []
    for col in assigned:
        acc = None
        for i, inst in enumerate(col):
            ls = pchip.scalar_mul_constant(inst, vk["g_lagrange"](i))
            acc = ls if acc is None else pchip.add(acc, ls)
        commitments.append(pchip.assign_const(None) if acc is None else pchip.normalize(acc))
    return plain, commitments


def verify_single_proof_no_eval(chips, assigned_instances, vk, transcript, key):
    """verify.rs:651-688"""
    p = build_params(chips, vk, assigned_instances, transcript, key)
    return batch_multi_open_proofs(p, chips), list(p.advice_commitments)


def evaluate_multiopen_proof(chips, w_x, w_g):
    """verify.rs:690-745 (the pairing check itself is the caller's: it needs the verifier's G2 elements)"""
    schip, pchip = chips.schip, chips.pchip
    one = schip.assign_one()
    left_s, left_e, names_x = schema_eval(w_x, chips, one)
    right_s, right_e, names_g = schema_eval(w_g, chips, one)
    generator = pchip.assign_one()
    left = left_s if left_e is None else pchip.add(left_s, pchip.scalar_mul(left_e, generator))
    right = right_s if right_e is None else pchip.sub(right_s, pchip.scalar_mul(right_e, generator))
    return left, right, names_x + names_g


def verify_aggregation_proofs_in_chip(chips, circuits, transcript):
    """verify.rs:835-942.  circuits: [dict(name, vk, proofs=[dict(instances, transcript, key)])]
    -> (w_x, w_g, plain assigned instances, advice commitments per proof, multi_exp point names)"""
    plain = []
    proofs = []
    for c in circuits:
        for pr in c["proofs"]:
            assigned, commitments = assign_instance_commitment(chips, pr["instances"], c["vk"])
            plain += assigned
            proofs.append(verify_single_proof_no_eval(chips, commitments, c["vk"], pr["transcript"], pr["key"]))
        for pr in c["proofs"]:                                   # update aggregation challenge :910-914
            s = pr["transcript"].squeeze_challenge_scalar(chips)
            transcript.common_scalar(chips, s)
    ch = transcript.squeeze_challenge_scalar(chips)
    acc, commits = None, []
    for (w_x, w_g), c in proofs:
        acc = (w_x, w_g) if acc is None else (q_add(q_mul(acc[0], q_scalar(ch)), w_x), q_add(q_mul(acc[1], q_scalar(ch)), w_g))
        commits.append(c)
    left, right, names = evaluate_multiopen_proof(chips, acc[0], acc[1])
    return left, right, plain, commits, names


def synthesize(chips, circuits_data, coherent=()):
    """Halo2VerifierCircuits::synthesize (verify_circuit.rs:242-371) + synthesize_proof (:380-504).
    circuits_data: [dict(name, vk, nproofs, proofs=[dict(instances, transcript_bytes)])]
    -> dict(w_x, w_g, instance_cells: the AssignedValues bound to the instance column rows 0.., names)"""
    schip, pchip = chips.schip, chips.pchip
    schip.assign_const(1)                                    # in_shape_mode probe: one_line([(1, -1)], 1)  five/base_gate.rs:16-25
    circuits = []
    for ci, c in enumerate(circuits_data):                   # transcripts are created first, per proof (:436-449)
        proofs = [dict(instances=pr["instances"], transcript=PoseidonTranscriptRead(pr["transcript_bytes"], chips.nchip),
                       key="%s_p%d" % (c["name"], i)) for i, pr in enumerate(c["proofs"])]
        circuits.append(dict(name=c["name"], vk=c["vk"], proofs=proofs))
    transcript = PoseidonTranscriptRead(b"", chips.nchip)    # :470-477
    w_x, w_g, plain, commits, names = verify_aggregation_proofs_in_chip(chips, circuits, transcript)
    for a, b in coherent:                                    # :487-493
        pchip.assert_equal(commits[a[0]][a[1]], commits[b[0]][b[1]])
    pchip.assert_not_identity(w_x)                           # :495-496
    pchip.assert_not_identity(w_g)
    cells = pchip.expose_final_pair(w_x, w_g)                # second region :264-344
    return dict(w_x=w_x, w_g=w_g, instance_cells=list(cells) + list(plain), names=names)
