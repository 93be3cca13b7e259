"""Resident prover replay (halo2_snark_aggregator_b200/prover.py) against the CPU oracles, stage by stage:
commit rounds (K1-K3) -> quotient (N1 + K3 + K1) -> evaluation round (N2) -> GWC multi-opening (fold + N2 + K1).
Every polynomial stays in HBM between the stages; what comes back (commitments, evaluations) must be bit-identical
to the composition of the oracle's best_multiexp / ifft / coeff_to_extended / evaluate_h / extended_to_coeff /
eval_polynomial / kate_division on the same inputs."""
import random

import numpy as np
import pytest

import oracle_binding as ob
import quotient_util as qu
import quotient_ref as qr
from halo2_snark_aggregator_b200 import plonk
from halo2_snark_aggregator_b200.prover import ResidentProver, create_proof_queries, transcript_eval_order
from util import domain_consts

pytestmark = pytest.mark.gpu
R = qr.R


def _small_cs():
    E = plonk.Expression
    cs = plonk.ConstraintSystem(num_fixed=3, num_advice=3, num_instance=1)
    cs.create_gate("g0", [E.fixed(0) * (E.advice(0) * E.advice(1) - E.advice(2, 1)),
                          E.fixed(1) * (E.advice(0) + 5) * (E.advice(0, -1) - E.constant(3))])
    cs.lookup("l", [(E.advice(0) * E.fixed(0), E.fixed(2))])
    cs.enable_equality("advice", 0)
    cs.enable_equality("advice", 1)
    cs.enable_equality("instance", 0)
    return cs


@pytest.mark.parametrize("which,k", [("small", 5), ("aggregation", 6)])
def test_resident_prover_matches_oracle_composition(ctx, which, k):
    cs = _small_cs() if which == "small" else plonk.aggregation_circuit_cs()
    n = 1 << k
    g = ob.gen_bases(0x6000 + k, n)
    gl = ob.gen_bases(0x7000 + k, n)
    sid_g, sid_gl = ctx.srs_register(g), ctx.srs_register(gl)
    pr = ResidentProver(ctx, cs, k, sid_gl, sid_g)
    ext_k = pr.ext_k
    d = domain_consts(k, ext_k)
    lag = qu.random_lagrange_columns(pr.plan, k, seed=77 + k)
    lag[("random", 0)] = [random.Random(5).randrange(R) for _ in range(n)]
    names = list(pr.plan.columns)
    # --- commit rounds: everything the quotient touches, in three rounds like the prover's phases; "random" has no extended form
    thirds = [names[0::3], names[1::3], names[2::3]]
    o_coeff, o_ext = {}, {}
    for rnd in thirds:
        got = pr.commit_columns(rnd, [qu.pack(lag[nm]) for nm in rnd])
        for j, nm in enumerate(rnd):
            col = qu.pack(lag[nm])
            want = ob.best_multiexp(col, gl)
            assert np.array_equal(got[j], want[:8]) or (not want[8:].any() and not got[j].any()), nm
            o_coeff[nm] = ob.ifft(col.copy(), d["omega_inv"], d["n_inv"], k)
            o_ext[nm] = ob.coeff_to_extended(o_coeff[nm], k, ext_k, d["zeta"], d["omega_ext"])
            assert np.array_equal(ctx.d2h(pr.coeff[nm], 4 * n), o_coeff[nm]), nm
            assert np.array_equal(ctx.d2h(pr.ext[nm], 4 << ext_k), o_ext[nm]), nm
    pr.commit_columns([("random", 0)], [qu.pack(lag[("random", 0)])], extended=False)
    o_coeff[("random", 0)] = ob.ifft(qu.pack(lag[("random", 0)]), d["omega_inv"], d["n_inv"], k)
    # --- quotient
    rng = random.Random(k)
    y, beta, gamma, theta, x, v = [rng.randrange(R) for _ in range(6)]
    h_comms = pr.quotient(y, beta, gamma, theta)
    ext_ints = {nm: qu.unpack(o_ext[nm]) for nm in names}
    h = qr.divide_by_vanishing_poly(qr.evaluate_h(qu.oracle_desc(cs), ext_ints, k, ext_k, y, beta, gamma, theta), k, ext_k)
    q = cs.degree() - 1
    h_coeff = ob.extended_to_coeff(qu.pack(h), ext_k, d["omega_ext_inv"], d["ext_n_inv"], d["zeta"], n * q)
    assert len(h_comms) == q
    for i in range(q):
        piece = np.ascontiguousarray(h_coeff[4 * n * i: 4 * n * (i + 1)])
        o_coeff[("h_piece", i)] = piece
        assert np.array_equal(h_comms[i], ob.best_multiexp(piece, g)[:8]), "h piece %d" % i
    # --- evaluation round
    pr.fold_h(x)
    xn = pow(x, n, R)
    hp = [qu.unpack(o_coeff[("h_piece", i)]) for i in range(q)]
    folded = [sum(hp[i][j] * pow(xn, i, R) for i in range(q)) % R for j in range(n)]
    o_coeff[("h", 0)] = qu.pack(folded)
    queries = create_proof_queries(cs)
    evals = pr.evaluate(queries, x)
    w = qr.omega(k)
    for (nm, rot), e in zip(queries, evals):
        pt = x * pow(w, rot, R) % R
        assert np.array_equal(e, ob.eval_polynomial(o_coeff[nm], qu.pack([pt]))), (nm, rot)
    # --- GWC multi-opening
    order, ws = pr.open(queries, x, v)
    assert order == list(dict.fromkeys(rot for _, rot in queries))
    for rot, wpt in zip(order, ws):
        acc = [0] * n
        for nm, r2 in reversed(queries):      # query i of a point is weighted v^i
            if r2 == rot:
                c = qu.unpack(o_coeff[nm])
                acc = [(a * v + b) % R for a, b in zip(acc, c)]
        quo = ob.kate_division(qu.pack(acc), qu.pack([x * pow(w, rot, R) % R]))
        quo = np.concatenate([quo, np.zeros(4, dtype=np.uint64)])
        assert np.array_equal(wpt, ob.best_multiexp(quo, g)[:8]), "W at rotation %d" % rot
    pr.close()
    ctx.srs_release(sid_g)
    ctx.srs_release(sid_gl)


def test_resident_prover_at_k16_matches_the_cpp_oracle_composition(ctx):
    """The same stage-by-stage check for the aggregation circuit's constraint system at k = 16 (extended domain 2^18:
    multi-pass NTT plans, table-mode MSM with c = 14, a 55-column evaluate_h over 262 144 coset rows), with the C++
    oracle on packed arrays throughout: best_multiexp / ifft / coeff_to_extended / the plan-driven evaluate_h (itself
    equal to the big-int restatement, tests/test_quotient_cpu.py) / extended_to_coeff / eval_polynomial / kate_division."""
    cs = plonk.aggregation_circuit_cs()
    k = 16
    n = 1 << k
    g = ob.gen_bases(0x6100 + k, n)
    gl = ob.gen_bases(0x7100 + k, n)
    sid_g, sid_gl = ctx.srs_register(g), ctx.srs_register(gl)
    pr = ResidentProver(ctx, cs, k, sid_gl, sid_g)
    ext_k = pr.ext_k
    d = domain_consts(k, ext_k)
    names = list(pr.plan.columns)
    lag = {nm: ob.gen_scalars(0x9000 + i, 0, n) for i, nm in enumerate(names)}
    l0, l_last, l_active = qu.lagrange_selectors(k, cs.blinding_factors())
    lag[("l0", 0)], lag[("l_last", 0)], lag[("l_active_row", 0)] = qu.pack(l0), qu.pack(l_last), qu.pack(l_active)
    lag[("random", 0)] = ob.gen_scalars(0x9100, 0, n)
    o_coeff, o_ext = {}, {}
    for rnd in (names[0::3], names[1::3], names[2::3]):
        got = pr.commit_columns(rnd, [lag[nm] for nm in rnd])
        for j, nm in enumerate(rnd):
            want = ob.best_multiexp(lag[nm], gl)
            assert np.array_equal(got[j], want[:8]), nm
            o_coeff[nm] = ob.ifft(lag[nm].copy(), d["omega_inv"], d["n_inv"], k)
            o_ext[nm] = ob.coeff_to_extended(o_coeff[nm], k, ext_k, d["zeta"], d["omega_ext"])
            assert np.array_equal(ctx.d2h(pr.coeff[nm], 4 * n), o_coeff[nm]), nm
            assert np.array_equal(ctx.d2h(pr.ext[nm], 4 << ext_k), o_ext[nm]), nm
    pr.commit_columns([("random", 0)], [lag[("random", 0)]], extended=False)
    o_coeff[("random", 0)] = ob.ifft(lag[("random", 0)].copy(), d["omega_inv"], d["n_inv"], k)
    rng = random.Random(k)
    y, beta, gamma, theta, x, v = [rng.randrange(R) for _ in range(6)]
    fr = plonk.fr_mont
    h_comms = pr.quotient(y, beta, gamma, theta)
    t_ev = np.concatenate([fr(t) for t in plonk.t_evaluations(k, ext_k)])
    h = ob.evaluate_h(pr.plan.words, pr.plan.consts, [o_ext[nm] for nm in names], k, ext_k, fr(y), fr(beta), fr(gamma), fr(theta),
                      d["omega_ext"], d["zeta"], fr(plonk.DELTA), t_ev)
    q = cs.degree() - 1
    h_coeff = ob.extended_to_coeff(h, ext_k, d["omega_ext_inv"], d["ext_n_inv"], d["zeta"], n * q)
    for i in range(q):
        piece = np.ascontiguousarray(h_coeff[4 * n * i: 4 * n * (i + 1)])
        o_coeff[("h_piece", i)] = piece
        assert np.array_equal(h_comms[i], ob.best_multiexp(piece, g)[:8]), "h piece %d" % i
        assert np.array_equal(ctx.d2h(pr.coeff[("h_piece", i)], 4 * n), piece), "h piece %d coefficients" % i
    pr.fold_h(x)
    xn = pow(x, n, R)

    def scale_add(acc, c, poly):     # acc * c + poly on packed arrays through the oracle's field ops
        return ob.field_op(0, 0, ob.field_op(0, 3, acc, np.tile(fr(c), n)), poly)

    folded = np.zeros(4 * n, dtype=np.uint64)
    for i in reversed(range(q)):
        folded = scale_add(folded, xn, o_coeff[("h_piece", i)])
    o_coeff[("h", 0)] = folded
    queries = create_proof_queries(cs)
    evals = pr.evaluate(queries, x)
    w = qr.omega(k)
    for (nm, rot), e in zip(queries, evals):
        assert np.array_equal(e, ob.eval_polynomial(o_coeff[nm], fr(x * pow(w, rot, R) % R))), (nm, rot)
    order, ws = pr.open(queries, x, v)
    for rot, wpt in zip(order, ws):
        acc = np.zeros(4 * n, dtype=np.uint64)
        for nm, r2 in reversed(queries):      # query i of a point is weighted v^i
            if r2 == rot:
                acc = scale_add(acc, v, o_coeff[nm])
        quo = np.concatenate([ob.kate_division(acc, fr(x * pow(w, rot, R) % R)), np.zeros(4, dtype=np.uint64)])
        assert np.array_equal(wpt, ob.best_multiexp(quo, g)[:8]), "W at rotation %d" % rot
    pr.close()
    ctx.srs_release(sid_g)
    ctx.srs_release(sid_gl)


def _satisfiable_columns(cs, which, k, seed):
    """Lagrange columns whose lookups can be satisfied: range tables i mod 16, 0/1 selectors, small advice values."""
    rng = random.Random(seed)
    n = 1 << k
    lag = {}
    for i in range(cs.num_fixed):
        lag[("fixed", i)] = [rng.randrange(R) for _ in range(n)]
    for i in range(cs.num_advice):
        lag[("advice", i)] = [rng.randrange(16) for _ in range(n)]
    for i in range(cs.num_instance):
        lag[("instance", i)] = [rng.randrange(R) for _ in range(n)]
    for j in range(len(cs.permutation_columns)):
        lag[("sigma", j)] = [rng.randrange(R) for _ in range(n)]
    sel_tab = [(9, 10), (11, 12), (13, 14), (15, 16)] if which == "aggregation" else [(0, 2)]
    for sel, tab in sel_tab:
        lag[("fixed", sel)] = [rng.randrange(2) for _ in range(n)]
        lag[("fixed", tab)] = [i % 16 for i in range(n)]
    return lag


@pytest.mark.parametrize("which,k", [("small", 5), ("aggregation", 6)])
def test_lookup_and_product_rounds_from_resident_columns(ctx, which, k):
    """Rounds 2 and 3 computed on the device from the resident advice / fixed / sigma columns: permuted lookup
    columns and grand products (values, commitments, coefficient and extended forms) against the Python restatement."""
    import lookup_ref as lr

    cs = _small_cs() if which == "small" else plonk.aggregation_circuit_cs()
    n = 1 << k
    gl = ob.gen_bases(0x7100 + k, n)
    sid = ctx.srs_register(gl)
    pr = ResidentProver(ctx, cs, k, sid, sid)
    d = domain_consts(k, pr.ext_k)
    lag = _satisfiable_columns(cs, which, k, 9 + k)
    first = [nm for nm in lag]
    pr.commit_columns(first, [qu.pack(lag[nm]) for nm in first], keep_lagrange=True)
    rng = random.Random(4)
    theta, beta, gamma = [rng.randrange(R) for _ in range(3)]
    blinds = {}

    def blind(name, rows):
        blinds[name] = [rng.randrange(R) for _ in range(rows)]
        return qu.pack(blinds[name])

    bf = cs.blinding_factors()
    u = n - bf - 1
    w = qr.omega(k)

    def check(names, comms, want_cols):
        for nm, c in zip(names, comms):
            col = qu.pack(want_cols[nm])
            assert np.array_equal(ctx.d2h(pr.lag[nm], 4 * n), col), nm
            assert np.array_equal(c, ob.best_multiexp(col, gl)[:8]), nm
            co = ob.ifft(col.copy(), d["omega_inv"], d["n_inv"], k)
            assert np.array_equal(ctx.d2h(pr.coeff[nm], 4 * n), co), nm
            assert np.array_equal(ctx.d2h(pr.ext[nm], 4 << pr.ext_k), ob.coeff_to_extended(co, k, pr.ext_k, d["zeta"], d["omega_ext"])), nm

    # --- round 2
    comms = pr.lookup_round(theta, blind)
    want, names, comp = {}, [], {}
    for i, (_, ins, tabs) in enumerate(cs.lookups):
        A = lr.compress_expressions([e.to_tuple() for e in ins], lag, n, theta)
        S = lr.compress_expressions([e.to_tuple() for e in tabs], lag, n, theta)
        comp[i] = (A, S)
        pa, ps = lr.permute_expression_pair(A[:u], S[:u])
        want[("lookup_input", i)] = pa + blinds[("lookup_input", i)]
        want[("lookup_table", i)] = ps + blinds[("lookup_table", i)]
        names += [("lookup_input", i), ("lookup_table", i)]
    check(names, comms, want)
    # --- round 3
    comms = pr.product_round(beta, gamma, blind)
    names = []
    chunk = cs.chunk_len()
    last = 1
    for s in range(cs.num_permutation_sets()):
        cols = cs.permutation_columns[s * chunk:(s + 1) * chunk]
        z = lr.permutation_product([lag[c] for c in cols], [lag[("sigma", s * chunk + j)] for j in range(len(cols))], k, w, beta,
                                   gamma, s * chunk, last)
        last = z[u]
        want[("perm_z", s)] = z[: n - bf] + blinds[("perm_z", s)]
        names.append(("perm_z", s))
    for i in range(len(cs.lookups)):
        z = lr.lookup_product(comp[i][0], comp[i][1], want[("lookup_input", i)], want[("lookup_table", i)], beta, gamma)
        assert z[u] == 1
        want[("lookup_z", i)] = z[: n - bf] + blinds[("lookup_z", i)]
        names.append(("lookup_z", i))
    check(names, comms, want)
    pr.close()
    ctx.srs_release(sid)


def _drive_proof(ctx, cs, k, lag, sid, sid_g=None):
    """witness in -> proof elements out, challenges from the ShaWrite transcript (T1) in create_proof's order.
    Returns (desc of what the verifier needs)."""
    from halo2_snark_aggregator_b200.transcript import ShaWrite

    n = 1 << k
    pr = ResidentProver(ctx, cs, k, sid, sid if sid_g is None else sid_g)
    rng = random.Random(31)
    t = ShaWrite()
    comm = {}
    # (the vk digest halo2 absorbs first is an external-crate format: not restated)
    pk = [nm for nm in lag if nm[0] in ("fixed", "sigma", "l0", "l_last", "l_active_row")]
    comm.update(zip(pk, pr.commit_columns(pk, [qu.pack(lag[nm]) for nm in pk], keep_lagrange=True)))   # keygen: not part of the proof
    wit = [("instance", 0)] + [("advice", i) for i in range(cs.num_advice)]
    cw = pr.commit_columns(wit, [qu.pack(lag[nm]) for nm in wit], keep_lagrange=True)
    comm.update(zip(wit, cw))
    t.common_point(cw[0])          # instance commitment: absorbed, not written (API/systems/halo2/verify.rs:74-92)
    for c in cw[1:]:
        t.write_point(c)                                                                  # advice commitments
    theta = t.squeeze_challenge()

    def blind(name, rows):
        return qu.pack([rng.randrange(R) for _ in range(rows)])

    names2 = [(w_, i) for i in range(len(cs.lookups)) for w_ in ("lookup_input", "lookup_table")]
    for nm, c in zip(names2, pr.lookup_round(theta, blind)):
        comm[nm] = c
        t.write_point(c)
    beta, gamma = t.squeeze_challenge(), t.squeeze_challenge()
    names3 = [("perm_z", s_) for s_ in range(cs.num_permutation_sets())] + [("lookup_z", i) for i in range(len(cs.lookups))]
    for nm, c in zip(names3, pr.product_round(beta, gamma, blind)):
        comm[nm] = c
        t.write_point(c)
    comm[("random", 0)] = pr.commit_coeff_columns([("random", 0)], [qu.pack([rng.randrange(R) for _ in range(n)])])[0]
    t.write_point(comm[("random", 0)])
    y = t.squeeze_challenge()
    for i, c in enumerate(pr.quotient(y, beta, gamma, theta)):
        comm[("h_piece", i)] = c
        t.write_point(c)
    x = t.squeeze_challenge()
    pr.fold_h(x)
    queries = create_proof_queries(cs)
    evals = pr.evaluate(queries, x)
    ev_limbs = dict(zip(queries, evals))
    for q in transcript_eval_order(cs):      # halo2's write order; h(x) is not part of the proof
        t.write_scalar(ev_limbs[q])
    v = t.squeeze_challenge()
    order, ws = pr.open(queries, x, v)
    for c in ws:
        t.write_point(c)
    proof = t.finalize()
    ev = {q: qu.unpack(e)[0] for q, e in zip(queries, evals)}
    pr.close()
    return dict(x=x, y=y, beta=beta, gamma=gamma, theta=theta, v=v, ev=ev, proof=proof, n_w=len(ws), comm=comm, order=order,
                ws=ws, queries=queries)


def _valid_aggregation_witness(k, seed):
    """Columns that SATISFY the aggregation circuit's constraint system: gate coefficients zero (the base gate has no
    selector: all-zero coefficient rows are how unused rows look), range tables i mod 16 with 0/1 selectors and small
    advice cells, and sigma columns encoding one real copy cycle between equal cells."""
    cs = plonk.aggregation_circuit_cs()
    lag = _satisfiable_columns(cs, "aggregation", k, seed)
    n = 1 << k
    for i in range(9):
        lag[("fixed", i)] = [0] * n
    w = qr.omega(k)
    ident = [[pow(plonk.DELTA, j, R) * pow(w, i, R) % R for i in range(n)] for j in range(len(cs.permutation_columns))]
    for j in range(len(cs.permutation_columns)):
        lag[("sigma", j)] = list(ident[j])
    cells = [(0, 3), (2, 17), (5, 9), (4, 20)]        # (permutation column index, row): a0[3] = a2[17] = instance[9] = a4[20]
    val = 11
    for j, i in cells:
        lag[cs.permutation_columns[j]][i] = val
    for (j, i), (j2, i2) in zip(cells, cells[1:] + cells[:1]):
        lag[("sigma", j)][i] = ident[j2][i2]
    l0, l_last, l_active = qu.lagrange_selectors(k, cs.blinding_factors())
    lag[("l0", 0)], lag[("l_last", 0)], lag[("l_active_row", 0)] = l0, l_last, l_active
    return cs, lag


def test_device_proof_satisfies_the_reference_verifiers_equation(ctx):
    """End to end: a satisfying witness goes in, every round runs on the device with Fiat-Shamir challenges from the
    ShaWrite transcript, and the evaluations that come out satisfy the equation the REFERENCE's verifier checks --
    h(x) (x^n - 1) = fold_y(gates, permutation, lookups)(x), restated from
    halo2-snark-aggregator-api/src/systems/halo2/{params,permutation,lookup,vanish}.rs (oracle/py/quotient_ref.py).
    The identity holds only if the permuted columns, the grand products and the quotient are all right."""
    k = 6
    n = 1 << k
    cs, lag = _valid_aggregation_witness(k, 3)
    sid = ctx.srs_register(ob.gen_bases(0x7200 + k, n))
    out = _drive_proof(ctx, cs, k, lag, sid)
    desc = qu.oracle_desc(cs)

    def ev(name, rot):
        return out["ev"][(name, rot)]

    want = qr.verifier_h_eval(desc, ev, k, out["x"], out["y"], out["beta"], out["gamma"], out["theta"])
    assert out["ev"][(("h", 0), 0)] == want
    # proof bytes: 5 + 14 + 9 + 1 + 4 + 4 points of 64 B and 71 scalars of 32 B (1 instance, 6 advice, 17 fixed,
    # 1 random, 6 sigma, 5 permutation z, 35 lookup: SURVEY.md App. C)
    assert out["n_w"] == 4 and len(out["proof"]) == 64 * (5 + 14 + 9 + 1 + 4 + 4) + 32 * 71

    # a broken copy constraint: same flow, the verifier's equation must fail
    lag2 = {nm: list(v) for nm, v in lag.items()}
    lag2[("advice", 0)][3] = 12
    out2 = _drive_proof(ctx, cs, k, lag2, sid)
    want2 = qr.verifier_h_eval(desc, lambda nm, rot: out2["ev"][(nm, rot)], k, out2["x"], out2["y"], out2["beta"], out2["gamma"],
                               out2["theta"])
    assert out2["ev"][(("h", 0), 0)] != want2

    # a cell outside the range table: the lookup round reports it like halo2 (ConstraintSystemFailure)
    lag3 = {nm: list(v) for nm, v in lag.items()}
    lag3[("advice", 1)][5] = 16
    lag3[("fixed", 9)][5] = 1
    from halo2_snark_aggregator_b200 import H2aggError
    with pytest.raises(H2aggError) as e:
        _drive_proof(ctx, cs, k, lag3, sid)
    assert "error 4" in str(e.value)
    ctx.srs_release(sid)


def test_device_proof_openings_verify_against_a_trapdoor_srs(ctx):
    """The other half of verify_proof: the GWC opening check e(W, [s - z]_2) = e(F - [e]_1, [1]_2) per point, done in
    the group with the trapdoor s of a toy SRS (g_i = s^i G, g_lagrange_i = L_i(s) G) and textbook affine arithmetic on
    Python integers: (s - z) W = sum_i v^i (C_i - e_i G).  It ties together what the device produced -- the
    commitments (MSM against both SRS forms), the coefficient forms behind the evaluations (iNTT), the fold, the Kate
    quotients and their commitments."""
    import bn254_ref as ref

    k = 6
    n = 1 << k
    cs, lag = _valid_aggregation_witness(k, 5)
    s_trap = 0x1234567890ABCDEF1234567890ABCDEF % R
    w = qr.omega(k)
    G = ref.G1_GEN
    g = [ref.g1_mul(pow(s_trap, i, R), G) for i in range(n)]
    # L_i(s) = (s^n - 1) w^i / (n (s - w^i))
    sn1 = (pow(s_trap, n, R) - 1) % R
    gl = [ref.g1_mul(sn1 * pow(w, i, R) % R * pow(n * (s_trap - pow(w, i, R)) % R, -1, R) % R, G) for i in range(n)]
    sid_g = ctx.srs_register(np.array(ref.pack_points(g), dtype=np.uint64))
    sid_gl = ctx.srs_register(np.array(ref.pack_points(gl), dtype=np.uint64))
    out = _drive_proof(ctx, cs, k, lag, sid_gl, sid_g)
    x, v = out["x"], out["v"]
    pt = lambda c: ref.unpack_point([int(t_) for t_ in c])
    C = {nm: pt(c) for nm, c in out["comm"].items()}
    # h(X) = sum_i x^(n i) h_i(X): its commitment is the same combination of the piece commitments
    xn = pow(x, n, R)
    C[("h", 0)] = None
    for i in reversed(range(cs.degree() - 1)):
        C[("h", 0)] = ref.g1_add(ref.g1_mul(xn, C[("h", 0)]), C[("h_piece", i)])
    assert all(ref.on_curve(p) for p in C.values())
    for rot, wpt in zip(out["order"], out["ws"]):
        z = x * pow(w, rot, R) % R
        F, E = None, 0
        for nm, r2 in reversed(out["queries"]):   # query i of a point is weighted v^i
            if r2 == rot:
                F = ref.g1_add(ref.g1_mul(v, F), C[nm])
                E = (E * v + out["ev"][(nm, r2)]) % R
        lhs = ref.g1_mul((s_trap - z) % R, pt(wpt))
        rhs = ref.g1_add(F, ref.g1_neg(ref.g1_mul(E, G)))
        assert lhs == rhs, "opening at rotation %d" % rot
    ctx.srs_release(sid_g)
    ctx.srs_release(sid_gl)


def _write_sections(path, sections):
    with open(path, "wb") as f:
        for tag, arr in sections:
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64))
            f.write(np.array([tag, a.size], dtype=np.uint64).tobytes())
            f.write(a.tobytes())


def test_cpp_resident_prover_equals_python_twin(ctx, tmp_path):
    """include/h2agg_prover.hpp (the C++ host driver a compiled prover would link) against prover.py on the same
    inputs: every commitment, evaluation and W point of the whole pipeline, bit for bit.  The Python twin is itself
    held against the CPU oracles above."""
    import os
    import subprocess

    k = 6
    n = 1 << k
    cs, lag = _valid_aggregation_witness(k, 8)
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    idx = plan.index
    bf = cs.blinding_factors()
    rng = random.Random(77)
    g, gl = ob.gen_bases(0x7300, n), ob.gen_bases(0x7301, n)
    theta, beta, gamma, y, x, v = [rng.randrange(R) for _ in range(6)]
    w = qr.omega(k)
    pk = [nm for nm in plan.columns if nm[0] in ("fixed", "sigma", "l0", "l_last", "l_active_row")]
    wit = [("instance", 0)] + [("advice", i) for i in range(cs.num_advice)]
    blinds2 = {nm: [rng.randrange(R) for _ in range(bf + 1)] for i in range(len(cs.lookups)) for nm in (("lookup_input", i), ("lookup_table", i))}
    znames = [("perm_z", s) for s in range(cs.num_permutation_sets())] + [("lookup_z", i) for i in range(len(cs.lookups))]
    blinds3 = {nm: [rng.randrange(R) for _ in range(bf)] for nm in znames}
    random_poly = [rng.randrange(R) for _ in range(n)]
    queries = create_proof_queries(cs)

    # ---- Python twin
    sid_g, sid_gl = ctx.srs_register(g), ctx.srs_register(gl)
    pr = ResidentProver(ctx, cs, k, sid_gl, sid_g)
    want = []
    want.append(pr.commit_columns(pk, [qu.pack(lag[nm]) for nm in pk], keep_lagrange=True))
    want.append(pr.commit_columns(wit, [qu.pack(lag[nm]) for nm in wit], keep_lagrange=True))
    blind = lambda nm, rows: qu.pack((blinds2 if nm in blinds2 else blinds3)[nm])
    want.append(pr.lookup_round(theta, blind))
    want.append(pr.product_round(beta, gamma, blind))
    want.append(pr.commit_coeff_columns([("random", 0)], [qu.pack(random_poly)]))
    want.append(pr.quotient(y, beta, gamma, theta))
    pr.fold_h(x)
    want.append(pr.evaluate(queries, x))
    want.append(pr.open(queries, x, v)[1])
    pr.close()
    ctx.srs_release(sid_g)
    ctx.srs_release(sid_gl)
    ctx.synchronize()
    want = np.concatenate([np.asarray(a, dtype=np.uint64).ravel() for a in want])

    # ---- the same job for the C++ driver
    fr = plonk.fr_mont
    w_ext = qr.omega(ext_k)
    ids = dict(idx)
    ids[("random", 0)] = len(plan.columns)
    ids[("h", 0)] = len(plan.columns) + 1
    expr_names = [nm for nm in plan.columns if nm[0] in ("fixed", "advice", "instance")]
    eidx = {nm: i for i, nm in enumerate(expr_names)}
    sec = [(1, [k, ext_k, bf, cs.chunk_len(), cs.degree() - 1, len(plan.columns)]),
           (2, plan.words.astype(np.uint64)), (3, plan.consts),
           (4, np.concatenate([fr(c) for c in (w, pow(w, -1, R), pow(n, -1, R), w_ext, pow(w_ext, -1, R), pow(1 << ext_k, -1, R),
                                               plonk.ZETA, plonk.DELTA)])),
           (5, np.concatenate([fr(t) for t in plonk.t_evaluations(k, ext_k)])),
           (6, [idx[nm] for nm in expr_names])]
    for i, (_, ins, tabs) in enumerate(cs.lookups):
        pi, pt = plonk.ExpressionList(ins, eidx), plonk.ExpressionList(tabs, eidx)
        sec += [(7, pi.words.astype(np.uint64)), (8, pi.consts), (9, pt.words.astype(np.uint64)), (10, pt.consts),
                (11, [idx[("lookup_z", i)], idx[("lookup_input", i)], idx[("lookup_table", i)]])]
    sec += [(12, [idx[c] for c in cs.permutation_columns]), (13, [idx[("sigma", j)] for j in range(len(cs.permutation_columns))]),
            (14, [idx[("perm_z", s)] for s in range(cs.num_permutation_sets())]), (15, gl), (16, g), (17, [idx[nm] for nm in pk])]
    sec += [(18, qu.pack(lag[nm])) for nm in pk]
    sec += [(19, [idx[nm] for nm in wit])] + [(20, qu.pack(lag[nm])) for nm in wit]
    sec += [(21, np.concatenate([fr(c) for c in (theta, beta, gamma, y, v)])),
            (22, np.concatenate([fr(beta * pow(plonk.DELTA, s * cs.chunk_len(), R)) for s in range(cs.num_permutation_sets())]))]
    for i in range(len(cs.lookups)):
        sec += [(23, qu.pack(blinds2[("lookup_input", i)])), (23, qu.pack(blinds2[("lookup_table", i)]))]
    sec += [(24, qu.pack(blinds3[nm])) for nm in znames]
    sec += [(25, qu.pack(random_poly)), (26, fr(pow(x, n, R)))]
    m64 = (1 << 64) - 1
    sec.append((27, [val for nm, rot in queries for val in (ids[nm], rot & m64)]))
    rots = list(dict.fromkeys(rot for _, rot in queries))
    pts = []
    for rot in rots:
        pts.append(np.array([rot & m64], dtype=np.uint64))
        pts.append(fr(x * pow(w, rot, R) % R))
    sec.append((28, np.concatenate(pts)))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_sections(fin, sec)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "prover_main")
    assert os.path.exists(exe), "tests/cpp/prover_main is not built (python -c 'import __graft_entry__ as g; g.build()')"
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    got = np.fromfile(fout, dtype=np.uint64)
    assert got.size == want.size
    assert np.array_equal(got, want)


def test_distributed_prover_world1_equals_resident_prover():
    """halo2_snark_aggregator_b200/dist_prover.py with one rank issues no collective and must reproduce the
    ResidentProver bit for bit (tests/dist_prover_main.py; the same script runs under torchrun for 2 / 4 / 8 GPUs)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "dist_prover_main.py"), "11"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_PROVER_OK world=1" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_keygen_pk_and_proving_key_cache(ctx):
    """ResidentProver.keygen_pk: the vk's fixed / permutation commitments = oracle best_multiexp of the Lagrange columns;
    l_0 / l_last / l_active_row on the extended domain = oracle transforms of the halo2 keygen definitions; and the
    ProvingKeyCache hands the same resident prover back instead of recomputing the key (the reference recomputes
    keygen_pk in every verify_run, verify_circuit.rs:974-979)."""
    from halo2_snark_aggregator_b200.prover import ProvingKeyCache

    cs = plonk.aggregation_circuit_cs()
    k = 7
    n = 1 << k
    gl = ob.gen_bases(0x7400, n)
    sid = ctx.srs_register(gl)
    cols = {}
    calls = []

    def fill(nm, d_l):
        calls.append(nm)
        cols[nm] = ob.gen_scalars(0x7500 + len(cols), 0, n)
        ctx.h2d(d_l, cols[nm])

    cache = ProvingKeyCache()
    pr, comm, cached = cache.get(("vk-digest", k), lambda: ResidentProver(ctx, cs, k, sid, sid), fill)
    assert not cached and len(calls) == cs.num_fixed + len(cs.permutation_columns)
    d = domain_consts(k, pr.ext_k)
    for i in range(cs.num_fixed):
        assert np.array_equal(comm["fixed"][i], ob.best_multiexp(cols[("fixed", i)], gl)[:8])
    for j in range(len(cs.permutation_columns)):
        assert np.array_equal(comm["sigma"][j], ob.best_multiexp(cols[("sigma", j)], gl)[:8])
    for nm, vals in zip((("l0", 0), ("l_last", 0), ("l_active_row", 0)), qu.lagrange_selectors(k, cs.blinding_factors())):
        co = ob.ifft(qu.pack(vals), d["omega_inv"], d["n_inv"], k)
        assert np.array_equal(ctx.d2h(pr.ext[nm], 4 << pr.ext_k), ob.coeff_to_extended(co, k, pr.ext_k, d["zeta"], d["omega_ext"])), nm
    pr2, comm2, cached2 = cache.get(("vk-digest", k), lambda: 1 / 0, fill)
    assert cached2 and pr2 is pr and len(calls) == cs.num_fixed + len(cs.permutation_columns)
    cache.drop(("vk-digest", k))
    ctx.srs_release(sid)
