"""N1 (evaluate_h) on the CPU: the oracle is pinned to the reference's VERIFIER equations, and the plan builder
is checked against the oracle through a plain-Python interpreter of the plan words.  No GPU."""
import random

import numpy as np
import pytest

import quotient_util as qu
import quotient_ref as qr
from halo2_snark_aggregator_b200 import plonk

R = qr.R


def test_aggregation_circuit_shape():
    """degree 5, chunk_len 3, 2 permutation sets, 5 blinding factors, extended_k = k + 2"""
    cs = plonk.aggregation_circuit_cs()
    assert (cs.num_fixed, cs.num_advice, cs.num_instance) == (17, 5, 1)
    assert cs.degree() == 5 and cs.chunk_len() == 3 and cs.num_permutation_sets() == 2
    assert cs.blinding_factors() == 5
    assert cs.extended_k(22) == 24
    assert len(cs.lookups) == 7 and len(cs.permutation_columns) == 6
    assert cs.gates[0][1][0].degree() == 3
    plan = plonk.build_quotient_plan(cs)
    # 17 fixed + 5 advice + 1 instance + 6 sigma + 3 + 2 z + 7 * 3
    assert len(plan.columns) == 55
    # the gate: 1 constant term + next + 5 linear + 2 cubic
    assert plan.words[9] == 9


def test_expression_expand_matches_tree():
    rng = random.Random(5)
    E = plonk.Expression
    e = (E.advice(0) + E.fixed(1) * 3 - E.constant(7)) * (E.advice(0, 1) - E.instance(0)) * E.fixed(2, -1) + E.advice(1) * E.advice(1)
    vals = {}

    def query(kind, c, r):
        return vals.setdefault((kind, c, r), rng.randrange(R))

    want = qr.eval_expr(e.to_tuple(), query)
    got = 0
    for mono, c in e.expand().items():
        p = c
        for q in mono:
            p = p * query(*q) % R
        got = (got + p) % R
    assert got == want
    assert e.degree() == 3


def _setup(cs, k, seed):
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    lag = qu.random_lagrange_columns(plan, k, seed)
    coeffs = {name: qr.lagrange_to_coeff(v, k) for name, v in lag.items()}
    ext = {name: qr.coeff_to_extended(c, k, ext_k) for name, c in coeffs.items()}
    rng = random.Random(seed + 1)
    ch = [rng.randrange(R) for _ in range(4)]
    return plan, ext_k, coeffs, ext, ch


@pytest.mark.parametrize("k", [3, 4])
def test_oracle_matches_reference_verifier_equations(k):
    """evaluate_h / (X^n - 1) at row i == the reference verifier's expected_h_eval at X_i = zeta omega_ext^i,
    with every polynomial evaluated by Horner on its coefficients (independent of the coset FFT)."""
    cs = plonk.aggregation_circuit_cs()
    plan, ext_k, coeffs, ext, (y, beta, gamma, theta) = _setup(cs, k, 100 + k)
    desc = qu.oracle_desc(cs)
    h = qr.divide_by_vanishing_poly(qr.evaluate_h(desc, ext, k, ext_k, y, beta, gamma, theta), k, ext_k)
    w, w_ext = qr.omega(k), qr.omega(ext_k)
    for i in range(1 << ext_k):
        x = qr.ZETA * pow(w_ext, i, R) % R

        def ev(name, rot, x=x):
            return qr.horner(coeffs[name], x * pow(w, rot, R) % R)

        assert h[i] == qr.verifier_h_eval(desc, ev, k, x, y, beta, gamma, theta), "row %d" % i


def test_verifier_equations_at_a_random_point_after_extended_to_coeff():
    """With constraints that HOLD (all-zero witness: gate 0 = 0, z = 1 products, equal permuted columns) the
    quotient is a polynomial: interpolate h from the coset and compare at a fresh point (what the proof opens)."""
    k = 3
    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    n = 1 << k
    rng = random.Random(9)
    lag = {name: [0] * n for name in plan.columns}
    l0, l_last, l_active = qu.lagrange_selectors(k, cs.blinding_factors())
    lag[("l0", 0)], lag[("l_last", 0)], lag[("l_active_row", 0)] = l0, l_last, l_active
    # identity permutation: sigma_j(omega^i) = delta^j omega^i ; z = 1 everywhere; a' = s' ; lookups z = 1
    w = qr.omega(k)
    for j in range(len(cs.permutation_columns)):
        lag[("sigma", j)] = [pow(qr.DELTA, j, R) * pow(w, i, R) % R for i in range(n)]
    for s in range(cs.num_permutation_sets()):
        lag[("perm_z", s)] = [1] * n
    for i in range(len(cs.lookups)):
        lag[("lookup_z", i)] = [1] * n
        perm = [rng.randrange(R) for _ in range(n)]
        lag[("lookup_input", i)] = perm
        lag[("lookup_table", i)] = list(perm)
    # advice with arbitrary values where the gate is switched off (all fixed = 0) and lookups see input 0 = table 0
    for a in range(cs.num_advice):
        lag[("advice", a)] = [rng.randrange(R) for _ in range(n)]
    lag[("instance", 0)] = [rng.randrange(R) for _ in range(n)]
    # lookup product relation needs (a'+beta)(s'+gamma) == (0+beta)(0+gamma): take a' = s' = 0 instead
    for i in range(len(cs.lookups)):
        lag[("lookup_input", i)] = [0] * n
        lag[("lookup_table", i)] = [0] * n
    coeffs = {name: qr.lagrange_to_coeff(v, k) for name, v in lag.items()}
    ext = {name: qr.coeff_to_extended(c, k, ext_k) for name, c in coeffs.items()}
    y, beta, gamma, theta = [rng.randrange(R) for _ in range(4)]
    desc = qu.oracle_desc(cs)
    h = qr.divide_by_vanishing_poly(qr.evaluate_h(desc, ext, k, ext_k, y, beta, gamma, theta), k, ext_k)
    # interpolate on the coset (extended_to_coeff), big-int DFT
    size = 1 << ext_k
    w_ext_inv = pow(qr.omega(ext_k), -1, R)
    hc = []
    for j in range(size):
        acc = 0
        for i in range(size):
            acc = (acc + h[i] * pow(w_ext_inv, i * j, R)) % R
        hc.append(acc * pow(size, -1, R) % R * pow(qr.ZETA, -j, R) % R)
    assert all(c == 0 for c in hc[n * (cs.degree() - 1):]), "quotient degree exceeds the truncation bound"
    x = rng.randrange(R)

    def ev(name, rot):
        return qr.horner(coeffs[name], x * pow(w, rot, R) % R)

    assert qr.horner(hc, x) == qr.verifier_h_eval(desc, ev, k, x, y, beta, gamma, theta)


@pytest.mark.parametrize("k,seed", [(3, 1), (4, 2)])
def test_plan_interpreter_matches_oracle(k, seed):
    cs = plonk.aggregation_circuit_cs()
    plan, ext_k, _, ext, (y, beta, gamma, theta) = _setup(cs, k, seed)
    desc = qu.oracle_desc(cs)
    want = qr.divide_by_vanishing_poly(qr.evaluate_h(desc, ext, k, ext_k, y, beta, gamma, theta), k, ext_k)
    got = qu.interpret_plan(plan, [ext[name] for name in plan.columns], k, ext_k, y, beta, gamma, theta)
    assert got == want


def test_plan_of_a_generic_system():
    """not only the aggregation shape: multi-polynomial gates, constants, multi-expression lookups, one set"""
    E = plonk.Expression
    cs = plonk.ConstraintSystem(num_fixed=3, num_advice=3, num_instance=1)
    cs.create_gate("g0", [E.fixed(0) * (E.advice(0) * E.advice(1) - E.advice(2, 1)),
                          E.fixed(1) * (E.advice(0) + 5) * (E.advice(0, -1) - E.constant(3))])
    cs.create_gate("g1", [E.instance(0) - E.advice(2) * 7])
    cs.lookup("two columns", [(E.advice(0) * E.fixed(0), E.fixed(2)), (E.advice(1) + E.constant(1), E.fixed(1, 1))])
    cs.enable_equality("advice", 0)
    cs.enable_equality("fixed", 2)
    k = 4
    plan, ext_k, _, ext, (y, beta, gamma, theta) = _setup(cs, k, 77)
    desc = qu.oracle_desc(cs)
    want = qr.evaluate_h(desc, ext, k, ext_k, y, beta, gamma, theta)
    got = qu.interpret_plan(plan, [ext[name] for name in plan.columns], k, ext_k, y, beta, gamma, theta, divide=False)
    assert got == want
    assert np.array_equal(qu.pack(qu.unpack(qu.pack(want[:5]))), qu.pack(want[:5]))


@pytest.mark.parametrize("k,divide", [(3, True), (5, False), (6, True)])
def test_cpp_oracle_matches_python_oracle(k, divide):
    """the multi-threaded C++ restatement (bench cpu_baseline, mid-size parity) against the big-int tree oracle"""
    import oracle_binding as ob

    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    size = 1 << ext_k
    rng = random.Random(40 + k)
    ext = {name: [rng.randrange(R) for _ in range(size)] for name in plan.columns}
    y, beta, gamma, theta = [rng.randrange(R) for _ in range(4)]
    want = qr.evaluate_h(qu.oracle_desc(cs), ext, k, ext_k, y, beta, gamma, theta)
    t = None
    if divide:
        want = qr.divide_by_vanishing_poly(want, k, ext_k)
        t = qu.pack(plonk.t_evaluations(k, ext_k))
    cols = [qu.pack(ext[name]) for name in plan.columns]
    got = ob.evaluate_h(plan.words, plan.consts, cols, k, ext_k, qu.pack([y]), qu.pack([beta]), qu.pack([gamma]),
                        qu.pack([theta]), qu.pack([qr.omega(ext_k)]), qu.pack([qr.ZETA]), qu.pack([qr.DELTA]), t, nthreads=3)
    assert qu.unpack(got) == want


def test_expression_list_words_evaluate_like_the_oracle():
    """plonk.ExpressionList (the word program h2agg_compress_expressions_dev runs) interpreted in plain Python equals
    the tree evaluation of oracle/py/lookup_ref.compress_expressions, rotations wrapping mod n."""
    import lookup_ref as lr

    E = plonk.Expression
    k = 4
    n = 1 << k
    rng = random.Random(21)
    names = [("advice", 0), ("advice", 1), ("fixed", 0), ("instance", 0)]
    index = {nm: i for i, nm in enumerate(names)}
    cols = {nm: [rng.randrange(R) for _ in range(n)] for nm in names}
    exprs = [E.advice(0) * E.fixed(0), E.advice(1, 1) * 3 + E.constant(9), E.instance(0, -2) * E.advice(0) * E.advice(0, 1) - E.fixed(0)]
    prog = plonk.ExpressionList(exprs, index)
    w = [int(x) for x in prog.words]
    consts = qu.unpack(prog.consts) if prog.consts.size else []
    theta = rng.randrange(R)
    got = []
    for i in range(n):
        pc = 1
        acc = 0
        for _ in range(w[0]):
            nt = w[pc]
            pc += 1
            s = 0
            for _ in range(nt):
                ci, nf = w[pc], w[pc + 1]
                pc += 2
                prod = 1 if ci == plonk.NOCONST else consts[ci]
                for _ in range(nf):
                    word = w[pc]
                    pc += 1
                    rot = word >> 16
                    rot = rot - 65536 if rot >= 32768 else rot
                    prod = prod * cols[names[word & 0xFFFF]][(i + rot) % n] % R
                s = (s + prod) % R
            acc = (acc * theta + s) % R
        assert pc == len(w)
        got.append(acc)
    assert got == lr.compress_expressions([e.to_tuple() for e in exprs], cols, n, theta)
