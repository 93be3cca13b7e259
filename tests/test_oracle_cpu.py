"""CPU suite: the C++ oracle against the Python big-int golden vectors and against itself
(bucket method vs double-and-add, radix-2 recursion vs O(n^2) DFT), plus generator agreement."""
import numpy as np
import pytest

import bn254_ref as ref
import oracle_binding as ob
from util import affine_of, arr, domain_consts, fr_limbs, golden, omega


def test_field_golden():
    for case in golden("field.json"):
        f = case["field"]
        a, b = arr(case["a"]), arr(case["b"])
        assert np.array_equal(ob.field_op(f, 0, a, b), arr(case["add"]))
        assert np.array_equal(ob.field_op(f, 1, a, b), arr(case["sub"]))
        assert np.array_equal(ob.field_op(f, 3, a, b), arr(case["mul"]))
        assert np.array_equal(ob.field_op(f, 2, a), arr(case["inv"]))


@pytest.mark.parametrize("case", golden("msm.json"), ids=lambda c: c["name"])
def test_msm_golden(case):
    s, b = arr(case["scalars"]), arr(case["bases"])
    want = arr(case["affine"])
    for nthreads in (1, 3):
        assert np.array_equal(affine_of(ob.best_multiexp(s, b, nthreads)), want)
    assert np.array_equal(affine_of(ob.msm_naive(s, b)), want)


@pytest.mark.parametrize("case", golden("ntt.json"), ids=lambda c: c["name"])
def test_ntt_golden(case):
    a, want = arr(case["input"]), arr(case["output"])
    name = case["name"]
    if name.startswith("fft"):
        got = ob.best_fft(a.copy(), arr(case["omega"]), case["k"], 1)
        assert np.array_equal(got, want)
        assert np.array_equal(ob.best_fft(a.copy(), arr(case["omega"]), case["k"], 4), want)
    elif name.startswith("ifft"):
        assert np.array_equal(ob.ifft(a.copy(), arr(case["omega_inv"]), arr(case["n_inv"]), case["k"]), want)
    elif name.startswith("coeff_to_extended"):
        assert np.array_equal(ob.coeff_to_extended(a, case["k"], case["ext_k"], arr(case["zeta"]), arr(case["omega_ext"])), want)
    else:
        got = ob.extended_to_coeff(a.copy(), case["ext_k"], arr(case["omega_ext_inv"]), arr(case["ext_n_inv"]), arr(case["zeta"]), case["out_len"])
        assert np.array_equal(got, want)


def test_generators_match_python():
    for kind in range(4):
        got = ob.gen_scalars(0x1234 + kind, kind, 40, first=7)
        want = np.array(ref.pack_fr([ref.gen_scalar(0x1234 + kind, kind, 7 + i) for i in range(40)]), dtype=np.uint64)
        assert np.array_equal(got, want)
    got = ob.gen_bases(0x5352_5300, 12, first=3)
    pts = [ref.gen_base(0x5352_5300, 3 + i) for i in range(12)]
    assert all(ref.on_curve(p) for p in pts)
    assert np.array_equal(got, np.array(ref.pack_points(pts), dtype=np.uint64))
    for i in range(12):
        assert ob.on_curve(got[8 * i:8 * i + 8])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 31, 32, 33, 100, 1000])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_best_multiexp_vs_naive(n, kind):
    s = ob.gen_scalars(0xA660000 + n, kind, n)
    b = ob.gen_bases(0x53525300, n)
    want = ob.msm_naive(s, b)
    for nthreads in (1, 2, 8):
        assert np.array_equal(ob.best_multiexp(s, b, nthreads), want)


@pytest.mark.parametrize("k", range(0, 11))
def test_best_fft_vs_dft(k):
    a = ob.gen_scalars(0xF00 + k, 0, 1 << k)
    w = fr_limbs(omega(k))
    want = ob.dft_naive(a, w, k)
    for nthreads in (1, 4, 8):
        assert np.array_equal(ob.best_fft(a.copy(), w, k, nthreads), want)


@pytest.mark.parametrize("k", [3, 8, 12, 15])
def test_domain_round_trips(k):
    d = domain_consts(k)
    a = ob.gen_scalars(0xD0 + k, 0, 1 << k)
    f = ob.best_fft(a.copy(), d["omega"], k)
    assert np.array_equal(ob.ifft(f, d["omega_inv"], d["n_inv"], k), a)
    ext = ob.coeff_to_extended(a, k, k + 2, d["zeta"], d["omega_ext"])
    back = ob.extended_to_coeff(ext, k + 2, d["omega_ext_inv"], d["ext_n_inv"], d["zeta"], 3 << k)
    assert np.array_equal(back[: 4 << k], a)
    assert not back[4 << k:].any()


def test_msm_linearity_and_concatenation():
    n = 300
    s1, s2 = ob.gen_scalars(1, 0, n), ob.gen_scalars(2, 0, n)
    b = ob.gen_bases(3, n)
    lhs = ob.best_multiexp(ob.field_op(0, 0, s1, s2), b)
    rhs = ob.g1_sum(np.concatenate([ob.best_multiexp(s1, b), ob.best_multiexp(s2, b)]))
    assert np.array_equal(lhs, rhs)
    whole = ob.best_multiexp(s1, b)
    parts = ob.g1_sum(np.concatenate([ob.best_multiexp(s1[: 4 * 100], b[: 8 * 100]), ob.best_multiexp(s1[4 * 100:], b[8 * 100:])]))
    assert np.array_equal(whole, parts)


def test_eval_and_kate_division_vs_python_bigint():
    import random

    from util import R_MOD, fr_limbs

    rng = random.Random(5)
    for n in (1, 2, 3, 17, 64, 65, 200):
        coeffs = [rng.randrange(R_MOD) for _ in range(n)]
        b = rng.randrange(R_MOD)
        a = np.array(ref.pack_fr(coeffs), dtype=np.uint64)
        want_eval = sum(c * pow(b, i, R_MOD) for i, c in enumerate(coeffs)) % R_MOD
        assert ref.unpack_fr(list(ob.eval_polynomial(a, fr_limbs(b)))) == [want_eval]
        q = ref.unpack_fr(list(ob.kate_division(a, fr_limbs(b))))
        # a(X) = q(X) (X - b) + a(b): compare coefficients
        prod = [0] * n
        for i, qi in enumerate(q):
            prod[i + 1] = (prod[i + 1] + qi) % R_MOD
            prod[i] = (prod[i] - qi * b) % R_MOD
        prod[0] = (prod[0] + want_eval) % R_MOD
        assert prod == coeffs


def test_batch_invert_and_grand_product_vs_python_bigint():
    import random

    from util import R_MOD

    rng = random.Random(9)
    n = 150
    vals = [rng.randrange(R_MOD) for _ in range(n)]
    vals[7] = vals[64] = 0  # BatchInvert leaves zeros alone
    a = np.array(ref.pack_fr(vals), dtype=np.uint64)
    got = ref.unpack_fr(list(ob.batch_invert(a)))
    assert got == [pow(v, -1, R_MOD) if v else 0 for v in vals]
    num = [rng.randrange(R_MOD) for _ in range(n)]
    den = [rng.randrange(1, R_MOD) for _ in range(n)]
    z = ref.unpack_fr(list(ob.grand_product(np.array(ref.pack_fr(num), dtype=np.uint64), np.array(ref.pack_fr(den), dtype=np.uint64))))
    run, want = 1, []
    for x, y in zip(num, den):
        want.append(run)
        run = run * x * pow(y, -1, R_MOD) % R_MOD
    assert z == want


def test_permute_expression_pair_oracle_vs_python_pin():
    """C++ restatement of halo2's permute_expression_pair == the Python big-int restatement, and both satisfy the
    properties the reference's verifier enforces on (a', s') (systems/halo2/lookup.rs:58-119)."""
    import random

    import lookup_ref as lr
    import quotient_util as qu

    rng = random.Random(3)
    for u in (1, 2, 3, 17, 256, 1500):
        pool = [rng.randrange(qu.R) for _ in range(max(1, u // 4))] + list(range(5))
        table = [rng.choice(pool) for _ in range(u)]
        inp = [rng.choice(table) for _ in range(u)]
        a, s = lr.permute_expression_pair(inp, table)
        lr.check_lookup_constraints(inp, table, a, s)
        rc, ga, gs = ob.permute_expression_pair(qu.pack(inp), qu.pack(table))
        assert rc == 0 and qu.unpack(ga) == a and qu.unpack(gs) == s
    assert ob.permute_expression_pair(qu.pack([5, 6]), qu.pack([5, 5]))[0] == 1
    with pytest.raises(ValueError):
        lr.permute_expression_pair([5, 6], [5, 5])
    # hand-checked example: leftovers ascending go to repeated rows from the back
    a, s = lr.permute_expression_pair([2, 1, 2, 2, 1], [1, 2, 7, 3, 9])
    assert a == [1, 1, 2, 2, 2] and s == [1, 9, 2, 7, 3]


def test_lookup_and_permutation_product_restatements_telescope():
    """oracle/py/lookup_ref.py: with a valid permuted pair the lookup product returns to 1 at the last usable row, and
    with sigma columns that encode a real copy-constraint cycle the chained permutation products do too -- the two
    facts the reference's verifier relies on (systems/halo2/lookup.rs:58-119, permutation.rs:54-136)."""
    import random

    import lookup_ref as lr
    import quotient_ref as qr

    R = lr.R
    rng = random.Random(12)
    k, bf = 6, 5
    n = 1 << k
    u = n - bf - 1
    table = [i % 16 for i in range(n)]
    inp = [rng.randrange(16) for _ in range(n)]
    pa, ps = lr.permute_expression_pair(inp[:u], table[:u])
    pa += [rng.randrange(R) for _ in range(bf + 1)]
    ps += [rng.randrange(R) for _ in range(bf + 1)]
    beta, gamma = rng.randrange(R), rng.randrange(R)
    z = lr.lookup_product(inp, table, pa, ps, beta, gamma)
    assert z[0] == 1 and z[u] == 1 and any(v != 1 for v in z[1:u])
    # permutation: 4 columns in 2 sets of 2 (chunk_len 2); a 3-cycle of equal cells across columns, rows < u
    w = qr.omega(k)
    m, chunk = 4, 2
    values = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    ident = [[pow(lr.DELTA, j, R) * pow(w, i, R) % R for i in range(n)] for j in range(m)]
    sig = [list(c) for c in ident]
    cells = [(0, 3), (2, 17), (3, 40)]
    v = rng.randrange(R)
    for j, i in cells:
        values[j][i] = v
    for (j, i), (j2, i2) in zip(cells, cells[1:] + cells[:1]):
        sig[j][i] = ident[j2][i2]
    last = 1
    zs = []
    for s in range(2):
        zz = lr.permutation_product(values[s * chunk:(s + 1) * chunk], sig[s * chunk:(s + 1) * chunk], k, w, beta, gamma, s * chunk, last)
        last = zz[u]
        zs.append(zz)
    assert zs[0][0] == 1 and zs[1][0] == zs[0][u] and zs[0][u] != 1 and zs[1][u] == 1
    # a broken copy constraint does not telescope
    values[0][3] = (v + 1) % R
    last = 1
    for s in range(2):
        last = lr.permutation_product(values[s * chunk:(s + 1) * chunk], sig[s * chunk:(s + 1) * chunk], k, w, beta, gamma, s * chunk, last)[u]
    assert last != 1


def test_cpp_argument_restatements_match_the_python_ones():
    """oracle/cpu_halo2.cpp compress_expressions / lookup_product / permutation_product (the ones that scale to 2^22 rows)
    == oracle/py/lookup_ref.py on big ints."""
    import random

    import lookup_ref as lr
    import quotient_ref as qr
    import quotient_util as qu
    from halo2_snark_aggregator_b200 import plonk

    R = lr.R
    rng = random.Random(41)
    E = plonk.Expression
    for k in (1, 5):
        n = 1 << k
        names = [("advice", 0), ("advice", 1), ("fixed", 0), ("instance", 0)]
        cols = {nm: [rng.randrange(R) for _ in range(n)] for nm in names}
        index = {nm: i for i, nm in enumerate(names)}
        exprs = [E.advice(0) * E.fixed(0), E.advice(1, 1) * 3 + E.constant(9), E.instance(0, -2) * E.advice(0) * E.advice(0, 1) - E.fixed(0)]
        prog = plonk.ExpressionList(exprs, index)
        theta = rng.randrange(R)
        got = ob.compress_expressions(prog.words, prog.consts, [qu.pack(cols[nm]) for nm in names], k, qu.pack([theta]), 3)
        assert qu.unpack(got) == lr.compress_expressions([e.to_tuple() for e in exprs], cols, n, theta)
    for n in (1, 2, 65, 300):
        A, S, Ap, Sp = ([rng.randrange(R) for _ in range(n)] for _ in range(4))
        beta, gamma = rng.randrange(R), rng.randrange(R)
        got = ob.lookup_product(*(qu.pack(v) for v in (A, S, Ap, Sp)), qu.pack([beta]), qu.pack([gamma]))
        assert qu.unpack(got) == lr.lookup_product(A, S, Ap, Sp, beta, gamma)
    for k, m, first, lz in ((1, 1, 0, None), (4, 3, 3, 12345), (7, 2, 0, None)):
        n = 1 << k
        values = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
        sigmas = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
        beta, gamma = rng.randrange(R), rng.randrange(R)
        w = qr.omega(k)
        got = ob.permutation_product([qu.pack(v) for v in values], [qu.pack(v) for v in sigmas], k, qu.pack([w]),
                                     qu.pack([beta * pow(lr.DELTA, first, R) % R]), qu.pack([lr.DELTA]), qu.pack([beta]), qu.pack([gamma]),
                                     qu.pack([lz]) if lz is not None else None)
        assert qu.unpack(got) == lr.permutation_product(values, sigmas, k, w, beta, gamma, first, lz if lz is not None else 1)


# ---------------------------------------------------------------- external known answers (not minted by this repository)
def _kat_point(x_hex, y_hex):
    from util import fq_limbs
    return np.concatenate([fq_limbs(int(x_hex, 16)), fq_limbs(int(y_hex, 16))])


def test_external_kat_ecmul_ecadd_pin_both_oracles():
    """EIP-196 ecMul / ecAdd vectors (tests/golden/make_external_kat.py): the precompiles the reference's generated
    verifier calls (halo2-snark-aggregator-solidity/templates/verifier.sol:159-215).  An MSM of one pair is ecMul, an
    MSM of two pairs with unit scalars is ecAdd -- through best_multiexp (bucket method), the naive double-and-add and
    the Python big-int group law."""
    kat = golden("external_kat.json")
    one = fr_limbs(1)
    for c in kat["ecmul"]:
        s = int(c["scalar"], 16) % ref.R
        base = _kat_point(c["x"], c["y"])
        want = _kat_point(c["out_x"], c["out_y"])
        assert np.array_equal(affine_of(ob.best_multiexp(fr_limbs(s), base)), want), c["name"]
        assert np.array_equal(affine_of(ob.msm_naive(fr_limbs(s), base)), want), c["name"]
        assert ref.g1_mul(s, (int(c["x"], 16), int(c["y"], 16))) == (int(c["out_x"], 16), int(c["out_y"], 16))
    for c in kat["ecadd"]:
        bases = np.concatenate([_kat_point(c["ax"], c["ay"]), _kat_point(c["bx"], c["by"])])
        want = _kat_point(c["out_x"], c["out_y"])
        assert np.array_equal(affine_of(ob.best_multiexp(np.concatenate([one, one]), bases)), want), c["name"]


def test_moduli_equal_the_reference_templates_constants():
    """q_mod / p_mod as the reference's Solidity template holds them (verifier.sol:41, :144, :292): the oracle's -1
    in either field must be modulus - 1."""
    m = golden("external_kat.json")["moduli"]
    r, p = int(m["q_mod_decimal"]), int(m["p_mod_decimal"])
    assert r == int(m["r_hex"], 16) == ref.R and p == ref.P
    for field, mod in ((0, r), (1, p)):
        lim = fr_limbs if field == 0 else __import__("util").fq_limbs
        minus_one = ob.field_op(field, 1, lim(0), lim(1))          # 0 - 1
        assert np.array_equal(minus_one, lim(mod - 1))
        assert np.array_equal(ob.field_op(field, 0, minus_one, lim(1)), lim(0))  # (mod - 1) + 1 wraps to 0
