"""N1 GPU parity: h2agg_evaluate_h_dev (through the C ABI) against the big-int oracle, which is itself pinned to
the reference's verifier equations (tests/test_quotient_cpu.py).  evaluate_h is pointwise in the coset row, so
random extended columns exercise it fully; at the BASELINE size (k = 22, 2^24 rows, 55 resident columns) sampled
rows are recomputed by the oracle from values read back from HBM."""
import ctypes
import random

import numpy as np
import pytest

import quotient_util as qu
import quotient_ref as qr
from halo2_snark_aggregator_b200 import H2aggError, plonk
from halo2_snark_aggregator_b200._lib import QuotientArgs, c_vp

pytestmark = pytest.mark.gpu
R = qr.R


def _run(ctx, cs, k, seed, divide):
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    size = 1 << ext_k
    rng = random.Random(seed)
    ext = {name: [rng.randrange(R) for _ in range(size)] for name in plan.columns}
    # a few structured columns: zeros, ones, r - 1
    names = list(plan.columns)
    ext[names[0]] = [0] * size
    ext[names[1]] = [1] * size
    ext[names[2]] = [R - 1] * size
    y, beta, gamma, theta = [rng.randrange(R) for _ in range(4)]
    d_cols = []
    for name in plan.columns:
        a = qu.pack(ext[name])
        p = ctx.dev_alloc(a.nbytes)
        ctx.h2d(p, a)
        d_cols.append(p)
    d_out = ctx.dev_alloc(size * 32)
    ctx.evaluate_h_dev(plan, d_cols, k, ext_k, qu.pack([y]), qu.pack([beta]), qu.pack([gamma]), qu.pack([theta]), d_out,
                       divide=divide)
    got = qu.unpack(ctx.d2h(d_out, 4 * size))
    for p in d_cols + [d_out]:
        ctx.dev_free(p)
    want = qr.evaluate_h(qu.oracle_desc(cs), ext, k, ext_k, y, beta, gamma, theta)
    if divide:
        want = qr.divide_by_vanishing_poly(want, k, ext_k)
    return got, want


@pytest.mark.parametrize("k,divide", [(3, True), (5, True), (5, False), (9, True)])
def test_aggregation_circuit_quotient_matches_oracle(ctx, k, divide):
    got, want = _run(ctx, plonk.aggregation_circuit_cs(), k, 1000 + k, divide)
    assert got == want


def test_generic_constraint_system(ctx):
    E = plonk.Expression
    cs = plonk.ConstraintSystem(num_fixed=3, num_advice=3, num_instance=1)
    cs.create_gate("g0", [E.fixed(0) * (E.advice(0) * E.advice(1) - E.advice(2, 1)),
                          E.fixed(1) * (E.advice(0) + 5) * (E.advice(0, -1) - E.constant(3))])
    cs.create_gate("g1", [E.instance(0) - E.advice(2) * 7])
    cs.lookup("two columns", [(E.advice(0) * E.fixed(0), E.fixed(2)), (E.advice(1) + E.constant(1), E.fixed(1, 1))])
    cs.enable_equality("advice", 0)
    cs.enable_equality("fixed", 2)
    got, want = _run(ctx, cs, 6, 31, True)
    assert got == want


def test_no_permutation_no_lookup(ctx):
    E = plonk.Expression
    cs = plonk.ConstraintSystem(num_fixed=1, num_advice=2, num_instance=0)
    cs.create_gate("mul", [E.fixed(0) * (E.advice(0) * E.advice(0, 1) - E.advice(1))])
    got, want = _run(ctx, cs, 4, 8, True)
    assert got == want


def _raw_args(plan, words, d_cols, k, ext_k, keep):
    vals = [qu.pack([3]), qu.pack([5]), qu.pack([7]), qu.pack([11]),
            plonk.fr_mont(qr.omega(ext_k)), plonk.fr_mont(plonk.ZETA), plonk.fr_mont(plonk.DELTA)]
    cols = (c_vp * len(d_cols))(*d_cols)
    keep += vals + [cols, words]
    a = QuotientArgs()
    a.k, a.ext_k = k, ext_k
    a.plan, a.n_plan_words = words.ctypes.data, words.size
    a.d_columns, a.n_columns = ctypes.cast(cols, c_vp).value, len(d_cols)
    a.consts, a.n_consts = (plan.consts.ctypes.data if plan.consts.size else None), plan.consts.size // 4
    a.y, a.beta, a.gamma, a.theta, a.omega_ext, a.zeta, a.delta = [v.ctypes.data for v in vals]
    a.t_evaluations, a.t_len = None, 0
    return a


def test_invalid_plans_are_rejected(ctx):
    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    k, ext_k = 3, 5
    buf = ctx.dev_alloc(32 << ext_k)
    d_cols = [buf] * len(plan.columns)
    d_out = ctx.dev_alloc(32 << ext_k)
    keep = []

    def call(words, cols=d_cols):
        a = _raw_args(plan, np.ascontiguousarray(words, dtype=np.uint32), cols, k, ext_k, keep)
        return ctx.lib.h2agg_evaluate_h_dev(ctx.h, ctypes.byref(a), c_vp(d_out))

    assert call(plan.words) == 0
    bad = plan.words.copy()
    bad[0] ^= 1
    assert call(bad) == 1
    assert call(plan.words[:-1]) == 1                               # truncated
    assert call(np.append(plan.words, np.uint32(0))) == 1           # trailing words
    bad = plan.words.copy()
    bad[6] = len(plan.columns)                                      # l_0 column out of range
    assert call(bad) == 1
    bad = plan.words.copy()
    bad[-1] = 60000                                                 # lookup column out of range
    assert call(bad) == 1
    assert call(plan.words, d_cols[:-1]) == 1                       # fewer columns than the plan uses
    assert b"evaluate_h" in ctx.lib.h2agg_last_error(ctx.h)
    ctx.synchronize()
    ctx.dev_free(buf)
    ctx.dev_free(d_out)
    with pytest.raises(H2aggError):
        ctx.check(1)


def test_poly_fold_matches_horner(ctx):
    rng = random.Random(4)
    n, m = 1000, 70  # more than one launch worth of polynomials
    polys = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    v = rng.randrange(R)
    d = []
    for p in polys:
        a = qu.pack(p)
        q = ctx.dev_alloc(a.nbytes)
        ctx.h2d(q, a)
        d.append(q)
    d_out = ctx.dev_alloc(n * 32)
    ctx.poly_fold_dev(d, n, qu.pack([v]), d_out)
    got = qu.unpack(ctx.d2h(d_out, 4 * n))
    want = [0] * n
    for p in polys:
        want = [(w * v + c) % R for w, c in zip(want, p)]
    assert got == want
    for q in d + [d_out]:
        ctx.dev_free(q)


def test_full_size_k22_sampled_rows(ctx):
    """BASELINE size: 55 extended columns of 2^24 rows resident in HBM (27.5 GB); 48 sampled rows (incl. the wrap-
    around rows at both ends) are recomputed by the oracle from the values the kernel read."""
    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    k, ext_k = 22, 24
    size = 1 << ext_k
    d_cols = []
    for i, _ in enumerate(plan.columns):
        p = ctx.dev_alloc(size * 32)
        ctx.synth_scalars_dev(0x9100 + i, 0, 0, size, p)
        d_cols.append(p)
    d_out = ctx.dev_alloc(size * 32)
    rng = random.Random(22)
    y, beta, gamma, theta = [rng.randrange(R) for _ in range(4)]
    ctx.evaluate_h_dev(plan, d_cols, k, ext_k, qu.pack([y]), qu.pack([beta]), qu.pack([gamma]), qu.pack([theta]), d_out)
    ctx.synchronize()
    rows = [0, 1, 3, 4, 23, 24, size - 1, size - 4, size - 5, size - 24] + [rng.randrange(size) for _ in range(38)]

    class Fetch:
        def __init__(self, p):
            self.p, self.cache = p, {}

        def __getitem__(self, idx):
            if idx not in self.cache:
                self.cache[idx] = qu.unpack(ctx.d2h(self.p + 32 * idx, 4))[0]
            return self.cache[idx]

    cols = {name: Fetch(p) for name, p in zip(plan.columns, d_cols)}
    want = qr.evaluate_h(qu.oracle_desc(cs), cols, k, ext_k, y, beta, gamma, theta, rows=rows)
    t = plonk.t_evaluations(k, ext_k)
    out = Fetch(d_out)
    for r, w in zip(rows, want):
        assert out[r] == w * t[r % len(t)] % R, "row %d" % r
    for p in d_cols + [d_out]:
        ctx.dev_free(p)
