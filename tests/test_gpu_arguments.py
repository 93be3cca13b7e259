"""N3 (row-wise steps) GPU parity: h2agg_compress_expressions_dev / h2agg_lookup_product_dev /
h2agg_permutation_product_dev through the C ABI against the Python big-int restatement of halo2's loops
(oracle/py/lookup_ref.py), plus the telescoping property z[usable_rows] = 1 at the BASELINE size."""
import random

import numpy as np
import pytest

import lookup_ref as lr
import oracle_binding as ob
import quotient_util as qu
import quotient_ref as qr
from halo2_snark_aggregator_b200 import H2aggError, plonk
from util import fr_limbs

pytestmark = pytest.mark.gpu
R = qu.R


def _upload(ctx, vals):
    a = qu.pack(vals)
    p = ctx.dev_alloc(a.nbytes)
    ctx.h2d(p, a)
    return p


@pytest.mark.parametrize("k", [1, 4, 7])
def test_compress_expressions_matches_oracle(ctx, k):
    E = plonk.Expression
    n = 1 << k
    rng = random.Random(k)
    names = [("advice", 0), ("advice", 1), ("fixed", 0), ("fixed", 1), ("instance", 0)]
    cols = {nm: [rng.randrange(R) for _ in range(n)] for nm in names}
    index = {nm: i for i, nm in enumerate(names)}
    exprs = [E.advice(0) * E.fixed(0), E.advice(1, 1) + E.constant(7), E.fixed(1, -1) * E.advice(0, 2) * 5 - E.instance(0),
             E.constant(0) + E.advice(0)]
    prog = plonk.ExpressionList(exprs, index)
    d_cols = [_upload(ctx, cols[nm]) for nm in names]
    d_out = ctx.dev_alloc(n * 32)
    theta = rng.randrange(R)
    for sub in (exprs, exprs[:1]):
        prog = plonk.ExpressionList(sub, index)
        ctx.compress_expressions_dev(prog.words, prog.consts, d_cols, k, fr_limbs(theta), d_out)
        got = qu.unpack(ctx.d2h(d_out, 4 * n))
        assert got == lr.compress_expressions([e.to_tuple() for e in sub], cols, n, theta)
    # invalid lists are rejected before anything is launched
    bad = prog.words.copy()
    bad[-1] = 99
    with pytest.raises(H2aggError):
        ctx.compress_expressions_dev(bad, prog.consts, d_cols, k, fr_limbs(theta), d_out)
    with pytest.raises(H2aggError):
        ctx.compress_expressions_dev(prog.words[:-1], prog.consts, d_cols, k, fr_limbs(theta), d_out)
    for p in d_cols + [d_out]:
        ctx.dev_free(p)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, 5000])
def test_lookup_product_matches_oracle(ctx, n):
    rng = random.Random(n)
    A, S, Ap, Sp = ([rng.randrange(R) for _ in range(n)] for _ in range(4))
    beta, gamma = rng.randrange(R), rng.randrange(R)
    d = [_upload(ctx, v) for v in (A, S, Ap, Sp)]
    d_z = ctx.dev_alloc(n * 32)
    ctx.lookup_product_dev(*d, n, fr_limbs(beta), fr_limbs(gamma), d_z)
    assert qu.unpack(ctx.d2h(d_z, 4 * n)) == lr.lookup_product(A, S, Ap, Sp, beta, gamma)
    for p in d + [d_z]:
        ctx.dev_free(p)


@pytest.mark.parametrize("k,m,first", [(1, 1, 0), (3, 2, 0), (5, 3, 3), (7, 4, 6), (10, 3, 0)])
def test_permutation_product_matches_oracle(ctx, k, m, first):
    n = 1 << k
    rng = random.Random(100 * k + m)
    values = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    sigmas = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    beta, gamma, last_z = rng.randrange(R), rng.randrange(R), rng.randrange(R)
    w = qr.omega(k)
    dv = [_upload(ctx, v) for v in values]
    ds = [_upload(ctx, v) for v in sigmas]
    d_last = _upload(ctx, [last_z])
    d_z = ctx.dev_alloc(n * 32)
    for lz, dl in ((1, 0), (last_z, d_last)):
        ctx.permutation_product_dev(dv, ds, k, fr_limbs(w), fr_limbs(beta * pow(plonk.DELTA, first, R)), fr_limbs(plonk.DELTA),
                                    fr_limbs(beta), fr_limbs(gamma), dl, d_z)
        assert qu.unpack(ctx.d2h(d_z, 4 * n)) == lr.permutation_product(values, sigmas, k, w, beta, gamma, first, lz)
    for p in dv + ds + [d_last, d_z]:
        ctx.dev_free(p)


def test_identity_permutation_gives_constant_z(ctx):
    """sigma_j(omega^i) = delta^j omega^i (nothing is copy-constrained): every fraction is 1."""
    k, m = 6, 3
    n = 1 << k
    rng = random.Random(2)
    w = qr.omega(k)
    values = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    sigmas = [[pow(plonk.DELTA, j, R) * pow(w, i, R) % R for i in range(n)] for j in range(m)]
    beta, gamma = rng.randrange(R), rng.randrange(R)
    dv = [_upload(ctx, v) for v in values]
    ds = [_upload(ctx, v) for v in sigmas]
    d_z = ctx.dev_alloc(n * 32)
    ctx.permutation_product_dev(dv, ds, k, fr_limbs(w), fr_limbs(beta), fr_limbs(plonk.DELTA), fr_limbs(beta), fr_limbs(gamma), 0, d_z)
    assert qu.unpack(ctx.d2h(d_z, 4 * n)) == [1] * n
    for p in dv + ds + [d_z]:
        ctx.dev_free(p)


def test_lookup_argument_full_size_telescopes(ctx):
    """BASELINE size (k = 22): 17-bit inputs against a 17-bit table; the permuted pair comes from
    h2agg_permute_expression_pair_dev, random blinding rows are appended, and the running product must return to 1 at
    row usable_rows (what the verifier's l_last * (z^2 - z) and the product rule enforce)."""
    k = 22
    n = 1 << k
    u = n - 6
    d_a, d_s, d_ap, d_sp, d_z = (ctx.dev_alloc(n * 32) for _ in range(5))
    ctx.synth_scalars_dev(0x8801, 3, 0, n, d_a)
    ctx.synth_scalars_dev(0x8802, 3, 0, n, d_s)
    ctx.synth_scalars_dev(0x8803, 0, 0, n, d_ap)   # the tails stay random: blinding rows
    ctx.synth_scalars_dev(0x8804, 0, 0, n, d_sp)
    ctx.permute_expression_pair_dev(d_a, d_s, u, d_ap, d_sp)
    beta, gamma = 3 ** 120 % R, 5 ** 100 % R
    ctx.lookup_product_dev(d_a, d_s, d_ap, d_sp, n, fr_limbs(beta), fr_limbs(gamma), d_z)
    assert qu.unpack(ctx.d2h(d_z, 4)) == [1]
    assert qu.unpack(ctx.d2h(d_z + 32 * u, 4)) == [1]
    assert qu.unpack(ctx.d2h(d_z + 32 * (u // 2), 4)) != [1]
    # and the rows of the permuted pair satisfy the verifier's lookup rule on a sample
    ap = qu.unpack(ctx.d2h(d_ap + 32 * 1000, 4 * 64))
    sp = qu.unpack(ctx.d2h(d_sp + 32 * 1000, 4 * 64))
    for i in range(1, 64):
        assert ap[i] == sp[i] or ap[i] == ap[i - 1]
    for p in (d_a, d_s, d_ap, d_sp, d_z):
        ctx.dev_free(p)


@pytest.mark.parametrize("n,m", [(64, 1), (1000, 3), (4096, 11)])
def test_lookup_products_batch_equals_single_calls(ctx, n, m):
    """h2agg_lookup_products_dev (lookup i on lane i mod 8, private scratch per lane) == m single calls == the oracle;
    m = 11 makes lanes carry two products each."""
    rng = random.Random(n + m)
    beta, gamma = rng.randrange(R), rng.randrange(R)
    cols = [[[rng.randrange(R) for _ in range(n)] for _ in range(4)] for _ in range(m)]
    d = [[_upload(ctx, v) for v in c] for c in cols]
    dz = [ctx.dev_alloc(n * 32) for _ in range(m)]
    for _ in range(2):   # twice: the second round reuses the lanes' scratch while nothing else is pending
        ctx.lookup_products_dev([c[0] for c in d], [c[1] for c in d], [c[2] for c in d], [c[3] for c in d], n, fr_limbs(beta),
                                fr_limbs(gamma), dz)
        for i in range(m):
            assert qu.unpack(ctx.d2h(dz[i], 4 * n)) == lr.lookup_product(*cols[i], beta, gamma), i
    for c in d:
        for p in c:
            ctx.dev_free(p)
    for p in dz:
        ctx.dev_free(p)
