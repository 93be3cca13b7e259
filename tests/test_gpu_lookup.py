"""N3 (sorting half) GPU parity: h2agg_sort_fr / h2agg_permute_expression_pair through the C ABI against the C++ oracle
(restatement of halo2's loop, pinned by oracle/py/lookup_ref.py) and against the properties the reference's verifier
enforces on the permuted columns (halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119)."""
import random

import numpy as np
import pytest

import lookup_ref as lr
import oracle_binding as ob
import quotient_util as qu
from halo2_snark_aggregator_b200 import H2aggError

pytestmark = pytest.mark.gpu
R = qu.R


def _canon_sorted(limbs):
    return sorted(qu.unpack(limbs))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 256, 257, 2047, 2048, 2049, 5000, 70001])
def test_sort_fr_small(ctx, n):
    rng = random.Random(n)
    vals = [rng.randrange(R) for _ in range(n)]
    if n > 4:
        vals[1] = vals[0]               # duplicates
        vals[2], vals[3] = 0, R - 1     # extremes
    got = qu.unpack(ctx.sort_fr(qu.pack(vals)))
    assert got == sorted(vals)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_sort_fr_witness_like_columns(ctx, kind):
    """the synthetic column kinds of SURVEY.md 8d: uniform, a0..a3-like, a4-like, 17-bit (few passes run for the small ones)"""
    n = 1 << 16
    a = ob.gen_scalars(0x5011 + kind, kind, n)
    got = ob.from_mont(0, ctx.sort_fr(a)).reshape(n, 4)
    want = ob.from_mont(0, a).reshape(n, 4)
    order = np.lexsort((want[:, 0], want[:, 1], want[:, 2], want[:, 3]))
    assert np.array_equal(got, want[order])


def test_sort_fr_constant_and_sorted_inputs(ctx):
    n = 4099
    same = qu.pack([12345] * n)
    assert np.array_equal(ctx.sort_fr(same), same)       # every pass skipped
    asc = list(range(n))
    assert qu.unpack(ctx.sort_fr(qu.pack(asc[::-1]))) == asc
    assert ctx.sort_fr(np.zeros(0, dtype=np.uint64)).size == 0


def _lookup_case(rng, u, table_kind):
    if table_kind == "range":          # the aggregation circuit's tables: 0..2^b-1 padded with zeros, 17-bit inputs
        b = max(1, min(17, u.bit_length() - 1))
        table = [i if i < (1 << b) else 0 for i in range(u)]
        inp = [rng.randrange(min(u, 1 << b)) if rng.random() < 0.8 else 0 for _ in range(u)]
    elif table_kind == "wide":         # full-width values, heavy repetition
        pool = [rng.randrange(R) for _ in range(max(1, u // 5))]
        table = [rng.choice(pool) for _ in range(u)]
        inp = [rng.choice(table) for _ in range(u)]
    else:                              # input is a permutation of the table (no repeated rows unless the table repeats)
        table = [rng.randrange(R) for _ in range(u)]
        inp = list(table)
        rng.shuffle(inp)
    return inp, table


@pytest.mark.parametrize("u", [1, 2, 5, 64, 257, 2048, 2049, 10000])
@pytest.mark.parametrize("table_kind", ["range", "wide", "perm"])
def test_permute_expression_pair_matches_oracle(ctx, u, table_kind):
    rng = random.Random(u * 7 + len(table_kind))
    inp, table = _lookup_case(rng, u, table_kind)
    a, s = ctx.permute_expression_pair(qu.pack(inp), qu.pack(table))
    rc, wa, ws = ob.permute_expression_pair(qu.pack(inp), qu.pack(table))
    assert rc == 0
    assert np.array_equal(a, wa) and np.array_equal(s, ws)
    if u <= 2049:
        pa, ps = lr.permute_expression_pair(inp, table)
        assert qu.unpack(a) == pa and qu.unpack(s) == ps
        lr.check_lookup_constraints(inp, table, qu.unpack(a), qu.unpack(s))


def test_missing_input_value_is_an_error(ctx):
    inp, table = [1, 2, 3, 3], [1, 2, 4, 4]
    with pytest.raises(H2aggError) as e:
        ctx.permute_expression_pair(qu.pack(inp), qu.pack(table))
    assert "error 4" in str(e.value) and "ConstraintSystemFailure" in str(e.value)
    assert ob.permute_expression_pair(qu.pack(inp), qu.pack(table))[0] == 1
    # the context stays usable
    a, s = ctx.permute_expression_pair(qu.pack([2, 1]), qu.pack([1, 2]))
    assert qu.unpack(a) == [1, 2] and qu.unpack(s) == [1, 2]


def test_permute_expression_pair_full_size(ctx):
    """BASELINE size: usable rows of a k = 22 column (2^22 - 6): 17-bit inputs against the padded range table, and a
    full-width pair; bit-exact against the oracle, on device buffers."""
    u = (1 << 22) - 6
    n = 1 << 22
    for kind in ("range", "uniform"):
        if kind == "range":
            inp = ob.gen_scalars(0x7711, 3, n)[: 4 * u]                   # uniform 17-bit
            t = np.arange(u, dtype=np.uint64)
            t[t >= (1 << 17)] = 0
            canon = np.zeros((u, 4), dtype=np.uint64)
            canon[:, 0] = t
            table = ob.to_mont(0, canon.reshape(-1))
        else:
            table = ob.gen_scalars(0x7712, 0, n)[: 4 * u]
            perm = np.random.default_rng(5).permutation(u)
            inp = np.ascontiguousarray(table.reshape(u, 4)[perm].reshape(-1))
            inp.reshape(u, 4)[: u // 2] = inp.reshape(u, 4)[u // 2: 2 * (u // 2)]   # half the rows repeat
        d = [ctx.dev_alloc(u * 32) for _ in range(4)]
        ctx.h2d(d[0], inp)
        ctx.h2d(d[1], table)
        ctx.permute_expression_pair_dev(d[0], d[1], u, d[2], d[3])
        a, s = ctx.d2h(d[2], 4 * u), ctx.d2h(d[3], 4 * u)
        rc, wa, ws = ob.permute_expression_pair(inp, table)
        assert rc == 0
        assert np.array_equal(a, wa), kind
        assert np.array_equal(s, ws), kind
        for p in d:
            ctx.dev_free(p)
