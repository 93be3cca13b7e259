"""CPU suite for the witness path: (1) the Python oracle is a faithful chip restatement -- its rows
satisfy the gate polynomial, range lookups and copy constraints (mini-MockProver), its results equal
native G1 arithmetic and its row counts reproduce SURVEY.md App. G / the reference's estimator;
(2) the product's host recorder reproduces the oracle's row layout op for op (no GPU needed for
that); (3) kernel constants."""
import os
import random
import re

import numpy as np
import pytest

import bn254_ref as ref
import ecc_chip_ref as E
import halo2_snark_aggregator_b200 as h2
import witness_scenarios as ws

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integer_chip_vs_native_fq_and_row_counts():
    rng = random.Random(3)
    ctx = E.Context()
    counts = {}
    for _ in range(6):
        x, y = rng.randrange(E.P), rng.randrange(E.P)
        a, b = E.assign_w(ctx, x), E.assign_w(ctx, y)
        o = ctx.offset; r = E.int_mul(ctx, a, b); counts["mul"] = ctx.offset - o
        assert r.w() == x * y % E.P
        a = E.assign_w(ctx, x); o = ctx.offset; r = E.int_square(ctx, a); counts["square"] = ctx.offset - o
        assert r.w() == x * x % E.P
        a, b = E.assign_w(ctx, x), E.assign_w(ctx, y)
        o = ctx.offset; z, c = E.int_div(ctx, a, b); counts["div"] = ctx.offset - o
        assert c.w() * y % E.P == x and z.value == 0
        a, b = E.assign_w(ctx, x), E.assign_w(ctx, y)
        s = E.int_add(ctx, a, b); assert s.w() == (x + y) % E.P
        d = E.int_sub(ctx, a, b); assert d.w() == (x - y) % E.P
        n = E.int_neg(ctx, a); assert n.w() == (-x) % E.P
        o = ctx.offset; E.reduce(ctx, s); counts["reduce"] = ctx.offset - o
        assert s.w() == (x + y) % E.P and s.overflows == 0
        a = E.assign_w(ctx, x); o = ctx.offset; z = E.int_is_zero(ctx, a); counts["is_zero"] = ctx.offset - o
        assert z.value == 0
    z0 = E.assign_w(ctx, 0)
    assert E.int_is_zero(ctx, z0).value == 1
    a, b = E.assign_w(ctx, 5), E.assign_w(ctx, 0)  # division by zero: c = 0, flag set (:745-782)
    zf, c = E.int_div(ctx, a, b)
    assert zf.value == 1 and c.w() == 0
    a = E.assign_w(ctx, 12345); bit = E.int_get_last_bit(ctx, a); assert bit.value == 1
    big = E.assign_w(ctx, E.P - 1)
    acc = big
    for _ in range(40):  # drive overflows past the threshold: conditional reduce must kick in
        acc = E.int_add(ctx, acc, big)
        assert acc.overflows < E.OVERFLOW_LIMIT
    assert acc.w() == 41 * (E.P - 1) % E.P
    assert counts == {"mul": 31, "square": 30, "div": 47, "reduce": 9, "is_zero": 12}
    E.check(ctx)


@pytest.mark.parametrize("name", ws.SCENARIOS)
def test_oracle_scenarios_mockprover_and_native_results(name):
    b, res = ws.run(name)
    for handle, want in res:
        assert b.value(handle) == want
    E.check(b.ctx)


def test_oracle_row_counts_match_reference_estimator():
    """shamir with 1 / 2 points = 73 463 / 110 064 rows, i.e. the marginal cost per
    point brackets the reference's hard-coded ecmul_rows = 32196 (evaluation.rs:132)."""
    rows = {}
    for n in (1, 2):
        rng = random.Random(n)
        ctx = E.Context()
        pts = [E.assign_point(ctx, ref.g1_mul(rng.randrange(1, ref.R), ref.G1_GEN)) for _ in range(n)]
        sc = [E.bg_assign(ctx, rng.randrange(ref.R)) for _ in range(n)]
        o = ctx.offset
        E.ecc_shamir(ctx, pts, sc)
        rows[n] = ctx.offset - o
    assert rows == {1: 73463, 2: 110064}
    assert 30000 < rows[2] - rows[1] < 40000


@pytest.mark.parametrize("name", ws.SCENARIOS)
def test_recorder_reproduces_oracle_layout_and_values(name):
    chip = h2.B200EccChip()
    b, res = ws.run(name, chip)
    assert chip.rows() == b.ctx.offset, "row layout diverged from the reference restatement"
    for (o, h), want in res:
        xy, ident = chip.to_value(h)
        if want is None:
            assert ident
        else:
            assert not ident and np.array_equal(xy, ws.xy_mont(want))
    assert chip.ops() > 0
    chip.close()


def test_recorder_rejects_point_off_curve():
    chip = h2.B200EccChip()
    with pytest.raises(h2.H2aggError):
        chip.assign_var(ws.xy_mont((1, 3)))
    chip.close()


def test_kernel_constants():
    src = open(os.path.join(ROOT, "halo2_snark_aggregator_b200", "csrc", "witness.cu")).read()

    def words(name):
        m = re.search(name + r"\[[^=]*=\s*\{(.*?)\};", src, re.S)
        vals = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", m.group(1))]
        return vals

    def val(ws_):
        return sum(w << (32 * i) for i, w in enumerate(ws_))

    assert val(words("W_PINV288")) == pow(E.P, -1, 1 << 288)
    negw = words("W_NEGW")
    want = E.Helper.bn_to_limb_le(E.Helper.integer_modulus - E.P)
    assert [val(negw[3 * i:3 * i + 3]) for i in range(4)] == want
    assert val(words("W_P_LIMB0")) == E.P % (1 << 68)
    assert val(words("W_NATIVE")) == E.P % E.R
    assert val(words("W_INV_2_136")) == pow(1 << 136, -1, E.R)


def test_codec_square_root_exponent_constant():
    src = open(os.path.join(ROOT, "halo2_snark_aggregator_b200", "csrc", "codec.cu")).read()
    m = re.search(r"FQ_SQRT_EXP\[8\]\s*=\s*\{(.*?)\};", src, re.S)
    words = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", m.group(1))]
    assert sum(w << (32 * i) for i, w in enumerate(words)) == (E.P + 1) // 4 and E.P % 4 == 3


def test_recorder_multi_exp_layout_does_not_depend_on_host_threads():
    """The candidate tables and the inner window sums of a multi_exp are recorded on host threads and stitched in at
    their row offsets: rows, record count, result and the cells of the final pair must not depend on the thread count."""
    import ctypes

    lib = h2._lib.load()
    rng = random.Random(11)
    pts = [ref.g1_mul(rng.randrange(1, ref.R), ref.G1_GEN) for _ in range(5)]
    pts[3] = pts[1]                                     # P + P inside the inner sums
    scalars = [rng.randrange(ref.R) for _ in range(5)]
    scalars[2] = 0
    seen = []
    prev = lib.h2agg_wit_set_threads(0)
    try:
        for threads in (1, 3, 8):
            lib.h2agg_wit_set_threads(threads)
            chip = h2.B200EccChip()
            hp = [chip.assign_var(ws.xy_mont(p)) for p in pts]
            hs = [chip.assign_scalar(s) for s in scalars]
            r = chip.multi_exp(hp, hs)
            r2 = chip.multi_exp(hp[:2], hs[:2])
            cells = [chip.scalar_chip.cell(h) for h in chip.expose_final_pair(r2, r)]
            xy, ident = chip.to_value(r)
            seen.append((chip.rows(), chip.ops(), xy.tobytes(), ident, tuple(cells)))
            chip.close()
    finally:
        lib.h2agg_wit_set_threads(prev)
    assert seen[0] == seen[1] == seen[2]
    want = None
    for p, s in zip(pts, scalars):
        want = ref.g1_add(want, ref.g1_mul(s, p))
    assert np.array_equal(np.frombuffer(seen[0][2], dtype=np.uint64), ws.xy_mont(want))


def test_host_modular_inverse_through_div():
    """ScalarChip::div goes through the safegcd inverse (csrc/host_modinv.hpp): a / b * b == a for edge values."""
    from halo2_snark_aggregator_b200.witness import B200Context, B200ScalarChip

    w = B200Context()
    s = B200ScalarChip(w)
    rng = random.Random(5)
    vals = [1, 2, ref.R - 1, ref.R - 2, (1 << 253) + 5, 1 << 62, (1 << 62) - 1, (1 << 124) + 1] + [rng.randrange(1, ref.R) for _ in range(200)]
    for b in vals:
        a = rng.randrange(ref.R)
        q = s.to_value(s.div(s.assign_var(a), s.assign_var(b)))
        assert q == a * pow(b, -1, ref.R) % ref.R
    w.close()


def test_compiled_host_driver_records_the_same_layout_as_the_python_chips(tmp_path):
    """tests/cpp/witness_main.cpp drives the aggregation op mix through the C ABI from C++; the Python twin
    (witness_workload.py) makes the same calls through ctypes: same rows, same record count (no GPU needed)."""
    import json
    import subprocess

    import oracle_binding as ob
    from halo2_snark_aggregator_b200.witness import B200Context, B200EccChip, B200EncodeChip, B200ScalarChip
    from halo2_snark_aggregator_b200.witness_workload import record_aggregation_like

    exe = os.path.join(ROOT, "tests", "cpp", "witness_main")
    if not os.path.exists(exe):   # built by __graft_entry__.build(); a fresh checkout that only built the library gets it here
        lib_dir = os.path.join(ROOT, "halo2_snark_aggregator_b200")
        subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), exe + ".cpp", "-o", exe,
                               "-L" + lib_dir, "-lh2agg", "-Wl,-rpath," + lib_dir])
    pts = ob.gen_bases(0x77, 2048)
    path = tmp_path / "pts.bin"
    pts.tofile(str(path))
    r = subprocess.run([exe, str(path), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout.strip().split("\n")[-1])
    w = B200Context()
    record_aggregation_like(B200ScalarChip(w), B200EccChip(w), B200EncodeChip(w), lambda i: pts.reshape(-1, 8)[i % 2048], 1)
    assert (got["rows"], got["op_records"]) == (w.rows(), w.ops())
    w.close()


def test_recorder_rejects_an_empty_multi_exp_and_bad_handles():
    chip = h2.B200EccChip()
    with pytest.raises(h2.H2aggError):
        chip.multi_exp([], [])
    p = chip.assign_var(ws.xy_mont(ref.G1_GEN))
    with pytest.raises(h2.H2aggError):
        chip.multi_exp([p], [12345])          # no such scalar handle
    with pytest.raises(h2.H2aggError):
        chip.add(p, 99)                       # no such point handle
    chip.close()


def test_scalar_mul_constant_values_with_the_batched_constant_table():
    """constant_mul builds its 127 x {B, 2B, 3B} table on a Jacobian chain with batched inversions: results equal s * B for
    random and edge scalars, the identity base gives the identity, and the row count does not depend on the values."""
    rng = random.Random(77)
    rows = set()
    for s in [0, 1, 2, 3, ref.R - 1, 1 << 200] + [rng.randrange(ref.R) for _ in range(4)]:
        base = ref.g1_mul(rng.randrange(1, ref.R), ref.G1_GEN)
        chip = h2.B200EccChip()
        hs = chip.assign_scalar(s)
        o = chip.rows()
        h = chip.scalar_mul_constant(hs, ws.xy_mont(base))
        rows.add(chip.rows() - o)
        xy, ident = chip.to_value(h)
        want = ref.g1_mul(s, base)
        if want is None:
            assert ident
        else:
            assert not ident and np.array_equal(xy, ws.xy_mont(want))
        chip.close()
    assert len(rows) == 1
    chip = h2.B200EccChip()
    h = chip.scalar_mul_constant(chip.assign_scalar(12345), np.zeros(8, dtype=np.uint64))   # identity base
    assert chip.to_value(h)[1]
    chip.close()
