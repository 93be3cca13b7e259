"""GPU parity for the witness path (W1-W5): the 5 advice columns produced by the recorder + the
B200 expansion kernel (through the C ABI) equal, bit for bit, the advice cells of the Python
restatement of halo2-ecc-circuit-lib -- whose rows are themselves MockProver-checked."""
import numpy as np
import pytest

import ecc_chip_ref as E
import halo2_snark_aggregator_b200 as h2
import witness_scenarios as ws

pytestmark = pytest.mark.gpu


def _compare(b, cols):
    want = E.advice_columns(b.ctx)
    n = b.ctx.offset
    assert cols.shape[1] >= n
    for c in range(5):
        w = ws.fr_mont_rows(want[c])
        got = cols[c, :n]
        if not np.array_equal(got, w):
            bad = np.nonzero((got != w).any(axis=1))[0]
            raise AssertionError("advice column %d differs at %d rows, first row %d" % (c, bad.size, bad[0]))
    assert not cols[:, n:].any()  # rows past the layout stay zero


@pytest.mark.parametrize("name", ws.SCENARIOS)
def test_advice_columns_match_reference_restatement(ctx, name):
    chip = h2.B200EccChip()
    b, res = ws.run(name, chip)
    E.check(b.ctx)  # the oracle's rows satisfy gate + lookups + copy constraints
    assert chip.rows() == b.ctx.offset
    cols = chip.expand(ctx, n_rows=b.ctx.offset + 17)
    _compare(b, cols)
    chip.close()


def test_expand_dev_feeds_commit_without_leaving_hbm(ctx):
    """Witness columns stay in HBM and go straight into the MSM (commit_lagrange of an advice column)."""
    import oracle_binding as ob

    chip = h2.B200EccChip()
    b, _ = ws.run("multi_exp_1", chip)
    n = 1 << 17
    assert chip.rows() <= n
    d_cols = [ctx.dev_alloc(n * 32) for _ in range(5)]
    d_b = ctx.dev_alloc(n * 64)
    d_o = ctx.dev_alloc(5 * 160)
    try:
        chip.expand_dev(ctx, d_cols, n)
        ctx.synth_bases_dev(0x53525300, 0, n, d_b)
        ctx.msm_g1_batch_dev(d_cols, n, d_o, d_bases=d_b)
        ctx.synchronize()
        got = ctx.d2h(d_o, 100).reshape(5, 20)
        bases = ob.gen_bases(0x53525300, n)
        want_cols = E.advice_columns(b.ctx)
        for c in (0, 4):
            col = np.zeros((n, 4), dtype=np.uint64)
            col[: b.ctx.offset] = ws.fr_mont_rows(want_cols[c])
            assert np.array_equal(ctx.d2h(d_cols[c], 4 * n).reshape(n, 4), col)
            assert np.array_equal(got[c, 8:], ob.best_multiexp(np.ascontiguousarray(col).ravel(), bases))
    finally:
        for d in d_cols + [d_b, d_o]:
            ctx.dev_free(d)
        chip.close()


def test_multi_exp_22_points_realistic_w_g(ctx):
    """The size of one proof's `w_g` multi_exp in the reference's own example:
    22 points -> 711 724 rows, every advice cell bit-exact, result = native MSM."""
    import random

    chip = h2.B200EccChip()
    b = ws.Both(chip)
    res, want, rows = ws.scenario_multi_exp_n(b, random.Random(22), 22)
    assert rows == 711724  # SURVEY.md App. G row model
    assert b.value(res) == want
    xy, ident = chip.to_value(res[1])
    assert not ident and np.array_equal(xy, ws.xy_mont(want))
    assert chip.rows() == b.ctx.offset
    cols = chip.expand(ctx)
    _compare(b, cols)
    chip.close()


def test_witness_kernel_feeds_the_resident_prover(ctx):
    """W -> K hand-off without host memory: the recorder's op list is expanded straight into the prover's resident
    Lagrange slots, and round 1 (commit + lagrange_to_coeff + coeff_to_extended) runs from there."""
    import oracle_binding as ob
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.prover import ResidentProver
    from util import domain_consts

    chip = h2.B200EccChip()
    b, _ = ws.run("multi_exp_1", chip)
    k = 17
    n = 1 << k
    assert chip.rows() <= n
    bases = ob.gen_bases(0x53525311, n)
    sid = ctx.srs_register(bases)
    pr = ResidentProver(ctx, plonk.aggregation_circuit_cs(), k, sid, sid)
    names = [("advice", i) for i in range(5)]
    chip.expand_dev(ctx, [pr.lagrange_slot(nm) for nm in names], n)
    comms = pr.commit_device_columns(names)
    d = domain_consts(k, pr.ext_k)
    want_cols = E.advice_columns(b.ctx)
    for c in (1, 4):
        col = np.zeros((n, 4), dtype=np.uint64)
        col[: b.ctx.offset] = ws.fr_mont_rows(want_cols[c])
        col = np.ascontiguousarray(col).ravel()
        assert np.array_equal(comms[c], ob.best_multiexp(col, bases)[:8])
        co = ob.ifft(col.copy(), d["omega_inv"], d["n_inv"], k)
        assert np.array_equal(ctx.d2h(pr.coeff[names[c]], 4 * n), co)
        assert np.array_equal(ctx.d2h(pr.ext[names[c]], 4 << pr.ext_k), ob.coeff_to_extended(co, k, pr.ext_k, d["zeta"], d["omega_ext"]))
    pr.close()
    ctx.srs_release(sid)
    chip.close()


def test_full_aggregation_witness_all_five_advice_columns(ctx):
    """W6: the advice columns of a COMPLETE aggregation witness -- Poseidon transcript + ScalarChip expression evaluation +
    instance commitment (scalar_mul_constant) + both multi_exps + the final-pair packing -- for one tiny inner proof
    (oracle/py/mini_prover.py), driven by the restated verify_aggregation_proofs_in_chip (oracle/py/verifier_ref.py,
    api/src/systems/halo2/verify.rs:835-942) through the product's ArithEccChip / ArithFieldChip / Encode chips, against the
    same driver over the circuit-chip oracle.  ~0.9 M rows; the oracle's rows pass gate + lookups + copy constraints; and
    the columns then go through commit_lagrange on the device like any advice column."""
    import aggregation_util as au
    import oracle_binding as ob
    import verifier_ref as V

    inner = au.tiny_inner_proof()
    ochips, octx = V.ref_chips()
    oout = V.synthesize(ochips, au.circuits_data(inner))
    E.check(octx)
    chips, w = au.b200_chips()
    out = V.synthesize(chips, au.circuits_data(inner))
    assert w.rows() == octx.offset
    k = 20
    n = 1 << k
    assert w.rows() <= n - 6
    cols = w.expand(ctx, n_rows=n)

    class _B:
        pass

    b = _B()
    b.ctx = octx
    _compare(b, cols)
    # the exposed cells really hold what constrain_instance binds (verify_circuit.rs:357-367)
    rinv = pow(1 << 256, -1, E.R)
    for h, c in zip(out["instance_cells"], oout["instance_cells"]):
        col, row = chips.schip.cell(h)
        limbs = cols[col, row]
        got = sum(int(x) << (64 * i) for i, x in enumerate(limbs)) * rinv % E.R
        assert got == c.value == chips.schip.to_value(h)
    # and one column through the commit path: commit_lagrange(a4) on the device == oracle best_multiexp
    bases = ob.gen_bases(0x53525320, n)
    a4 = np.ascontiguousarray(cols[4]).ravel()
    assert np.array_equal(ctx.msm_g1(a4, bases), ob.best_multiexp(a4, bases))
    w.close()


def test_multi_exp_columns_do_not_depend_on_host_threads(ctx):
    """Sections of a multi_exp recorded on 1 / 8 host threads (chunked, page-locked record store) expand to the same
    five advice columns, bit for bit."""
    import random

    import bn254_ref as ref

    lib = h2._lib.load()
    rng = random.Random(23)
    pts = [ws.xy_mont(ref.g1_mul(rng.randrange(1, ref.R), ref.G1_GEN)) for _ in range(6)]
    scalars = [rng.randrange(ref.R) for _ in range(6)]
    prev = lib.h2agg_wit_set_threads(0)
    cols = []
    try:
        for threads in (1, 8):
            lib.h2agg_wit_set_threads(threads)
            chip = h2.B200EccChip()
            hp = [chip.assign_var(p) for p in pts]
            hs = [chip.assign_scalar(s) for s in scalars]
            chip.multi_exp(hp, hs)
            cols.append(chip.expand(ctx, n_rows=chip.rows() + 5))
            chip.close()
    finally:
        lib.h2agg_wit_set_threads(prev)
    assert cols[0].shape == cols[1].shape and np.array_equal(cols[0], cols[1])
    assert cols[0][:, :-5].any()
