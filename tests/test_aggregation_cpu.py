"""W6 / B2 on the CPU: the restated verifier-as-a-schema (oracle/py/verifier_ref.py) drives three chip families over
ONE tiny inner proof (oracle/py/mini_prover.py):
  * plain values (the reference's Mock chips): the proof is accepted -- final pairing holds -- and tampered proofs are not;
  * the circuit-chip oracle (ecc_chip_ref.py): every row satisfies the gate, the range lookups and the copy constraints,
    the final pair equals the plain-value one and the instance cells hold final_pair_to_instances;
  * the product's recording chips (ArithEccChip + ArithFieldChip + Encode over the C ABI): same row layout, same values,
    same cells -- no GPU needed for that (the advice columns themselves are compared in tests/test_gpu_witness.py)."""
import pytest

import aggregation_util as au
import bn254_ref as ref
import ecc_chip_ref as E
import verifier_ref as V
from halo2_snark_aggregator_b200 import fs

R = ref.R


@pytest.fixture(scope="module")
def inner():
    return au.tiny_inner_proof()


@pytest.fixture(scope="module")
def oracle_run(inner):
    chips, ctx = V.ref_chips()
    out = V.synthesize(chips, au.circuits_data(inner))
    return chips, ctx, out


def test_plain_value_chips_accept_the_proof_and_reject_tampering(inner):
    chips = V.mock_chips()
    out = V.synthesize(_NoExpose(chips), au.circuits_data(inner))
    assert au.pairing_holds(inner, out["w_x"], out["w_g"])
    # 37 written points for the aggregation circuit; this tiny one: 2 advice + 2 permuted + 1 perm z (3 columns, chunk 3)
    # + 1 lookup z + 1 random + 4 h + 3 W (x, omega x, omega^-1 x: one permutation set, so no omega^last x)
    assert len(inner["proof"]) == 32 * (2 + 2 + 1 + 1 + 1 + 4 + 3) + 32 * len(__import__("mini_prover").eval_write_order(inner["cs"]))
    for where in (40, len(inner["proof"]) - 200):        # an advice commitment's x, an evaluation
        bad = bytearray(inner["proof"])
        bad[where] ^= 1
        tampered = dict(inner, proof=bytes(bad))
        try:
            o2 = V.synthesize(_NoExpose(V.mock_chips()), au.circuits_data(tampered))
        except AssertionError:
            continue                                       # not a curve point / not a field element any more
        assert not au.pairing_holds(inner, o2["w_x"], o2["w_g"])
    wrong_instance = dict(inner, instances=[[(inner["instances"][0][0] + 1) % R, inner["instances"][0][1]]])
    o3 = V.synthesize(_NoExpose(V.mock_chips()), au.circuits_data(wrong_instance))
    assert not au.pairing_holds(inner, o3["w_x"], o3["w_g"])


class _NoExpose:
    """plain-value run: the second region (limb packing) has no value-level counterpart -- the reference computes the
    instances natively with final_pair_to_instances (verify_circuit.rs:768-804)"""

    def __init__(self, chips):
        self.nchip, self.schip, self.encode = chips.nchip, chips.schip, chips.encode
        self.pchip = _MockPchip(chips.pchip)


class _MockPchip:
    def __init__(self, p):
        self.p = p

    def __getattr__(self, name):
        return getattr(self.p, name)

    def assert_equal(self, a, b): assert a == b
    def assert_not_identity(self, p): assert p is not None
    def expose_final_pair(self, w_x, w_g): return []


def test_circuit_chip_oracle_witness_is_valid_and_matches_plain_values(inner, oracle_run):
    chips, ctx, out = oracle_run
    rows = E.check(ctx)                                     # gate + range lookups + copy constraints on every row
    mock = V.synthesize(_NoExpose(V.mock_chips()), au.circuits_data(inner))
    w_x, w_g = chips.pchip.to_value(out["w_x"]), chips.pchip.to_value(out["w_g"])
    assert (w_x, w_g) == (mock["w_x"], mock["w_g"]) and au.pairing_holds(inner, w_x, w_g)
    # the four exposed cells + the inner proof's instances = final_pair_to_instances (verify_circuit.rs:768-804)
    want = fs.final_pair_to_instances((w_x, w_g, inner["instances"][0]))
    assert [c.value for c in out["instance_cells"]] == want
    # row budget: the reference's estimator says 32196 rows per ecmul (evaluation.rs:132)
    n_points = len([nm for nm in out["names"] if nm])
    assert 0.8 * 32196 * n_points < rows < 1.6 * 32196 * n_points, (rows, n_points)


def test_recording_chips_reproduce_layout_values_and_cells(inner, oracle_run):
    ochips, octx, oout = oracle_run
    chips, w = au.b200_chips()
    out = V.synthesize(chips, au.circuits_data(inner))
    assert w.rows() == octx.offset, "row layout diverged from the reference restatement"
    assert chips.pchip.to_value(out["w_x"]) == ochips.pchip.to_value(oout["w_x"])
    assert chips.pchip.to_value(out["w_g"]) == ochips.pchip.to_value(oout["w_g"])
    assert out["names"] == oout["names"]
    for h, c in zip(out["instance_cells"], oout["instance_cells"]):
        assert chips.schip.to_value(h) == c.value
        assert chips.schip.cell(h) == (c.col, c.row)
    w.close()
