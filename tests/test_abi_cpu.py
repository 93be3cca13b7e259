"""CPU suite: the C-ABI library loads and exports every symbol include/h2agg.h declares; host-side
constants agree with Python big ints; the PTX generator's self-check passes; without a GPU the
product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

import halo2_snark_aggregator_b200 as h2
from util import P_MOD, R_MOD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(h2.LIB_PATH)
    names = h2.declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libh2agg.so does not export %s" % name
    # and the ctypes layer binds exactly the declared surface
    bound = set(h2.load()._h2agg_signatures)
    assert bound == set(names), bound ^ set(names)


def test_version_string():
    assert b"sm_100a" in h2.load().h2agg_version()


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(h2.H2aggError):
        h2.Context(0)


def test_device_field_constants():
    src = open(os.path.join(ROOT, "halo2_snark_aggregator_b200", "csrc", "bn254_field.cuh")).read()
    arrays = re.findall(r"constexpr uint32_t v\[8\] = \{([^}]*)\}", src)
    vals = [sum(int(x.strip().rstrip("u"), 16) << (32 * i) for i, x in enumerate(a.split(","))) for a in arrays]
    assert vals == [R_MOD, (1 << 256) % R_MOD, (1 << 512) % R_MOD, P_MOD, (1 << 256) % P_MOD, (1 << 512) % P_MOD]
    invs = [int(x, 16) for x in re.findall(r"INV = (0x[0-9a-f]+)u", src)]
    assert invs == [(-pow(R_MOD, -1, 1 << 32)) % (1 << 32), (-pow(P_MOD, -1, 1 << 32)) % (1 << 32)]


def test_ptx_generator_selfcheck_and_inc_is_current(tmp_path):
    tool = os.path.join(ROOT, "halo2_snark_aggregator_b200", "csrc", "tools", "gen_mont_ptx.py")
    inc = os.path.join(ROOT, "halo2_snark_aggregator_b200", "csrc", "gen", "mont_mul_bn254.inc")
    before = open(inc).read()
    subprocess.check_call([sys.executable, tool], stdout=subprocess.DEVNULL)
    assert open(inc).read() == before, "generated PTX is stale: re-run tools/gen_mont_ptx.py"


def test_cpp_host_header_compiles(tmp_path):
    hpp = os.path.join(ROOT, "include", "h2agg.hpp")
    if not os.path.exists(hpp):
        pytest.skip("no C++ host header yet")
    src = tmp_path / "t.cpp"
    src.write_text('#include "h2agg.hpp"\nint main(){return 0;}\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_create_proof_query_list_of_the_aggregation_circuit():
    """71 written evaluations; 4 opening points (x, omega x, omega^-1 x, omega^last x)."""
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.prover import create_proof_queries, transcript_eval_order

    cs = plonk.aggregation_circuit_cs()
    q = create_proof_queries(cs)
    evals = [x for x in q if x[0] != ("h", 0)]
    assert len(evals) == 1 + 6 + 17 + 1 + 6 + 5 + 35   # instance, advice, fixed, random, sigma, permutation z, lookups
    assert len(set(q)) == len(q)
    rots = list(dict.fromkeys(r for _, r in q))
    assert sorted(rots) == sorted([0, 1, -1, -(cs.blinding_factors() + 1)])
    assert sum(1 for _, r in q if r == 0) == 54 and sum(1 for _, r in q if r == 1) == 10 and sum(1 for _, r in q if r == -1) == 7
    t = transcript_eval_order(cs)
    assert sorted(t) == sorted(evals) and ("h", 0) not in [nm for nm, _ in t]


def test_query_registration_order_follows_the_reference_configure():
    """halo2 registers a (column, rotation) query the first time configure() asks for it; create_proof writes the
    evaluations, and GWC folds the polynomials, in that order.  For Halo2VerifierCircuits::configure
    (halo2-snark-aggregator-circuit/src/verify_circuit.rs:225-241) that is:
      BaseGate::configure (halo2-ecc-circuit-lib/src/gates/base_gate.rs:692-720): enable_equality(base[0..5]) ->
        advice (a_i, cur); the gate closure queries constant (fixed 8), base[4] at next, next_coeff (fixed 7), then per
        column a_i / coeff_i (fixed 0..4), then mul_coeff (fixed 5, 6);
      RangeGate::configure (five/range_gate.rs:38-93): per lookup the selector then the table column: 9, 10, 11, ... 16;
      instance column: enable_equality -> (instance 0, cur)."""
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.prover import create_proof_queries

    cs = plonk.aggregation_circuit_cs()
    assert cs.queries["advice"] == [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (4, 1)]
    assert cs.queries["fixed"] == [(c, 0) for c in [8, 7, 0, 1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16]]
    assert cs.queries["instance"] == [(0, 0)]
    assert cs.blinding_factors() == 5
    q = create_proof_queries(cs)
    # verifier order (API/systems/halo2/params.rs:156-224): instance, advice, permutation, lookups, fixed, sigma, h, random
    assert q[0] == (("instance", 0), 0)
    assert q[1:7] == [(("advice", i), 0) for i in range(5)] + [(("advice", 4), 1)]
    assert q[7:12] == [(("perm_z", 0), 0), (("perm_z", 0), 1), (("perm_z", 1), 0), (("perm_z", 1), 1), (("perm_z", 0), -6)]
    assert q[12:17] == [(("lookup_z", 0), 0), (("lookup_input", 0), 0), (("lookup_table", 0), 0), (("lookup_input", 0), -1), (("lookup_z", 0), 1)]
    assert [nm for nm, _ in q[47:64]] == [("fixed", c) for c in [8, 7, 0, 1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16]]
    assert q[64:70] == [(("sigma", j), 0) for j in range(6)]
    assert q[70:] == [(("h", 0), 0), (("random", 0), 0)]


def test_permutation_last_queries_run_in_reverse_set_order():
    """halo2 and the reference's verifier (API/systems/halo2/permutation.rs:138-181) open (x, omega x) for every set
    first and then omega^last x with sets.iter().rev().skip(1): with three or more sets the fold order at omega^last x
    is the REVERSE set order."""
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.prover import create_proof_queries, transcript_eval_order

    E = plonk.Expression
    cs = plonk.ConstraintSystem(num_fixed=1, num_advice=7, num_instance=0)
    cs.create_gate("g", [E.fixed(0) * E.advice(0) * E.advice(1)])   # degree 3 -> chunk_len 1 -> one set per column
    for i in range(4):
        cs.enable_equality("advice", i)
    assert cs.num_permutation_sets() == 4
    last = -(cs.blinding_factors() + 1)
    q = create_proof_queries(cs)
    perm = [x for x in q if x[0][0] == "perm_z"]
    assert perm == [(("perm_z", s), r) for s in range(4) for r in (0, 1)] + [(("perm_z", s), last) for s in (2, 1, 0)]
    # the transcript, in contrast, carries the evaluations set by set (API/systems/halo2/verify.rs:198-230)
    t = [x for x in transcript_eval_order(cs) if x[0][0] == "perm_z"]
    assert t == [(("perm_z", 0), 0), (("perm_z", 0), 1), (("perm_z", 0), last), (("perm_z", 1), 0), (("perm_z", 1), 1), (("perm_z", 1), last),
                 (("perm_z", 2), 0), (("perm_z", 2), 1), (("perm_z", 2), last), (("perm_z", 3), 0), (("perm_z", 3), 1)]
