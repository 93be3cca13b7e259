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
    """~70 evaluations; 4 opening points (x, omega x, omega^-1 x, omega^last x)."""
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.prover import create_proof_queries

    cs = plonk.aggregation_circuit_cs()
    q = create_proof_queries(cs)
    evals = [x for x in q if x[0] != ("h", 0)]
    assert len(evals) == 6 + 17 + 1 + 6 + 5 + 35   # advice, fixed, random, sigma, permutation z, lookups
    assert len(set(q)) == len(q)
    rots = list(dict.fromkeys(r for _, r in q))
    assert sorted(rots) == sorted([0, 1, -1, -(cs.blinding_factors() + 1)])
    assert sum(1 for _, r in q if r == 0) == 53 and sum(1 for _, r in q if r == 1) == 10 and sum(1 for _, r in q if r == -1) == 7
