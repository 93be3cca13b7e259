"""TEST INFRASTRUCTURE: one tiny inner circuit, proved by oracle/py/mini_prover.py, for the aggregation-witness tests
(the stand-in for the reference's sample_setup / sample_run stages, halo2-snark-aggregator-circuit/src/sample_circuit.rs:32-124)."""
import random

import numpy as np

import bn254_ref as ref
import mini_prover as mp
import verifier_ref as V
import witness_scenarios as ws

R = ref.R
TRAPDOOR = 0x1A2B3C4D5E6F708192A3B4C5D6E7F8091A2B3C4D5E6F7081 % R


def tiny_cs():
    """2 advice, 2 fixed (selector q, table t), 1 instance; one gate q (a0 a1 + 3 a0(wX) - inst - 7), one lookup
    (q a1) in t, copy constraints over a0, a1, inst.  Queries in halo2's registration order for a configure() that
    enables equality first, then builds the gate, then the lookup."""
    a0, a1, a0n = ("advice", 0, 0), ("advice", 1, 0), ("advice", 0, 1)
    q, t, inst = ("fixed", 0, 0), ("fixed", 1, 0), ("instance", 0, 0)
    gate = ("product", q, ("sum", ("sum", ("sum", ("product", a0, a1), ("scaled", a0n, 3)), ("neg", inst)), ("neg", ("const", 7))))
    return dict(num_advice=2, num_fixed=2, num_instance=1, gates=[gate], lookups=[([("product", q, a1)], [t])],
                perm_columns=[("advice", 0), ("advice", 1), ("instance", 0)],
                advice_queries=[(0, 0), (1, 0), (0, 1)], fixed_queries=[(0, 0), (1, 0)], instance_queries=[(0, 0)],
                degree=5, blinding_factors=5)


def tiny_inner_proof(seed=1, k=4):
    """-> dict(vk, proof bytes, instances [[...]], srs)"""
    cs = tiny_cs()
    rng = random.Random(seed)
    n = 1 << k
    u = n - cs["blinding_factors"] - 1
    srs = mp.ToySrs(k, TRAPDOOR)
    qsel = [1 if i < 2 else 0 for i in range(n)]
    table = [i % 8 for i in range(n)]
    a0 = [rng.randrange(R) for _ in range(u)]
    a1 = [rng.randrange(8) if qsel[i] else rng.randrange(R) for i in range(u)]
    inst = [(a0[i] * a1[i] + 3 * a0[i + 1] - 7) % R for i in range(2)]
    a1[3] = inst[0]            # copy: a1[3] = inst[0]
    a1[8] = a0[7]              # copy: a0[7] = a1[8]
    vk, pk = mp.keygen(cs, k, [qsel, table], [[(1, 3), (2, 0)], [(0, 7), (1, 8)]], srs)
    proof = mp.create_proof(pk, [a0, a1], [inst], rng)
    return dict(vk=vk, proof=proof, instances=[inst], srs=srs, cs=cs)


def circuits_data(inner, nproofs=1):
    return [dict(name="tiny", vk=inner["vk"], nproofs=nproofs,
                 proofs=[dict(instances=inner["instances"], transcript_bytes=inner["proof"]) for _ in range(nproofs)])]


def pairing_holds(inner, w_x, w_g):
    """e(w_x, [s]_2) e(w_g, -[1]_2) = 1 (verify.rs:733-740) with the toy SRS's trapdoor: s w_x = w_g"""
    return ref.g1_mul(inner["srs"].s, w_x) == w_g


# ---- the product's chips behind the tuple-valued interface verifier_ref.py drives
class _TupleEccChip:
    def __init__(self, chip):
        self.c = chip

    def __getattr__(self, name):
        return getattr(self.c, name)

    def assign_var(self, pt): return self.c.assign_var(ws.xy_mont(pt))
    def assign_const(self, pt): return self.c.assign_const(ws.xy_mont(pt))
    def assign_one(self): return self.c.assign_const(ws.xy_mont(ref.G1_GEN))
    def scalar_mul_constant(self, s, pt): return self.c.scalar_mul_constant(s, ws.xy_mont(pt))

    def to_value(self, h):
        xy, ident = self.c.to_value(h)
        return None if ident else ref.unpack_point([int(v) for v in xy])


def b200_chips():
    import halo2_snark_aggregator_b200 as h2

    w = h2.B200Context()
    f = h2.B200ScalarChip(w)
    return V.Chips(f, f, _TupleEccChip(h2.B200EccChip(w)), h2.B200EncodeChip(w)), w


def fr_mont_columns(ctx):
    """oracle Context -> (5, rows, 4) uint64 Montgomery limbs (the layout the expansion kernel writes)"""
    import ecc_chip_ref as E

    cols = E.advice_columns(ctx)
    return np.stack([ws.fr_mont_rows(c) for c in cols])
