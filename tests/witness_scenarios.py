"""The same chip-op sequences driven through the Python oracle (oracle/py/ecc_chip_ref.py) and the
product recorder (halo2_snark_aggregator_b200.B200EccChip).  TEST INFRASTRUCTURE."""
import random

import numpy as np

import bn254_ref as ref
import ecc_chip_ref as E

R_MONT = 1 << 256


def fq_mont(x):
    v = x % ref.P * R_MONT % ref.P
    return [(v >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]


def xy_mont(pt):
    if pt is None:
        return np.zeros(8, dtype=np.uint64)
    return np.array(fq_mont(pt[0]) + fq_mont(pt[1]), dtype=np.uint64)


def fr_mont_rows(col):
    """list of canonical ints -> (n, 4) uint64 Montgomery limbs"""
    out = np.zeros((len(col), 4), dtype=np.uint64)
    for i, v in enumerate(col):
        m = v * R_MONT % ref.R
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & ((1 << 64) - 1)
    return out


class Both:
    """Runs every op on the oracle and (optionally) on the recorder, keeping paired handles."""

    def __init__(self, chip=None):
        self.ctx = E.Context()
        self.chip = chip

    def assign_var(self, pt):
        o = E.assign_point(self.ctx, pt)
        h = self.chip.assign_var(xy_mont(pt)) if self.chip else None
        return (o, h)

    def assign_const(self, pt):
        o = E.assign_constant_point(self.ctx, pt)
        h = self.chip.assign_const(xy_mont(pt)) if self.chip else None
        return (o, h)

    def assign_scalar(self, s):
        o = E.bg_assign(self.ctx, s)
        h = self.chip.assign_scalar(s) if self.chip else None
        return (o, h)

    # The oracle side takes its operands exactly as the reference's adapter does
    # (halo2-snark-aggregator-circuit/src/chips/ecc_chip.rs:34-52, 99-131): add(&mut a.clone(), &mut b.clone()),
    # sub(&mut a.clone(), b), mul(&mut rhs.clone(), lhs), reduce(&mut v.clone()) -- caches filled in by an op die with the clone.
    def add(self, a, b):
        return (E.ecc_add(self.ctx, a[0].clone(), b[0].clone()), self.chip.add(a[1], b[1]) if self.chip else None)

    def sub(self, a, b):
        return (E.ecc_sub(self.ctx, a[0].clone(), b[0]), self.chip.sub(a[1], b[1]) if self.chip else None)

    def double(self, a):
        return (E.ecc_double(self.ctx, a[0].clone()), self.chip.double(a[1]) if self.chip else None)

    def normalize(self, a):
        return (E.ecc_reduce(self.ctx, a[0].clone()), self.chip.normalize(a[1]) if self.chip else None)

    def scalar_mul(self, s, a):
        return (E.ecc_mul(self.ctx, a[0].clone(), s[0]), self.chip.scalar_mul(s[1], a[1]) if self.chip else None)

    def multi_exp(self, pts, scalars):
        o = E.ecc_shamir(self.ctx, [p[0].clone() for p in pts], [s[0] for s in scalars])
        h = self.chip.multi_exp([p[1] for p in pts], [s[1] for s in scalars]) if self.chip else None
        return (o, h)

    def scalar_mul_constant(self, s, base):
        o = E.ecc_constant_mul(self.ctx, base, s[0], ref.g1_add)
        h = self.chip.scalar_mul_constant(s[1], xy_mont(base)) if self.chip else None
        return (o, h)

    def value(self, a):
        o = a[0]
        return None if o.z.value == 1 else (o.x.w(), o.y.w())


def scenario(name, b, rng):
    """Returns [(handle pair, expected native point)] for result checks."""
    G = ref.G1_GEN
    rp = lambda: ref.g1_mul(rng.randrange(1, ref.R), G)
    out = []
    if name == "add_double":
        p1, p2 = rp(), rp()
        A, B = b.assign_var(p1), b.assign_var(p2)
        C = b.add(A, B); out.append((C, ref.g1_add(p1, p2)))
        D = b.double(C); out.append((D, ref.g1_add(ref.g1_add(p1, p2), ref.g1_add(p1, p2))))
        E2 = b.add(A, b.assign_var(p1)); out.append((E2, ref.g1_add(p1, p1)))          # P + P through the add path
        F = b.sub(A, b.assign_var(p1)); out.append((F, None))                              # P - P = identity
        I = b.assign_const(None)
        H = b.add(B, I); out.append((H, p2))                                               # P + O
        H2 = b.add(I, B); out.append((H2, p2))                                             # O + P
        K = b.sub(D, A); out.append((K, ref.g1_add(ref.g1_add(ref.g1_add(p1, p2), ref.g1_add(p1, p2)), ref.g1_neg(p1))))
        N = b.normalize(K); out.append((N, b.value(K)))
        cur = A
        for _ in range(12):  # long add/sub chains drive `overflows` through the conditional reduces
            cur = b.sub(b.add(cur, B), A)
        out.append((cur, ref.g1_add(p1, ref.g1_mul(12, ref.g1_add(p2, ref.g1_neg(p1))))))
    elif name == "scalar_mul":
        p1 = rp()
        s = rng.randrange(ref.R)
        A = b.assign_var(p1)
        S = b.assign_scalar(s)
        out.append((b.scalar_mul(S, A), ref.g1_mul(s, p1)))
    elif name == "multi_exp_1":
        p1 = rp()
        s = rng.randrange(ref.R)
        out.append((b.multi_exp([b.assign_var(p1)], [b.assign_scalar(s)]), ref.g1_mul(s, p1)))
    elif name == "multi_exp_2":
        p1, p2 = rp(), rp()
        s1, s2 = rng.randrange(ref.R), rng.randrange(1 << 20)
        out.append((b.multi_exp([b.assign_var(p1), b.assign_var(p2)], [b.assign_scalar(s1), b.assign_scalar(s2)]),
                    ref.g1_add(ref.g1_mul(s1, p1), ref.g1_mul(s2, p2))))
    elif name == "multi_exp_zero_and_identity":
        p1 = rp()
        out.append((b.multi_exp([b.assign_var(p1), b.assign_var(None)], [b.assign_scalar(0), b.assign_scalar(5)]), None))
    elif name == "constant_mul":
        base = rp()
        s = rng.randrange(ref.R)
        out.append((b.scalar_mul_constant(b.assign_scalar(s), base), ref.g1_mul(s, base)))
    else:
        raise KeyError(name)
    return out


SCENARIOS = ["add_double", "scalar_mul", "multi_exp_1", "multi_exp_2", "multi_exp_zero_and_identity", "constant_mul"]


def run(name, chip=None, seed=7):
    b = Both(chip)
    res = scenario(name, b, random.Random(seed))
    return b, res


def scenario_multi_exp_n(b, rng, npts):
    G = ref.G1_GEN
    pts = [ref.g1_mul(rng.randrange(1, ref.R), G) for _ in range(npts)]
    scs = [rng.randrange(ref.R) for _ in range(npts)]
    hp = [b.assign_var(p) for p in pts]
    hs = [b.assign_scalar(s) for s in scs]
    start = b.ctx.offset
    res = b.multi_exp(hp, hs)
    return res, ref.msm(scs, pts), b.ctx.offset - start
