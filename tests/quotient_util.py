"""Shared helpers of the quotient (evaluate_h) tests: random column sets, the oracle's view of a constraint
system, and a plain-Python interpreter of the plan words (checks the plan builder without a GPU)."""
import random

import numpy as np

import oracle_binding  # noqa: F401  (puts oracle/py on sys.path)
import quotient_ref as qr
from halo2_snark_aggregator_b200 import plonk

R = qr.R


def oracle_desc(cs):
    """constraint system -> the dict oracle/py/quotient_ref.py takes (expression TREES)"""
    return dict(
        gates=[p.to_tuple() for _, polys in cs.gates for p in polys],
        lookups=[([e.to_tuple() for e in ins], [e.to_tuple() for e in tabs]) for _, ins, tabs in cs.lookups],
        perm_columns=list(cs.permutation_columns),
        chunk_len=cs.chunk_len(),
        last_rotation=-(cs.blinding_factors() + 1),
    )


def lagrange_selectors(k, blinding):
    """l_0, l_last, l_active_row as Lagrange-basis columns (halo2 keygen): row 0; row n - blinding - 1; 1 on rows
    before it"""
    n = 1 << k
    last = n - blinding - 1
    l0 = [1 if i == 0 else 0 for i in range(n)]
    l_last = [1 if i == last else 0 for i in range(n)]
    l_active = [1 if i < last else 0 for i in range(n)]
    return l0, l_last, l_active


def random_lagrange_columns(plan, k, seed):
    """name -> n random field elements (the identity under test holds for any contents); real l_0/l_last/l_active"""
    rng = random.Random(seed)
    n = 1 << k
    cols = {}
    for name in plan.columns:
        cols[name] = [rng.randrange(R) for _ in range(n)]
    l0, l_last, l_active = lagrange_selectors(k, plan.cs.blinding_factors())
    cols[("l0", 0)], cols[("l_last", 0)], cols[("l_active_row", 0)] = l0, l_last, l_active
    return cols


def pack(vals):
    """canonical ints -> Montgomery limb array (n*4 u64)"""
    out = np.empty(4 * len(vals), dtype=np.uint64)
    m64 = (1 << 64) - 1
    for i, v in enumerate(vals):
        x = v * (1 << 256) % R
        out[4 * i] = x & m64
        out[4 * i + 1] = (x >> 64) & m64
        out[4 * i + 2] = (x >> 128) & m64
        out[4 * i + 3] = x >> 192
    return out


def unpack(limbs):
    rinv = pow(1 << 256, -1, R)
    a = limbs.reshape(-1, 4)
    return [((int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) * rinv) % R for r in a]


def interpret_plan(plan, ext_cols, k, ext_k, y, beta, gamma, theta, divide=True):
    """Execute the plan words exactly as include/h2agg.h documents them, on Python ints.
    ext_cols: list (plan.columns order) of 2^ext_k ints."""
    w = [int(x) for x in plan.words]
    consts = unpack(plan.consts) if plan.consts.size else []
    size, rs = 1 << ext_k, 1 << (ext_k - k)
    w_ext = qr.omega(ext_k)
    assert w[0] == plonk.PLAN_MAGIC
    n_gates, n_pcols, chunk, n_lookups = w[1], w[2], w[3], w[5]
    last_rot = w[4] - (1 << 32) if w[4] >> 31 else w[4]
    l0c, llc, lac = ext_cols[w[6]], ext_cols[w[7]], ext_cols[w[8]]
    t = plonk.t_evaluations(k, ext_k)
    out = []
    for idx in range(size):
        pc = 9

        def load(col, rot):
            return ext_cols[col][(idx + rot * rs) % size]

        def poly():
            nonlocal pc
            nt = w[pc]
            pc += 1
            acc = 0
            for _ in range(nt):
                ci, nf = w[pc], w[pc + 1]
                pc += 2
                prod = 1 if ci == plonk.NOCONST else consts[ci]
                for _ in range(nf):
                    word = w[pc]
                    pc += 1
                    rot = word >> 16
                    rot = rot - 65536 if rot >= 32768 else rot
                    prod = prod * load(word & 0xFFFF, rot) % R
                acc = (acc + prod) % R
            return acc

        def compress():
            nonlocal pc
            ne = w[pc]
            pc += 1
            acc = 0
            for _ in range(ne):
                acc = (acc * theta + poly()) % R
            return acc

        v = 0
        for _ in range(n_gates):
            v = (v * y + poly()) % R
        if n_pcols:
            n_sets = (n_pcols + chunk - 1) // chunk
            pcols, zcols = pc, pc + 2 * n_pcols
            pc = zcols + n_sets
            z0, zl = load(w[zcols], 0), load(w[zcols + n_sets - 1], 0)
            v = (v * y + (1 - z0) * l0c[idx]) % R
            v = (v * y + (zl * zl - zl) * llc[idx]) % R
            for s in range(1, n_sets):
                v = (v * y + (load(w[zcols + s], 0) - load(w[zcols + s - 1], last_rot)) * l0c[idx]) % R
            cd = beta * qr.ZETA % R * pow(w_ext, idx, R) % R
            for s in range(n_sets):
                left, right = load(w[zcols + s], 1), load(w[zcols + s], 0)
                for j in range(s * chunk, min((s + 1) * chunk, n_pcols)):
                    val, sg = load(w[pcols + 2 * j], 0), load(w[pcols + 2 * j + 1], 0)
                    left = left * (val + beta * sg + gamma) % R
                    right = right * (val + cd + gamma) % R
                    cd = cd * qr.DELTA % R
                v = (v * y + (left - right) * lac[idx]) % R
        for _ in range(n_lookups):
            tv = (compress() + beta) * (compress() + gamma) % R
            zc, ac, sc = w[pc], w[pc + 1], w[pc + 2]
            pc += 3
            z, zn, a, ap, s_ = load(zc, 0), load(zc, 1), load(ac, 0), load(ac, -1), load(sc, 0)
            ams = (a - s_) % R
            v = (v * y + (1 - z) * l0c[idx]) % R
            v = (v * y + (z * z - z) * llc[idx]) % R
            v = (v * y + (zn * (a + beta) % R * (s_ + gamma) - z * tv) * lac[idx]) % R
            v = (v * y + ams * l0c[idx]) % R
            v = (v * y + ams * (a - ap) % R * lac[idx]) % R
        assert pc == len(w)
        out.append(v * t[idx % len(t)] % R if divide else v)
    return out
