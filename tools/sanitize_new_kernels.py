#!/usr/bin/env python3
"""Small run of the kernels added in round 2 for `compute-sanitizer --tool memcheck`: the NTT last pass with TMA bulk copies
+ mbarrier (k = 13: two passes), the witness expansion with device-side grouping and batched is_zero inversions (a 2-point
multi_exp), checked against the oracle / for non-zero output so that a clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import halo2_snark_aggregator_b200 as h2
import oracle_binding as ob
from util import domain_consts

ctx = h2.Context(0)
k = 13
d = domain_consts(k)
a = ob.gen_scalars(0x53, 0, 1 << k)
x = a.copy()
ctx.intt_fr(x, d["omega_inv"], d["n_inv"], k)
assert np.array_equal(x, ob.ifft(a.copy(), d["omega_inv"], d["n_inv"], k)), "iNTT mismatch"
ext = ctx.coeff_to_extended(a, k, k + 2, d["zeta"], d["omega_ext"])
assert np.array_equal(ext, ob.coeff_to_extended(a, k, k + 2, d["zeta"], d["omega_ext"])), "coset NTT mismatch"
pts = ob.gen_bases(0x77, 2).reshape(2, 8)
chip = h2.B200EccChip()
hp = [chip.assign_var(pts[i]) for i in range(2)]
hs = [chip.assign_scalar(12345 + 77 * i) for i in range(2)]
chip.multi_exp(hp, hs)
cols = chip.expand(ctx, n_rows=chip.rows() + 3)
assert cols[:, :-3].any() and not cols[:, -3:].any()
print("sanitize run ok: iNTT + coset NTT 2^13 (bulk-copy last pass) bit-exact, witness rows", chip.rows())
chip.close()
ctx.close()
