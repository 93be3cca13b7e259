#!/usr/bin/env python3
"""MSM cost on the lookup argument's permuted columns (a' sorted, s' = first occurrences + leftovers from the back)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200 import plonk

k = 22
n = 1 << k
u = n - 6
ctx = h2.Context(0)
d_b = ctx.dev_alloc(n * 64)
ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
srs = ctx.srs_register_dev(d_b, n)
d_out = ctx.dev_alloc(160)
d_a, d_s, d_ap, d_sp = (ctx.dev_alloc(n * 32) for _ in range(4))
ctx.synth_scalars_dev(0x1001, 1, 0, n, d_a)
canon = np.zeros((n, 4), dtype=np.uint64)
canon[:, 0] = np.arange(n, dtype=np.uint64) & np.uint64((1 << 17) - 1)
ctx.h2d(d_s, ctx.field_op(0, 3, canon.reshape(-1), np.tile(plonk.fr_mont(1 << 256), n)))
ctx.synth_scalars_dev(0x1002, 0, 0, n, d_ap)
ctx.synth_scalars_dev(0x1003, 0, 0, n, d_sp)
ctx.permute_expression_pair_dev(d_a, d_s, u, d_ap, d_sp)
for label, p in (("input a", d_a), ("table s", d_s), ("permuted a'", d_ap), ("permuted s'", d_sp)):
    ctx.msm_g1_dev(p, n, d_out, srs_id=srs)
    ctx.synchronize()
    ctx.kernel_timing(True)
    for _ in range(3):
        ctx.msm_g1_dev(p, n, d_out, srs_id=srs)
    t = ctx.kernel_times()
    ctx.kernel_timing(False)
    print("%-12s total %.3f ms  digits %.3f  accumulate %.3f  reduce %.3f" % (
        label, t["msm_total"][0] / 3, t["msm_digits_sort"][0] / 3, t["msm_accumulate"][0] / 3, t["msm_reduce"][0] / 3))
