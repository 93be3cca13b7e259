#!/usr/bin/env python3
"""One of each hot kernel at k (for `ncu --set full`): table-mode MSM of a uniform column, iNTT(n),
coset NTT(n -> 4n), witness expansion of an 8-point multi_exp."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200.domain import EvaluationDomain

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
ctx = h2.Context(0)
dom = EvaluationDomain(5, k, ctx)
d_b = ctx.dev_alloc(n * 64)
d_s = ctx.dev_alloc(n * 32)
d_e = ctx.dev_alloc(n * 4 * 32)
d_o = ctx.dev_alloc(160)
ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
ctx.synth_scalars_dev(0x1000, 0, 0, n, d_s)
sid = ctx.srs_register_dev(d_b, n)
ctx.msm_g1_dev(d_s, n, d_o, srs_id=sid)
dom.lagrange_to_coeff_dev(d_s)
dom.coeff_to_extended_dev(d_s, d_e)
ctx.synchronize()
npts = 8
pts = ctx.d2h(d_b, 8 * npts).reshape(npts, 8)
chip = h2.B200EccChip()
hp = [chip.assign_var(pts[i]) for i in range(npts)]
hs = [chip.assign_scalar((0x1234567 * (i + 3)) ** 9 % ((1 << 253) - 1)) for i in range(npts)]
chip.multi_exp(hp, hs)
d_cols = [ctx.dev_alloc((1 << 20) * 32) for _ in range(5)]
chip.expand_dev(ctx, d_cols, 1 << 20)
ctx.synchronize()
print("done", chip.rows())
