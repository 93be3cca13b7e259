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

# ---- the "next" rows (N1, N3) at the same size: evaluate_h over 55 resident extended columns, the lookup
# argument's sort / permutation / product, one permutation product
if os.environ.get("NEXT_ROWS", "1") == "1":
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.domain import fr_to_limbs

    for p in d_cols:
        ctx.dev_free(p)
    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    ext_k = cs.extended_k(k)
    size = 1 << ext_k
    cols = []
    for i in range(len(plan.columns)):
        p = ctx.dev_alloc(size * 32)
        ctx.synth_scalars_dev(0x9100 + i, 0, 0, size, p)
        cols.append(p)
    d_h = ctx.dev_alloc(size * 32)
    ch = [plonk.fr_mont(v) for v in (3 ** 100, 5 ** 90, 7 ** 80, 11 ** 70)]
    ctx.evaluate_h_dev(plan, cols, k, ext_k, *ch, d_h)
    ctx.synchronize()
    u = n - 6
    d_a, d_t, d_ap, d_tp, d_z = cols[:5]  # reuse: only the first n rows of these 2^ext_k-row buffers are touched
    ctx.synth_scalars_dev(0x8801, 3, 0, n, d_a)
    ctx.synth_scalars_dev(0x8802, 3, 0, n, d_t)
    ctx.permute_expression_pair_dev(d_a, d_t, u, d_ap, d_tp)
    ctx.lookup_product_dev(d_a, d_t, d_ap, d_tp, n, ch[1], ch[2], d_z)
    w = pow(plonk.ROOT_OF_UNITY, 1 << (28 - k), plonk.R_MOD)
    ctx.permutation_product_dev(cols[5:8], cols[8:11], k, fr_to_limbs(w), ch[1], plonk.fr_mont(plonk.DELTA), ch[1], ch[2], 0, d_z)
    ctx.synchronize()
    print("next rows done")
