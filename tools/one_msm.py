"""One table-mode MSM of a 17-bit column at k = 22 (twice), for `ncu --metrics gpu__time_duration.sum -k regex:msm_` captures
of the per-launch times of the reduce phase."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
k = 22; n = 1 << k
ctx = h2.Context(0)
d_b = ctx.dev_alloc(n * 64); ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
srs = ctx.srs_register_dev(d_b, n)
d_c = ctx.dev_alloc(n * 32); ctx.synth_scalars_dev(0x1000, 3, 0, n, d_c)
d_o = ctx.dev_alloc(160)
for _ in range(2):
    ctx.msm_g1_dev(d_c, n, d_o, srs_id=srs)
ctx.synchronize()
