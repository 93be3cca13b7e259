#!/usr/bin/env python3
"""MSM cost on a SORTED column (what the lookup argument's permuted columns are) against the same values unsorted:
per-kernel-class CUDA-event times.  Usage: python tools/prof_sorted_msm.py [--k 22]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=22)
a = ap.parse_args()
n = 1 << a.k
ctx = h2.Context(0)
d_b = ctx.dev_alloc(n * 64)
ctx.synth_bases_dev(0x53525300 + a.k, 0, n, d_b)
srs = ctx.srs_register_dev(d_b, n)
d_out = ctx.dev_alloc(160)
for kind, label in ((3, "17-bit"), (1, "witness a0..a3"), (0, "uniform")):
    d_c = ctx.dev_alloc(n * 32)
    ctx.synth_scalars_dev(0x1000 + kind, kind, 0, n, d_c)
    for state in ("unsorted", "sorted"):
        if state == "sorted":
            ctx.sort_fr_dev(d_c, n)
        ctx.msm_g1_dev(d_c, n, d_out, srs_id=srs)
        ctx.synchronize()
        ctx.kernel_timing(True)
        for _ in range(3):
            ctx.msm_g1_dev(d_c, n, d_out, srs_id=srs)
        t = ctx.kernel_times()
        ctx.kernel_timing(False)
        print("%-16s %-9s total %.3f ms  digits %.3f  accumulate %.3f  reduce %.3f" % (
            label, state, t["msm_total"][0] / 3, t["msm_digits_sort"][0] / 3, t["msm_accumulate"][0] / 3, t["msm_reduce"][0] / 3))
    ctx.dev_free(d_c)
