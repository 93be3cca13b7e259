#!/usr/bin/env python3
"""Tiny driver for ncu captures: a few MSMs (uniform + witness-like) and NTTs at k, nothing else.
usage: python tools/prof_kernels.py [k] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200.domain import EvaluationDomain

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
warm = True
n = 1 << k
ctx = h2.Context(0)
dom = EvaluationDomain(5, k, ctx)
d_b = ctx.dev_alloc(n * 64)
d_s = [ctx.dev_alloc(n * 32) for _ in range(3)]
d_e = ctx.dev_alloc(n * 4 * 32)
d_o = ctx.dev_alloc(1600)
ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
for i, kind in enumerate((0, 1, 2)):
    ctx.synth_scalars_dev(0x1000 + i, kind, 0, n, d_s[i])
ctx.synchronize()
sid = ctx.srs_register_dev(d_b, n) if os.environ.get("TABLE", "1") == "1" else 0
print("srs config", ctx.srs_config(sid) if sid else None)
ctx.kernel_timing(True)
if sid:
    for i in range(3):
        t0 = time.perf_counter()
        ctx.msm_g1_dev(d_s[i], n, d_o + 160 * i, srs_id=sid)
        ctx.synchronize()
        print("table-mode msm kind", i, "ms", (time.perf_counter() - t0) * 1e3)
    print(ctx.kernel_times())
for r in range(reps):
    for i in range(3):
        t0 = time.perf_counter()
        ctx.msm_g1_dev(d_s[i], n, d_o + 160 * i, d_bases=d_b)
        ctx.synchronize()
        print("msm kind", i, "ms", (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    dom.lagrange_to_coeff_dev(d_s[0])
    ctx.synchronize()
    t1 = time.perf_counter()
    dom.coeff_to_extended_dev(d_s[0], d_e)
    ctx.synchronize()
    t2 = time.perf_counter()
    dom.extended_to_coeff_dev(d_e)
    ctx.synchronize()
    t3 = time.perf_counter()
    print("intt ms", (t1 - t0) * 1e3, "coset ms", (t2 - t1) * 1e3, "ext_intt ms", (t3 - t2) * 1e3)
    ctx.synth_scalars_dev(0x1000, 0, 0, n, d_s[0])
    ctx.synchronize()
print(ctx.kernel_times())
