#!/usr/bin/env python3
"""Time iNTT(2^k) and the coset NTT 2^k -> 2^(k+2) under different pass plans (radix cap) -- experiment harness."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200.domain import EvaluationDomain

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << k
ctx = h2.Context(0)
dom = EvaluationDomain(5, k, ctx)
d_s = ctx.dev_alloc(n * 32)
d_e = ctx.dev_alloc(n * 4 * 32)
ctx.synth_scalars_dev(0x1000, 0, 0, n, d_s)
for cap in [int(c) for c in (sys.argv[2:] or ["8", "9", "10", "11"])]:
    ctx.set_ntt_radix_cap(cap)
    for what in ("intt", "coset", "ext_intt"):
        f = {"intt": lambda: dom.lagrange_to_coeff_dev(d_s), "coset": lambda: dom.coeff_to_extended_dev(d_s, d_e),
             "ext_intt": lambda: dom.extended_to_coeff_dev(d_e)}[what]
        for _ in range(3):
            f()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        ctx.synchronize()
        print("cap %d %s %.3f ms" % (cap, what, (time.perf_counter() - t0) * 100), flush=True)
