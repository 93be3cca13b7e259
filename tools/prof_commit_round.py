#!/usr/bin/env python3
"""commit_round_dev (MSM + iNTT + coset NTT per resident column, over the lanes) on sorted / unsorted / uniform columns:
wall time per round and per-kernel-class CUDA-event sums.  Usage: python tools/prof_commit_round.py [--k 22] [--cols 14]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200.domain import EvaluationDomain

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=22)
ap.add_argument("--cols", type=int, default=14)
a = ap.parse_args()
k, n, m = a.k, 1 << a.k, a.cols
ctx = h2.Context(0)
dom = EvaluationDomain(5, k, ctx)
d_b = ctx.dev_alloc(n * 64)
ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
srs = ctx.srs_register_dev(d_b, n)
lag = [ctx.dev_alloc(n * 32) for _ in range(m)]
co = [ctx.dev_alloc(n * 32) for _ in range(m)]
ex = [ctx.dev_alloc(n * 32 * 4) for _ in range(m)]
for label, kind, srt in (("17-bit unsorted", 3, False), ("17-bit sorted", 3, True), ("witness-like sorted", 1, True), ("uniform", 0, False)):
    for i, p in enumerate(lag):
        ctx.synth_scalars_dev(0x2000 + 16 * kind + i, kind, 0, n, p)
        if srt:
            ctx.sort_fr_dev(p, n)
    for with_ext in (True, False):
        args = dict(ext_k=k + 2, zeta=dom.g_coset, omega_ext=dom.extended_omega, d_ext_out=ex) if with_ext else {}
        ctx.commit_round_dev(srs, lag, k, dom.omega_inv, dom.ifft_divisor, co, **args)
        ctx.synchronize()
        ctx.kernel_timing(True)
        t0 = time.perf_counter()
        ctx.commit_round_dev(srs, lag, k, dom.omega_inv, dom.ifft_divisor, co, **args)
        ctx.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = ctx.kernel_times()
        ctx.kernel_timing(False)
        print("%-20s ext=%d  %7.2f ms / %d columns   msm_total %.1f (digits %.1f accumulate %.1f reduce %.1f)  ntt passes %.1f" % (
            label, with_ext, wall, m, t["msm_total"][0], t["msm_digits_sort"][0], t["msm_accumulate"][0], t["msm_reduce"][0], t["ntt_pass"][0]))
