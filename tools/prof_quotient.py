#!/usr/bin/env python3
"""evaluate_h at k (default 22): 55 resident extended columns, the aggregation circuit's plan; prints the
CUDA-event time per launch (for ncu: run with `--reps 1`)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_snark_aggregator_b200 as h2
from halo2_snark_aggregator_b200 import plonk

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=22)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
cs = plonk.aggregation_circuit_cs()
plan = plonk.build_quotient_plan(cs)
k, ext_k = a.k, cs.extended_k(a.k)
size = 1 << ext_k
ctx = h2.Context(0)
d_cols = []
for i in range(len(plan.columns)):
    p = ctx.dev_alloc(size * 32)
    ctx.synth_scalars_dev(0x9100 + i, 0, 0, size, p)
    d_cols.append(p)
d_out = ctx.dev_alloc(size * 32)
ch = [plonk.fr_mont(v) for v in (3 ** 100, 5 ** 90, 7 ** 80, 11 ** 70)]
ctx.evaluate_h_dev(plan, d_cols, k, ext_k, *ch, d_out)
ctx.synchronize()
ctx.kernel_timing(True)
for _ in range(a.reps):
    ctx.evaluate_h_dev(plan, d_cols, k, ext_k, *ch, d_out)
ms, cnt = ctx.kernel_times()["evaluate_h"]
ctx.kernel_timing(False)
rows = size
print("evaluate_h k=%d rows=%d: %.3f ms/launch over %d launches; columns read %d x %.0f MiB = %.1f GB -> %.0f GB/s" % (
    k, rows, ms / cnt, cnt, len(d_cols), size * 32 / 2**20, len(d_cols) * size * 32 / 1e9,
    (len(d_cols) + 1) * size * 32 / 1e9 / (ms / cnt / 1e3)))
