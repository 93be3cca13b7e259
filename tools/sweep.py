#!/usr/bin/env python3
"""BASELINE config 5: standalone BN254 MSM + NTT sweep 2^16 .. 2^26 on one B200 (device resident),
with size-independent correctness checks at every size (NTT: iNTT(NTT(a)) == a; MSM: table mode ==
plain mode, and both halves add up to the whole).  Writes one JSON line per size.
usage: python tools/sweep.py [kmin] [kmax] > gpurun_out/sweep.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import halo2_snark_aggregator_b200 as h2
import oracle_binding as ob
from util import R_MOD, fr_limbs, omega

kmin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 26
ctx = h2.Context(0)


def timed(fn, reps):
    fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for k in range(kmin, kmax + 1):
    n = 1 << k
    reps = 5 if k <= 22 else 2
    d_b, d_s, d_a, d_o = ctx.dev_alloc(n * 64), ctx.dev_alloc(n * 32), ctx.dev_alloc(n * 32), ctx.dev_alloc(4 * 160)
    rec = {"k": k, "n": n}
    try:
        ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
        ctx.synth_scalars_dev(0xA660000 + k, 0, 0, n, d_s)
        # ---- NTT
        w = omega(k)
        w_l, wi_l, ni_l = fr_limbs(w), fr_limbs(pow(w, -1, R_MOD)), fr_limbs(pow(n, -1, R_MOD))
        ctx.synth_scalars_dev(0xF00 + k, 0, 0, n, d_a)
        head = ctx.d2h(d_a, 4 * 4096)
        ctx.ntt_fr_dev(d_a, w_l, k)
        ctx.ntt_fr_dev(d_a, wi_l, k, scale=ni_l)
        ctx.synchronize()
        rec["ntt_round_trip_ok"] = bool(np.array_equal(ctx.d2h(d_a, 4 * 4096), head))
        rec["ntt_ms"] = timed(lambda: ctx.ntt_fr_dev(d_a, w_l, k), reps)
        rec["ntt_hbm_gbs"] = 64.0 * n / (rec["ntt_ms"] * 1e-3) / 1e9
        # ---- MSM plain mode (bases per call)
        ctx.msm_g1_dev(d_s, n, d_o, d_bases=d_b)
        ctx.msm_g1_dev(d_s, n // 2, d_o + 160, d_bases=d_b)
        ctx.msm_g1_dev(d_s + (n // 2) * 32, n // 2, d_o + 320, d_bases=d_b + (n // 2) * 64)
        ctx.synchronize()
        out = ctx.d2h(d_o, 60).reshape(3, 20)
        rec["msm_halves_add_up"] = bool(np.array_equal(ob.g1_sum(np.concatenate([out[1, 8:], out[2, 8:]])), out[0, 8:]))
        rec["msm_plain_ms"] = timed(lambda: ctx.msm_g1_dev(d_s, n, d_o, d_bases=d_b), reps)
        # ---- MSM table mode (registered SRS)
        sid = ctx.srs_register_dev(d_b, n)
        table, c, nwin = ctx.srs_config(sid)
        ctx.msm_g1_dev(d_s, n, d_o + 480, srs_id=sid)
        ctx.synchronize()
        rec["msm_table_equals_plain"] = bool(np.array_equal(ctx.d2h(d_o + 480, 20), out[0]))
        rec["msm_table_ms"] = timed(lambda: ctx.msm_g1_dev(d_s, n, d_o + 480, srs_id=sid), reps)
        rec["msm_table_mode"] = [bool(table), c, nwin]
        rec["msm_pairs_per_s"] = n / (rec["msm_table_ms"] * 1e-3)
        rec["msm_hbm_gbs"] = (96.0 * n + 96) / (rec["msm_table_ms"] * 1e-3) / 1e9
        ctx.srs_release(sid)
    finally:
        for d in (d_b, d_s, d_a, d_o):
            ctx.dev_free(d)
    print(json.dumps(rec), flush=True)
