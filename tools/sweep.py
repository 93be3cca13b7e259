#!/usr/bin/env python3
"""BASELINE config 5: standalone BN254 MSM + NTT sweep 2^16 .. 2^26 at 1 / 2 / 4 / 8 B200 against the CPU port of
halo2's best_multiexp / best_fft, with ORACLE EQUALITY at every size the oracle finishes (--oracle-kmax, default 24)
and size-independent checks above that.  One JSON line per size (rank 0).

    python tools/sweep.py [--kmin 16] [--kmax 26] [--oracle-kmax 24] > profiles/r02_sweep_n1.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \\
        tools/sweep.py ... > profiles/r02_sweep_nN.jsonl

Multi-GPU: the MSM is WINDOW-SHARDED -- rank g takes windows [W g / N, W (g+1) / N) of the signed-digit
decomposition against the replicated fixed-base table, the 96-byte Jacobian partials are all-gathered over NCCL and added
on every rank (EC addition is not an NCCL reduce op).  A single NTT does not shard below 2^26 (a 4-step exchange does not
pay): the N ranks transform N independent columns, i.e. the per-column time is the single-GPU one and
the throughput is N columns per that time -- reported as such, not as a speed-up of one transform.
CPU timings use all host threads of the box (`cores`); they are measured at sizes <= --cpu-kmax (default 24)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmin", type=int, default=16)
    ap.add_argument("--kmax", type=int, default=26)
    ap.add_argument("--oracle-kmax", type=int, default=24)
    ap.add_argument("--cpu-kmax", type=int, default=24)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import halo2_snark_aggregator_b200 as h2
    import oracle_binding as ob
    from halo2_snark_aggregator_b200 import parallel as par
    from util import R_MOD, ZETA, fr_limbs, omega

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = h2.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    cores = ob.threads()

    def timed(fn, reps):
        """device time of `reps` calls (CUDA events on the launching stream), max over ranks, ms per call"""
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for k in range(args.kmin, args.kmax + 1):
        n = 1 << k
        reps = 5 if k <= 22 else 2
        bufs = [torch.empty(sz, dtype=torch.uint8, device=dev) for sz in (n * 64, n * 32, n * 32, 4 * n * 32 if k <= 24 else 32, 1600)]
        d_b, d_s, d_a, d_e, d_o = [b.data_ptr() for b in bufs]
        rec = {"k": k, "n": n, "n_gpus": world, "cpu_cores": cores}
        ctx.synth_bases_dev(0x53525300 + k, 0, n, d_b)
        ctx.synth_scalars_dev(0xA660000 + k, 0, 0, n, d_s)
        ctx.synth_scalars_dev(0xF00 + k, 0, 0, n, d_a)
        use_oracle = rank == 0 and k <= args.oracle_kmax
        w = omega(k)
        w_l, wi_l, ni_l = fr_limbs(w), fr_limbs(pow(w, -1, R_MOD)), fr_limbs(pow(n, -1, R_MOD))
        # ---- NTT: forward, inverse, coset 4x (device resident, one column per rank)
        a_host = ctx.d2h(d_a, 4 * n) if use_oracle else None
        ctx.ntt_fr_dev(d_a, w_l, k)
        ctx.synchronize()
        if use_oracle:
            t0 = time.perf_counter()
            want = ob.best_fft(a_host.copy(), w_l, k, cores)
            if k <= args.cpu_kmax:
                rec["cpu_fft_s"] = time.perf_counter() - t0
            rec["ntt_equals_oracle"] = bool(np.array_equal(ctx.d2h(d_a, 4 * n), want))
            del want
        head = ctx.d2h(d_a, 4 * 4096)
        ctx.ntt_fr_dev(d_a, wi_l, k, scale=ni_l)
        ctx.ntt_fr_dev(d_a, w_l, k)
        ctx.synchronize()
        rec["ntt_round_trip_ok"] = bool(np.array_equal(ctx.d2h(d_a, 4 * 4096), head))
        rec["ntt_ms"] = timed(lambda: ctx.ntt_fr_dev(d_a, w_l, k), reps)
        rec["intt_ms"] = timed(lambda: ctx.ntt_fr_dev(d_a, wi_l, k, scale=ni_l), reps)
        rec["ntt_columns_per_s_all_gpus"] = world / (rec["ntt_ms"] * 1e-3)
        rec["ntt_hbm_gbs_per_gpu"] = 64.0 * n / (rec["ntt_ms"] * 1e-3) / 1e9
        if k <= 24:
            we_l, z_l = fr_limbs(omega(k + 2)), fr_limbs(ZETA)
            ctx.coeff_to_extended_dev(d_a, k, k + 2, z_l, we_l, d_e)
            ctx.synchronize()
            if use_oracle and k <= 22:
                coeffs = ctx.d2h(d_a, 4 * n)
                rec["coset_equals_oracle"] = bool(np.array_equal(ctx.d2h(d_e, 16 * n), ob.coeff_to_extended(coeffs, k, k + 2, z_l, we_l, cores)))
            rec["coset_ntt_ms"] = timed(lambda: ctx.coeff_to_extended_dev(d_a, k, k + 2, z_l, we_l, d_e), reps)
        # ---- MSM: registered SRS (fixed-base table), window-sharded over the ranks
        sid = ctx.srs_register_dev(d_b, n)
        table, c, nwin = ctx.srs_config(sid)
        shard = par.window_shards(nwin, world)[rank]
        t_part = torch.zeros(160, dtype=torch.uint8, device=dev)
        t_all = torch.zeros(world * 160, dtype=torch.uint8, device=dev)
        t_sum = torch.zeros(160, dtype=torch.uint8, device=dev)

        def msm_step():
            if world == 1:
                ctx.msm_g1_dev(d_s, n, t_sum.data_ptr(), srs_id=sid)
                return
            if shard[1] > shard[0]:
                ctx.msm_g1_dev(d_s, n, t_part.data_ptr(), srs_id=sid, windows=shard)
            dist.all_gather_into_tensor(t_all, t_part)
            ctx.g1_sum_dev(t_all.data_ptr() + 64, world, 160, 1, t_sum.data_ptr())

        msm_step()
        ctx.synchronize()
        got = ctx.d2h(t_sum.data_ptr(), 20)
        if use_oracle:
            s_host, b_host = ctx.d2h(d_s, 4 * n), ctx.d2h(d_b, 8 * n)
            t0 = time.perf_counter()
            want = ob.best_multiexp(s_host, b_host, cores)
            if k <= args.cpu_kmax:
                rec["cpu_best_multiexp_s"] = time.perf_counter() - t0
            rec["msm_equals_oracle"] = bool(np.array_equal(got[8:], want))
            del s_host, b_host
        else:
            # size-independent check: the two halves add up to the whole (plain mode, per-call bases)
            ctx.msm_g1_dev(d_s, n // 2, d_o + 160, d_bases=d_b)
            ctx.msm_g1_dev(d_s + (n // 2) * 32, n // 2, d_o + 320, d_bases=d_b + (n // 2) * 64)
            ctx.synchronize()
            out = ctx.d2h(d_o + 160, 40).reshape(2, 20)
            rec["msm_halves_add_up_to_the_sharded_result"] = bool(np.array_equal(ob.g1_sum(np.concatenate([out[0, 8:], out[1, 8:]])), got[8:]))
        rec["msm_ms"] = timed(msm_step, reps)
        # ---- the same MSM on the witness-like mixture (a0..a3-like cells -- 70 % 17-bit, 10 % {0,1},
        # 20 % zero -- ending in 6 full-width blinding rows as real halo2 columns do)
        ctx.synth_scalars_dev(0xA770000 + k, 1, 0, n, d_s)
        if n > 6:
            ctx.synth_scalars_dev(0xA780000 + k, 0, 0, 6, d_s + 32 * (n - 6))
        msm_step()
        ctx.synchronize()
        got_w = ctx.d2h(t_sum.data_ptr(), 20)
        if use_oracle:
            s_host, b_host = ctx.d2h(d_s, 4 * n), ctx.d2h(d_b, 8 * n)
            t0 = time.perf_counter()
            want = ob.best_multiexp(s_host, b_host, cores)
            if k <= args.cpu_kmax:
                rec["cpu_best_multiexp_witness_like_s"] = time.perf_counter() - t0
            rec["msm_witness_like_equals_oracle"] = bool(np.array_equal(got_w[8:], want))
            del s_host, b_host
        rec["msm_witness_like_ms"] = timed(msm_step, reps)
        rec["msm_mode"] = {"table": bool(table), "window_bits": c, "windows": nwin, "my_windows": list(shard)}
        rec["msm_pairs_per_s"] = n / (rec["msm_ms"] * 1e-3)
        rec["msm_hbm_gbs"] = (96.0 * n + 96) / (rec["msm_ms"] * 1e-3) / 1e9
        if "cpu_best_multiexp_s" in rec:
            rec["msm_speedup_vs_cpu_port"] = rec["cpu_best_multiexp_s"] / (rec["msm_ms"] * 1e-3)
        if "cpu_fft_s" in rec:
            rec["ntt_speedup_vs_cpu_port_per_column"] = rec["cpu_fft_s"] / (rec["ntt_ms"] * 1e-3)
        ctx.srs_release(sid)
        ctx.synchronize()
        del bufs
        torch.cuda.empty_cache()
        if rank == 0:
            print(json.dumps(rec), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
