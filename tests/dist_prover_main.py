"""Multi-GPU resident prover check: the DistributedProver (halo2_snark_aggregator_b200/dist_prover.py) on WORLD_SIZE
ranks against the single-GPU ResidentProver on the same inputs -- every commitment, evaluation and W point, bit for bit.
The single-GPU prover is itself held against the CPU oracles in tests/test_gpu_prover.py.

    python tests/dist_prover_main.py [k]                                              (world = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_prover_main.py [k]                                                 (NCCL, one rank per GPU)
"""
import os
import random
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))


def build_inputs(cs, k, seed):
    """Lagrange columns that satisfy the lookups (range tables i mod 2^t, 0/1 selectors, small advice cells)."""
    import oracle_binding as ob
    from halo2_snark_aggregator_b200 import plonk

    n = 1 << k
    rng = np.random.default_rng(seed)
    tbits = min(17, k - 1)

    def small(vals):
        canon = np.zeros((n, 4), dtype=np.uint64)
        canon[:, 0] = vals
        return ob.to_mont(0, np.ascontiguousarray(canon).ravel())

    cols = {}
    for i in range(cs.num_fixed):
        cols[("fixed", i)] = ob.gen_scalars(0xD100 + i, 0, n)
    for j in range(len(cs.permutation_columns)):
        cols[("sigma", j)] = ob.gen_scalars(0xD200 + j, 0, n)
    rows = np.arange(n, dtype=np.uint64)
    for sel, tab in ((9, 10), (11, 12), (13, 14), (15, 16)):
        cols[("fixed", tab)] = small(rows & np.uint64((1 << tbits) - 1))
        cols[("fixed", sel)] = small((rows % np.uint64(3) != 0).astype(np.uint64))
    for i in range(cs.num_advice):
        cols[("advice", i)] = small(rng.integers(0, 1 << tbits, n, dtype=np.uint64))
    cols[("instance", 0)] = ob.gen_scalars(0xD300, 0, n)
    cols[("random", 0)] = ob.gen_scalars(0xD400, 0, n)
    return cols


def main():
    import torch
    import torch.distributed as dist

    import halo2_snark_aggregator_b200 as h2
    import oracle_binding as ob
    from halo2_snark_aggregator_b200 import plonk
    from halo2_snark_aggregator_b200.dist_prover import DistributedProver
    from halo2_snark_aggregator_b200.prover import ResidentProver, create_proof_queries

    k = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = h2.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    cs = plonk.aggregation_circuit_cs()
    n = 1 << k
    g = ob.gen_bases(0xE000 + k, n)
    gl = ob.gen_bases(0xE100 + k, n)
    sid_g, sid_gl = ctx.srs_register(g), ctx.srs_register(gl)
    cols = build_inputs(cs, k, 5)
    R = plonk.R_MOD
    ch = {"theta": pow(5, 77, R), "beta_gamma": (pow(7, 55, R), pow(11, 44, R)), "y": pow(13, 33, R), "x": pow(17, 22, R), "v": pow(19, 11, R)}
    brng = random.Random(9)
    blind_vals = {}

    def blind(name, nrows):
        if (name, nrows) not in blind_vals:
            blind_vals[(name, nrows)] = ob.gen_scalars(0xB000 + (zlib.crc32(str(name).encode()) & 0xFFF), 0, nrows)   # the same in every process
        return blind_vals[(name, nrows)]

    # proving-key side: the same on every rank
    def fill(nm, d_l):
        ctx.h2d(d_l, cols[nm])

    def selectors(pr):
        import quotient_util as qu
        l0, l_last, l_active = qu.lagrange_selectors(k, cs.blinding_factors())
        names = [("l0", 0), ("l_last", 0), ("l_active_row", 0)]
        pr.commit_columns(names, [qu.pack(v) for v in (l0, l_last, l_active)])

    dp = DistributedProver(ctx, cs, k, sid_gl, sid_g, torch, dist if world > 1 else None, rank, world, dev)
    dp.load_proving_key(fill)
    host = {nm: cols[nm] for nm in dp.witness}
    out = dp.prove(host, cols[("random", 0)], blind, lambda stage, _o: ch[stage])
    ctx.synchronize()

    # the single-GPU prover on this rank, same inputs
    pr = ResidentProver(ctx, cs, k, sid_gl, sid_g)
    pr.keygen_pk(fill)
    want = {}
    want["round1"] = pr.commit_columns(dp.witness, [cols[nm] for nm in dp.witness], keep_lagrange=True)
    want["round2"] = pr.lookup_round(ch["theta"], blind)
    want["round3"] = pr.product_round(*ch["beta_gamma"], blind)
    want["random"] = pr.commit_coeff_columns([("random", 0)], [cols[("random", 0)]])[0]
    want["h"] = pr.quotient(ch["y"], *ch["beta_gamma"], ch["theta"])
    pr.fold_h(ch["x"])
    queries = create_proof_queries(cs)
    want["evals"] = pr.evaluate([q for q in queries if q[0] != ("h", 0)], ch["x"])
    want["order"], want["w"] = pr.open(queries, ch["x"], ch["v"])
    bad = [key for key in want if not np.array_equal(np.asarray(out[key]), np.asarray(want[key]))]
    ok = not bad
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = bool(flag.item())
    else:
        ok_all = ok
    print("rank %d/%d k=%d: %s%s  nvlink bytes received %d" % (rank, world, k, "MATCH" if ok else "MISMATCH in ", "" if ok else bad, dp.nvlink_bytes), flush=True)
    pr.close()
    dp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    if not ok_all:
        raise SystemExit(1)
    if rank == 0:
        print("DIST_PROVER_OK world=%d k=%d" % (world, k), flush=True)


if __name__ == "__main__":
    main()
