"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: unit sharding, the per-round
all-gather of commitments, and window-sharded partial MSMs adding up to the whole (the compute is
stood in for by the CPU oracle; on a GPU box the same logic drives libh2agg, see bench.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding as ob
    from halo2_snark_aggregator_b200 import parallel as par
    from util import R_MOD, fr_limbs

    dev = torch.device("cpu")
    n = 256
    bases = ob.gen_bases(0x53525300, n)
    # --- column-parallel round: 5 columns dealt round-robin, gathered on every rank
    cols = [ob.gen_scalars(300 + i, i % 3, n) for i in range(5)]
    mine = par.shard_units(5, world, rank)
    local = {}
    for i in mine:
        jac = ob.best_multiexp(cols[i], bases, 1)
        pt = np.concatenate([jac[:8] if jac[8:].any() else np.zeros(8, dtype=np.uint64), jac]).view(np.uint8)
        local[i] = torch.from_numpy(pt.copy())
    slots = max(len(par.shard_units(5, world, r)) for r in range(world))
    allpts = par.gather_round(dist, torch, local, slots, world, dev)
    ok_gather = sorted(allpts) == list(range(5))
    for i in range(5):
        want = ob.best_multiexp(cols[i], bases, 1)
        got = np.frombuffer(bytes(allpts[i].tolist()), dtype=np.uint64)
        ok_gather = ok_gather and np.array_equal(got[8:], want)
    # --- window-sharded MSM: rank g takes a contiguous window range of the unsigned 16-bit digits
    c, nwin = 16, 16
    lo, hi = par.window_shards(nwin, world)[rank]
    canon = ob.from_mont(0, cols[0]).reshape(-1, 4)
    part = np.zeros_like(canon)
    for i in range(n):
        v = sum(int(canon[i, k]) << (64 * k) for k in range(4))
        masked = sum(((v >> (c * w)) & 0xFFFF) << (c * w) for w in range(lo, hi))
        part[i] = fr_limbs(masked % R_MOD) if False else [(masked >> (64 * k)) & (2**64 - 1) for k in range(4)]
    partial = ob.best_multiexp(ob.to_mont(0, np.ascontiguousarray(part).ravel()), bases, 1)
    t = torch.from_numpy(partial.view(np.uint8).copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    parts = np.concatenate([np.frombuffer(bytes(g.tolist()), dtype=np.uint64) for g in gathered])
    ok_win = np.array_equal(ob.g1_sum(parts), ob.best_multiexp(cols[0], bases, 1))
    q.put((rank, bool(ok_gather), bool(ok_win)))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gather_and_window_shards():
    world = 2
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "per-round all-gather of commitments"
    assert all(r[2] for r in res), "window-sharded partials must add up to the full MSM"


def test_shard_helpers_cover_everything_once():
    from halo2_snark_aggregator_b200 import parallel as par

    for world in (1, 2, 3, 4, 8):
        seen = sorted(i for r in range(world) for i in par.shard_units(97, world, r))
        assert seen == list(range(97))
        for nwin in (13, 16, 24):
            sh = par.window_shards(nwin, world)
            assert sh[0][0] == 0 and sh[-1][1] == nwin and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))


def test_plan_phase_is_complete_and_balanced():
    from halo2_snark_aggregator_b200 import parallel as par

    nwin = 13
    for world in (1, 2, 4, 8):
        for n_msm in (1, 4, 6, 9, 14):
            mc = {i: 12.5 for i in range(n_msm)}
            oc = {100 + i: 1.1 for i in range(n_msm)}
            plan = par.plan_phase(mc, oc, world, nwin)
            # every MSM is covered exactly once over all windows, every other unit exactly once
            for u in mc:
                parts = sorted((w if w else (0, nwin)) for (x, rk, w) in plan if x == u)
                assert parts[0][0] == 0 and parts[-1][1] == nwin and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
                assert len(set(rk for (x, rk, w) in plan if x == u)) == len(parts)
            for u in oc:
                assert sum(1 for (x, rk, w) in plan if x == u) == 1
            load = [0.0] * world
            for (x, rk, w) in plan:
                load[rk] += (mc[x] * ((w[1] - w[0]) / nwin if w else 1.0)) if x in mc else oc[x]
            ideal = (sum(mc.values()) + sum(oc.values())) / world
            assert max(load) <= ideal + 12.5 * (2.0 / nwin) + 12.5 * (0 if n_msm % world == 0 or world // max(n_msm % world, 1) > 1 else 1) + 1.2


def _quotient_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle", "py"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random

    import quotient_ref as qr
    import quotient_util as qu
    from halo2_snark_aggregator_b200 import parallel as par
    from halo2_snark_aggregator_b200 import plonk

    cs = plonk.aggregation_circuit_cs()
    plan = plonk.build_quotient_plan(cs)
    k = 4
    ext_k = cs.extended_k(k)
    ext_n = 1 << ext_k
    rng = random.Random(5)  # same stream on every rank: the full columns are known everywhere only to CHECK the result
    full = {nm: [rng.randrange(qr.R) for _ in range(ext_n)] for nm in plan.columns}
    y, beta, gamma, theta = [rng.randrange(qr.R) for _ in range(4)]
    names = list(plan.columns)
    owner_of = {nm: i % world for i, nm in enumerate(names)}          # column-parallel ownership
    to_t = lambda col: torch.from_numpy(qu.pack(col).view(np.int64).reshape(ext_n, 4).copy())
    owned = {nm: to_t(full[nm]) for nm in names if owner_of[nm] == rank}
    shards = par.quotient_row_shards(ext_n, world)
    halo = par.rotation_halo(par.constraint_system_rotations(cs), 1 << (ext_k - k))
    got = par.exchange_row_windows(dist, torch, owned, owner_of, names, shards, halo, ext_n, world, rank)
    begin, end = shards[rank]
    cols = {}
    for nm in names:
        vals = qu.unpack(got[nm].numpy().view(np.uint64).reshape(-1))
        cols[nm] = par.RowWindow(vals, begin, end, halo[0], halo[1], ext_n)
    mine = qr.evaluate_h(qu.oracle_desc(cs), cols, k, ext_k, y, beta, gamma, theta, rows=list(range(begin, end)))
    want = qr.evaluate_h(qu.oracle_desc(cs), full, k, ext_k, y, beta, gamma, theta, rows=list(range(begin, end)))
    # a row outside the window must not be readable (the halo is exactly what the rotations need)
    outside_ok = True
    if halo[0] + halo[1] + (end - begin) < ext_n:
        try:
            cols[names[0]][(end + halo[1]) % ext_n]
            outside_ok = False
        except IndexError:
            pass
    q.put((rank, mine == want, outside_ok, halo))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_row_sharded_quotient_exchange():
    """Groundwork for the multi-GPU resident prover: columns owned column-parallel, one point-to-point exchange of row
    windows (shard + rotation halo, wrapping), and every rank's shard of h equals the single-process evaluate_h."""
    world = 2
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_quotient_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "row-sharded evaluate_h must equal the full one on the shard"
    assert all(r[2] for r in res)
    assert res[0][3] == (24, 4)   # aggregation circuit: last rotation -6 and next rotation +1 at 4 coset rows per row


def test_rotation_halo_of_the_aggregation_circuit():
    from halo2_snark_aggregator_b200 import parallel as par
    from halo2_snark_aggregator_b200 import plonk

    cs = plonk.aggregation_circuit_cs()
    assert par.constraint_system_rotations(cs) == [-6, -1, 0, 1]
    assert par.rotation_halo([-6, -1, 0, 1], 4) == (24, 4)
    for world in (1, 2, 4, 8):
        sh = par.quotient_row_shards(1 << 24, world)
        assert sh[0][0] == 0 and sh[-1][1] == 1 << 24 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    w = par.window_rows(0, 8, 3, 2, 16)
    assert w == [13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
    rw = par.RowWindow(list(range(100, 113)), 0, 8, 3, 2, 16)
    assert rw[13] == 100 and rw[0] == 103 and rw[9] == 112


def test_plan_phase_covers_every_window_once_and_uses_the_fixed_cost_model():
    """Every MSM of a phase is either whole on one rank or cut into window ranges that tile [0, n_windows) exactly;
    spare ranks go to the MSM whose shards are the most expensive; a phase of many transforms and one MSM puts the MSM
    on the ranks that carry fewer transforms."""
    from halo2_snark_aggregator_b200 import parallel as par

    nw = 13
    tot = {0: 10.5, 1: 2.08, 2: 6.11, 3: 2.09}
    fx = {0: 2.45, 1: 1.48, 2: 3.25, 3: 1.39}

    def check(kinds, others, world):
        mc = {i: tot[k] for i, k in enumerate(kinds)}
        mf = {i: fx[k] for i, k in enumerate(kinds)}
        oc = {100 + i: c for i, c in enumerate(others)}
        plan = par.plan_phase(mc, oc, world, nw, mf)
        seen = {}
        for u, r, w in plan:
            assert 0 <= r < world
            seen.setdefault(u, []).append(w)
        assert set(seen) == set(mc) | set(oc)
        for u, ws in seen.items():
            if u in oc or ws == [None]:
                assert ws == [None]
                continue
            ws = sorted(ws)
            assert ws[0][0] == 0 and ws[-1][1] == nw and all(a[1] == b[0] for a, b in zip(ws, ws[1:]))
        return plan

    for world in (1, 2, 3, 4, 8):
        for kinds, others in (([1, 1, 1, 1, 1, 2], [0.9] * 6), ([3] * 14, [0.9] * 14), ([0] * 9, [0.9] * 9), ([0], [3.4] * 29), ([0] * 4, [])):
            check(kinds, others, world)
    # 6 MSMs on 8 ranks: the two spare ranks both go to the expensive `a4`-like column (unit 5)
    plan = check([1, 1, 1, 1, 1, 2], [0.9] * 6, 8)
    assert sorted(w for u, r, w in plan if u == 5) == [(0, 4), (4, 8), (8, 13)]
    assert all(w is None for u, r, w in plan if u < 5)
    # 29 transforms + 1 MSM on 8 ranks: the MSM's shards sit on ranks that got 3 transforms, not 4
    plan = check([0], [3.4] * 29, 8)
    n_other = {r: sum(1 for u, rr, w in plan if rr == r and u >= 100) for r in range(8)}
    assert all(n_other[r] == 3 for u, r, w in plan if u == 0)
    # 4 MSMs on 8 ranks: two shards each, disjoint rank pairs
    plan = check([0] * 4, [], 8)
    assert sorted(r for u, r, w in plan) == list(range(8))


def test_pooled_window_ranges_tile_every_item_exactly_once():
    from halo2_snark_aggregator_b200 import parallel as par

    for item_windows in ([13], [13, 13], [13, 13, 13, 13], [13, 13, 12], [16] * 9, [5]):
        for world in (1, 2, 3, 4, 8):
            cover = {i: [] for i in range(len(item_windows))}
            sizes = []
            for rank in range(world):
                mine = par.pooled_window_ranges(item_windows, world, rank)
                sizes.append(sum(w1 - w0 for _, w0, w1 in mine))
                for i, w0, w1 in mine:
                    assert 0 <= w0 < w1 <= item_windows[i]
                    cover[i].append((w0, w1))
            assert max(sizes) - min(sizes) <= 1
            for i, ws in cover.items():
                ws.sort()
                assert ws[0][0] == 0 and ws[-1][1] == item_windows[i] and all(a[1] == b[0] for a, b in zip(ws, ws[1:]))
