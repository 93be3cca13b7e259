"""Multi-GPU resident prover: `create_proof` for ONE aggregation proof spread over the GPUs of a node.

What the reference runs under halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994 is a sequence of commit
rounds separated by Fiat-Shamir challenges.  Inside a round the units are independent, so:

  rounds 1-3   COLUMN-PARALLEL: every committed column has an owner rank that holds its Lagrange / coefficient /
               extended forms in HBM (parallel.prover_plan).  MSMs that would unbalance a round are WINDOW-SHARDED over
               all ranks (`parallel.window_units`): the owner broadcasts the 2^k scalars over NVLink, every rank runs its
               window range against the replicated SRS table, the 96-byte partials are all-gathered and added locally
               (EC addition is not an NCCL reduction operator: gather + h2agg_g1_sum_dev, never all-reduce).
  quotient     ROW-SHARDED: evaluate_h is pointwise in the coset row up to the query rotations, so rank g takes rows
               [g 2^ext_k / N, (g+1) 2^ext_k / N).  One grouped send/recv (NCCL turns the batch into an all-to-all over
               NVSwitch) moves, of every proof column, each rank's row window + rotation halo; the proving-key columns are
               replicated.  The row slices of h are all-gathered, every rank runs the (cheap, 4.4 ms) extended iNTT, and
               the h pieces are committed from the window-unit pool.
  evaluations  by the owner of each polynomial; 32-byte results all-reduced as integers (one non-zero contribution each).
  GWC          per opening point every rank folds the polynomials it owns with the powers of v they carry
               (h2agg_poly_lincomb_dev); the partial folds go point-to-point to the ranks that commit that point's
               quotient (again the window-unit pool), which add them, run kate_division and their window range.

Every rank returns the same commitments / evaluations (they feed the transcript, which every rank drives identically),
bit-identical to the single-GPU ResidentProver: all arithmetic is exact, only the schedule differs.  With world == 1
no collective is issued and the class degenerates to ResidentProver's flow.

torch is used for what the task assigns to it: device buffers NCCL can address, and torch.distributed for the plumbing.
"""
import numpy as np

from . import parallel as par
from . import plonk
from .domain import fr_to_limbs
from .prover import ResidentProver, create_proof_queries

_R = plonk.R_MOD


class DistributedProver:
    def __init__(self, ctx, cs, k, srs_lagrange, srs_g, torch, dist=None, rank=0, world=1, device=None):
        self.torch, self.dist, self.rank, self.world, self.dev = torch, dist, rank, world, device
        self._tensors = {}
        self.ctx, self.cs, self.k, self.n = ctx, cs, k, 1 << k
        self.pr = ResidentProver(ctx, cs, k, srs_lagrange, srs_g, alloc=self._alloc)
        pr = self.pr
        self.ext_n, self.ext_k = pr.ext_n, pr.ext_k
        assert self.ext_n % world == 0, "the world size must divide the extended domain"
        self.nwin_l = ctx.srs_config(srs_lagrange)[2]
        self.nwin_g = ctx.srs_config(srs_g)[2]
        self.witness = [("instance", i) for i in range(cs.num_instance)] + [("advice", i) for i in range(cs.num_advice)]
        self.n_sets = cs.num_permutation_sets()
        self.plan = par.prover_plan(len(self.witness), len(cs.lookups), self.n_sets, world)
        self.owner = {nm: self.plan["witness"][j] for j, nm in enumerate(self.witness)}
        for i, r in enumerate(self.plan["lookup"]):
            for kind in ("lookup_input", "lookup_table", "lookup_z"):
                self.owner[(kind, i)] = r
        for s in range(self.n_sets):
            self.owner[("perm_z", s)] = self.plan["perm_rank"]
        self.owner[("random", 0)] = world - 1
        # proving-key polynomials are replicated; their evaluation / fold work is dealt round-robin
        self.pk_names = [("fixed", i) for i in range(cs.num_fixed)] + [("sigma", j) for j in range(len(cs.permutation_columns))]
        for j, nm in enumerate(self.pk_names):
            self.owner.setdefault(nm, j % world)
        self.owner[("h", 0)] = 0
        self.halo = par.rotation_halo(par.constraint_system_rotations(cs), 1 << (self.ext_k - k))
        self.shards = par.quotient_row_shards(self.ext_n, world)
        self._win = {}        # name -> window buffer pointer (non-owned proof columns)
        self._scratch = {}
        self.nvlink_bytes = 0  # bytes this rank RECEIVED over NVLink in the last proof (payload of the collectives)

    # ---- memory: torch tensors, addressed by pointer inside the library --------------------------------------------
    def _alloc(self, nbytes):
        t = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.dev)
        self._tensors[t.data_ptr()] = t
        return t.data_ptr()

    def tensor(self, ptr, nbytes=None, offset=0):
        t = self._tensors[ptr]
        return t[offset:offset + nbytes] if nbytes is not None else t

    def _buf(self, key, nbytes):
        if key not in self._scratch:
            self._scratch[key] = self._alloc(nbytes)
        return self._scratch[key]

    def close(self):
        self.ctx.synchronize()
        self._tensors.clear()

    # ---- small collectives ---------------------------------------------------------------------------------------------
    def _share_rows(self, n_rows, width, mine):
        """mine: {row index -> uint64 array of `width`}; every row is produced by exactly one rank -> all rows, everywhere"""
        t = self.torch
        host = np.zeros((n_rows, width), dtype=np.uint64)
        for i, v in mine.items():
            host[i] = v
        if self.world == 1:
            return host
        d = t.from_numpy(host.view(np.int64)).to(self.dev)
        self.dist.all_reduce(d)        # integer sum; all other contributions are zero
        return d.cpu().numpy().view(np.uint64)

    def _broadcast(self, ptr, nbytes, src):
        if self.world > 1:
            self.dist.broadcast(self.tensor(ptr, nbytes), src=src)
            if self.rank != src:
                self.nvlink_bytes += nbytes

    def _pool_msm(self, items):
        """items: [(device pointer of n scalars, srs id, n windows)] present on EVERY rank -> affine commitments (len, 8).
        The window units of all items are cut into `world` contiguous ranges (parallel.pooled_window_ranges); everything this
        rank owes goes out in ONE batched call per SRS with a window range per column (h2agg_msm_g1_batch_ranges_dev),
        so the latency-bound ends of one shard run beside the accumulation of the next; one all-gather + local add + one
        read-back for the whole list."""
        ctx = self.ctx
        m = len(items)
        if m == 0:
            return np.zeros((0, 8), dtype=np.uint64)
        d_part = self._buf(("pool_part", m), m * 160)
        self.tensor(d_part, m * 160).zero_()
        # the window units are laid out item after item; items may differ in their window count (two SRS forms)
        by_srs = {}
        for (i, w0, w1) in par.pooled_window_ranges([it[2] for it in items], self.world, self.rank):
            by_srs.setdefault(items[i][1], []).append((i, w0, w1, items[i][2]))
        d_stage = self._buf(("pool_stage", m), m * 160)
        for srs, mine in sorted(by_srs.items()):
            wins = [None if (w0 == 0 and w1 == nwin) else (w0, w1) for (_, w0, w1, nwin) in mine]
            ctx.msm_g1_batch_dev([items[i][0] for (i, _, _, _) in mine], self.n, d_stage, srs_id=srs,
                                 windows=wins if any(w is not None for w in wins) else None)
            for q, (i, _, _, _) in enumerate(mine):
                self.tensor(d_part, 160, 160 * i).copy_(self.tensor(d_stage, 160, 160 * q))
        if self.world == 1:
            return ctx.d2h(d_part, 20 * m).reshape(m, 20)[:, :8].copy()
        d_all = self._buf(("pool_all", m), self.world * m * 160)
        self.dist.all_gather_into_tensor(self.tensor(d_all, self.world * m * 160), self.tensor(d_part, m * 160))
        d_sum = self._buf(("pool_sum", m), m * 160)
        ctx.g1_sum_dev(d_all + 64, self.world, m * 160, m, d_sum)
        return ctx.d2h(d_sum, 20 * m).reshape(m, 20)[:, :8].copy()

    def mine(self, names):
        return [nm for nm in names if self.owner[nm] == self.rank]

    # ---- proving key: replicated on every rank ------------------------------------------------------------------------
    def load_proving_key(self, fill):
        """ResidentProver.keygen_pk on every rank (the key is replicated): fixed / sigma commitments + their coefficient and
        extended forms + l_0, l_last, l_active_row; fill(name, d_lagrange) writes a Lagrange column into HBM."""
        return self.pr.keygen_pk(fill)

    # ---- round 1: witness columns --------------------------------------------------------------------------------------
    def round1(self, host_cols):
        """host_cols: {name -> pinned host array} for (at least) the columns this rank owns -> commitments (6, 8)"""
        pr = self.pr
        mine = self.mine(self.witness)
        got = {}
        if mine:
            c = pr.commit_columns(mine, [host_cols[nm] for nm in mine], keep_lagrange=True)
            got = {self.witness.index(nm): c[j] for j, nm in enumerate(mine)}
        comm = self._share_rows(len(self.witness), 8, got)
        for nm in self.witness:       # rounds 2 and 3 read every witness column in Lagrange form
            self._broadcast(pr.lagrange_slot(nm), self.n * 32, self.owner[nm])
        return comm

    # ---- round 2: lookup arguments -------------------------------------------------------------------------------------
    def round2(self, theta, blind):
        pr, cs = self.pr, self.cs
        L = len(cs.lookups)
        names = [(w_, i) for i in range(L) for w_ in ("lookup_input", "lookup_table")]
        my = [i for i in range(L) if self.plan["lookup"][i] == self.rank]
        got = {}
        if my:
            c = pr.lookup_round(theta, blind, only=my)
            for j, i in enumerate(my):
                got[2 * i], got[2 * i + 1] = c[2 * j], c[2 * j + 1]
        return self._share_rows(2 * L, 8, got)

    # ---- round 3: grand products + the vanishing argument's random polynomial -------------------------------------------
    def round3(self, beta, gamma, blind, h_random=None):
        """-> (commitments of perm_z 0.., lookup_z 0.., shape (sets + lookups, 8); commitment of the random polynomial (8,))
        h_random: pinned host coefficients of the random polynomial (needed on its owner rank only)."""
        pr, cs, ctx, n = self.pr, self.cs, self.ctx, self.n
        L = len(cs.lookups)
        znames = [("perm_z", s) for s in range(self.n_sets)] + [("lookup_z", i) for i in range(L)]
        pooled = [nm for nm in znames if nm in self.plan["pooled_z"]] if self.world > 1 else []
        if self.rank == self.plan["perm_rank"] and self.n_sets:
            pr.permutation_products(beta, gamma, blind)
        my_l = [i for i in range(L) if self.plan["lookup"][i] == self.rank]
        if my_l:
            pr.lookup_products(beta, gamma, blind, only=my_l)
        d_rand, _ = pr.slot(("random", 0), extended=False)
        if self.rank == self.owner[("random", 0)]:
            ctx.h2d(d_rand, h_random)
        # columns whose MSM is pooled travel to every rank (Lagrange form; the random polynomial in coefficient form)
        for nm in pooled:
            self._broadcast(pr.lagrange_slot(nm), n * 32, self.owner[nm])
        self._broadcast(d_rand, n * 32, self.owner[("random", 0)])
        # owned, not pooled: MSM + lagrange_to_coeff + coeff_to_extended on the lanes
        whole = [nm for nm in self.mine(znames) if nm not in pooled]
        got = {}
        if whole:
            c = pr._commit_resident(whole)
            got = {znames.index(nm): c[j] for j, nm in enumerate(whole)}
        if self.mine(pooled):          # owned and pooled: the transforms only (background stream)
            pr.transform_resident(self.mine(pooled))
        pool_all = self._pool_msm([(pr.lag[nm], pr.srs_lagrange, self.nwin_l) for nm in pooled] + [(d_rand, pr.srs_g, self.nwin_g)])
        pool_c, rand_c = pool_all[:len(pooled)], pool_all[len(pooled)]
        comm = self._share_rows(len(znames), 8, got)
        for j, nm in enumerate(pooled):
            comm[znames.index(nm)] = pool_c[j]
        return comm, rand_c

    # ---- quotient --------------------------------------------------------------------------------------------------------
    def _exchange_windows(self):
        """Every proof column's row window (+ halo) reaches every rank: one batch of point-to-point sends / receives."""
        pr, dist, rank, world, ext_n = self.pr, self.dist, self.rank, self.world, self.ext_n
        hl, hh = self.halo
        cols, row0 = [], []
        ops = []
        for nm in pr.plan.columns:
            src = self.owner.get(nm)
            if world == 1 or nm in self.pk_names or nm[0] in ("l0", "l_last", "l_active_row") or src == rank:
                cols.append(pr.ext[nm])
                row0.append(0)
            else:
                lo, hi = self.shards[rank]
                cnt = (hi - lo) + hl + hh
                if nm not in self._win:
                    self._win[nm] = self._alloc(cnt * 32)
                cols.append(self._win[nm])
                row0.append((lo - hl) % ext_n)
            if world > 1 and src is not None and nm not in self.pk_names and nm[0] not in ("l0", "l_last", "l_active_row"):
                for dst in range(world):
                    if dst == src or rank not in (src, dst):
                        continue
                    lo, hi = self.shards[dst]
                    first, cnt = (lo - hl) % ext_n, (hi - lo) + hl + hh
                    # the window may wrap around the end of the domain: at most two contiguous pieces
                    pieces = [(first, min(cnt, ext_n - first), 0)]
                    if pieces[0][1] < cnt:
                        pieces.append((0, cnt - pieces[0][1], pieces[0][1]))
                    for (g0, c, off) in pieces:
                        if rank == src:
                            ops.append(dist.P2POp(dist.isend, self.tensor(pr.ext[nm], c * 32, g0 * 32), dst))
                        else:
                            ops.append(dist.P2POp(dist.irecv, self.tensor(self._win[nm], c * 32, off * 32), src))
                            self.nvlink_bytes += c * 32
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return cols, row0

    def quotient(self, y, beta, gamma, theta):
        pr, ctx, n = self.pr, self.ctx, self.n
        pr.join_transforms()           # the exchange reads the extended forms the background stream produced
        cols, row0 = self._exchange_windows()
        lo, hi = self.shards[self.rank]
        d_h = pr._buf("h", self.ext_n * 32)
        lim = [fr_to_limbs(v) for v in (y, beta, gamma, theta)]
        if self.world == 1:
            ctx.evaluate_h_dev(pr.plan, cols, self.k, self.ext_k, *lim, d_h, divide=True)
        else:
            d_slice = self._buf("h_slice", (hi - lo) * 32)
            ctx.evaluate_h_dev(pr.plan, cols, self.k, self.ext_k, *lim, d_slice, divide=True, rows=(lo, hi - lo), col_row0=row0)
            self.dist.all_gather_into_tensor(self.tensor(d_h, self.ext_n * 32), self.tensor(d_slice, (hi - lo) * 32))
            self.nvlink_bytes += (self.ext_n - (hi - lo)) * 32
        pr.dom.extended_to_coeff_dev(d_h)          # every rank: 4.4 ms, and the pieces are then everywhere
        pieces = [d_h + i * n * 32 for i in range(pr.n_pieces)]
        for i in range(pr.n_pieces):
            pr.coeff[("h_piece", i)] = pieces[i]
        # the h pieces live inside one allocation: address them through (base, offset) views
        self._tensors.update({p: self.tensor(d_h, n * 32, i * n * 32) for i, p in enumerate(pieces) if i})
        return self._pool_msm([(p, pr.srs_g, self.nwin_g) for p in pieces])

    # ---- evaluation round ---------------------------------------------------------------------------------------------
    def fold_h(self, x):
        self.pr.fold_h(x)

    def evaluate(self, queries, x):
        mine = [i for i, (nm, _) in enumerate(queries) if self.owner[nm] == self.rank]
        got = {}
        if mine:
            ev = self.pr.evaluate([queries[i] for i in mine], x)
            got = {i: ev[j] for j, i in enumerate(mine)}
        return self._share_rows(len(queries), 4, got)

    # ---- GWC multi-opening ---------------------------------------------------------------------------------------------
    def open(self, queries, x, v):
        pr, ctx, n, dist, rank, world = self.pr, self.ctx, self.n, self.dist, self.rank, self.world
        pr.join_transforms()
        order, groups = [], {}
        for nm, rot in queries:
            if rot not in groups:
                groups[rot] = []
                order.append(rot)
            groups[rot].append(nm)
        P = len(order)
        units = par.window_units(P, self.nwin_g, world)
        committers = [[r for r in range(world) if any(it == p for (it, _, _) in units[r])] for p in range(P)]
        # my share of every fold: query i of a point carries v^i (multiopen.rs:55-61)
        d_part = []
        for p, rot in enumerate(order):
            polys, weights = [], []
            for i, nm in enumerate(groups[rot]):
                if self.owner[nm] == rank:
                    polys.append(pr.coeff[nm])
                    weights.append(fr_to_limbs(pow(v, i, _R)))
            d = self._buf(("fold_part", p), n * 32)
            ctx.poly_lincomb_dev(polys, np.concatenate(weights) if weights else None, n, d)
            d_part.append(d)
        recv = {}
        if world > 1:
            ops = []
            for p in range(P):
                for dst in committers[p]:
                    if dst != rank:
                        ops.append(dist.P2POp(dist.isend, self.tensor(d_part[p], n * 32), dst))
                if rank in committers[p]:
                    for src in range(world):
                        if src != rank:
                            recv[(p, src)] = self._buf(("fold_recv", p, src), n * 32)
                            ops.append(dist.P2POp(dist.irecv, self.tensor(recv[(p, src)], n * 32), src))
                            self.nvlink_bytes += n * 32
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        # the committers of a point add the shares, divide, and run their window range
        d_out = self._buf(("w_part", P), P * 160)
        self.tensor(d_out, P * 160).zero_()
        one = fr_to_limbs(1)
        d_stage = self._buf("w_stage", 160 * max(1, len(units[rank])))
        d_ws, wins = [], []
        for (p, w0, w1) in units[rank]:
            d_fold = self._buf(("fold_sum", p), n * 32)
            parts = [d_part[p]] + [recv[(p, s)] for s in range(world) if s != rank]
            ctx.poly_lincomb_dev(parts, np.tile(one, len(parts)), n, d_fold)
            d_w = self._buf(("w_poly", p), n * 32)
            ctx.kate_division_dev(d_fold, n, fr_to_limbs(pr.rotate_omega(x, order[p])), d_w)
            d_ws.append(d_w)
            wins.append(None if (w0 == 0 and w1 == self.nwin_g) else (w0, w1))
        if d_ws:   # all of this rank's window ranges in one batched call
            ctx.msm_g1_batch_dev(d_ws, n, d_stage, srs_id=pr.srs_g, windows=wins if any(w is not None for w in wins) else None)
            for q, (p, _, _) in enumerate(units[rank]):
                self.tensor(d_out, 160, 160 * p).copy_(self.tensor(d_stage, 160, 160 * q))
        if world == 1:
            return order, ctx.d2h(d_out, 20 * P).reshape(P, 20)[:, :8].copy()
        d_all = self._buf(("w_all", P), world * P * 160)
        dist.all_gather_into_tensor(self.tensor(d_all, world * P * 160), self.tensor(d_out, P * 160))
        d_sum = self._buf(("w_sum", P), P * 160)
        ctx.g1_sum_dev(d_all + 64, world, P * 160, P, d_sum)
        return order, ctx.d2h(d_sum, 20 * P).reshape(P, 20)[:, :8].copy()

    # ---- the whole proof ---------------------------------------------------------------------------------------------
    def prove(self, host_cols, h_random, blind, challenges):
        """witness in (pinned host columns on their owner ranks), proof elements out (the same on every rank).
        challenges(stage, outputs so far) -> the challenge(s) the transcript yields at that point:
        "theta" -> theta, "beta_gamma" -> (beta, gamma), "y" -> y, "x" -> x, "v" -> v."""
        self.nvlink_bytes = 0
        out = {}
        out["round1"] = self.round1(host_cols)
        theta = challenges("theta", out)
        out["round2"] = self.round2(theta, blind)
        beta, gamma = challenges("beta_gamma", out)
        out["round3"], out["random"] = self.round3(beta, gamma, blind, h_random)
        y = challenges("y", out)
        out["h"] = self.quotient(y, beta, gamma, theta)
        x = challenges("x", out)
        self.fold_h(x)
        queries = create_proof_queries(self.cs)
        out["evals"] = self.evaluate([q for q in queries if q[0] != ("h", 0)], x)
        v = challenges("v", out)
        out["order"], out["w"] = self.open(queries, x, v)
        return out
