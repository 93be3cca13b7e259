"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

The path shards without a data-path exchange; the only collective is an all-gather of the
commitments (64-byte affine + 96-byte Jacobian = 160 B each) per commit round so every rank can
drive the Fiat-Shamir transcript.  EC addition is not an NCCL reduction operator, so combining
window-sharded partial MSMs is all-gather + a local add (`h2agg_g1_sum`), never all-reduce."""


def shard_units(n_units, world, rank):
    """Column-parallel sharding (8e-1): unit i belongs to rank i % world."""
    return [i for i in range(n_units) if i % world == rank]


def window_shards(n_windows, world):
    """Window-sharded MSM (8e-2): contiguous window ranges [begin, end) per rank, covering all windows."""
    return [(n_windows * g // world, n_windows * (g + 1) // world) for g in range(world)]


def gather_round(dist, torch, local, slots, world, device):
    """All-gather one commit round.

    local: {unit index -> uint8 tensor of 160 bytes (on `device`)} computed by this rank;
    slots: the maximum number of units any rank owns in this round.
    Returns {unit index -> 160-byte tensor} for ALL units of the round, identical on every rank.
    Message layout per rank: slots x (8-byte little-endian unit index + 1, 160-byte point)."""
    rec = 168
    send = torch.zeros(slots * rec, dtype=torch.uint8, device=device)
    for slot, (idx, pt) in enumerate(sorted(local.items())):
        tag = torch.tensor([(idx + 1 >> (8 * b)) & 0xFF for b in range(8)], dtype=torch.uint8, device=device)
        send[slot * rec:slot * rec + 8] = tag
        send[slot * rec + 8:(slot + 1) * rec] = pt
    if world == 1:
        recv = send
    else:
        recv = torch.empty(world * slots * rec, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(recv, send)
    out = {}
    flat = recv.cpu()
    for k in range(world * slots):
        chunk = flat[k * rec:(k + 1) * rec]
        tag = int.from_bytes(bytes(chunk[:8].tolist()), "little")
        if tag:
            out[tag - 1] = chunk[8:]
    return out


def plan_phase(msm_costs, other_costs, world, n_windows, msm_fixed=None):
    """Balance one phase of a commit round over `world` ranks (hybrid of 8e-1 and 8e-2).

    msm_costs / other_costs: {unit -> estimated cost}; msm_fixed: {unit -> the part of an MSM's cost every window shard
    pays again} (digit passes over all scalars, scans, the bucket tree; default 0).  Whole MSMs are dealt longest-first
    to the least loaded rank while at least `world` of them remain.  The remainder (fewer MSMs than ranks) is
    window-sharded: every MSM gets world // remainder ranks, spare ranks go one at a time to the MSM whose shards are
    the most expensive, and within that allowance the number of shards is the one that minimises the phase's makespan
    given what the ranks already carry (a shard costs fixed + variable * its windows / n_windows, so more shards are
    not always better).  NTT-like units fill the least loaded ranks, before or after the shards -- whichever order
    ends sooner (a phase of 29 transforms and one MSM wants the MSM on the ranks that got fewer transforms).
    Returns [(unit, rank, None | (win_begin, win_end))]."""
    fixed = dict(msm_fixed or {})
    msms = sorted(msm_costs, key=lambda u: (-msm_costs[u], u))
    n_whole = (len(msms) // world) * world
    whole, rest = msms[:n_whole], msms[n_whole:]

    def shard_cost(u, windows):
        f = min(fixed.get(u, 0.0), msm_costs[u])
        return f + (msm_costs[u] - f) * windows / n_windows

    # ranks each leftover MSM may use
    allow = {}
    if rest:
        g0 = max(1, world // len(rest))
        allow = {u: min(g0, n_windows) for u in rest}
        spare = world - g0 * len(rest) if world >= len(rest) else 0
        for _ in range(max(0, spare)):
            cand = [u for u in rest if allow[u] < n_windows]
            if not cand:
                break
            u = max(cand, key=lambda x: (shard_cost(x, -(-n_windows // allow[x])), -x if isinstance(x, int) else 0))
            allow[u] += 1

    def build(others_first):
        load = [0.0] * world
        plan = []
        for u in whole:
            r = min(range(world), key=lambda k: (load[k], k))
            plan.append((u, r, None))
            load[r] += msm_costs[u]

        def place_others():
            for u in sorted(other_costs, key=lambda x: (-other_costs[x], x)):
                r = min(range(world), key=lambda k: (load[k], k))
                plan.append((u, r, None))
                load[r] += other_costs[u]

        def place_rest():
            for u in rest:
                best = None
                for g in range(1, allow[u] + 1):
                    ranks = sorted(range(world), key=lambda k: (load[k], k))[:g]
                    shards = window_shards(n_windows, g)
                    trial = list(load)
                    for r, (lo, hi) in zip(ranks, shards):
                        trial[r] += shard_cost(u, hi - lo) if g > 1 else msm_costs[u]
                    mk = max(trial)
                    if best is None or mk < best[0] - 1e-9:
                        best = (mk, g, ranks, shards)
                _, g, ranks, shards = best
                if g == 1:
                    plan.append((u, ranks[0], None))
                    load[ranks[0]] += msm_costs[u]
                else:
                    for r, (lo, hi) in zip(ranks, shards):
                        if hi > lo:
                            plan.append((u, r, (lo, hi)))
                            load[r] += shard_cost(u, hi - lo)

        if others_first:
            place_others()
            place_rest()
        else:
            place_rest()
            place_others()
        return max(load) if load else 0.0, plan

    mk_a, plan_a = build(False)
    if not rest or not other_costs:
        return plan_a
    mk_b, plan_b = build(True)
    return plan_b if mk_b < mk_a - 1e-9 else plan_a


# ---- row-sharded quotient (groundwork for the multi-GPU resident prover) --------------------------
# evaluate_h is pointwise in the coset row, except that a query with rotation r reads row idx + r * 2^(ext_k - k).  If
# rank g owns the rows [begin, end) of h it needs, of EVERY column, those rows plus a halo of max|negative rotation| rows
# below and max positive rotation rows above (wrapping around the domain).  The columns are produced column-parallel (each
# by the rank that committed it), so one exchange of row slices precedes the quotient.


def quotient_row_shards(ext_n, world):
    """Contiguous row ranges [begin, end) of the extended coset, one per rank."""
    return [(ext_n * g // world, ext_n * (g + 1) // world) for g in range(world)]


def rotation_halo(rotations, rot_scale):
    """(rows below, rows above) a shard needs for the given query rotations; rot_scale = 2^(ext_k - k)."""
    lo = max([0] + [-r for r in rotations]) * rot_scale
    hi = max([0] + [r for r in rotations]) * rot_scale
    return lo, hi


def constraint_system_rotations(cs):
    """Every rotation evaluate_h applies for a plonk.ConstraintSystem: the expressions' queries, z(omega X), a'(omega^-1 X)
    and the permutation argument's z(omega^last X)."""
    rots = {0}
    exprs = [p for _, polys in cs.gates for p in polys]
    for _, ins, tabs in cs.lookups:
        exprs += ins + tabs
    for e in exprs:
        rots |= {q[2] for q in e.queries()}
    if cs.lookups:
        rots |= {1, -1}
    if cs.permutation_columns:
        rots.add(1)
        if cs.num_permutation_sets() > 1:
            rots.add(-(cs.blinding_factors() + 1))
    return sorted(rots)


class RowWindow:
    """Rows [begin - halo_lo, end + halo_hi) (mod ext_n) of one column, addressed by GLOBAL row index."""

    def __init__(self, data, begin, end, halo_lo, halo_hi, ext_n):
        self.data, self.first, self.ext_n = data, (begin - halo_lo) % ext_n, ext_n
        self.count = min(ext_n, (end - begin) + halo_lo + halo_hi)
        assert len(data) == self.count

    def __getitem__(self, idx):
        off = (idx - self.first) % self.ext_n
        if off >= self.count:
            raise IndexError("row %d is outside this rank's window" % idx)
        return self.data[off]


def window_rows(begin, end, halo_lo, halo_hi, ext_n):
    """Global row indices of a RowWindow, in storage order."""
    count = min(ext_n, (end - begin) + halo_lo + halo_hi)
    first = (begin - halo_lo) % ext_n
    return [(first + i) % ext_n for i in range(count)]


def exchange_row_windows(dist, torch, owned, owner_of, names, shards, halo, ext_n, world, rank, limbs=4):
    """One exchange before the quotient: every rank sends, of each column it owns, the row window of every other rank.

    owned: {name -> int64 tensor (ext_n, limbs)} for the columns this rank produced; owner_of: {name -> rank} for all
    `names`.  Returns {name -> int64 tensor (window rows, limbs)} for ALL names.  Point-to-point (batch_isend_irecv: the
    grouped send/recv that NCCL turns into one all-to-all over NVLink; gloo runs it as is)."""
    halo_lo, halo_hi = halo
    rows = [torch.tensor(window_rows(b, e, halo_lo, halo_hi, ext_n), dtype=torch.long) for (b, e) in shards]
    out, ops, keep = {}, [], []
    for nm in names:
        src = owner_of[nm]
        if src == rank:
            col = owned[nm]
            out[nm] = col[rows[rank]].clone()
            for dst in range(world):
                if dst != rank:
                    buf = col[rows[dst]].contiguous()
                    keep.append(buf)
                    ops.append(dist.P2POp(dist.isend, buf, dst))
        else:
            buf = torch.empty((len(rows[rank]), limbs), dtype=torch.int64)
            out[nm] = buf
            ops.append(dist.P2POp(dist.irecv, buf, src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


# ---- multi-GPU resident prover: static work plan (dist_prover.py) ------------------------------------------------------
# Measured single-B200 costs at k = 22 (ms); only the RATIOS matter for the balance.
COST = {"msm_small": 2.6, "msm_uniform": 12.5, "intt": 1.0, "coset": 3.4, "sort": 1.6, "product": 1.2, "shard_fixed": 0.5}


def window_units(n_items, n_windows, world):
    """Spread n_items MSMs of n_windows windows each over `world` ranks at WINDOW granularity (McNaughton's wrap-around
    rule): the n_items * n_windows window units are laid out item after item and cut into `world` contiguous ranges, so
    every rank gets the same number of windows (+-1), at most two MSMs per rank are partial and an MSM is cut into at most
    ceil(world / n_items) + 1 shards.  Returns per rank a list of (item, win_begin, win_end)."""
    total = n_items * n_windows
    out = []
    for r in range(world):
        lo, hi = total * r // world, total * (r + 1) // world
        mine = []
        u = lo
        while u < hi:
            item, w = divmod(u, n_windows)
            w_end = min(n_windows, w + (hi - u))
            mine.append((item, w, w_end))
            u += w_end - w
        out.append(mine)
    return out


def pooled_window_ranges(item_windows, world, rank):
    """The window units of a list of pooled MSMs (item i has item_windows[i] windows; the two SRS forms may differ) laid
    out item after item and cut into `world` contiguous ranges: -> [(item, win_begin, win_end)] of `rank`.  Every window of
    every item belongs to exactly one rank, ranks differ by at most one window, an item is cut at most
    ceil(world / len(items)) + 1 times."""
    total = sum(item_windows)
    lo, hi = total * rank // world, total * (rank + 1) // world
    out, base = [], 0
    for i, nwin in enumerate(item_windows):
        w0, w1 = max(lo, base) - base, min(hi, base + nwin) - base
        if w1 > w0:
            out.append((i, w0, w1))
        base += nwin
    return out


def prover_plan(n_witness, n_lookups, n_perm_sets, world, cost=None):
    """Who does what in the column-parallel rounds of the multi-GPU prover.  Everything here is a pure function of the
    circuit shape and the world size, so every rank computes the same plan without talking.

    round 1   witness column j (instance + advice) -> rank j mod world
    round 2   lookup i (compress, sort / permute, 2 commits, 2 x (iNTT + coset NTT)) -> rank i mod world
    round 3   lookup product i stays with lookup i (its inputs are resident there); the permutation products (a chain
              through z_s[u]) all go to the least loaded rank; then MSMs are moved, largest first from the most loaded
              rank, into a POOL that is window-sharded over all ranks (`window_units`) as long as that lowers the
              makespan -- so 9 grand products on 8 ranks no longer cost two MSMs on one of them.
    Returns dict(witness=[rank], lookup=[rank], perm_rank=int, pooled_z=[names], load_ms=[per rank estimate of round 3])."""
    c = dict(COST)
    if cost:
        c.update(cost)
    witness = [j % world for j in range(n_witness)]
    lookup = [i % world for i in range(n_lookups)]
    transforms = c["intt"] + c["coset"]
    fixed = [0.0] * world          # work that cannot move: products + transforms
    msms = {}                      # name -> owner
    for i, r in enumerate(lookup):
        fixed[r] += c["product"] + transforms
        msms[("lookup_z", i)] = r
    base = [fixed[r] + c["msm_uniform"] * sum(1 for o in msms.values() if o == r) for r in range(world)]
    perm_rank = min(range(world), key=lambda r: (base[r], -r)) if n_perm_sets else 0   # ties: the highest rank (lookups fill the low ones)
    for s in range(n_perm_sets):
        fixed[perm_rank] += c["product"] + transforms
        msms[("perm_z", s)] = perm_rank
    pooled = []

    def makespan(pool):
        per = [fixed[r] + c["msm_uniform"] * sum(1 for nm, o in msms.items() if o == r and nm not in pool) for r in range(world)]
        share = (len(pool) * c["msm_uniform"]) / world + (c["shard_fixed"] * 2 if pool else 0.0)
        return max(per) + share, per

    best, _ = makespan(pooled)
    while world > 1:
        _, per = makespan(pooled)
        r = max(range(world), key=lambda q: (per[q], q))
        cand = sorted(nm for nm, o in msms.items() if o == r and nm not in pooled)
        if not cand:
            break
        trial, _ = makespan(pooled + [cand[-1]])
        if trial + 1e-9 < best:
            pooled.append(cand[-1])
            best = trial
        else:
            break
    _, per = makespan(pooled)
    return dict(witness=witness, lookup=lookup, perm_rank=perm_rank, pooled_z=sorted(pooled), load_ms=per, makespan_ms=best)
