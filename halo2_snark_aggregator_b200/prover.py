"""Device-resident replay of the numeric core of halo2's `create_proof` for one circuit shape.

What the reference runs under halo2-snark-aggregator-circuit/src/verify_circuit.rs:986-994 (external crate
halo2_proofs, plonk/prover.rs; order restated in SURVEY.md App. B4) is, once the witness columns exist:

    1-3  commit rounds       commit_lagrange + lagrange_to_coeff (+ coeff_to_extended)   K1 / K2 / K3
         (2: permuted lookup columns = compress_expressions + permute_expression_pair;    N3
          3: permutation / lookup grand products)                                        N3
    4    quotient            evaluate_h, divide_by_vanishing_poly, extended_to_coeff,     N1 / K3 / K1
                             commit the n-coefficient pieces of h
    5    evaluation round    eval_polynomial of every queried polynomial at x * omega^rot  N2
    6    GWC multi-opening   per point: fold with v, kate_division, commit                N1 (fold) / N2 / K1

`ResidentProver` strings the C-ABI entry points together so that every polynomial stays in HBM between those
stages: columns arrive from host memory once (h2agg_commit_round_resident) and only commitments and evaluations
travel back.  It is a host-side orchestration layer -- all arithmetic runs in the CUDA kernels; challenges are
inputs (in the reference they come from the Rust transcript between the stages) and so are the blinding values
(halo2 draws them from its RNG).  The second and third round can either be uploaded like the first
(`commit_columns`) or be computed from the resident advice / fixed columns (`lookup_round`, `product_round`).
"""
import numpy as np

from . import plonk
from .domain import EvaluationDomain, fr_to_limbs

_R = plonk.R_MOD


class ResidentProver:
    def __init__(self, ctx, cs, k, srs_lagrange, srs_g, alloc=None):
        """srs_lagrange / srs_g: ids of the registered ParamsKZG::g_lagrange / g (h2agg_srs_register).
        alloc(nbytes) -> device pointer: caller-owned memory instead of h2agg_dev_alloc (the multi-GPU driver allocates
        torch tensors so that NCCL can address the same buffers)."""
        self.ctx, self.cs, self.k, self.n = ctx, cs, k, 1 << k
        self._ext_alloc = alloc
        self.plan = plonk.build_quotient_plan(cs)
        self.dom = EvaluationDomain(cs.degree(), k, ctx)
        self.ext_k, self.ext_n = self.dom.extended_k, self.dom.extended_len()
        self.srs_lagrange, self.srs_g = srs_lagrange, srs_g
        self.coeff, self.ext, self.lag = {}, {}, {}      # column name -> device pointer (coefficient / extended / Lagrange form)
        self._owned = []
        self.n_pieces = self.dom.quotient_poly_degree
        self._scratch = {}

    # The commitments of a round gate the next Fiat-Shamir challenge; its coefficient / extended forms are only read by
    # the quotient and the evaluation round.  With defer_transforms the NTT passes of every commit round run on the
    # context's background stream (h2agg_set_defer_transforms) and fill the latency-bound stretches of the following
    # rounds; `join_transforms` is called where their results are first needed.
    defer_transforms = True

    def _deferred(self, on):
        if self.defer_transforms:
            self.ctx.set_defer_transforms(on)

    def join_transforms(self):
        self.ctx.transforms_join()

    # -- optional stage trace (bench / profiling): set self.trace = {} to collect synchronised stage times in ms
    trace = None
    trace_kernels = False

    def _mark(self, label):
        if self.trace is None:
            return
        import time

        self.ctx.synchronize()
        now = time.perf_counter()
        if getattr(self, "_t_last", None) is not None:
            self.trace[label] = self.trace.get(label, 0.0) + (now - self._t_last) * 1e3
            if self.trace_kernels:  # per-kernel-class CUDA-event sums of the stage (needs ctx.kernel_timing(True))
                self.trace[label + ".kernel_class_ms"] = {kk: round(vv[0], 3) for kk, vv in self.ctx.kernel_times().items() if vv[1]}
        self._t_last = time.perf_counter()

    # -- memory ------------------------------------------------------------------------------------
    def _alloc(self, nbytes):
        if self._ext_alloc is not None:
            return self._ext_alloc(nbytes)
        p = self.ctx.dev_alloc(nbytes)
        self._owned.append(p)
        return p

    def _buf(self, key, nbytes):
        if key not in self._scratch:
            self._scratch[key] = self._alloc(nbytes)
        return self._scratch[key]

    def slot(self, name, extended=True):
        """Device buffers of a column's coefficient (and extended-coset) form, allocated on first use."""
        if name not in self.coeff:
            self.coeff[name] = self._alloc(self.n * 32)
        if extended and name not in self.ext:
            self.ext[name] = self._alloc(self.ext_n * 32)
        return self.coeff[name], self.ext.get(name)

    def adopt(self, name, d_coeff, d_ext=None):
        """Use caller-owned device buffers for a column (e.g. proving-key polynomials kept across proofs)."""
        self.coeff[name] = d_coeff
        if d_ext is not None:
            self.ext[name] = d_ext

    def close(self):
        self.ctx.synchronize()
        for p in self._owned:
            self.ctx.dev_free(p)
        self._owned = []

    # -- keygen: the proving key's polynomials, once per (params, vk) ----------------------------------
    def pk_names(self):
        cs = self.cs
        return [("fixed", i) for i in range(cs.num_fixed)] + [("sigma", j) for j in range(len(cs.permutation_columns))]

    def keygen_pk(self, fill):
        """The numeric part of `keygen_vk` + `keygen_pk` (halo2_proofs plonk/keygen.rs; the reference runs keygen_pk inside
        EVERY verify_run, halo2-snark-aggregator-circuit/src/verify_circuit.rs:974-979, and keygen_vk in verify_setup,
        :760-761): for the fixed columns and the permutation's sigma columns -- commit (the vk's fixed / permutation
        commitments), lagrange_to_coeff, coeff_to_extended -- plus l_0, l_last and l_active_row on the extended domain.
        fill(name, d_lagrange) writes the Lagrange form of ("fixed", i) / ("sigma", j) into the given device buffer
        (2^k x 32 B, Montgomery): the assignment of those cells is circuit synthesis and stays with the caller.
        Everything stays resident under the same names, so the key is computed once per process and reused by every proof
        (`ProvingKeyCache`).  -> {"fixed": (num_fixed, 8), "sigma": (num_perm_columns, 8)} affine commitments."""
        names = self.pk_names()
        for nm in names:
            fill(nm, self.lagrange_slot(nm))
        comm = self._commit_resident(names)
        nf = self.cs.num_fixed
        # l_0 = L_0, l_last = L_(n - blinding_factors - 1), l_active_row = 1 - (l_last + l_blind) = sum of L_i over the usable rows
        one = plonk.fr_mont(1)
        n, last = self.n, self.n - self.cs.blinding_factors() - 1
        col = np.zeros((n, 4), dtype=np.uint64)
        sel = [("l0", 0), ("l_last", 0), ("l_active_row", 0)]
        for nm, rows in zip(sel, (slice(0, 1), slice(last, last + 1), slice(0, last))):
            col[:] = 0
            col[rows] = one
            self.ctx.h2d(self.lagrange_slot(nm), col.reshape(-1))
        self._commit_resident(sel)      # (their commitments are not part of the vk; the transforms are what is needed)
        return {"fixed": comm[:nf], "sigma": comm[nf:]}

    # -- stages 1-3: commit rounds -------------------------------------------------------------------
    def lagrange_slot(self, name):
        if name not in self.lag:
            self.lag[name] = self._alloc(self.n * 32)
        return self.lag[name]

    def commit_columns(self, names, lagrange_cols, extended=True, keep_lagrange=False):
        """One commit round from HOST Lagrange columns -> affine commitments (len, 8).  The coefficient and
        extended forms (and, on request, the Lagrange form the lookup / permutation arguments read) stay resident
        under `names`."""
        slots = [self.slot(nm, extended) for nm in names]
        d = self.dom
        # deferred transforms read the Lagrange column after this call returns: it must live in its own buffer then
        keep = keep_lagrange or self.defer_transforms
        self._deferred(True)
        try:
            return self.ctx.commit_round_resident(
                self.srs_lagrange, list(lagrange_cols), self.k, d.omega_inv, d.ifft_divisor, [s[0] for s in slots],
                ext_k=self.ext_k if extended else 0, zeta=d.g_coset if extended else None,
                omega_ext=d.extended_omega if extended else None, d_ext_out=[s[1] for s in slots] if extended else None,
                d_lagrange_out=[self.lagrange_slot(nm) for nm in names] if keep else None)
        finally:
            self._deferred(False)

    def commit_device_columns(self, names):
        """Commit round for Lagrange columns that are ALREADY in HBM under `lagrange_slot(name)` -- e.g. the five advice
        columns the witness-expansion kernel wrote (h2agg_witness_expand_dev): the witness never visits host memory."""
        return self._commit_resident(names)

    def _commit_resident(self, names):
        """Commit round for Lagrange columns the device produced itself (self.lag[name])."""
        slots = [self.slot(nm) for nm in names]
        d = self.dom
        self._deferred(True)
        try:
            return self.ctx.commit_round_dev(self.srs_lagrange, [self.lag[nm] for nm in names], self.k, d.omega_inv, d.ifft_divisor,
                                             [s[0] for s in slots], ext_k=self.ext_k, zeta=d.g_coset, omega_ext=d.extended_omega,
                                             d_ext_out=[s[1] for s in slots])
        finally:
            self._deferred(False)

    def transform_resident(self, names):
        """lagrange_to_coeff + coeff_to_extended of resident Lagrange columns whose commitment is computed elsewhere."""
        slots = [self.slot(nm) for nm in names]
        d = self.dom
        self._deferred(True)
        try:
            self.ctx.transforms_dev([self.lag[nm] for nm in names], self.k, d.omega_inv, d.ifft_divisor, [s[0] for s in slots],
                                    ext_k=self.ext_k, zeta=d.g_coset, omega_ext=d.extended_omega, d_ext_out=[s[1] for s in slots])
        finally:
            self._deferred(False)

    # -- stage 2: lookup arguments (commit_permuted) -----------------------------------------------------
    def usable_rows(self):
        return self.n - (self.cs.blinding_factors() + 1)

    def lookup_round(self, theta, blind, only=None):
        """Per lookup: compress the input / table expressions over the resident Lagrange columns, sort + permute them
        over the usable rows (permute_expression_pair), append the caller's blinding rows, commit.
        blind(name, rows) -> uint64 array rows*4 (the values halo2 draws from its RNG).
        only: lookup indices to run (the multi-GPU driver gives every rank its share); default all."""
        names = [nm for nm in self.lag if nm[0] in ("fixed", "advice", "instance")]
        index = {nm: i for i, nm in enumerate(names)}
        cols = [self.lag[nm] for nm in names]
        u, tail = self.usable_rows(), self.cs.blinding_factors() + 1
        th = fr_to_limbs(theta)
        out_names = []
        self._mark("-")
        for i, (_, ins, tabs) in enumerate(self.cs.lookups):
            if only is not None and i not in only:
                continue
            for side, exprs in (("input", ins), ("table", tabs)):
                prog = plonk.ExpressionList(exprs, index)
                self.ctx.compress_expressions_dev(prog.words, prog.consts, cols, self.k, th,
                                                  self.lagrange_slot(("lookup_%s_compressed" % side, i)))
            self._mark("lookup.compress_expressions")
            pa, ps = self.lagrange_slot(("lookup_input", i)), self.lagrange_slot(("lookup_table", i))
            self.ctx.permute_expression_pair_dev(self.lag[("lookup_input_compressed", i)], self.lag[("lookup_table_compressed", i)],
                                                 u, pa, ps)
            self._mark("lookup.permute_expression_pair")
            for nm, p in ((("lookup_input", i), pa), (("lookup_table", i), ps)):
                self.ctx.h2d(p + 32 * u, np.ascontiguousarray(blind(nm, tail)))
                out_names.append(nm)
            self._mark("lookup.blinding_rows")
        out = self._commit_resident(out_names)
        self._mark("lookup.commit_round_dev")
        return out

    # -- stage 3: grand products (permutation::commit, lookup commit_product) ------------------------------
    def permutation_products(self, beta, gamma, blind):
        """The permutation argument's running products, one per column set, chained through z_s[u] -> names"""
        bf = self.cs.blinding_factors()
        u = self.usable_rows()
        b, g = fr_to_limbs(beta), fr_to_limbs(gamma)
        chunk = self.cs.chunk_len()
        pcs = self.cs.permutation_columns
        last = None
        names = []
        for s in range(self.cs.num_permutation_sets()):
            cols = pcs[s * chunk:(s + 1) * chunk]
            z = self.lagrange_slot(("perm_z", s))
            self.ctx.permutation_product_dev([self.lag[c] for c in cols], [self.lag[("sigma", s * chunk + j)] for j in range(len(cols))],
                                             self.k, self.dom.omega, fr_to_limbs(beta * pow(plonk.DELTA, s * chunk, _R)),
                                             fr_to_limbs(plonk.DELTA), b, g, last, z)
            self.ctx.h2d(z + 32 * (self.n - bf), np.ascontiguousarray(blind(("perm_z", s), bf)))
            last = z + 32 * u
            names.append(("perm_z", s))
        return names

    def lookup_products(self, beta, gamma, blind, only=None):
        """The lookup arguments' running products (all selected lookups in one call: their latency chains overlap on the lanes)"""
        bf = self.cs.blinding_factors()
        b, g = fr_to_limbs(beta), fr_to_limbs(gamma)
        L = [i for i in range(len(self.cs.lookups)) if only is None or i in only]
        zs = [self.lagrange_slot(("lookup_z", i)) for i in L]
        if zs:
            self.ctx.lookup_products_dev([self.lag[("lookup_input_compressed", i)] for i in L], [self.lag[("lookup_table_compressed", i)] for i in L],
                                         [self.lag[("lookup_input", i)] for i in L], [self.lag[("lookup_table", i)] for i in L],
                                         self.n, b, g, zs)
        for z, i in zip(zs, L):
            self.ctx.h2d(z + 32 * (self.n - bf), np.ascontiguousarray(blind(("lookup_z", i), bf)))
        return [("lookup_z", i) for i in L]

    def product_round(self, beta, gamma, blind):
        self._mark("-")
        out_names = self.permutation_products(beta, gamma, blind)
        self._mark("products.permutation")
        out_names += self.lookup_products(beta, gamma, blind)
        self._mark("products.lookup")
        out = self._commit_resident(out_names)
        self._mark("products.commit_round_dev")
        return out

    def commit_coeff_columns(self, names, coeff_cols):
        """Polynomials the prover creates in COEFFICIENT form (the vanishing argument's random polynomial):
        uploaded as they are and committed against `g` (ParamsKZG::commit)."""
        ptrs = []
        for nm, col in zip(names, coeff_cols):
            d, _ = self.slot(nm, extended=False)
            self.ctx.h2d(d, col)
            ptrs.append(d)
        return self._commit_dev(ptrs)

    # -- stage 4: quotient -----------------------------------------------------------------------------
    def quotient(self, y, beta, gamma, theta):
        """h = evaluate_h / (X^n - 1) on the coset -> coefficients -> commitments of its n-coefficient pieces."""
        self.join_transforms()
        cols = [self.ext[nm] for nm in self.plan.columns]
        d_h = self._buf("h", self.ext_n * 32)
        lim = [fr_to_limbs(v) for v in (y, beta, gamma, theta)]
        self.ctx.evaluate_h_dev(self.plan, cols, self.k, self.ext_k, *lim, d_h, divide=True)
        self.dom.extended_to_coeff_dev(d_h)
        pieces = [d_h + i * self.n * 32 for i in range(self.n_pieces)]
        for i, p in enumerate(pieces):
            self.coeff[("h_piece", i)] = p
        return self._commit_dev(pieces)

    def _commit_dev(self, d_polys):
        d_out = self._buf("pts", 160 * 64)
        assert len(d_polys) <= 64
        self.ctx.msm_g1_batch_dev(d_polys, self.n, d_out, srs_id=self.srs_g)
        pts = self.ctx.d2h(d_out, 20 * len(d_polys)).reshape(len(d_polys), 20)
        return pts[:, :8].copy()

    # -- stage 5: evaluation round -------------------------------------------------------------------
    def rotate_omega(self, x, rot):
        w = self.dom._omega if rot >= 0 else self.dom._omega_inv
        return x * pow(w, abs(rot), _R) % _R

    def fold_h(self, x):
        """vanishing::Constructed::evaluate: h(X) = sum_i piece_i(X) * x^(n i), kept under ("h", 0)."""
        xn = pow(x, self.n, _R)
        d = self._buf("h_folded", self.n * 32)
        self.ctx.poly_fold_dev([self.coeff[("h_piece", i)] for i in reversed(range(self.n_pieces))], self.n,
                               fr_to_limbs(xn), d)
        self.coeff[("h", 0)] = d

    def evaluate(self, queries, x):
        """queries: [(column name, rotation)] -> evaluations at x * omega^rotation, shape (len, 4) Montgomery limbs."""
        assert len(queries) <= 1024
        self.join_transforms()
        d_ev = self._buf("evals", 32 * 1024)
        groups = {}  # one batched call per opening point
        for i, (nm, rot) in enumerate(queries):
            groups.setdefault(rot, []).append(i)
        off, where = 0, [0] * len(queries)
        for rot, idxs in groups.items():
            self.ctx.eval_polynomials_dev([self.coeff[queries[i][0]] for i in idxs], self.n,
                                          fr_to_limbs(self.rotate_omega(x, rot)), d_ev + 32 * off)
            for j, i in enumerate(idxs):
                where[i] = off + j
            off += len(idxs)
        flat = self.ctx.d2h(d_ev, 4 * len(queries)).reshape(len(queries), 4)
        return flat[where]

    # -- stage 6: GWC multi-opening ---------------------------------------------------------------------
    def open(self, queries, x, v):
        """halo2_proofs poly/kzg/multiopen/gwc/prover.rs: queries are grouped by point in order of first appearance;
        per point  poly_batch = sum_i v^i poly_i  (the queries of a point are zipped with powers(v); the reference's
        verifier rebuilds exactly this combination, halo2-snark-aggregator-api/src/systems/halo2/multiopen.rs:55-61:
        `.rev().reduce(|acc, q| v * acc + q)`),  W = commit(kate_division(poly_batch - eval_batch, point)).
        The constant eval_batch only changes the remainder kate_division drops, so it is not subtracted here.
        Returns (rotations in point order, affine W commitments (len, 8))."""
        self.join_transforms()
        order, groups = [], {}
        for nm, rot in queries:
            if rot not in groups:
                groups[rot] = []
                order.append(rot)
            groups[rot].append(self.coeff[nm])
        d_fold = self._buf("fold", self.n * 32)
        ws = []
        for j, rot in enumerate(order):
            d_w = self._buf(("w", j), self.n * 32)
            self.ctx.poly_fold_dev(groups[rot][::-1], self.n, fr_to_limbs(v), d_fold)   # poly_fold: first argument = highest power
            self.ctx.kate_division_dev(d_fold, self.n, fr_to_limbs(self.rotate_omega(x, rot)), d_w)
            ws.append(d_w)
        return order, self._commit_dev(ws)


class ProvingKeyCache:
    """The reference recomputes keygen_pk in every verify_run (verify_circuit.rs:974-979: the proving key is never
    persisted).  Here a prover whose key is resident is kept per (device context, circuit shape, k,
    caller's key id -- e.g. the vk digest) for the life of the process, so the 26 x (MSM + iNTT + coset NTT) of the key are
    paid once, not per proof."""

    def __init__(self):
        self._provers = {}

    def get(self, key_id, make_prover, fill):
        """make_prover() -> a fresh ResidentProver; fill as for keygen_pk.  -> (prover, commitments, cached?)"""
        if key_id in self._provers:
            pr, comm = self._provers[key_id]
            return pr, comm, True
        pr = make_prover()
        comm = pr.keygen_pk(fill)
        self._provers[key_id] = (pr, comm)
        return pr, comm, False

    def drop(self, key_id):
        pr, _ = self._provers.pop(key_id)
        pr.close()


def create_proof_queries(cs):
    """The (polynomial, rotation) list create_proof opens for a constraint system, in halo2's order -- the order the
    reference's verifier rebuilds it in (halo2-snark-aggregator-api/src/systems/halo2/params.rs:156-224): instance
    queries, advice queries (both in REGISTRATION order, `ConstraintSystem.queries`), permutation products
    (permutation.rs:138-181: (x, omega x) for every set first, then omega^last x for all but the last set in REVERSE set
    order), lookups (lookup.rs:119-165: z, a', s' at x, a' at omega^-1 x, z at omega x), fixed queries, permutation
    sigmas, vanishing (h, random; vanish.rs:60-75).  GWC folds poly_batch = poly_batch * v + poly in this order per
    opening point, so the W commitments depend on it."""
    out = [(("instance", c), r) for c, r in cs.queries["instance"]]
    out += [(("advice", c), r) for c, r in cs.queries["advice"]]
    last = -(cs.blinding_factors() + 1)
    sets = cs.num_permutation_sets()
    for s in range(sets):
        out += [(("perm_z", s), 0), (("perm_z", s), 1)]
    for s in reversed(range(sets - 1)):
        out.append((("perm_z", s), last))
    for i in range(len(cs.lookups)):
        out += [(("lookup_z", i), 0), (("lookup_input", i), 0), (("lookup_table", i), 0),
                (("lookup_input", i), -1), (("lookup_z", i), 1)]
    out += [(("fixed", c), r) for c, r in cs.queries["fixed"]]
    out += [(("sigma", j), 0) for j in range(len(cs.permutation_columns))]
    out += [(("h", 0), 0), (("random", 0), 0)]
    return out


def transcript_eval_order(cs):
    """The order create_proof WRITES its evaluations to the transcript (halo2_proofs plonk/prover.rs), which is the
    order the reference's verifier reads them back (halo2-snark-aggregator-api/src/systems/halo2/verify.rs:446-462,
    :198-230 permutation sets, :294-312 lookups): instance evals, advice evals, fixed evals, random_eval, the
    permutation sigmas, per permutation set z(x), z(omega x) and -- except for the last set -- z(omega^last x), per
    lookup z(x), z(omega x), a'(x), a'(omega^-1 x), s'(x).  h(x) is never written: the verifier derives it."""
    out = [(("instance", c), r) for c, r in cs.queries["instance"]]
    out += [(("advice", c), r) for c, r in cs.queries["advice"]]
    out += [(("fixed", c), r) for c, r in cs.queries["fixed"]]
    out.append((("random", 0), 0))
    out += [(("sigma", j), 0) for j in range(len(cs.permutation_columns))]
    last = -(cs.blinding_factors() + 1)
    sets = cs.num_permutation_sets()
    for s in range(sets):
        out += [(("perm_z", s), 0), (("perm_z", s), 1)]
        if s + 1 < sets:
            out.append((("perm_z", s), last))
    for i in range(len(cs.lookups)):
        out += [(("lookup_z", i), 0), (("lookup_z", i), 1), (("lookup_input", i), 0), (("lookup_input", i), -1),
                (("lookup_table", i), 0)]
    return out
