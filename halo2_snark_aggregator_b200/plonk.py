"""Host mirror of the slice of halo2's constraint-system model that `evaluate_h` needs.

halo2_proofs is an external crate; the names below follow its `plonk::Expression`,
`ConstraintSystem`, `lookup::Argument` and `permutation::Argument`.  The aggregation circuit's own shape is
taken from the reference:
    gate      halo2-ecc-circuit-lib/src/gates/base_gate.rs:692-729
    lookups   halo2-ecc-circuit-lib/src/five/range_gate.rs:38-93
    columns   halo2-snark-aggregator-circuit/src/verify_circuit.rs:225-241

`build_quotient_plan` turns a constraint system into the word program `h2agg_evaluate_h_dev` executes
(layout in include/h2agg.h): every polynomial is expanded into a sum of products of column queries.
Nothing here computes field arithmetic on columns -- that happens on the GPU.
"""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_M64 = (1 << 64) - 1
PLAN_MAGIC = 0x31485148
NOCONST = 0xFFFFFFFF
# halo2curves bn256::Fr constants
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
DELTA = pow(7, 1 << 28, R_MOD)
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)


def fr_mont(x):
    """canonical integer -> 4 x u64 Montgomery limbs (the Rust in-memory form)"""
    v = (x % R_MOD) * (1 << 256) % R_MOD
    return np.array([(v >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


class Expression:
    """plonk::Expression: Constant | Fixed | Advice | Instance | Negated | Sum | Product | Scaled."""
    __slots__ = ("kind", "a", "b")

    def __init__(self, kind, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b

    # constructors -----------------------------------------------------------------------------
    @staticmethod
    def constant(v):
        return Expression("const", int(v) % R_MOD)

    @staticmethod
    def fixed(col, rot=0):
        return Expression("fixed", int(col), int(rot))

    @staticmethod
    def advice(col, rot=0):
        return Expression("advice", int(col), int(rot))

    @staticmethod
    def instance(col, rot=0):
        return Expression("instance", int(col), int(rot))

    # operators ---------------------------------------------------------------------------------
    def __add__(self, o):
        return Expression("sum", self, _expr(o))

    def __sub__(self, o):
        return Expression("sum", self, -_expr(o))

    def __neg__(self):
        return Expression("neg", self)

    def __mul__(self, o):
        if isinstance(o, int):
            return Expression("scaled", self, o % R_MOD)
        return Expression("product", self, o)

    def degree(self):
        k = self.kind
        if k == "const":
            return 0
        if k in ("fixed", "advice", "instance"):
            return 1
        if k in ("neg", "scaled"):
            return self.a.degree()
        if k == "sum":
            return max(self.a.degree(), self.b.degree())
        return self.a.degree() + self.b.degree()

    def to_tuple(self):
        """nested tuples, the form oracle/py/quotient_ref.py evaluates"""
        k = self.kind
        if k == "const":
            return ("const", self.a)
        if k in ("fixed", "advice", "instance"):
            return (k, self.a, self.b)
        if k == "neg":
            return ("neg", self.a.to_tuple())
        if k == "scaled":
            return ("scaled", self.a.to_tuple(), self.b)
        return (k, self.a.to_tuple(), self.b.to_tuple())

    def queries(self):
        k = self.kind
        if k == "const":
            return set()
        if k in ("fixed", "advice", "instance"):
            return {(k, self.a, self.b)}
        if k in ("neg", "scaled"):
            return self.a.queries()
        return self.a.queries() | self.b.queries()

    def leaves(self):
        """(kind, col, rot) of every column query, left to right -- the order in which the closure that built the
        expression called meta.query_*, which is the order halo2 registers the queries in."""
        k = self.kind
        if k == "const":
            return []
        if k in ("fixed", "advice", "instance"):
            return [(k, self.a, self.b)]
        if k in ("neg", "scaled"):
            return self.a.leaves()
        return self.a.leaves() + self.b.leaves()

    def expand(self):
        """{sorted tuple of (kind, col, rot) queries: coefficient mod r}: the sum-of-products form"""
        k = self.kind
        if k == "const":
            return {(): self.a} if self.a else {}
        if k in ("fixed", "advice", "instance"):
            return {((k, self.a, self.b),): 1}
        if k == "neg":
            return {m: (-c) % R_MOD for m, c in self.a.expand().items()}
        if k == "scaled":
            out = {m: c * self.b % R_MOD for m, c in self.a.expand().items()}
            return {m: c for m, c in out.items() if c}
        if k == "sum":
            out = dict(self.a.expand())
            for m, c in self.b.expand().items():
                out[m] = (out.get(m, 0) + c) % R_MOD
            return {m: c for m, c in out.items() if c}
        out = {}
        ea, eb = self.a.expand(), self.b.expand()
        for ma, ca in ea.items():
            for mb, cb in eb.items():
                m = tuple(sorted(ma + mb))
                out[m] = (out.get(m, 0) + ca * cb) % R_MOD
        return {m: c for m, c in out.items() if c}


def _expr(o):
    return o if isinstance(o, Expression) else Expression.constant(o)


class ConstraintSystem:
    """The part of plonk::ConstraintSystem that shapes the quotient."""

    def __init__(self, num_fixed, num_advice, num_instance):
        self.num_fixed, self.num_advice, self.num_instance = num_fixed, num_advice, num_instance
        self.gates = []        # list of (name, [Expression])
        self.lookups = []      # list of (name, [input Expression], [table Expression])
        self.permutation_columns = []  # [(kind, index)] in enable_equality order
        # ConstraintSystem::{advice,fixed,instance}_queries: (column, rotation) in REGISTRATION order.  create_proof
        # writes its evaluations and GWC folds its polynomials in this order, and the reference's verifier walks the
        # same lists (halo2-snark-aggregator-api/src/systems/halo2/params.rs:156-205), so it is part of the format.
        self.queries = {"advice": [], "fixed": [], "instance": []}

    def _query(self, kind, col, rot):
        """ConstraintSystem::query_{advice,fixed,instance}_index: append (column, rotation) unless already there."""
        q = (col, rot)
        if q not in self.queries[kind]:
            self.queries[kind].append(q)

    def create_gate(self, name, polys):
        """The caller's closure queries cells while it builds the polynomials: leaves left to right (write the
        Python expression in the order the Rust closure calls meta.query_*)."""
        polys = list(polys)
        for p in polys:
            for kind, col, rot in p.leaves():
                self._query(kind, col, rot)
        self.gates.append((name, polys))

    def lookup(self, name, pairs):
        """ConstraintSystem::lookup: table_map runs first (it queries the input cells), then every TableColumn is
        queried as a fixed column at the current row."""
        pairs = list(pairs)
        for inp, _ in pairs:
            for kind, col, rot in inp.leaves():
                self._query(kind, col, rot)
        for _, tab in pairs:
            for kind, col, rot in tab.leaves():
                self._query(kind, col, rot)
        self.lookups.append((name, [p[0] for p in pairs], [p[1] for p in pairs]))

    def enable_equality(self, kind, index):
        """ConstraintSystem::enable_equality: query_any_index(column, Rotation::cur()) then permutation.add_column."""
        self._query(kind, index, 0)
        if (kind, index) not in self.permutation_columns:
            self.permutation_columns.append((kind, index))

    def degree(self):
        """ConstraintSystem::degree: permutation needs 3, a lookup max(4, 2 + deg(input) + deg(table))."""
        d = 3 if self.permutation_columns else 1
        for _, ins, tabs in self.lookups:
            di = max([1] + [e.degree() for e in ins])
            dt = max([1] + [e.degree() for e in tabs])
            d = max(d, 4, 2 + di + dt)
        for _, polys in self.gates:
            for p in polys:
                d = max(d, p.degree())
        return d

    def blinding_factors(self):
        """max(3, most queries on one advice column) + 2 (ConstraintSystem::blinding_factors)."""
        per_col = {}
        for col, _ in self.queries["advice"]:
            per_col[col] = per_col.get(col, 0) + 1
        return max([3] + list(per_col.values())) + 2

    def chunk_len(self):
        return self.degree() - 2

    def num_permutation_sets(self):
        n = len(self.permutation_columns)
        return (n + self.chunk_len() - 1) // self.chunk_len() if n else 0

    def extended_k(self, k):
        """EvaluationDomain::new(j = degree, k): quotient degree j - 1, extended_k = k + ceil(log2(j - 1))"""
        q = self.degree() - 1
        e = 0
        while (1 << e) < q:
            e += 1
        return k + e


def aggregation_circuit_cs():
    """Constraint system of Halo2VerifierCircuit(s): FiveColumnBaseGate + FiveColumnRangeGate + 1 instance column."""
    V, M = 5, 2
    cs = ConstraintSystem(num_fixed=17, num_advice=5, num_instance=1)
    coeff = list(range(0, 5))
    mul_coeff = [5, 6]
    next_coeff, constant = 7, 8
    for i in range(V):
        cs.enable_equality("advice", i)
    E = Expression
    # base_gate.rs:701-720
    acc = E.fixed(constant) + E.advice(V - 1, 1) * E.fixed(next_coeff)
    for i in range(V):
        acc = acc + E.advice(i) * E.fixed(coeff[i])
    for i in range(M):
        acc = acc + E.advice(2 * i) * E.advice(2 * i + 1) * E.fixed(mul_coeff[i])
    cs.create_gate("base_gate", [acc])
    # five/range_gate.rs:42-80: selector 9 / table 10 on base[0..4], then three leading-limb lookups on base[0]
    for col in range(V - 1):
        cs.lookup("common range", [(E.advice(col) * E.fixed(9), E.fixed(10))])
    for sel, tab, name in ((11, 12, "w ceil leading limb range"), (13, 14, "n floor leading limb range"),
                           (15, 16, "d leading limb range")):
        cs.lookup(name, [(E.advice(0) * E.fixed(sel), E.fixed(tab))])
    # verify_circuit.rs:233-234
    cs.enable_equality("instance", 0)
    return cs


class QuotientPlan:
    """Word program + constants + the column order h2agg_evaluate_h_dev expects."""

    def __init__(self, cs):
        self.cs = cs
        names = [("fixed", i) for i in range(cs.num_fixed)]
        names += [("advice", i) for i in range(cs.num_advice)]
        names += [("instance", i) for i in range(cs.num_instance)]
        names += [("sigma", i) for i in range(len(cs.permutation_columns))]
        names += [("l0", 0), ("l_last", 0), ("l_active_row", 0)]
        names += [("perm_z", i) for i in range(cs.num_permutation_sets())]
        for i in range(len(cs.lookups)):
            names += [("lookup_z", i), ("lookup_input", i), ("lookup_table", i)]
        self.columns = names
        self.index = {n: i for i, n in enumerate(names)}
        self._consts = []
        self._const_index = {}
        w = [PLAN_MAGIC, sum(len(p) for _, p in cs.gates), len(cs.permutation_columns), cs.chunk_len(),
             (-(cs.blinding_factors() + 1)) & 0xFFFFFFFF, len(cs.lookups),
             self.index[("l0", 0)], self.index[("l_last", 0)], self.index[("l_active_row", 0)]]
        for _, polys in cs.gates:
            for p in polys:
                w += self._poly(p)
        if cs.permutation_columns:
            for j, (kind, idx) in enumerate(cs.permutation_columns):
                w += [self.index[(kind, idx)], self.index[("sigma", j)]]
            w += [self.index[("perm_z", s)] for s in range(cs.num_permutation_sets())]
        for i, (_, ins, tabs) in enumerate(cs.lookups):
            for side in (ins, tabs):
                w.append(len(side))
                for e in side:
                    w += self._poly(e)
            w += [self.index[("lookup_z", i)], self.index[("lookup_input", i)], self.index[("lookup_table", i)]]
        self.words = np.array(w, dtype=np.uint32)
        self.consts = (np.concatenate([fr_mont(c) for c in self._consts]) if self._consts
                       else np.zeros(0, dtype=np.uint64))

    def _const(self, c):
        if c not in self._const_index:
            self._const_index[c] = len(self._consts)
            self._consts.append(c)
        return self._const_index[c]

    def _poly(self, expr):
        terms = expr.expand()
        w = [len(terms)]
        for mono in sorted(terms):
            c = terms[mono]
            w += [NOCONST if c == 1 else self._const(c), len(mono)]
            for kind, col, rot in mono:
                assert -32768 <= rot < 32768
                w.append(self.index[(kind, col)] | ((rot & 0xFFFF) << 16))
        return w


class ExpressionList:
    """{ n_exprs, POLY x n_exprs } word list + constants for h2agg_compress_expressions_dev; `index` maps
    (kind, column) to a position in the caller's column array."""

    def __init__(self, exprs, index):
        self._consts, self._const_index = [], {}
        w = [len(exprs)]
        for e in exprs:
            terms = e.expand()
            w.append(len(terms))
            for mono in sorted(terms):
                c = terms[mono]
                if c == 1:
                    w.append(NOCONST)
                else:
                    if c not in self._const_index:
                        self._const_index[c] = len(self._consts)
                        self._consts.append(c)
                    w.append(self._const_index[c])
                w.append(len(mono))
                for kind, col, rot in mono:
                    assert -32768 <= rot < 32768
                    w.append(index[(kind, col)] | ((rot & 0xFFFF) << 16))
        self.words = np.array(w, dtype=np.uint32)
        self.consts = (np.concatenate([fr_mont(c) for c in self._consts]) if self._consts
                       else np.zeros(0, dtype=np.uint64))


def build_quotient_plan(cs):
    return QuotientPlan(cs)


def t_evaluations(k, ext_k):
    """EvaluationDomain::t_evaluations: 1 / ((zeta omega_ext^i)^n - 1) for i < 2^(ext_k - k), canonical ints"""
    n = 1 << k
    w_ext = pow(ROOT_OF_UNITY, 1 << (28 - ext_k), R_MOD)
    return [pow((pow(ZETA * pow(w_ext, i, R_MOD) % R_MOD, n, R_MOD) - 1) % R_MOD, -1, R_MOD)
            for i in range(1 << (ext_k - k))]
