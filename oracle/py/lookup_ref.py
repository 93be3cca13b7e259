"""Python big-int pin for the lookup argument's `permute_expression_pair`.  TEST INFRASTRUCTURE.

halo2_proofs plonk/lookup/prover.rs (external crate; scroll-tech/halo2 @ 3370852d, Cargo.lock:1549-1551) -- restated
from its published algorithm.  What the reference itself holds about this step is the verifier's view of its result
(halo2-snark-aggregator-api/src/systems/halo2/lookup.rs:58-119): the permuted columns must satisfy
    (a'(X) - s'(X)) * (a'(X) - a'(omega^-1 X)) = 0   on active rows,   a'(1) = s'(1),
and be permutations of the inputs (the grand product z).  `check_lookup_constraints` states exactly those properties;
the tests hold both the restated algorithm and the GPU result against them.
"""


def permute_expression_pair(inp, tab):
    """ints (canonical) -> (permuted_input, permuted_table); raises ValueError if an input value is not in the table."""
    u = len(inp)
    assert len(tab) == u
    a = sorted(inp)
    leftover = {}
    for v in tab:
        leftover[v] = leftover.get(v, 0) + 1
    s = [0] * u
    repeated = []
    for row, v in enumerate(a):
        if row == 0 or v != a[row - 1]:
            s[row] = v
            if leftover.get(v, 0) == 0:
                raise ValueError("ConstraintSystemFailure: input value not in table")
            leftover[v] -= 1
        else:
            repeated.append(row)
    for v in sorted(leftover):          # BTreeMap iteration = ascending
        for _ in range(leftover[v]):
            s[repeated.pop()] = v       # Vec::pop = from the back
    assert not repeated
    return a, s


def check_lookup_constraints(inp, tab, a, s):
    """The properties the verifier enforces on (a', s') -- independent of how they were produced."""
    assert sorted(a) == sorted(inp), "a' is not a permutation of the input"
    assert sorted(s) == sorted(tab), "s' is not a permutation of the table"
    assert a[0] == s[0]
    for i in range(1, len(a)):
        assert a[i] == s[i] or a[i] == a[i - 1], "row %d violates (a'-s')(a'-a'_prev) = 0" % i


# ---- the row-wise steps around the sort (halo2_proofs plonk/lookup/prover.rs, plonk/permutation/prover.rs) -------------
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
DELTA = pow(7, 1 << 28, R)


def compress_expressions(exprs, cols, n, theta):
    """lookup::Argument::commit_permuted's compress_expressions: every expression (nested tuples as in quotient_ref)
    is evaluated on the Lagrange domain -- rotation r of row i reads row (i + r) mod n -- and folded acc*theta + e."""
    import quotient_ref as qr

    out = []
    for i in range(n):
        acc = 0
        for e in exprs:
            acc = (acc * theta + qr.eval_expr(e, lambda kind, col, rot: cols[(kind, col)][(i + rot) % n])) % R
        out.append(acc)
    return out


def lookup_product(A, S, Ap, Sp, beta, gamma):
    """lookup::Permuted::commit_product before blinding: z[0] = 1, z[i+1] = z[i] (A+b)(S+g) / ((A'+b)(S'+g)); n values."""
    z = [1]
    for i in range(len(A) - 1):
        num = (A[i] + beta) * (S[i] + gamma) % R
        den = (Ap[i] + beta) * (Sp[i] + gamma) % R
        z.append(z[-1] * num % R * pow(den, -1, R) % R)
    return z


def permutation_product(values, sigmas, k, omega, beta, gamma, first_index, last_z=1):
    """permutation::Argument::commit for ONE column set, before blinding: z[0] = last_z,
    z[i+1] = z[i] prod_j (v_j[i] + beta delta^(first+j) omega^i + gamma) / (v_j[i] + beta sigma_j[i] + gamma)."""
    n = 1 << k
    z = [last_z % R]
    w = 1
    for i in range(n - 1):
        num = den = 1
        d = beta * pow(DELTA, first_index, R) % R
        for v, s in zip(values, sigmas):
            num = num * (v[i] + d * w + gamma) % R
            den = den * (v[i] + beta * s[i] + gamma) % R
            d = d * DELTA % R
        z.append(z[-1] * num % R * pow(den, -1, R) % R)
        w = w * omega % R
    return z
