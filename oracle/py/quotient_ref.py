"""CPU oracle for the quotient numerator.  TEST INFRASTRUCTURE: imported only by tests/.

Two independent halves, both on Python big integers:

* `evaluate_h` restates the PROVER side -- halo2_proofs plonk/evaluation.rs `Evaluator::evaluate_h`
  and `EvaluationDomain::divide_by_vanishing_poly` (external crate, pinned at
  scroll-tech/halo2 @ 3370852d, Cargo.lock:1549-1551; restated from its published algorithm, the source is not
  under /root/reference).  It walks expression TREES (nested tuples), not the sum-of-products plan the GPU runs.

* `verifier_h_eval` restates the VERIFIER side from the reference's own files, which pins every formula and the
  order of the y-fold:
      expression order: gates, permutation, lookups   halo2-snark-aggregator-api/src/systems/halo2/params.rs:95-150
      permutation terms                                .../permutation.rs:54-136
      lookup terms                                     .../lookup.rs:58-119
      l_0, l_last, l_blind                             .../lagrange.rs:16-39, params.rs:80-88
      fold with y (acc * y + v), divide by x^n - 1     .../vanish.rs:28-29, arith/field.rs:68-81
  Pin: for columns given as polynomials, evaluate_h at coset row i must equal verifier_h_eval at the point
  X_i = zeta * omega_ext^i (tests/test_quotient_cpu.py).  That identity holds for ANY column contents, so it
  needs no satisfying witness.
"""
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
DELTA = pow(7, 1 << 28, R)
ROOT_OF_UNITY = pow(7, (R - 1) >> 28, R)


def omega(k):
    return pow(ROOT_OF_UNITY, 1 << (28 - k), R)


def eval_expr(t, query):
    """plonk::Expression::evaluate over nested tuples; query(kind, col, rot) -> int"""
    k = t[0]
    if k == "const":
        return t[1] % R
    if k in ("fixed", "advice", "instance"):
        return query(k, t[1], t[2])
    if k == "neg":
        return (-eval_expr(t[1], query)) % R
    if k == "scaled":
        return eval_expr(t[1], query) * t[2] % R
    if k == "sum":
        return (eval_expr(t[1], query) + eval_expr(t[2], query)) % R
    if k == "product":
        return eval_expr(t[1], query) * eval_expr(t[2], query) % R
    raise ValueError(k)


def compress(exprs, theta, query):
    acc = 0
    for e in exprs:
        acc = (acc * theta + eval_expr(e, query)) % R
    return acc


def evaluate_h(desc, cols, k, ext_k, y, beta, gamma, theta, rows=None):
    """rows: only these coset rows (default all); cols[name] only needs __getitem__.
    desc: dict(gates=[tuple], lookups=[(inputs, tables)], perm_columns=[(kind, idx)], chunk_len, last_rotation)
    cols: dict name -> list of 2^ext_k ints, names ("fixed", i), ("advice", i), ("instance", i), ("sigma", j),
    ("l0", 0), ("l_last", 0), ("l_active_row", 0), ("perm_z", s), ("lookup_z" | "lookup_input" | "lookup_table", i)."""
    size = 1 << ext_k
    rot_scale = 1 << (ext_k - k)
    w_ext = omega(ext_k)
    l0, l_last, l_active = cols[("l0", 0)], cols[("l_last", 0)], cols[("l_active_row", 0)]
    pcols = desc["perm_columns"]
    chunk = desc["chunk_len"]
    n_sets = (len(pcols) + chunk - 1) // chunk if pcols else 0
    out = []
    for idx in (range(size) if rows is None else rows):
        beta_term = pow(w_ext, idx, R)
        def rot(r, idx=idx):
            return (idx + r * rot_scale) % size

        def query(kind, c, r, idx=idx):
            return cols[(kind, c)][(idx + r * rot_scale) % size]

        v = 0
        for g in desc["gates"]:
            v = (v * y + eval_expr(g, query)) % R
        if pcols:
            z = [cols[("perm_z", s)] for s in range(n_sets)]
            v = (v * y + (1 - z[0][idx]) * l0[idx]) % R
            v = (v * y + (z[-1][idx] * z[-1][idx] - z[-1][idx]) * l_last[idx]) % R
            for s in range(1, n_sets):
                v = (v * y + (z[s][idx] - z[s - 1][rot(desc["last_rotation"])]) * l0[idx]) % R
            current_delta = beta * ZETA % R * beta_term % R
            for s in range(n_sets):
                left = z[s][rot(1)]
                right = z[s][idx]
                for j in range(s * chunk, min((s + 1) * chunk, len(pcols))):
                    val = cols[pcols[j]][idx]
                    left = left * (val + beta * cols[("sigma", j)][idx] + gamma) % R
                    right = right * (val + current_delta + gamma) % R
                    current_delta = current_delta * DELTA % R
                v = (v * y + (left - right) * l_active[idx]) % R
        for i, (ins, tabs) in enumerate(desc["lookups"]):
            table_value = (compress(ins, theta, query) + beta) * (compress(tabs, theta, query) + gamma) % R
            zc, ac, sc = cols[("lookup_z", i)], cols[("lookup_input", i)], cols[("lookup_table", i)]
            a_minus_s = (ac[idx] - sc[idx]) % R
            v = (v * y + (1 - zc[idx]) * l0[idx]) % R
            v = (v * y + (zc[idx] * zc[idx] - zc[idx]) * l_last[idx]) % R
            v = (v * y + (zc[rot(1)] * (ac[idx] + beta) % R * (sc[idx] + gamma) - zc[idx] * table_value) * l_active[idx]) % R
            v = (v * y + a_minus_s * l0[idx]) % R
            v = (v * y + a_minus_s * (ac[idx] - ac[rot(-1)]) % R * l_active[idx]) % R
        out.append(v)
    return out


def divide_by_vanishing_poly(h, k, ext_k):
    """EvaluationDomain::divide_by_vanishing_poly: h[i] *= t_evaluations[i % 2^(ext_k - k)]"""
    n = 1 << k
    w_ext = omega(ext_k)
    m = 1 << (ext_k - k)
    t = [pow((pow(ZETA * pow(w_ext, i, R) % R, n, R) - 1) % R, -1, R) for i in range(m)]
    return [h[i] * t[i % m] % R for i in range(len(h))]


# ---- verifier side (restated from the reference) --------------------------------------------------

def lagrange_evals(k, l, x):
    """lagrange.rs:16-39: ls[i] = (w_i / n) (x^n - 1) / (x - w_i), w_i = omega^-i, i = 0..l"""
    n = 1 << k
    w_inv = pow(omega(k), -1, R)
    xn = pow(x, n, R)
    n_inv = pow(n, -1, R)
    out = []
    wi = 1
    for _ in range(l + 1):
        out.append(wi * n_inv % R * (xn - 1) % R * pow((x - wi) % R, -1, R) % R)
        wi = wi * w_inv % R
    return out


def verifier_h_eval(desc, ev, k, x, y, beta, gamma, theta):
    """expected_h_eval of vanish.rs:28-29.  ev(name, rot) -> evaluation at x * omega^rot of the polynomial `name`
    (names as in evaluate_h).  blinding_factors = -last_rotation - 1, l = blinding_factors + 1 (verify.rs:440)."""
    l = -desc["last_rotation"]
    ls = lagrange_evals(k, l, x)
    l_0, l_last = ls[0], ls[l]
    l_blind = sum(ls[1:l]) % R                                   # params.rs:83-87
    exprs = []

    def query(kind, c, r):
        return ev((kind, c), r)

    for g in desc["gates"]:                                      # params.rs:103-117
        exprs.append(eval_expr(g, query))
    pcols, chunk = desc["perm_columns"], desc["chunk_len"]
    if pcols:                                                    # permutation.rs:54-136
        n_sets = (len(pcols) + chunk - 1) // chunk
        z = [ev(("perm_z", s), 0) for s in range(n_sets)]
        exprs.append(l_0 * (1 - z[0]) % R)
        exprs.append(l_last * (z[-1] * z[-1] - z[-1]) % R)
        for s in range(1, n_sets):
            exprs.append((z[s] - ev(("perm_z", s - 1), desc["last_rotation"])) * l_0 % R)
        t0 = beta * x % R
        t1 = (1 - (l_last + l_blind)) % R
        for s in range(n_sets):
            left = ev(("perm_z", s), 1)
            right = z[s]
            d = t0 * (pow(DELTA, s * chunk, R) if s else 1) % R
            for j in range(s * chunk, min((s + 1) * chunk, len(pcols))):
                t2 = (ev(pcols[j], 0) + gamma) % R
                left = (t2 + beta * ev(("sigma", j), 0)) * left % R
                right = (t2 + d) * right % R
                d = DELTA * d % R
            exprs.append((left - right) * t1 % R)
    for i, (ins, tabs) in enumerate(desc["lookups"]):            # lookup.rs:58-119
        z_wx, z_x = ev(("lookup_z", i), 1), ev(("lookup_z", i), 0)
        a_x, s_x, a_invwx = ev(("lookup_input", i), 0), ev(("lookup_table", i), 0), ev(("lookup_input", i), -1)
        left = z_wx * (a_x + beta) % R * (s_x + gamma) % R
        input_eval = compress(ins, theta, query)
        table_eval = compress(tabs, theta, query)
        t0 = (1 - (l_last + l_blind)) % R
        t1 = (a_x - s_x) % R
        exprs += [l_0 * (1 - z_x) % R,
                  l_last * (z_x * z_x - z_x) % R,
                  (left - z_x * (input_eval + beta) % R * (table_eval + gamma)) * t0 % R,
                  l_0 * t1 % R,
                  t1 * (a_x - a_invwx) % R * t0 % R]
    acc = 0
    for v in exprs:                                              # vanish.rs:28, field.rs:68-81
        acc = (acc * y + v) % R
    return acc * pow((pow(x, 1 << k, R) - 1) % R, -1, R) % R     # vanish.rs:29


# ---- plain polynomial helpers for the tests ---------------------------------------------------------

def horner(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def lagrange_to_coeff(vals, k):
    """O(n^2)-free small iDFT by recursion is not needed: sizes are tiny, do the plain sum"""
    n = 1 << k
    w_inv = pow(omega(k), -1, R)
    n_inv = pow(n, -1, R)
    out = []
    for j in range(n):
        wj = pow(w_inv, j, R)
        acc, p = 0, 1
        for i in range(n):
            acc = (acc + vals[i] * p) % R
            p = p * wj % R
        out.append(acc * n_inv % R)
    return out


def coeff_to_extended(coeffs, k, ext_k):
    """evaluations at zeta * omega_ext^i (distribute_powers_zeta + zero-pad + FFT, as a plain sum)"""
    w_ext = omega(ext_k)
    return [horner(coeffs, ZETA * pow(w_ext, i, R) % R) for i in range(1 << ext_k)]
