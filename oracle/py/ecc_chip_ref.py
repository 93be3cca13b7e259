"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/cpu_halo2.cpp header).

Plain-Python restatement of the reference's witness synthesis for the aggregation circuit
(SURVEY.md 8a rows W1-W5), following halo2-ecc-circuit-lib line by line:
  W5  BaseGate::one_line / gate polynomial        gates/base_gate.rs:701-720, 739-807
      BaseGateOps (sum_with_constant, mul, ...)   gates/base_gate.rs:193-672, five/base_gate.rs:16-129
      range selectors                             gates/range_gate.rs:86-182, tables :198-294
  W1  FiveColumnIntegerChip::{mul, square}        five/integer_chip.rs:104-320, 709-743
  W2  reduce/div/is_zero/add/sub/neg/...          five/integer_chip.rs:31-102, 324-901
  W3  EccChipOps::{add, double, curvature, ...}   chips/ecc_chip.rs:280-580
  W4  shamir / mul / constant_mul                 chips/ecc_chip.rs:86-279
      NativeEccChip::decompose_scalar             chips/native_ecc_chip.rs:42-132
(paths relative to /root/reference/halo2-ecc-circuit-lib/src/).

Every advice AND fixed cell is produced, together with the copy constraints, so `check()` can
play MockProver: gate polynomial on every row, the four range lookups, copy constraints.
The reference pins no concrete witness bytes (its tests are MockProver runs on time-seeded inputs,
SURVEY.md section 4); what pins this restatement is (i) that checker, (ii) native Fq/G1 results,
(iii) the row counts of SURVEY.md App. G, which reproduce the reference's own estimator constant
`ecmul_rows = 32196` (halo2-snark-aggregator-api/src/systems/halo2/evaluation.rs:132).
"""
import copy

P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # W = Fq
R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # N = Fr

VAR_COLUMNS, MUL_COLUMNS = 5, 2                      # five/config.rs:1-2
LIMBS, COMMON_RANGE_BITS = 4, 17                     # five/integer_chip.rs:16-19
LIMB_WIDTH = 4 * COMMON_RANGE_BITS                   # 68
OVERFLOW_LIMIT, OVERFLOW_THRESHOLD = 64, 32          # five/integer_chip.rs:21-25
CONFIG_WINDOW_SIZE = 4                               # chips/ecc_chip.rs:70


class Cell:
    """AssignedValue / AssignedCondition: (column, row, value)."""
    __slots__ = ("col", "row", "value")

    def __init__(self, col, row, value):
        self.col, self.row, self.value = col, row, value


class Context:
    def __init__(self):
        self.rows = []      # dict(adv[5], coeff[5], mul[2], const, nxt, sel)
        self.copies = []    # ((col,row),(col,row))

    @property
    def offset(self):
        return len(self.rows)


# ------------------------------------------------------------------------------------ base gate
def one_line(ctx, pairs, constant, mul_next=((), 0), sel=None):
    """gates/base_gate.rs:739-807. pairs: [(Cell | int, coeff)]"""
    assert len(pairs) <= VAR_COLUMNS and len(mul_next[0]) <= MUL_COLUMNS
    row = len(ctx.rows)
    adv, coeff, cells = [0] * 5, [0] * 5, []
    for i, (base, c) in enumerate(pairs):
        v = base.value if isinstance(base, Cell) else base % R
        adv[i], coeff[i] = v, c % R
        if isinstance(base, Cell):
            ctx.copies.append(((base.col, base.row), (i, row)))
    for i in range(5):
        cells.append(Cell(i, row, adv[i]))
    mul = [m % R for m in mul_next[0]] + [0] * (2 - len(mul_next[0]))
    ctx.rows.append(dict(adv=adv, coeff=coeff, mul=mul, const=constant % R, nxt=mul_next[1] % R, sel=sel))
    return cells


def one_line_with_last_base(ctx, pairs, last, constant, mul_next):
    assert len(pairs) < VAR_COLUMNS
    pairs = list(pairs) + [(0, 0)] * (VAR_COLUMNS - 1 - len(pairs)) + [last]
    return one_line(ctx, pairs, constant, mul_next)


def sum_with_constant(ctx, elems, constant):
    """gates/base_gate.rs:193-262. elems: [(Cell, coeff)]"""
    acc, curr = None, 0
    while len(elems) - curr + (0 if acc is None else 1) + 1 > VAR_COLUMNS:
        line_len = VAR_COLUMNS - (0 if acc is None else 1)
        line = elems[curr:curr + line_len]
        curr += line_len
        line_sum = sum(v.value * c for v, c in line) % R
        if acc is None:
            one_line(ctx, [(v, c) for v, c in line], 0, ((), -1))
        else:
            one_line_with_last_base(ctx, [(v, c) for v, c in line], (acc, 1), 0, ((), -1))
        acc = ((acc or 0) + line_sum) % R
    s = (sum(v.value * c for v, c in elems[curr:]) + constant + (acc or 0)) % R
    pairs = [(s, -1)] + [(v, c) for v, c in elems[curr:]]
    if acc is None:
        cells = one_line(ctx, pairs, constant)
    else:
        cells = one_line_with_last_base(ctx, pairs, (acc, 1), constant, ((), 0))
    return cells[0]


def bg_add(ctx, a, b):
    return sum_with_constant(ctx, [(a, 1), (b, 1)], 0)


def bg_mul(ctx, a, b):  # :302-324
    return one_line(ctx, [(a, 0), (b, 0), (a.value * b.value % R, -1)], 0, ((1,), 0))[2]


def bg_mul_add_constant(ctx, a, b, c):  # :326-349
    d = (a.value * b.value + c) % R
    return one_line(ctx, [(a, 0), (b, 0), (d, -1)], c, ((1,), 0))[2]


def bg_div_unsafe(ctx, a, b):  # :478-497  b * c - a = 0, c = a / b (b = 0: `invert().unwrap()` panics)
    assert b.value % R, "div_unsafe: division by zero"
    c = pow(b.value, -1, R) * a.value % R
    return one_line(ctx, [(b, 0), (c, 0), (a, -1)], 0, ((1,), 0))[1]


def bg_mul_add(ctx, a, b, c, c_coeff):  # :351-380
    d = (a.value * b.value + c.value * c_coeff) % R
    return one_line(ctx, [(a, 0), (b, 0), (c, c_coeff), (d, -1)], 0, ((1,), 0))[3]


def bg_mul_add2(ctx, a, b, c, c_coeff, d, d_coeff):  # five/base_gate.rs:27-59
    e = (a.value * b.value + c.value * c_coeff + d.value * d_coeff) % R
    return one_line(ctx, [(a, 0), (b, 0), (c, c_coeff), (d, d_coeff), (e, -1)], 0, ((1,), 0))[4]


def bg_mul_add_with_next_line(ctx, ls):  # five/base_gate.rs:110-128
    a, b, c, cc = ls[0]
    acc = bg_mul_add(ctx, a, b, c, cc)
    for a, b, c, cc in ls[1:]:
        acc = bg_mul_add2(ctx, a, b, c, cc, acc, 1)
    return acc


def bg_invert(ctx, a):  # gates/base_gate.rs:439-476 -> (is_zero condition, inverse)
    b = pow(a.value, -1, R) if a.value else 0
    c = (1 - a.value * b) % R
    c = one_line(ctx, [(a, 0), (c, 0)], 0, ((1,), 0))[1]
    cells = one_line(ctx, [(a, 0), (b, 0), (c, 1)], -1, ((1,), 0))
    return cells[2], cells[1]


def bg_is_zero(ctx, a):
    return bg_invert(ctx, a)[0]


def bg_assign_constant(ctx, v):  # :499-505
    return one_line(ctx, [(v, -1)], v)[0]


def bg_assign(ctx, v):  # :507-511
    return one_line(ctx, [(v, 0)], 0)[0]


def bg_assert_constant(ctx, a, b):  # :525-538
    assert a.value % R == b % R, "assert_constant fails"
    one_line(ctx, [(a, -1)], b)


def bg_assert_bit(ctx, a):  # :540-552
    one_line(ctx, [(a, 1), (a, 0)], 0, ((-1,), 0))


def bg_and(ctx, a, b):
    return bg_mul(ctx, a, b)


def bg_not(ctx, a):  # :566-575
    return sum_with_constant(ctx, [(a, -1)], 1)


def bg_or(ctx, a, b):  # :577-596
    c = (a.value + b.value - a.value * b.value) % R
    return one_line(ctx, [(a, 1), (b, 1), (c, -1)], 0, ((-1,), 0))[2]


def bg_xnor(ctx, a, b):  # :620-639
    c = (1 - a.value - b.value + 2 * a.value * b.value) % R
    return one_line(ctx, [(a, -1), (b, -1), (c, -1)], 1, ((2,), 0))[2]


def bg_bisec(ctx, cond, a, b):  # five/base_gate.rs:82-108
    c = (cond.value * a.value + (1 - cond.value) * b.value) % R
    return one_line(ctx, [(cond, 0), (a, 0), (cond, 0), (b, 1), (c, -1)], 0, ((1, -1), 0))[4]


# ------------------------------------------------------------------------------------ integer chip
class Helper:  # chips/integer_chip.rs:94-128
    limb_modulus = 1 << LIMB_WIDTH
    integer_modulus = 1 << (LIMB_WIDTH * LIMBS)
    limb_modulus_on_n = (1 << LIMB_WIDTH) % R
    w_modulus, n_modulus = P, R
    w_native = P % R
    w_ceil_bits = P.bit_length()          # 254
    n_floor_bits = R.bit_length() - 1     # 253
    limb_modulus_exps = [pow(1 << LIMB_WIDTH, i, R) for i in range(LIMBS)]

    @staticmethod
    def bn_to_limb_le(bn):
        out = []
        for _ in range(LIMBS - 1):
            out.append(bn % Helper.limb_modulus)
            bn >>= LIMB_WIDTH
        out.append(bn)
        return out


def _lcm(a, b):
    from math import gcd
    return a // gcd(a, b) * b


# utils.rs:46-57
Helper.d_bits = ((_lcm(Helper.integer_modulus, R) >> Helper.w_ceil_bits) - 1).bit_length() - 1
assert (1 << Helper.d_bits) * P + P <= _lcm(Helper.integer_modulus, R) and Helper.d_bits == 271
Helper.w_modulus_limbs_le = Helper.bn_to_limb_le(P)

SEL_COMMON, SEL_W, SEL_N, SEL_D = "common", "w_ceil", "n_floor", "d"
LEADING_TABLE_BITS = {SEL_W: Helper.w_ceil_bits % 17 or 17, SEL_N: Helper.n_floor_bits % 17 or 17, SEL_D: Helper.d_bits % 17 or 17}


class AssignedInteger:
    def __init__(self, limbs_le, overflows):
        self.limbs_le, self.native, self.overflows = list(limbs_le), None, overflows

    def bn(self):
        v = 0
        for c in reversed(self.limbs_le):
            v = v * Helper.limb_modulus + c.value
        return v

    def w(self):
        return self.bn() % P

    def clone(self):
        return copy.copy(self) if False else _clone_int(self)


def _clone_int(a):
    b = AssignedInteger(a.limbs_le, a.overflows)
    b.native = a.native
    return b


def _decompose(bn, chunks):  # utils.rs:21-37, returned most significant first (":334 .rev()")
    return [(((bn >> (i * 17)) & 0x1ffff), (1 << (i * 17)) % R) for i in range(chunks)][::-1]


def assign_nonleading_limb(ctx, n):  # five/integer_chip.rs:324-341
    schema = _decompose(n, 4) + [(n, -1)]
    return one_line(ctx, schema, 0, sel=SEL_COMMON)[4]


def _assign_leading(ctx, n, bits, sel):
    leading = bits % LIMB_WIDTH
    if leading == 0:
        return assign_nonleading_limb(ctx, n)
    nchunks = (leading + 16) // 17
    schema = _decompose(n, nchunks)
    schema += [(0, 0)] * (4 - len(schema)) + [(n, -1)]
    return one_line(ctx, schema, 0, sel=sel)[4]


def assign_w_ceil_leading_limb(ctx, n):  # :377-403
    return _assign_leading(ctx, n, Helper.w_ceil_bits, SEL_W)


def assign_n_floor_leading_limb(ctx, n):  # :343-375
    return _assign_leading(ctx, n, Helper.n_floor_bits, SEL_N)


def assign_d_leading_limb(ctx, n):  # :405-426
    return _assign_leading(ctx, n, Helper.d_bits, SEL_D)


def assign_d(ctx, v):  # :428-445
    limbs = Helper.bn_to_limb_le(v)
    cells = [assign_d_leading_limb(ctx, l) if i == 0 else assign_nonleading_limb(ctx, l) for i, l in enumerate(limbs[::-1])]
    return cells[::-1]


def assign_w(ctx, w):  # :447-464
    limbs = Helper.bn_to_limb_le(w % P)
    cells = [assign_w_ceil_leading_limb(ctx, l) if i == 0 else assign_nonleading_limb(ctx, l) for i, l in enumerate(limbs[::-1])]
    return AssignedInteger(cells[::-1], 0)


def int_assign_constant(ctx, w):  # :784-794
    return AssignedInteger([bg_assign_constant(ctx, l) for l in Helper.bn_to_limb_le(w % P)], 0)


def native(ctx, a):  # :595-621
    if a.native is None:
        a.native = sum_with_constant(ctx, list(zip(a.limbs_le, Helper.limb_modulus_exps)), 0)
    return a.native


def find_w_modulus_ceil(a):  # :31-51
    max_a = (a.overflows + 1) << Helper.w_ceil_bits
    n, rem = divmod(max_a, P)
    if rem > 0:
        n += 1
    upper = n * P
    limbs = []
    for _ in range(LIMBS - 1):
        rem = upper % Helper.limb_modulus + (a.overflows + 1) * Helper.limb_modulus
        upper = (upper - rem) // Helper.limb_modulus
        limbs.append(rem)
    limbs.append(upper)
    return limbs


def reduce(ctx, a):  # :483-581 (in place)
    if a.overflows == 0:
        return
    assert a.overflows < OVERFLOW_LIMIT
    a_bn = a.bn()
    d, rem = divmod(a_bn, P)
    u = d * Helper.w_modulus_limbs_le[0] + Helper.bn_to_limb_le(rem)[0] + Helper.limb_modulus * OVERFLOW_LIMIT - a.limbs_le[0].value
    v = u // Helper.limb_modulus
    rem_i = assign_w(ctx, rem)
    cells = one_line(ctx, [(d % R, 0), (v % R, 0)], 0, sel=SEL_COMMON)
    d_c, v_c = cells[0], cells[1]
    rem_native = native(ctx, rem_i)
    a_native = native(ctx, a)
    one_line(ctx, [(a_native, -1), (d_c, Helper.w_native), (rem_native, 1)], 0)
    one_line(ctx, [(d_c, Helper.w_modulus_limbs_le[0] % R), (rem_i.limbs_le[0], 1), (a.limbs_le[0], -1), (v_c, -(Helper.limb_modulus % R))],
             (Helper.limb_modulus * OVERFLOW_LIMIT) % R)
    a.limbs_le, a.overflows, a.native = rem_i.limbs_le, rem_i.overflows, rem_i.native


def conditionally_reduce(ctx, a):  # :583-593
    if a.overflows >= OVERFLOW_THRESHOLD:
        reduce(ctx, a)


def int_add(ctx, a, b):  # :641-658
    res = AssignedInteger([bg_add(ctx, a.limbs_le[i], b.limbs_le[i]) for i in range(LIMBS)], a.overflows + b.overflows + 1)
    conditionally_reduce(ctx, res)
    return res


def int_sub(ctx, a, b):  # :660-683
    upper = find_w_modulus_ceil(b)
    limbs = [sum_with_constant(ctx, [(a.limbs_le[i], 1), (b.limbs_le[i], -1)], upper[i] % R) for i in range(LIMBS)]
    res = AssignedInteger(limbs, a.overflows + (b.overflows + 1) + 1)
    conditionally_reduce(ctx, res)
    return res


def int_neg(ctx, a):  # :685-707
    upper = find_w_modulus_ceil(a)
    limbs = [sum_with_constant(ctx, [(a.limbs_le[i], -1)], upper[i] % R) for i in range(LIMBS)]
    res = AssignedInteger(limbs, a.overflows + 1)
    conditionally_reduce(ctx, res)
    return res


def _mul_equation_on_limb0(ctx, a, b, d, rem):  # :104-252
    assert a.overflows < OVERFLOW_LIMIT and b.overflows < OVERFLOW_LIMIT and rem.overflows < OVERFLOW_LIMIT
    neg_w = [x % R for x in Helper.bn_to_limb_le(Helper.integer_modulus - P)]
    limbs = []
    for pos in range(LIMBS):
        limbs.append(bg_mul_add_with_next_line(ctx, [(a.limbs_le[i], b.limbs_le[pos - i], d[i], neg_w[pos - i]) for i in range(pos + 1)]))
    lm, e = Helper.limb_modulus_on_n, Helper.limb_modulus_exps
    inv_e2 = pow(e[2], -1, R)
    u0 = ((limbs[1].value - rem.limbs_le[1].value) * lm + limbs[0].value - rem.limbs_le[0].value + e[2]) % R
    v0 = u0 * inv_e2 % R
    v0_h, v0_l = divmod(v0, Helper.limb_modulus)
    u1 = (v0 - 1 + limbs[2].value - rem.limbs_le[2].value + (limbs[3].value - rem.limbs_le[3].value) * lm) % R
    v1 = u1 * inv_e2 % R
    v1_h, v1_l = divmod(v1, Helper.limb_modulus)
    v0_h = assign_n_floor_leading_limb(ctx, v0_h % R)
    v0_l = assign_nonleading_limb(ctx, v0_l)
    v1_h = assign_n_floor_leading_limb(ctx, v1_h % R)
    v1_l = assign_nonleading_limb(ctx, v1_l)
    u0_c = sum_with_constant(ctx, [(limbs[0], 1), (limbs[1], lm), (rem.limbs_le[0], -1), (rem.limbs_le[1], -lm)], e[2])
    one_line(ctx, [(u0_c, -1), (v0_l, e[2]), (v0_h, e[3])], 0)
    u1_c = sum_with_constant(ctx, [(limbs[2], 1), (limbs[3], lm), (rem.limbs_le[2], -1), (rem.limbs_le[3], -lm)], 0)
    one_line(ctx, [(u1_c, 1), (v0_l, e[0]), (v0_h, e[1]), (v1_l, -e[2]), (v1_h, -e[3])], -1)


def _mul_equation_on_native(ctx, a, b, d, rem):  # :254-286
    a_n = native(ctx, a)
    b_n = native(ctx, b)
    d_n = sum_with_constant(ctx, list(zip(d, Helper.limb_modulus_exps)), 0)
    rem_n = native(ctx, rem)
    one_line(ctx, [(a_n, 0), (b_n, 0), (d_n, -Helper.w_native), (rem_n, -1)], 0, ((1,), 0))


def _square_equation_on_native(ctx, a, d, rem):  # :288-320
    a_n = native(ctx, a)
    d_n = sum_with_constant(ctx, list(zip(d, Helper.limb_modulus_exps)), 0)
    rem_n = native(ctx, rem)
    one_line(ctx, [(a_n, 0), (a_n, 0), (d_n, -Helper.w_native), (rem_n, -1)], 0, ((1,), 0))


def int_mul(ctx, a, b):  # :709-726
    d, rem = divmod(a.bn() * b.bn(), P)
    rem_i = assign_w(ctx, rem)
    d_c = assign_d(ctx, d)
    _mul_equation_on_limb0(ctx, a, b, d_c, rem_i)
    _mul_equation_on_native(ctx, a, b, d_c, rem_i)
    return rem_i


def int_square(ctx, a):  # :728-743
    d, rem = divmod(a.bn() * a.bn(), P)
    rem_i = assign_w(ctx, rem)
    d_c = assign_d(ctx, d)
    _mul_equation_on_limb0(ctx, a, a, d_c, rem_i)
    _square_equation_on_native(ctx, a, d_c, rem_i)
    return rem_i


def is_pure_zero(ctx, a):  # :53-66
    s = sum_with_constant(ctx, [(v, 1) for v in a.limbs_le], 0)
    return bg_is_zero(ctx, s)


def is_pure_w_modulus(ctx, a):  # :68-102
    native_a = native(ctx, a)
    native_diff = sum_with_constant(ctx, [(native_a, 1)], -Helper.w_native)
    is_native_eq = bg_is_zero(ctx, native_diff)
    limb0_diff = sum_with_constant(ctx, [(a.limbs_le[0], 1)], -(Helper.w_modulus_limbs_le[0] % R))
    is_limb0_eq = bg_is_zero(ctx, limb0_diff)
    return bg_and(ctx, is_native_eq, is_limb0_eq)


def int_is_zero(ctx, a):  # :796-806
    reduce(ctx, a)
    z = is_pure_zero(ctx, a)
    w = is_pure_w_modulus(ctx, a)
    return bg_or(ctx, z, w)


def int_is_equal(ctx, a, b):  # chips/integer_chip.rs:198-206
    diff = int_sub(ctx, a, b)
    return int_is_zero(ctx, diff)


def int_div(ctx, a, b):  # :745-782 -> (is_b_zero, c)
    is_b_zero = int_is_zero(ctx, b)
    a_coeff = bg_not(ctx, is_b_zero)
    reduce(ctx, a)
    a2 = AssignedInteger([bg_mul(ctx, a.limbs_le[i], a_coeff) for i in range(LIMBS)], a.overflows)
    a_bn, b_bn = a2.bn(), b.bn()
    a_w, b_w = a2.w(), b.w()
    c = (pow(b_w, -1, P) if b_w else 0) * a_w % P
    d = (c * b_bn - a_bn) // P
    c_i = assign_w(ctx, c)
    d_c = assign_d(ctx, d)
    _mul_equation_on_limb0(ctx, b, c_i, d_c, a2)
    _mul_equation_on_native(ctx, b, c_i, d_c, a2)
    return is_b_zero, c_i


def int_mul_small_constant(ctx, a, b):  # :808-835
    assert b < OVERFLOW_LIMIT
    if a.overflows * b >= OVERFLOW_LIMIT:
        reduce(ctx, a)
    res = AssignedInteger([sum_with_constant(ctx, [(a.limbs_le[i], b)], 0) for i in range(LIMBS)], a.overflows * b)
    conditionally_reduce(ctx, res)
    return res


def int_bisec(ctx, cond, a, b):  # :845-866
    return AssignedInteger([bg_bisec(ctx, cond, a.limbs_le[i], b.limbs_le[i]) for i in range(LIMBS)], max(a.overflows, b.overflows))


def int_assert_equal(ctx, a, b):  # :623-639
    diff = int_sub(ctx, a, b)
    reduce(ctx, diff)
    bg_assert_constant(ctx, native(ctx, diff), 0)
    bg_assert_constant(ctx, diff.limbs_le[0], 0)


def int_get_last_bit(ctx, a):  # :874-901
    l0 = a.limbs_le[0].value
    d = assign_nonleading_limb(ctx, l0 // 2)
    cells = one_line(ctx, [(d, 2), (l0 & 1, 1), (a.limbs_le[0], -1)], 0)
    bg_assert_bit(ctx, cells[1])
    return cells[1]


# ------------------------------------------------------------------------------------ ecc chip
class AssignedCurvature:
    def __init__(self, v, z):
        self.v, self.z = v, z

    def clone(self):
        return AssignedCurvature(_clone_int(self.v), self.z)


class AssignedPoint:
    def __init__(self, x, y, z, curvature=None):
        self.x, self.y, self.z, self.curvature = x, y, z, curvature

    def clone(self):
        return AssignedPoint(_clone_int(self.x), _clone_int(self.y), self.z, self.curvature.clone() if self.curvature else None)


def ecc_curvature(ctx, a):  # chips/ecc_chip.rs:280-307
    if a.curvature is None:
        x_square = int_square(ctx, a.x)
        numerator = int_mul_small_constant(ctx, x_square, 3)
        denominator = int_mul_small_constant(ctx, a.y, 2)
        z, v = int_div(ctx, numerator, denominator)
        a.curvature = AssignedCurvature(v, z)
    return a.curvature


def bisec_curvature(ctx, cond, a, b):  # :308-321
    return AssignedCurvature(int_bisec(ctx, cond, a.v, b.v), bg_bisec(ctx, cond, a.z, b.z))


def bisec_point(ctx, cond, a, b):  # :322-336
    return AssignedPoint(int_bisec(ctx, cond, a.x, b.x), int_bisec(ctx, cond, a.y, b.y), bg_bisec(ctx, cond, a.z, b.z))


def bisec_point_with_curvature(ctx, cond, a, b):  # :337-355
    x = int_bisec(ctx, cond, a.x, b.x)
    y = int_bisec(ctx, cond, a.y, b.y)
    z = bg_bisec(ctx, cond, a.z, b.z)
    c_a = ecc_curvature(ctx, a)
    c_b = ecc_curvature(ctx, b)
    return AssignedPoint(x, y, z, bisec_curvature(ctx, cond, c_a, c_b))


def lambda_to_point(ctx, lam, a, b):  # :356-382
    l = lam.v
    l_square = int_square(ctx, l)
    t = int_sub(ctx, l_square, a.x)
    cx = int_sub(ctx, t, b.x)
    t = int_sub(ctx, a.x, cx)
    t = int_mul(ctx, t, l)
    cy = int_sub(ctx, t, a.y)
    return AssignedPoint(cx, cy, lam.z)


def ecc_add(ctx, a, b):  # :383-408
    diff_x = int_sub(ctx, a.x, b.x)
    diff_y = int_sub(ctx, a.y, b.y)
    x_eq, tangent = int_div(ctx, diff_y, diff_x)
    y_eq = int_is_zero(ctx, diff_y)
    eq = bg_and(ctx, x_eq, y_eq)
    tangent = AssignedCurvature(tangent, x_eq)
    curv = ecc_curvature(ctx, a)
    lam = bisec_curvature(ctx, eq, curv, tangent)
    p = lambda_to_point(ctx, lam, a, b)
    p = bisec_point(ctx, a.z, b, p)
    p = bisec_point(ctx, b.z, a, p)
    return p


def ecc_double(ctx, a):  # :409-419
    curv = ecc_curvature(ctx, a)
    p = lambda_to_point(ctx, curv.clone(), a, a)
    p.z = bg_bisec(ctx, a.z, a.z, p.z)
    return p


def _affine(pt):
    return (0, 0, 1) if pt is None else (pt[0], pt[1], 0)


def assign_constant_point(ctx, pt):  # :420-437   pt = None | (x, y)
    x, y, z = _affine(pt)
    return AssignedPoint(int_assign_constant(ctx, x), int_assign_constant(ctx, y), bg_assign_constant(ctx, z))


def assign_constant_point_with_curvature(ctx, pt):  # :438-472 (note: "curvature" = y/x, App. E1)
    x, y, z = _affine(pt)
    cv = int_assign_constant(ctx, y * (pow(x, -1, P) if x else 0) % P)
    cz = bg_assign_constant(ctx, 1 if x == 0 else 0)
    xi = int_assign_constant(ctx, x)
    yi = int_assign_constant(ctx, y)
    zi = bg_assign_constant(ctx, z)
    return AssignedPoint(xi, yi, zi, AssignedCurvature(cv, cz))


def assign_point(ctx, pt):  # :473-500
    x, y, z = _affine(pt)
    xi = assign_w(ctx, x)
    yi = assign_w(ctx, y)
    zi = bg_assign(ctx, z)
    b = int_assign_constant(ctx, 3)
    y2 = int_square(ctx, yi)
    x2 = int_square(ctx, xi)
    x3 = int_mul(ctx, x2, xi)
    right = int_add(ctx, x3, b)
    eq = int_is_equal(ctx, y2, right)
    eq_or_identity = bg_or(ctx, eq, zi)
    bg_assert_constant(ctx, eq_or_identity, 1)
    return AssignedPoint(xi, yi, zi)


def assign_identity(ctx):  # :517-527
    zero = int_assign_constant(ctx, 0)
    one = bg_assign_constant(ctx, 1)
    return AssignedPoint(_clone_int(zero), _clone_int(zero), one, AssignedCurvature(zero, one))


def ecc_neg(ctx, a):  # :549-559
    return AssignedPoint(_clone_int(a.x), int_neg(ctx, a.y), a.z)


def ecc_sub(ctx, a, b):  # :560-568
    return ecc_add(ctx, a, ecc_neg(ctx, b))


def ecc_reduce(ctx, a):  # :569-580
    reduce(ctx, a.x)
    reduce(ctx, a.y)
    identity = assign_identity(ctx)
    return bisec_point(ctx, a.z, identity, a)


def ecc_assert_equal(ctx, a, b):  # :528-548
    eq_x = int_is_equal(ctx, a.x, b.x)
    eq_y = int_is_equal(ctx, a.y, b.y)
    eq_z = bg_xnor(ctx, eq_x, eq_y)
    eq_xy = bg_and(ctx, eq_x, eq_y)
    eq_xyz = bg_and(ctx, eq_xy, eq_z)
    both = bg_and(ctx, a.z, b.z)
    eq = bg_or(ctx, eq_xyz, both)
    bg_assert_constant(ctx, eq, 1)


def decompose_scalar(ctx, s, window):  # chips/native_ecc_chip.rs:42-132; s: Cell; windows big-endian
    num_bits = 254
    windows = (num_bits - 1 + window) // window
    ret = []
    s_bn = s.value

    def bits_of(v):
        return [(v >> i) & 1 for i in range(window)], v >> window

    bits, s_bn = bits_of(s_bn)
    cells = one_line_with_last_base(ctx, [(b, 1 << i) for i, b in enumerate(bits)], (s, -1), 0, ((), 1 << window))
    ret.append(cells[0:window])
    for _ in range(1, windows - 1):
        s_n = s_bn % R
        bits, nxt = bits_of(s_bn)
        cells = one_line_with_last_base(ctx, [(b, 1 << i) for i, b in enumerate(bits)], (s_n, -1), 0, ((), 1 << window))
        ret.append(cells[0:window])
        s_bn = nxt
    s_n = s_bn % R
    bits, _ = bits_of(s_bn)
    cells = one_line_with_last_base(ctx, [(b, 1 << i) for i, b in enumerate(bits)], (s_n, -1), 0, ((), 0))
    ret.append(cells[0:window])
    ret.reverse()
    for w in ret:
        for bit in w:
            bg_assert_bit(ctx, bit)
    return ret


def _pick_candidate(ctx, candidates, bits_in_le):  # chips/ecc_chip.rs:100-120 / 165-186
    curr = [c.clone() for c in candidates]
    for bit in bits_in_le:
        nxt = []
        for k in range(len(curr) // 2):
            a0, a1 = curr[2 * k], curr[2 * k + 1]
            nxt.append(bisec_point_with_curvature(ctx, bit, a1, a0))
        curr = nxt
    return curr[0].clone()


def ecc_mul(ctx, a, s):  # :86-138
    windows_in_be = decompose_scalar(ctx, s, CONFIG_WINDOW_SIZE)
    identity = assign_identity(ctx)
    candidates = [identity, a.clone()]
    for i in range(2, 1 << CONFIG_WINDOW_SIZE):
        candidates.append(ecc_add(ctx, candidates[i - 1], a))
    acc = _pick_candidate(ctx, candidates, windows_in_be[0])
    for bits in windows_in_be[1:]:
        for _ in range(CONFIG_WINDOW_SIZE):
            acc = ecc_double(ctx, acc)
        curr = _pick_candidate(ctx, candidates, bits)
        acc = ecc_add(ctx, curr, acc)
    return acc


def ecc_shamir(ctx, points, scalars):  # :139-244 (assignment pass: shape-mode shortcut not taken)
    assert len(points) == len(scalars)
    windows_in_be = [decompose_scalar(ctx, s, CONFIG_WINDOW_SIZE) for s in scalars]
    identity = assign_identity(ctx)
    point_candidates = []
    for a in points:
        cands = [identity.clone(), a.clone()]
        for i in range(2, 1 << CONFIG_WINDOW_SIZE):
            ai = ecc_add(ctx, cands[i - 1], a)
            ecc_curvature(ctx, ai)
            cands.append(ai)
        point_candidates.append(cands)
    acc = None
    for wi in range(len(windows_in_be[0])):
        inner = None
        for pi in range(len(points)):
            ci = _pick_candidate(ctx, point_candidates[pi], windows_in_be[pi][wi])
            inner = ci if inner is None else ecc_add(ctx, ci, inner)
        if acc is None:
            acc = inner
        else:
            for _ in range(CONFIG_WINDOW_SIZE):
                acc = ecc_double(ctx, acc)
            acc = ecc_add(ctx, inner, acc)
    return acc


def ecc_constant_mul(ctx, base, s, g1_add):  # :245-279; base: affine tuple; g1_add: native group law
    bits_be = decompose_scalar(ctx, s, 2)
    identity = assign_constant_point_with_curvature(ctx, None)
    acc = None
    for bit_le in reversed(bits_be):
        b2 = g1_add(base, base)
        c01 = assign_constant_point_with_curvature(ctx, b2)
        c10 = assign_constant_point_with_curvature(ctx, base)
        c11 = assign_constant_point_with_curvature(ctx, g1_add(b2, base))
        c0 = bisec_point_with_curvature(ctx, bit_le[0], c10, identity)
        c1 = bisec_point_with_curvature(ctx, bit_le[0], c11, c01)
        slot = bisec_point_with_curvature(ctx, bit_le[1], c1, c0)
        acc = slot if acc is None else ecc_add(ctx, slot, acc)
        base = g1_add(g1_add(b2, base), base)
    return acc


def expose_final_pair(ctx, w_x, w_g):
    """Halo2VerifierCircuits::synthesize, second region (halo2-snark-aggregator-circuit/src/verify_circuit.rs:264-344):
    reduce the coordinates, take the parity bits of the y's, pack each point into two 136-bit halves -> the four cells
    constrain_instance binds to instance rows 0..3 (:357-360)."""
    pts = [w_x.clone(), w_g.clone()]
    for p in pts:
        reduce(ctx, p.x)
        reduce(ctx, p.y)
    bits = [int_get_last_bit(ctx, p.y) for p in pts]
    e = Helper.limb_modulus_exps
    out = []
    for p, bit in zip(pts, bits):
        out.append(sum_with_constant(ctx, [(p.x.limbs_le[0], e[0]), (p.x.limbs_le[1], e[1])], 0))
        out.append(sum_with_constant(ctx, [(p.x.limbs_le[2], e[0]), (p.x.limbs_le[3], e[1]), (bit, e[2])], 0))
    return out


# ------------------------------------------------------------------------------------ MockProver
def check(ctx):
    """Gate polynomial (gates/base_gate.rs:701-720), range lookups (five/range_gate.rs:45-81 with the
    table sizes of gates/range_gate.rs:198-294) and copy constraints. Returns the number of rows."""
    rows = ctx.rows
    for r, row in enumerate(rows):
        a = row["adv"]
        nxt_a4 = rows[r + 1]["adv"][4] if r + 1 < len(rows) else 0
        acc = row["const"] + nxt_a4 * row["nxt"]
        for i in range(5):
            acc += a[i] * row["coeff"][i]
        acc += a[0] * a[1] * row["mul"][0] + a[2] * a[3] * row["mul"][1]
        assert acc % R == 0, "gate violated at row %d: %r" % (r, row)
        sel = row["sel"]
        if sel is not None:
            for i in range(4):
                assert a[i] < (1 << 17), "common range violated at row %d col %d" % (r, i)
            if sel != SEL_COMMON:
                assert a[0] < (1 << LEADING_TABLE_BITS[sel]), "%s leading range violated at row %d" % (sel, r)
    for (c0, r0), (c1, r1) in ctx.copies:
        assert rows[r0]["adv"][c0] == rows[r1]["adv"][c1], "copy constraint violated (%d,%d)-(%d,%d)" % (c0, r0, c1, r1)
    return len(rows)


def advice_columns(ctx):
    """5 lists of Fr values (canonical ints), the only thing create_proof keeps."""
    return [[row["adv"][c] for row in ctx.rows] for c in range(5)]
