#!/usr/bin/env python3
"""Generate the inline-PTX Montgomery multiplication bodies for BN254 Fr / Fq.

The product is computed over 8 x 32-bit limbs with the "even/odd column" schedule:
products a[j]*b_i with even j accumulate into an `even` limb array, odd j into an
`odd` array that is offset by one limb, so every (lo,hi) pair of a 32x32 product lands
on an aligned register pair and ptxas fuses `mad.lo.cc` + `madc.hi.cc` into a single
IMAD.WIDE.U32(.X).  That is 16 wide MADs per row, 128 per multiplication, plus 8
`mul.lo` for the Montgomery quotient digits -- the minimum for 256-bit CIOS.

The modulus limbs and -p^-1 mod 2^32 are emitted as immediates (they are compile-time
constants of the curve), so a multiplication needs only its 16 input registers.

Before writing the .inc files the script *executes* the generated PTX with a tiny
interpreter (Python big ints, explicit carry flag) on random and edge-case operands and
checks  r == a*b*R^-1 (mod p), r < 2p  -- so a typo in the schedule can never reach
the GPU.  Run:  python gen_mont_ptx.py   (writes ../gen/mont_mul_bn254.inc)
"""
import os
import random
import re
import sys

FR = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
FQ = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
N = 8
M32 = 0xFFFFFFFF


class Gen:
    def __init__(self, P):
        self.P = P
        self.p = [(P >> (32 * i)) & M32 for i in range(N)]
        self.inv = (-pow(P, -1, 1 << 32)) & M32
        self.lines = []
        self.ntemps = 0

    def new(self):
        self.ntemps += 1
        return "t%d" % (self.ntemps - 1)

    def emit(self, s):
        self.lines.append(s)

    @staticmethod
    def imm(v):
        return "0x%08x" % v

    # acc[0..7] (+)= x[0,2,4,6] * y with one carry chain through the four aligned pairs
    def cmad(self, acc, xs, y):
        """xs may start with None entries (limbs known to be zero): their products are skipped and the carry chain
        starts at the first present limb.  Returns False if nothing was emitted."""
        started = False
        for k in range(4):
            if xs[k] is None:
                assert not started, "zero limbs must be leading"
                continue
            lo = "madc.lo.cc.u32" if started else "mad.lo.cc.u32"
            started = True
            self.emit("%s %s, %s, %s, %s;" % (lo, acc[2 * k], y, xs[k], acc[2 * k]))
            self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (acc[2 * k + 1], y, xs[k], acc[2 * k + 1]))
        return started

    def redc(self, even, odd):
        mi = self.new()
        self.emit("mul.lo.u32 %s, %s, %s;" % (mi, even[0], self.imm(self.inv)))
        self.cmad(odd, [self.imm(self.p[j]) for j in (1, 3, 5, 7)], mi)
        self.cmad(even, [self.imm(self.p[j]) for j in (0, 2, 4, 6)], mi)
        self.emit("addc.u32 %s, %s, 0;" % (odd[7], odd[7]))

    def row(self, even, odd, A, bi, first, C=None, di=None):
        """One CIOS row; returns the (even, odd) register-name lists after the row.
        With (C, di) the row also accumulates C * di before the reduction step (dual product a*b + c*d: the top
        pair holds A7*b + C7*d + m*p7 + carries < 3 * 2^62, so the odd chain still never carries out)."""
        if first:
            for k in range(4):
                odd[2 * k], odd[2 * k + 1] = self.new(), self.new()
                self.emit("mul.lo.u32 %s, %s, %s;" % (odd[2 * k], A[2 * k + 1], bi))
                self.emit("mul.hi.u32 %s, %s, %s;" % (odd[2 * k + 1], A[2 * k + 1], bi))
            for k in range(4):
                even[2 * k], even[2 * k + 1] = self.new(), self.new()
                self.emit("mul.lo.u32 %s, %s, %s;" % (even[2 * k], A[2 * k], bi))
                self.emit("mul.hi.u32 %s, %s, %s;" % (even[2 * k + 1], A[2 * k], bi))
        else:
            # the limb that fell to position 0 after the previous row's /2^32
            self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
            nodd = [None] * 8
            for k in range(3):  # shift the odd array down two limbs while accumulating
                nodd[2 * k], nodd[2 * k + 1] = self.new(), self.new()
                if A[2 * k + 1] is None:  # zero limb (squaring rows): the shift still carries
                    self.emit("addc.cc.u32 %s, %s, 0;" % (nodd[2 * k], odd[2 * k + 2]))
                    self.emit("addc.cc.u32 %s, %s, 0;" % (nodd[2 * k + 1], odd[2 * k + 3]))
                else:
                    self.emit("madc.lo.cc.u32 %s, %s, %s, %s;" % (nodd[2 * k], A[2 * k + 1], bi, odd[2 * k + 2]))
                    self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (nodd[2 * k + 1], A[2 * k + 1], bi, odd[2 * k + 3]))
            nodd[6], nodd[7] = self.new(), self.new()
            assert A[7] is not None
            self.emit("madc.lo.cc.u32 %s, %s, %s, 0;" % (nodd[6], A[7], bi))
            self.emit("madc.hi.u32 %s, %s, %s, 0;" % (nodd[7], A[7], bi))
            odd[:] = nodd
            if self.cmad(even, [A[j] for j in (0, 2, 4, 6)], bi):
                self.emit("addc.u32 %s, %s, 0;" % (odd[7], odd[7]))
        if C is not None:
            self.cmad(odd, [C[j] for j in (1, 3, 5, 7)], di)
            self.cmad(even, [C[j] for j in (0, 2, 4, 6)], di)
            self.emit("addc.u32 %s, %s, 0;" % (odd[7], odd[7]))
        self.redc(even, odd)

    def mont_mul(self):
        A = ["%%%d" % (8 + i) for i in range(8)]
        B = ["%%%d" % (16 + i) for i in range(8)]
        even, odd = [None] * 8, [None] * 8
        for i in range(0, 8, 2):
            self.row(even, odd, A, B[i], i == 0)
            self.row(odd, even, A, B[i + 1], False)
        # merge: result[j] = even[j] + odd[j+1]
        self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
        for j in range(1, 7):
            self.emit("addc.cc.u32 %s, %s, %s;" % (even[j], even[j], odd[j + 1]))
        self.emit("addc.u32 %s, %s, 0;" % (even[7], even[7]))
        for j in range(8):
            self.emit("mov.u32 %%%d, %s;" % (j, even[j]))
        return self.lines


    def mont_sqr(self):
        """r = a*a * 2^-256 mod p, r in [0, 2p), with 36 + 72 multiplier instructions instead of 136:
        a^2 = sum_i a_i 2^(32 i) * V_i,  V_i = a_i 2^(32 i) + 2 * sum_{j>i} a_j 2^(32 j), i.e. CIOS row i multiplies by
        a vector with i leading zero limbs (products skipped); the limbs of V_i above i are those of 2a, the one at
        i + 1 with its lowest bit (the top bit of a_i, which belongs to 2 a_i) cleared.  a < 2^254, so 2a has 8 limbs."""
        A = ["%%%d" % (8 + i) for i in range(8)]
        a2 = [self.new() for _ in range(8)]
        self.emit("add.cc.u32 %s, %s, %s;" % (a2[0], A[0], A[0]))
        for j in range(1, 7):
            self.emit("addc.cc.u32 %s, %s, %s;" % (a2[j], A[j], A[j]))
        self.emit("addc.u32 %s, %s, %s;" % (a2[7], A[7], A[7]))
        even, odd = [None] * 8, [None] * 8
        for i in range(8):
            V = [None] * 8
            V[i] = A[i]
            if i + 1 < 8:
                m = self.new()
                self.emit("and.b32 %s, %s, 0xfffffffe;" % (m, a2[i + 1]))
                V[i + 1] = m
            for j in range(i + 2, 8):
                V[j] = a2[j]
            if i % 2 == 0:
                self.row(even, odd, V, A[i], i == 0)
            else:
                self.row(odd, even, V, A[i], False)
        self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
        for j in range(1, 7):
            self.emit("addc.cc.u32 %s, %s, %s;" % (even[j], even[j], odd[j + 1]))
        self.emit("addc.u32 %s, %s, 0;" % (even[7], even[7]))
        for j in range(8):
            self.emit("mov.u32 %%%d, %s;" % (j, even[j]))
        return self.lines

    def mont_mul_add2(self):
        """r = (a*b + c*d) * 2^-256 mod p, r in [0, 2p): one reduction for two products (200 wide MADs + 8 mul.lo
        instead of 2 x 136).  Operands %8..%15 a, %16..%23 b, %24..%31 c, %32..%39 d."""
        A = ["%%%d" % (8 + i) for i in range(8)]
        B = ["%%%d" % (16 + i) for i in range(8)]
        C = ["%%%d" % (24 + i) for i in range(8)]
        D = ["%%%d" % (32 + i) for i in range(8)]
        even, odd = [None] * 8, [None] * 8
        for i in range(0, 8, 2):
            self.row(even, odd, A, B[i], i == 0, C, D[i])
            self.row(odd, even, A, B[i + 1], False, C, D[i + 1])
        self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
        for j in range(1, 7):
            self.emit("addc.cc.u32 %s, %s, %s;" % (even[j], even[j], odd[j + 1]))
        self.emit("addc.u32 %s, %s, 0;" % (even[7], even[7]))
        for j in range(8):
            self.emit("mov.u32 %%%d, %s;" % (j, even[j]))
        return self.lines



    # ------------------------------------------------------------------------------------------------ Karatsuba variant
    # a*b over two 128-bit halves: three 4x4-limb products (48 wide MADs instead of 64) + additions, then the Montgomery
    # reduction of the low half row by row (64 wide MADs + 8 mul.lo) and one addition of the high half:
    # 112 + 8 multiplier instructions instead of 128 + 8.  The extra work is additions, which issue on the ALU pipe while
    # the multiplier (the bound of every kernel here) is busy.
    def prod4(self, X, Y):
        """X, Y: 4 register names each (128-bit values) -> 8 fresh registers holding X * Y.  Row-wise product scanning with
        the even/odd column arrays of `row` (window of 5 limbs: X * y + carry-in < 2^160), one finished limb per row."""
        out = []
        E = [self.new() for _ in range(4)]      # positions 0..3
        O = [self.new() for _ in range(4)]      # positions 1..4
        for k in range(2):
            self.emit("mul.lo.u32 %s, %s, %s;" % (O[2 * k], X[2 * k + 1], Y[0]))
            self.emit("mul.hi.u32 %s, %s, %s;" % (O[2 * k + 1], X[2 * k + 1], Y[0]))
        for k in range(2):
            self.emit("mul.lo.u32 %s, %s, %s;" % (E[2 * k], X[2 * k], Y[0]))
            self.emit("mul.hi.u32 %s, %s, %s;" % (E[2 * k + 1], X[2 * k], Y[0]))
        out.append(E[0])
        for i in range(1, 4):
            E, O = O, E                           # the array at positions 1.. becomes the one at 0..
            self.emit("add.cc.u32 %s, %s, %s;" % (E[0], E[0], O[1]))
            nO = [self.new() for _ in range(4)]
            self.emit("madc.lo.cc.u32 %s, %s, %s, %s;" % (nO[0], X[1], Y[i], O[2]))
            self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (nO[1], X[1], Y[i], O[3]))
            self.emit("madc.lo.cc.u32 %s, %s, %s, 0;" % (nO[2], X[3], Y[i]))
            self.emit("madc.hi.u32 %s, %s, %s, 0;" % (nO[3], X[3], Y[i]))
            O = nO
            self.emit("mad.lo.cc.u32 %s, %s, %s, %s;" % (E[0], X[0], Y[i], E[0]))
            self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (E[1], X[0], Y[i], E[1]))
            self.emit("madc.lo.cc.u32 %s, %s, %s, %s;" % (E[2], X[2], Y[i], E[2]))
            self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (E[3], X[2], Y[i], E[3]))
            self.emit("addc.u32 %s, %s, 0;" % (O[3], O[3]))
            out.append(E[0])
        # remaining limbs 4..7 = E[1..3] + O[0..3] (O one position up)
        hi = [self.new() for _ in range(4)]
        self.emit("add.cc.u32 %s, %s, %s;" % (hi[0], E[1], O[0]))
        self.emit("addc.cc.u32 %s, %s, %s;" % (hi[1], E[2], O[1]))
        self.emit("addc.cc.u32 %s, %s, %s;" % (hi[2], E[3], O[2]))
        self.emit("addc.u32 %s, %s, 0;" % (hi[3], O[3]))
        return out + hi

    def add_n(self, dst, x, y, carry_out=None):
        """dst = x + y over len(x) limbs (y may be shorter: zero extended); carry_out: register that receives the carry"""
        n = len(x)
        for j in range(n):
            op = "add.cc.u32" if j == 0 else ("addc.cc.u32" if (j < n - 1 or carry_out) else "addc.u32")
            self.emit("%s %s, %s, %s;" % (op, dst[j], x[j], y[j] if j < len(y) else "0"))
        if carry_out:
            self.emit("addc.u32 %s, 0, 0;" % carry_out)

    def sub_n(self, dst, x, y):
        """dst = x - y over len(x) limbs (no final borrow by construction)"""
        n = len(x)
        for j in range(n):
            op = "sub.cc.u32" if j == 0 else ("subc.cc.u32" if j < n - 1 else "subc.u32")
            self.emit("%s %s, %s, %s;" % (op, dst[j], x[j], y[j] if j < len(y) else "0"))

    def product_karatsuba(self, A, B):
        """-> 16 registers holding A * B (A, B: 8 register names)"""
        A0, A1, B0, B1 = A[:4], A[4:], B[:4], B[4:]
        z0 = self.prod4(A0, B0)
        z2 = self.prod4(A1, B1)
        sa, sb = [self.new() for _ in range(4)], [self.new() for _ in range(4)]
        ca, cb = self.new(), self.new()
        self.add_n(sa, A0, A1, ca)
        self.add_n(sb, B0, B1, cb)
        z1 = self.prod4(sa, sb) + [self.new()]                       # 9 limbs
        # (sa + ca 2^128)(sb + cb 2^128) = sa sb + (ca sb + cb sa) 2^128 + ca cb 2^256
        ma, mb = self.new(), self.new()
        self.emit("sub.u32 %s, 0, %s;" % (ma, ca))                   # all-ones mask when the carry is set
        self.emit("sub.u32 %s, 0, %s;" % (mb, cb))
        ta, tb = [self.new() for _ in range(4)], [self.new() for _ in range(4)]
        for j in range(4):
            self.emit("and.b32 %s, %s, %s;" % (ta[j], sb[j], ma))
            self.emit("and.b32 %s, %s, %s;" % (tb[j], sa[j], mb))
        cc = self.new()
        self.emit("and.b32 %s, %s, %s;" % (cc, ca, cb))
        self.emit("mov.u32 %s, %s;" % (z1[8], cc))
        top = z1[4:9]
        self.add_n(top, top, ta)
        self.add_n(top, top, tb)
        # middle term = z1 - z0 - z2 (>= 0, < 2^257)
        self.sub_n(z1, z1, z0)
        self.sub_n(z1, z1, z2)
        # T = z0 + mid 2^128 + z2 2^256
        T = z0 + z2
        self.add_n(T[4:16], T[4:16], z1)
        return T

    def red_row(self, even, odd, first):
        """One row of the Montgomery reduction of an 8-limb window (the CIOS row without its a*b_i part): the position-0
        limb is completed, m = limb * (-p^-1) mod 2^32, m * p is added with the shift of the odd array fused into the MADs."""
        mi = self.new()
        if first:
            self.emit("mul.lo.u32 %s, %s, %s;" % (mi, even[0], self.imm(self.inv)))
            for k in range(4):
                odd[2 * k], odd[2 * k + 1] = self.new(), self.new()
                self.emit("mul.lo.u32 %s, %s, %s;" % (odd[2 * k], mi, self.imm(self.p[2 * k + 1])))
                self.emit("mul.hi.u32 %s, %s, %s;" % (odd[2 * k + 1], mi, self.imm(self.p[2 * k + 1])))
        else:
            self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
            self.emit("mul.lo.u32 %s, %s, %s;" % (mi, even[0], self.imm(self.inv)))
            nodd = [None] * 8
            for k in range(3):
                nodd[2 * k], nodd[2 * k + 1] = self.new(), self.new()
                self.emit("madc.lo.cc.u32 %s, %s, %s, %s;" % (nodd[2 * k], mi, self.imm(self.p[2 * k + 1]), odd[2 * k + 2]))
                self.emit("madc.hi.cc.u32 %s, %s, %s, %s;" % (nodd[2 * k + 1], mi, self.imm(self.p[2 * k + 1]), odd[2 * k + 3]))
            nodd[6], nodd[7] = self.new(), self.new()
            self.emit("madc.lo.cc.u32 %s, %s, %s, 0;" % (nodd[6], mi, self.imm(self.p[7])))
            self.emit("madc.hi.u32 %s, %s, %s, 0;" % (nodd[7], mi, self.imm(self.p[7])))
            odd[:] = nodd
        self.cmad(even, [self.imm(self.p[j]) for j in (0, 2, 4, 6)], mi)
        self.emit("addc.u32 %s, %s, 0;" % (odd[7], odd[7]))

    def reduce16(self, T):
        """T: 16 registers (a product < p^2 or a sum of two) -> result registers r[0..7] = T 2^-256 mod p, in [0, 2p)"""
        even, odd = list(T[:8]), [None] * 8
        for i in range(0, 8, 2):
            self.red_row(even, odd, i == 0)
            self.red_row(odd, even, False)
        self.emit("add.cc.u32 %s, %s, %s;" % (even[0], even[0], odd[1]))
        for j in range(1, 7):
            self.emit("addc.cc.u32 %s, %s, %s;" % (even[j], even[j], odd[j + 1]))
        self.emit("addc.u32 %s, %s, 0;" % (even[7], even[7]))
        self.add_n(even, even, T[8:16])           # + the high half: U + T_hi < p + 1 + p^2 / 2^256 < 2p
        return even

    def mont_mul_k(self):
        A = ["%%%d" % (8 + i) for i in range(8)]
        B = ["%%%d" % (16 + i) for i in range(8)]
        r = self.reduce16(self.product_karatsuba(A, B))
        for j in range(8):
            self.emit("mov.u32 %%%d, %s;" % (j, r[j]))
        return self.lines

    def mont_mul_add2_k(self):
        A = ["%%%d" % (8 + i) for i in range(8)]
        B = ["%%%d" % (16 + i) for i in range(8)]
        C = ["%%%d" % (24 + i) for i in range(8)]
        D = ["%%%d" % (32 + i) for i in range(8)]
        T = self.product_karatsuba(A, B)
        U = self.product_karatsuba(C, D)
        self.add_n(T, T, U)                        # < 2 p^2 < 2^509
        r = self.reduce16(T)
        for j in range(8):
            self.emit("mov.u32 %%%d, %s;" % (j, r[j]))
        return self.lines


def run_ptx(lines, a, b, c=0, d=0):
    """Minimal interpreter for exactly the instruction forms emitted above."""
    regs = {}
    for i in range(8):
        regs["%%%d" % (8 + i)] = (a >> (32 * i)) & M32
        regs["%%%d" % (16 + i)] = (b >> (32 * i)) & M32
        regs["%%%d" % (24 + i)] = (c >> (32 * i)) & M32
        regs["%%%d" % (32 + i)] = (d >> (32 * i)) & M32
    cc = 0
    bw = 0       # borrow flag of sub.cc / subc chains (kept apart from the carry so that a mixed-up chain is caught)
    pending = 0  # a carry of 1 written by a .cc instruction that no instruction has consumed yet

    def val(x):
        x = x.strip()
        if x.startswith("0x"):
            return int(x, 16)
        if x.isdigit():
            return int(x)
        return regs[x]

    for ln in lines:
        m = re.match(r"([a-z0-9.]+)\s+(.*);", ln)
        op, args = m.group(1), [s.strip() for s in m.group(2).split(",")]
        d = args[0]
        assert not (bw and not op.startswith("sub")), "borrow pending across a non-sub instruction: " + ln
        consumes = op.startswith("madc") or op.startswith("addc")
        if consumes:
            pending = 0
        elif ".cc" in op or op in ("and.b32", "sub.u32"):
            # a fresh chain starts here: a carry of 1 nobody consumed would be lost -- the schedule's bounds forbid that
            assert pending == 0, "carry dropped before: " + ln
        if op == "mov.u32":
            regs[d] = val(args[1])
        elif op == "and.b32":
            regs[d] = val(args[1]) & val(args[2])
        elif op == "mul.lo.u32":
            regs[d] = (val(args[1]) * val(args[2])) & M32
        elif op == "mul.hi.u32":
            regs[d] = ((val(args[1]) * val(args[2])) >> 32) & M32
        elif op.startswith("mad"):
            base, part = op.split(".")[0], op.split(".")[1]
            prod = val(args[1]) * val(args[2])
            prod = (prod & M32) if part == "lo" else (prod >> 32)
            s = prod + val(args[3]) + (cc if base == "madc" else 0)
            regs[d] = s & M32
            if ".cc." in op:
                cc = s >> 32
                pending = cc
        elif op.startswith("add"):
            base = op.split(".")[0]
            s = val(args[1]) + val(args[2]) + (cc if base == "addc" else 0)
            regs[d] = s & M32
            if ".cc." in op:
                cc = s >> 32
                pending = cc
        elif op.startswith("sub"):
            # PTX sub.cc / subc: CC.CF holds the BORROW of the subtraction chain
            base = op.split(".")[0]
            s = val(args[1]) - val(args[2]) - (bw if base == "subc" else 0)
            regs[d] = s & M32
            if ".cc." in op:
                bw = 1 if s < 0 else 0
            elif base == "subc":
                assert s >= 0, "borrow out of the last limb: " + ln
                bw = 0
        else:
            raise ValueError(op)
        assert cc in (0, 1)
    assert pending == 0, "carry out of the last chain"
    return sum(regs["%%%d" % i] << (32 * i) for i in range(8))


def selfcheck_add2(P, lines, rounds=400):
    rinv = pow(1 << 256, -1, P)
    rng = random.Random(0xADD2 ^ (P & 0xFFFF))
    top = (1 << 254) - 1  # also operands above p (the schedule only needs < 2^254)
    cases = [(0, 0, 0, 0), (P - 1, P - 1, P - 1, P - 1), (P - 1, P - 1, 0, 0), (0, 0, P - 1, P - 1), (1, 1, P - 1, 1),
             (top, top, top, top), (P - 1, 1 << 253, P - 1, (1 << 253) + 12345)]
    cases += [tuple(rng.randrange(P) for _ in range(4)) for _ in range(rounds)]
    cases += [tuple(rng.choice((P - 1, P - 2, top, (1 << 224) - 1, 0xFFFFFFFF << 192 | 0xFFFFFFFF)) for _ in range(4)) for _ in range(100)]
    for a, b, c, d in cases:
        r = run_ptx(lines, a, b, c, d)
        assert r < 2 * P, "bound violated"
        assert r % P == ((a * b + c * d) * rinv) % P, (hex(a), hex(b), hex(c), hex(d), hex(r))


def selfcheck_sqr(P, lines, rounds=600):
    rinv = pow(1 << 256, -1, P)
    rng = random.Random(0x5152 ^ (P & 0xFFFF))
    top = (1 << 254) - 1
    cases = [0, 1, 2, P - 1, P - 2, top, (1 << 253), (1 << 253) - 1, 0x80000000, 0xFFFFFFFF, sum(0x80000000 << (32 * i) for i in range(7)),
             sum(0xFFFFFFFF << (32 * i) for i in range(7)) | (0x3FFFFFFF << 224), (1 << 256) % P]
    cases += [rng.randrange(P) for _ in range(rounds)]
    cases += [rng.randrange(1 << 254) & ~(rng.randrange(1 << 254)) for _ in range(100)]
    for a in cases:
        r = run_ptx(lines, a, 0)
        assert r < 2 * P, "bound violated"
        assert r % P == (a * a * rinv) % P, (hex(a), hex(r))


def selfcheck(P, lines, rounds=400):
    rinv = pow(1 << 256, -1, P)
    rng = random.Random(0xB200 ^ (P & 0xFFFF))
    cases = [(0, 0), (1, 1), (P - 1, P - 1), (P - 1, 1), (0, P - 1), ((1 << 256) % P, (1 << 256) % P)]
    cases += [(rng.randrange(P), rng.randrange(P)) for _ in range(rounds)]
    for a, b in cases:
        r = run_ptx(lines, a, b)
        assert r < 2 * P, "row bound violated"
        assert r % P == (a * b * rinv) % P, (hex(a), hex(b), hex(r))


def c_body_add2(name, lines, ntemps):
    out = []
    out.append("// GENERATED by tools/gen_mont_ptx.py -- do not edit. %s: r = (a*b + c*d)*2^-256 mod p, r in [0,2p)" % name)
    out.append("#define H2AGG_MONT_MUL_ADD2_%s(r, a, b, c, d) \\" % name)
    out.append('  asm("{\\n\\t.reg .u32 t<%d>;\\n\\t" \\' % ntemps)
    for ln in lines:
        out.append('      "%s\\n\\t" \\' % ln)
    out.append('      "}" \\')
    out.append('      : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]) \\')
    for nm, last in (("a", False), ("b", False), ("c", False), ("d", True)):
        out.append('      %s "r"((%s)[0]), "r"((%s)[1]), "r"((%s)[2]), "r"((%s)[3]), "r"((%s)[4]), "r"((%s)[5]), "r"((%s)[6]), "r"((%s)[7])%s \\' % (
            ":" if nm == "a" else " ", nm, nm, nm, nm, nm, nm, nm, nm, ")" if last else ","))
    text = "\n".join(out)
    return text[:-2] + "\n"  # drop the continuation after the closing parenthesis


def c_body_sqr(name, lines, ntemps):
    out = []
    out.append("// GENERATED by tools/gen_mont_ptx.py -- do not edit. %s: r = a*a*2^-256 mod p, r in [0,2p), a < 2^254" % name)
    out.append("#define H2AGG_MONT_SQR_%s(r, a) \\" % name)
    out.append('  asm("{\\n\\t.reg .u32 t<%d>;\\n\\t" \\' % ntemps)
    for ln in lines:
        out.append('      "%s\\n\\t" \\' % ln)
    out.append('      "}" \\')
    out.append('      : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]) \\')
    out.append('      : "r"((a)[0]), "r"((a)[1]), "r"((a)[2]), "r"((a)[3]), "r"((a)[4]), "r"((a)[5]), "r"((a)[6]), "r"((a)[7]))')
    return "\n".join(out) + "\n"


def c_body(name, lines, ntemps):
    out = []
    out.append("// GENERATED by tools/gen_mont_ptx.py -- do not edit. %s: r = a*b*2^-256 mod p, r in [0,2p)" % name)
    out.append("#define H2AGG_MONT_MUL_%s(r, a, b) \\" % name)
    out.append('  asm("{\\n\\t.reg .u32 t<%d>;\\n\\t" \\' % ntemps)
    for ln in lines:
        out.append('      "%s\\n\\t" \\' % ln)
    out.append('      "}" \\')
    out.append('      : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]), "=r"((r)[7]) \\')
    out.append('      : "r"((a)[0]), "r"((a)[1]), "r"((a)[2]), "r"((a)[3]), "r"((a)[4]), "r"((a)[5]), "r"((a)[6]), "r"((a)[7]), \\')
    out.append('        "r"((b)[0]), "r"((b)[1]), "r"((b)[2]), "r"((b)[3]), "r"((b)[4]), "r"((b)[5]), "r"((b)[6]), "r"((b)[7]))')
    return "\n".join(out) + "\n"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "gen", "mont_mul_bn254.inc")
    text = "#pragma once\n"
    for name, P in (("FR", FR), ("FQ", FQ)):
        g = Gen(P)
        lines = g.mont_mul()
        selfcheck(P, lines)
        text += c_body(name, lines, g.ntemps) + "\n"
        print("%s: %d PTX instructions, %d temps, self-check ok" % (name, len(lines), g.ntemps))
        g3 = Gen(P)
        lines3 = g3.mont_sqr()
        selfcheck_sqr(P, lines3)
        text += c_body_sqr(name, lines3, g3.ntemps) + "\n"
        nmul = sum(1 for ln in lines3 if ln.startswith(("mad.lo", "madc.lo", "mul.lo")))
        print("%s sqr: %d PTX instructions (%d multiplier pairs/singles), %d temps, self-check ok" % (name, len(lines3), nmul, g3.ntemps))
        g2 = Gen(P)
        lines2 = g2.mont_mul_add2()
        selfcheck_add2(P, lines2)
        text += c_body_add2(name, lines2, g2.ntemps) + "\n"
        print("%s add2: %d PTX instructions, %d temps, self-check ok" % (name, len(lines2), g2.ntemps))
        if "--with-karatsuba" not in sys.argv:
            continue
        # Measured on B200 and REJECTED: 120 instead of 136 multiplier instructions per product, but the
        # ~100 extra additions (IADD3.X / IMAD.X) cost more issue slots than the 16 wide MADs save:
        # msm_accumulate 8.09 -> 9.82 ms, coset NTT 3.39 -> 3.74 ms.  Kept behind this flag (+ -DH2AGG_KARATSUBA) so the
        # experiment can be repeated.
        gk = Gen(P)
        lk = gk.mont_mul_k()
        selfcheck(P, lk, rounds=1500)
        text += c_body(name + "_K", lk, gk.ntemps) + "\n"
        nm = sum(1 for ln in lk if ln.startswith(("mad.lo", "madc.lo", "mul.lo")))
        print("%s karatsuba: %d PTX instructions (%d multiplier pairs/singles), %d temps, self-check ok" % (name, len(lk), nm, gk.ntemps))
        gk2 = Gen(P)
        lk2 = gk2.mont_mul_add2_k()
        selfcheck_add2(P, lk2, rounds=1000)
        text += c_body_add2(name + "_K", lk2, gk2.ntemps) + "\n"
        nm = sum(1 for ln in lk2 if ln.startswith(("mad.lo", "madc.lo", "mul.lo")))
        print("%s karatsuba add2: %d PTX instructions (%d multiplier pairs/singles), %d temps, self-check ok" % (name, len(lk2), nm, gk2.ntemps))
    with open(dst, "w") as f:
        f.write(text)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
