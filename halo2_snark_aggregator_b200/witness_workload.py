"""A synthetic chip-op stream with the SHAPE of the reference's aggregation witness, for measuring the witness path.

`verify_aggregation_proofs_in_chip` (halo2-snark-aggregator-api/src/systems/halo2/verify.rs:835-942) emits, per inner
proof: a Poseidon transcript (T = 9: ~27 permutations of 4 + 63 + 4 rounds, api/src/hash/poseidon.rs:196-230) fed by ~22
transcript points and ~71 scalars, a few hundred ScalarChip operations for the gate / permutation / lookup expressions,
one scalar_mul_constant per instance value; and once per aggregation two multi_exps over ~23 points per proof plus the
final-pair packing.  This module replays that op mix through the product's chips (witness.py) with arbitrary values --
constants and points are random, only the ROW LAYOUT and the arithmetic per row are those of the real thing (the real
stream, driven by the restated verifier, is what tests/test_aggregation_cpu.py and tests/test_gpu_witness.py check bit for
bit).  Used by bench.py's `witness` object."""
import random

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
T, R_F, R_P = 9, 8, 63


class _PoseidonShape:
    """PoseidonChip's op pattern (api/src/hash/poseidon.rs) with random constants: same calls, same rows."""

    def __init__(self, chip, rng):
        self.chip, self.rng = chip, rng
        self.s = [chip.assign_const(rng.randrange(R_MOD)) for _ in range(T)]
        self.absorbing = []

    def c(self):
        return self.rng.randrange(R_MOD)

    def update(self, elems):
        self.absorbing += list(elems)

    def _x5(self, x):
        x2 = self.chip.mul(x, x)
        x4 = self.chip.mul(x2, x2)
        return self.chip.mul_add_constant(x, x4, self.c())

    def _full(self):
        self.s = [self._x5(x) for x in self.s]
        self.s = [self.chip.sum_with_coeff_and_constant([(x, self.c()) for x in self.s], 0) for _ in range(T)]

    def _partial(self):
        self.s[0] = self._x5(self.s[0])
        res = [self.chip.sum_with_coeff_and_constant([(x, self.c()) for x in self.s], 0)]
        for x in self.s[1:]:
            res.append(self.chip.sum_with_coeff_and_constant([(self.s[0], self.c()), (x, 1)], 0))
        self.s = res

    def _permutation(self, inputs):
        s = self.s
        s[0] = self.chip.sum_with_constant([s[0]], self.c())
        for i, v in enumerate(inputs):
            s[i + 1] = self.chip.sum_with_constant([s[i + 1], v], self.c())
        for i in range(len(inputs) + 1, T):
            s[i] = self.chip.sum_with_constant([s[i]], self.c())
        for _ in range(R_F // 2):
            self._full()
        for _ in range(R_P):
            self._partial()
        for _ in range(R_F // 2):
            self._full()

    def squeeze(self):
        inputs, self.absorbing = self.absorbing, []
        pad = 0
        for i in range(0, len(inputs), T - 1):
            chunk = inputs[i:i + T - 1]
            pad = T - 1 - len(chunk)
            self._permutation(chunk)
        if pad == 0:
            self._permutation([])
        return self.s[1]


def record_aggregation_like(schip, pchip, encode, points, n_proofs, seed=1, points_per_proof=23, scalars_per_proof=71,
                            field_ops_per_proof=600, instances_per_proof=2):
    """points: callable i -> affine Montgomery limbs (8 x u64) of a curve point.  Returns the final-pair cells."""
    rng = random.Random(seed)
    all_pts, all_scs = [], []
    for p in range(n_proofs):
        tr = _PoseidonShape(schip, rng)
        inst = [schip.assign_var(rng.randrange(R_MOD)) for _ in range(instances_per_proof)]
        acc = None
        for i, v in enumerate(inst):                       # assign_instance_commitment: scalar_mul_constant + add
            ls = pchip.scalar_mul_constant(v, points(1000 * p + i))
            acc = ls if acc is None else pchip.add(acc, ls)
        pts = [pchip.normalize(acc)]
        tr.update(encode.encode_point(pts[0]))
        scs = []
        for i in range(points_per_proof - 1):              # read_point: assign_var (on-curve check) + encode
            pt = pchip.assign_var(points(1000 * p + 100 + i))
            tr.update(encode.encode_point(pt))
            pts.append(pt)
            if i % 5 == 4:
                scs.append(tr.squeeze())                   # challenges in between
        for i in range(scalars_per_proof):                 # read_scalar
            s = schip.assign_var(rng.randrange(R_MOD))
            tr.update([s])
            scs.append(s)
        scs.append(tr.squeeze())
        acc_s = scs[0]
        for i in range(field_ops_per_proof):               # expression evaluation: the mul / add / sub mix of params.rs
            o = scs[(7 * i + 3) % len(scs)]
            acc_s = schip.mul(acc_s, o) if i % 3 == 0 else (schip.add(acc_s, o) if i % 3 == 1 else schip.sub(acc_s, o))
        all_pts += pts
        all_scs += [schip.mul(acc_s, scs[i % len(scs)]) for i in range(len(pts))]
    w_g = pchip.multi_exp(all_pts, all_scs)
    w_x = pchip.multi_exp(all_pts[:4], all_scs[:4])
    return pchip.expose_final_pair(w_x, w_g)
