#!/usr/bin/env python3
"""Mint the golden vectors in this directory from the independent Python big-int reference
(oracle/py/bn254_ref.py).  The reference repository holds no vectors for this path,
so these are derived from the mathematical definition of its call-site semantics.
Run:  python tests/golden/make_golden.py     (deterministic; rewrites *.json)"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle", "py"))
import bn254_ref as ref  # noqa: E402


def hexs(limbs):
    return ["%016x" % x for x in limbs]


def msm_cases():
    rng = random.Random(0x4D534D)
    cases = []
    G = ref.G1_GEN
    pts = [ref.g1_mul(rng.randrange(1, ref.R), G) for _ in range(24)]

    def case(name, scalars, points):
        res = ref.msm(scalars, points)
        cases.append({"name": name, "scalars": hexs(ref.pack_fr(scalars)), "bases": hexs(ref.pack_points(points)),
                      "affine": hexs(ref.pack_points([res]))})

    case("single_one", [1], [G])
    case("single_zero", [0], [G])
    case("generator_times_r_minus_1", [ref.R - 1], [G])
    case("two_cancel", [5, ref.R - 5], [G, G])
    case("same_point_many", [3, 4, 5, 6, 7, 8, 9], [pts[0]] * 7)
    case("p_and_minus_p", [7, 7], [pts[1], ref.g1_neg(pts[1])])
    case("identity_bases", [rng.randrange(ref.R) for _ in range(4)], [None, pts[2], None, pts[3]])
    case("all_zero_scalars", [0] * 8, pts[:8])
    case("all_r_minus_1", [ref.R - 1] * 8, pts[:8])
    case("small_17bit", [rng.randrange(1 << 17) for _ in range(16)], pts[:16])
    case("booleans", [rng.randrange(2) for _ in range(24)], pts[:24])
    case("uniform_24", [rng.randrange(ref.R) for _ in range(24)], pts[:24])
    case("powers_of_two", [1 << (11 * i) for i in range(23)], pts[:23])
    case("window_boundaries", [(1 << 16) - 1, 1 << 15, (1 << 15) + 1, (1 << 254) % ref.R, (1 << 253) + 12345, (1 << 16)], pts[:6])
    for n in (2, 3, 5, 31, 33):
        # synthetic generator streams (pins oracle_gen_* and the device generator too)
        sc = [ref.gen_scalar(0xA660000 + n, 0, i) for i in range(n)]
        bs = [ref.gen_base(0x53525300 + n, i) for i in range(n)]
        case("synth_%d" % n, sc, bs)
    return cases


def ntt_cases():
    rng = random.Random(0x4E5454)
    cases = []
    for k in range(0, 7):
        n = 1 << k
        a = [rng.randrange(ref.R) for _ in range(n)]
        w = ref.omega(k)
        cases.append({"name": "fft_k%d" % k, "k": k, "omega": hexs(ref.to_mont_limbs(w, ref.R)),
                      "input": hexs(ref.pack_fr(a)), "output": hexs(ref.pack_fr(ref.dft(a, w)))})
    for k in range(1, 6):
        n = 1 << k
        a = [rng.randrange(ref.R) for _ in range(n)]
        cases.append({"name": "ifft_k%d" % k, "k": k,
                      "omega_inv": hexs(ref.to_mont_limbs(pow(ref.omega(k), -1, ref.R), ref.R)),
                      "n_inv": hexs(ref.to_mont_limbs(pow(n, -1, ref.R), ref.R)),
                      "input": hexs(ref.pack_fr(a)), "output": hexs(ref.pack_fr(ref.ifft(a, k)))})
    for k in range(1, 5):
        ext_k = k + 2
        a = [rng.randrange(ref.R) for _ in range(1 << k)]
        ext = ref.coeff_to_extended(a, k, ext_k)
        cases.append({"name": "coeff_to_extended_k%d" % k, "k": k, "ext_k": ext_k,
                      "zeta": hexs(ref.to_mont_limbs(ref.ZETA, ref.R)),
                      "omega_ext": hexs(ref.to_mont_limbs(ref.omega(ext_k), ref.R)),
                      "input": hexs(ref.pack_fr(a)), "output": hexs(ref.pack_fr(ext))})
        b = [rng.randrange(ref.R) for _ in range(1 << ext_k)]
        out_len = 3 << k
        cases.append({"name": "extended_to_coeff_k%d" % k, "k": k, "ext_k": ext_k, "out_len": out_len,
                      "zeta": hexs(ref.to_mont_limbs(ref.ZETA, ref.R)),
                      "omega_ext_inv": hexs(ref.to_mont_limbs(pow(ref.omega(ext_k), -1, ref.R), ref.R)),
                      "ext_n_inv": hexs(ref.to_mont_limbs(pow(1 << ext_k, -1, ref.R), ref.R)),
                      "input": hexs(ref.pack_fr(b)), "output": hexs(ref.pack_fr(ref.extended_to_coeff(b, ext_k, out_len)))})
    return cases


def field_cases():
    rng = random.Random(0x464C44)
    out = []
    for field, mod in ((0, ref.R), (1, ref.P)):
        vals = [0, 1, mod - 1, mod - 2, 2, (1 << 256) % mod, (1 << 128) - 1] + [rng.randrange(mod) for _ in range(25)]
        a = vals
        b = list(reversed(vals))
        out.append({"field": field,
                    "a": hexs(sum((ref.to_mont_limbs(x, mod) for x in a), [])),
                    "b": hexs(sum((ref.to_mont_limbs(x, mod) for x in b), [])),
                    "add": hexs(sum((ref.to_mont_limbs(x + y, mod) for x, y in zip(a, b)), [])),
                    "sub": hexs(sum((ref.to_mont_limbs(x - y, mod) for x, y in zip(a, b)), [])),
                    "mul": hexs(sum((ref.to_mont_limbs(x * y, mod) for x, y in zip(a, b)), [])),
                    "inv": hexs(sum((ref.to_mont_limbs(pow(x, -1, mod) if x else 0, mod) for x in a), []))})
    return out


def main():
    for name, fn in (("msm.json", msm_cases), ("ntt.json", ntt_cases), ("field.json", field_cases)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(), f, indent=0)
        print("wrote", name)


if __name__ == "__main__":
    main()
