import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
M64 = (1 << 64) - 1
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23


def golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def arr(hex_list):
    return np.array([int(x, 16) for x in hex_list], dtype=np.uint64)


def fr_limbs(x):
    v = (x % R_MOD) * (1 << 256) % R_MOD
    return np.array([(v >> (64 * i)) & M64 for i in range(4)], dtype=np.uint64)


def fq_limbs(x):
    v = (x % P_MOD) * (1 << 256) % P_MOD
    return np.array([(v >> (64 * i)) & M64 for i in range(4)], dtype=np.uint64)


def omega(k):
    return pow(ROOT_OF_UNITY, 1 << (28 - k), R_MOD)


def domain_consts(k, ext_k=None):
    ext_k = k + 2 if ext_k is None else ext_k
    w, we = omega(k), omega(ext_k)
    return dict(
        omega=fr_limbs(w), omega_inv=fr_limbs(pow(w, -1, R_MOD)), n_inv=fr_limbs(pow(1 << k, -1, R_MOD)),
        omega_ext=fr_limbs(we), omega_ext_inv=fr_limbs(pow(we, -1, R_MOD)),
        ext_n_inv=fr_limbs(pow(1 << ext_k, -1, R_MOD)), zeta=fr_limbs(ZETA))


def affine_of(jac12):
    """first 64 bytes of a NORMALISED jacobian point (z = 1) / zeros for the identity"""
    j = np.asarray(jac12, dtype=np.uint64)
    if not j[8:12].any():
        return np.zeros(8, dtype=np.uint64)
    return j[:8].copy()
