"""N4: the reference's stage files and the instance encoding of the aggregation proof (host side).

Mirrors halo2-snark-aggregator-circuit/src/fs.rs (file names and byte layouts that the reference itself defines) and
verify_circuit.rs:768-804 `final_pair_to_instances`.  Files whose content is defined by the external crate
(`ParamsKZG::write`, `VerifyingKey::write`) are only named here, not parsed.

    verify_circuit_instance.data    scalars, each `to_repr()` = 32 bytes little-endian canonical      fs.rs:169-180
    verify_circuit_final_pair.data  w_x.x, w_x.y, w_g.x, w_g.y reprs, then the instance scalars        fs.rs:182-197
    verify_circuit_proof.data       the transcript bytes (ShaWrite, transcript.py)                     fs.rs:199-201
Scalars and coordinates are canonical Python integers at this level; `Context.fr_to_repr / fr_from_repr`
(h2agg_fr_repr) convert whole vectors of Montgomery limbs on the device.
"""
import os

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
P_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
LIMBS, LIMB_WIDTH = 4, 68          # FiveColumnIntegerChipHelper (halo2-ecc-circuit-lib/src/five/integer_chip.rs)


def read_file(folder, filename):
    with open(os.path.join(folder, filename), "rb") as f:
        return f.read()


def write_file(folder, filename, buf):
    with open(os.path.join(folder, filename), "wb") as f:
        f.write(buf)


# ---- names (fs.rs:40-131) -------------------------------------------------------------------------------------------
def target_circuit_params_name(params_name):
    return "sample_circuit_%s.params" % params_name


def target_circuit_vk_name(params_name):
    return "sample_circuit_%s.vkey" % params_name


def target_circuit_instance_name(name, index):
    return "sample_circuit_instance_%s%d.data" % (name, index)


def target_circuit_proof_name(name, index):
    return "sample_circuit_proof_%s%d.data" % (name, index)


VERIFY_CIRCUIT_PARAMS = "verify_circuit.params"
VERIFY_CIRCUIT_VK = "verify_circuit.vkey"
VERIFY_CIRCUIT_INSTANCE = "verify_circuit_instance.data"
VERIFY_CIRCUIT_FINAL_PAIR = "verify_circuit_final_pair.data"
VERIFY_CIRCUIT_PROOF = "verify_circuit_proof.data"
VERIFIER_SOL = "verifier.sol"


# ---- scalars ----------------------------------------------------------------------------------------------------------
def to_repr(v):
    return int(v).to_bytes(32, "little")


def load_instances(buf):
    """fs.rs:134-146: 32-byte chunks until the data run out (a trailing partial chunk is ignored, as read_exact fails),
    each through from_repr(..).unwrap() -> ValueError for a non-canonical value; shaped vec![vec![ret]]."""
    ret = []
    for off in range(0, len(buf) - len(buf) % 32, 32):
        v = int.from_bytes(buf[off:off + 32], "little")
        if v >= R_MOD:
            raise ValueError("from_repr: scalar at offset %d is not canonical" % off)
        ret.append(v)
    return [[ret]]


def write_verify_circuit_instance(folder, scalars):
    write_file(folder, VERIFY_CIRCUIT_INSTANCE, b"".join(to_repr(s) for s in scalars))


def load_verify_circuit_instance(folder):
    return load_instances(read_file(folder, VERIFY_CIRCUIT_INSTANCE))


def write_verify_circuit_final_pair(folder, pair):
    """pair = ((w_x.x, w_x.y), (w_g.x, w_g.y), [instance scalars])"""
    (wxx, wxy), (wgx, wgy), inst = pair
    write_file(folder, VERIFY_CIRCUIT_FINAL_PAIR, b"".join(to_repr(v) for v in (wxx, wxy, wgx, wgy)) + b"".join(to_repr(s) for s in inst))


def write_verify_circuit_proof(folder, buf):
    write_file(folder, VERIFY_CIRCUIT_PROOF, bytes(buf))


def load_verify_circuit_proof(folder):
    return read_file(folder, VERIFY_CIRCUIT_PROOF)


# ---- verify_circuit.rs:768-804 ------------------------------------------------------------------------------------------
def w_to_limb_n_le(w):
    """IntegerChipHelper::w_to_limb_n_le: the four 68-bit limbs of a base-field element, little-endian"""
    w = int(w)
    return [(w >> (LIMB_WIDTH * i)) & ((1 << LIMB_WIDTH) - 1) for i in range(LIMBS)]


def final_pair_to_instances(pair):
    """The public inputs of the aggregation circuit: per point, x packed as two 136-bit halves, the parity of y folded
    into the second half at bit 136; then the forwarded instances."""
    (wxx, wxy), (wgx, wgy), inst = pair
    e = [pow(2, LIMB_WIDTH * i, R_MOD) for i in range(LIMBS)]      # limb_modulus_exps
    out = []
    for x, y in ((wxx, wxy), (wgx, wgy)):
        lx, ly = w_to_limb_n_le(x), w_to_limb_n_le(y)
        last_bit = e[2] if ly[0] & 1 else 0
        out.append((lx[0] * e[0] + lx[1] * e[1]) % R_MOD)
        out.append((lx[2] * e[0] + lx[3] * e[1] + last_bit) % R_MOD)
    return out + [int(s) % R_MOD for s in inst]
