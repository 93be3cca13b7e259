"""T1: the outer proof's transcript, `ShaWrite<_, G1Affine, Challenge255<_>, sha2::Sha256>`
(halo2-snark-aggregator-api/src/transcript/sha.rs:130-232, instantiated at
halo2-snark-aggregator-circuit/src/verify_circuit.rs:985).  Format contract only -- it defines the
bytes our commitments end up in:
  common_point   absorbs [0u8;31] || 0x01 || BE32(x) || BE32(y)        (:199-217; identity is an error)
  common_scalar  absorbs [0u8;31] || 0x02 || BE32(s)                   (:219-231)
  write_point    common_point, then emits LE32(x) || LE32(y)           (:156-173)
  write_scalar   common_scalar, then emits LE32(s)                     (:175-180)
  squeeze_challenge  absorbs 0x00, finalises a CLONE, re-seeds the state with the 32-byte digest,
                 challenge = digest as a little-endian integer (zero-extended to 64 bytes) mod r  (:186-197)
Points come straight from the MSM output (`h2agg_msm_g1*`: normalised Jacobian whose first 64 bytes
are affine Montgomery limbs); the Montgomery -> canonical conversion is host-side setup work on
64 bytes per commitment."""
import hashlib

_P = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
_R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_RINV_P = pow(1 << 256, -1, _P)
_RINV_R = pow(1 << 256, -1, _R)


def _canon(limbs4, mod, rinv):
    v = sum(int(l) << (64 * i) for i, l in enumerate(limbs4))
    return v * rinv % mod


class ShaWrite:
    def __init__(self):
        self.state = hashlib.sha256()
        self.out = bytearray()

    # -- raw canonical integers
    def common_point_xy(self, x, y):
        self.state.update(bytes(31) + b"\x01" + x.to_bytes(32, "big") + y.to_bytes(32, "big"))

    def common_scalar_int(self, s):
        self.state.update(bytes(31) + b"\x02" + s.to_bytes(32, "big"))

    # -- Montgomery limb inputs as they come out of the library
    @staticmethod
    def _affine(jac12_or_aff8):
        limbs = [int(v) for v in jac12_or_aff8]
        if (len(limbs) == 12 and not any(limbs[8:12])) or (len(limbs) == 8 and not any(limbs)):
            raise IOError("cannot write points at infinity to the transcript")
        return _canon(limbs[0:4], _P, _RINV_P), _canon(limbs[4:8], _P, _RINV_P)

    def common_point(self, jac12_or_aff8):
        """absorb without emitting (instance commitments: the verifier recomputes them)"""
        self.common_point_xy(*self._affine(jac12_or_aff8))

    def write_point(self, jac12_or_aff8):
        """jac12: the 12-limb normalised Jacobian `h2agg_msm_g1` returns (or 8 affine limbs)."""
        x, y = self._affine(jac12_or_aff8)
        self.common_point_xy(x, y)
        self.out += x.to_bytes(32, "little") + y.to_bytes(32, "little")

    def write_scalar(self, fr4):
        s = _canon(fr4, _R, _RINV_R)
        self.common_scalar_int(s)
        self.out += s.to_bytes(32, "little")

    def squeeze_challenge(self):
        self.state.update(b"\x00")
        digest = self.state.copy().digest()
        self.state = hashlib.sha256()
        self.state.update(digest)
        return int.from_bytes(digest + bytes(32), "little") % _R

    def finalize(self):
        return bytes(self.out)
