"""T1 format contract: ShaWrite bytes (halo2-snark-aggregator-api/src/transcript/sha.rs:130-232)."""
import hashlib

import numpy as np
import pytest

import bn254_ref as ref
from halo2_snark_aggregator_b200.transcript import ShaWrite


def test_sha_write_bytes_and_challenges():
    t = ShaWrite()
    g = ref.G1_GEN
    p2 = ref.g1_add(g, g)
    jac = np.array(ref.pack_points([p2]) + ref.to_mont_limbs(1, ref.P), dtype=np.uint64)
    t.write_point(jac)
    s = 0x1234567890ABCDEF << 100
    t.write_scalar(np.array(ref.to_mont_limbs(s, ref.R), dtype=np.uint64))
    c1 = t.squeeze_challenge()
    t.write_point(np.array(ref.pack_points([g]), dtype=np.uint64))
    c2 = t.squeeze_challenge()
    # independent assembly of the same byte stream
    h = hashlib.sha256()
    h.update(bytes(31) + b"\x01" + p2[0].to_bytes(32, "big") + p2[1].to_bytes(32, "big"))
    h.update(bytes(31) + b"\x02" + s.to_bytes(32, "big"))
    h.update(b"\x00")
    d1 = h.digest()
    assert c1 == int.from_bytes(d1, "little") % ref.R
    h = hashlib.sha256()
    h.update(d1)
    h.update(bytes(31) + b"\x01" + (1).to_bytes(32, "big") + (2).to_bytes(32, "big"))
    h.update(b"\x00")
    assert c2 == int.from_bytes(h.digest(), "little") % ref.R
    out = t.finalize()
    assert out == p2[0].to_bytes(32, "little") + p2[1].to_bytes(32, "little") + s.to_bytes(32, "little") + (1).to_bytes(32, "little") + (2).to_bytes(32, "little")
    assert len(out) == 64 + 32 + 64


def test_identity_is_rejected_like_the_reference():
    t = ShaWrite()
    ident = np.array([0] * 4 + ref.to_mont_limbs(1, ref.P) + [0] * 4, dtype=np.uint64)
    with pytest.raises(IOError):
        t.write_point(ident)
