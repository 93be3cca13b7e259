"""N4 host formats (halo2_snark_aggregator_b200/fs.py) against the byte layouts the reference defines in
halo2-snark-aggregator-circuit/src/fs.rs and verify_circuit.rs:768-804."""
import pytest

from halo2_snark_aggregator_b200 import fs


def test_instance_file_round_trip(tmp_path):
    vals = [0, 1, fs.R_MOD - 1, 0x1234567890ABCDEF << 130]
    fs.write_verify_circuit_instance(str(tmp_path), vals)
    raw = fs.read_file(str(tmp_path), "verify_circuit_instance.data")
    assert len(raw) == 128 and raw[32:64] == b"\x01" + bytes(31)          # to_repr is little-endian
    assert fs.load_verify_circuit_instance(str(tmp_path)) == [[vals]]
    assert fs.load_instances(raw + b"\x07" * 5) == [[vals]]                 # read_exact fails on the partial tail
    with pytest.raises(ValueError):
        fs.load_instances(fs.to_repr(fs.R_MOD))                             # from_repr(..).unwrap()


def test_final_pair_file_and_instances(tmp_path):
    x1, y1 = (1 << 253) + 12345, 2 * 999 + 1          # odd y
    x2, y2 = 0xABCDEF << 200, 2 * 777                 # even y
    pair = ((x1, y1), (x2, y2), [5, 6])
    fs.write_verify_circuit_final_pair(str(tmp_path), pair)
    raw = fs.read_file(str(tmp_path), "verify_circuit_final_pair.data")
    assert len(raw) == 6 * 32
    assert [int.from_bytes(raw[i:i + 32], "little") for i in range(0, 192, 32)] == [x1, y1, x2, y2, 5, 6]
    inst = fs.final_pair_to_instances(pair)
    m136 = (1 << 136) - 1
    assert inst == [x1 & m136, (x1 >> 136) + (1 << 136), x2 & m136, x2 >> 136, 5, 6]
    assert fs.w_to_limb_n_le(x1) == [(x1 >> (68 * i)) & ((1 << 68) - 1) for i in range(4)]


def test_names():
    assert fs.target_circuit_instance_name("simple", 3) == "sample_circuit_instance_simple3.data"
    assert fs.target_circuit_proof_name("simple", 0) == "sample_circuit_proof_simple0.data"
    assert fs.target_circuit_params_name("k8") == "sample_circuit_k8.params"
